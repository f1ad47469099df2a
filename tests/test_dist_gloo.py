"""CPU, world_size 2, gloo: the communication schedule of the data-parallel head (stil_tta_b200/distributed.py)
reproduces the single-process oracle on the concatenated batch.  The per-rank compute is injected (plain torch
restatement of what stil_infonce_fwd/bwd return) so that only the schedule is under test here; the CUDA compute
is covered by tests/test_gpu_parity.py and the 2-GPU test in tests/test_gpu_dist.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stil_head_oracle as O

T, LAM = 0.1, 0.3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _norm(x):
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def fwd_local(a_all, b_all, off, m):
    """What stil_infonce_fwd computes for the local rows [off, off+m) (include/stil_head.h)."""
    n = a_all.shape[0]
    an, bn = _norm(a_all.double()), _norm(b_all.double())
    L = an[off:off + m] @ bn.t() / T            # local rows of the logits
    Lc = bn[off:off + m] @ an.t() / T           # local columns (rows of the transpose)
    d = L[torch.arange(m), off + torch.arange(m)]
    lse_row, lse_col = torch.logsumexp(L, 1), torch.logsumexp(Lc, 1)
    loss = (LAM * (lse_row - d) + (1 - LAM) * (lse_col - d)).sum() / n
    return loss.reshape(1), lse_row, lse_col


def bwd_local(a_all, b_all, off, m, lse_row_all, lse_col_all):
    """What stil_infonce_bwd computes: d(global loss)/d(local rows of a and b)."""
    n = a_all.shape[0]
    a, b = a_all.double(), b_all.double()
    an, bn = _norm(a), _norm(b)
    eye = torch.zeros(m, n, dtype=torch.float64)
    eye[torch.arange(m), off + torch.arange(m)] = 1
    L = an[off:off + m] @ bn.t() / T
    G = (LAM * torch.exp(L - lse_row_all[off:off + m, None]) + (1 - LAM) * torch.exp(L - lse_col_all[None, :]) - eye) / n
    Lc = bn[off:off + m] @ an.t() / T
    Gc = ((1 - LAM) * torch.exp(Lc - lse_col_all[off:off + m, None]) + LAM * torch.exp(Lc - lse_row_all[None, :]) - eye) / n
    ga, gb = G @ bn / T, Gc @ an / T

    def through_norm(g, x):
        xn = _norm(x)
        return (g - xn * (xn * g).sum(1, keepdim=True)) / x.norm(dim=1, keepdim=True)
    return through_norm(ga, a[off:off + m]), through_norm(gb, b[off:off + m])


def _worker(rank, world, port, m, d, k):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stil_tta_b200.distributed import GlobalBatch, all_reduce_prototype_partials, sharded_infonce
        g = torch.Generator().manual_seed(0)
        a_full, b_full = torch.randn(world * m, d, generator=g), torch.randn(world * m, d, generator=g)
        a_loc, b_loc = a_full[rank * m:(rank + 1) * m], b_full[rank * m:(rank + 1) * m]
        gb = GlobalBatch()
        assert gb.world_size == world and gb.row_offset(m) == rank * m and gb.total_rows(m) == world * m
        loss, d_a, d_b = sharded_infonce(a_loc, b_loc, T, LAM, gb, fwd_local, bwd_local)
        # oracle: reference CLIPLoss on the concatenated batch in one process
        ar, br = a_full.double().requires_grad_(True), b_full.double().requires_grad_(True)
        loss_ref, _, _ = O.clip_loss_global([ar], [br], T, LAM)
        ga, gb_ = torch.autograd.grad(loss_ref, (ar, br))
        torch.testing.assert_close(loss.double().reshape(()), loss_ref.detach(), rtol=1e-9, atol=1e-12)
        torch.testing.assert_close(d_a, ga[rank * m:(rank + 1) * m], rtol=1e-8, atol=1e-12)
        torch.testing.assert_close(d_b, gb_[rank * m:(rank + 1) * m], rtol=1e-8, atol=1e-12)
        # prototype partials: one packed all-reduce == the reference's two (STiLModel.py:377-379)
        gs = torch.Generator().manual_seed(10 + rank)
        cs, cc = torch.randn(k, d, generator=gs), torch.rand(k, 1, generator=gs)
        rs, rc = all_reduce_prototype_partials(cs, cc)
        exp_s, exp_c = cs.clone(), cc.clone()
        dist.all_reduce(exp_s); dist.all_reduce(exp_c)
        assert torch.equal(rs, exp_s) and torch.equal(rc, exp_c)
        # sharded prototype accumulation == single process on the concatenated rows
        gl = torch.Generator().manual_seed(99)
        label = torch.softmax(torch.randn(world * m, k, generator=gl) * 5, 1)
        feat = torch.randn(world * m, d, generator=gl)
        loc = slice(rank * m, (rank + 1) * m)
        ps, pc = O.cal_prototypes(label[loc], feat[loc], 0.5)
        ps, pc = all_reduce_prototype_partials(ps, pc)
        fs, fc = O.cal_prototypes(label, feat, 0.5)
        torch.testing.assert_close(ps, fs, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(pc, fc)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,m,d", [(2, 24, 16), (2, 7, 8)])
def test_sharded_schedule_matches_single_process(world, m, d):
    mp.spawn(_worker, args=(world, _free_port(), m, d, 5), nprocs=world, join=True)


# ----------------------------------------------------------------------------------------------- column-sharded bank (C5)
def _bank_worker(rank, world, port, rows, kb, d, c):
    """The schedule of stil_tta_b200.ShardedSimMatchBank (gather, additive fixed-shift statistics, all-reduce,
    reduce-scatter of the gradient partials, bank writes to the owning shard) on CPU ranks, with the three shard kernels
    replaced by their torch restatement — against the oracle on the WHOLE bank in one process."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import bank_oracle as BO
        from stil_tta_b200.bank import ShardedSimMatchBank
        import contextlib

        class CpuSharded(ShardedSimMatchBank):
            def _check_device(self):          # test double: CPU ranks, torch kernels
                pass

            def _device_guard(self):
                return contextlib.nullcontext()

            def _alloc_ws(self, *a):
                pass

            def _k_stats(self, s, fk, fq, p, tt, st, out):
                out.copy_(BO.simmatch_shard_stats(fk, fq, p, self.bank[s], self.labels[s], tt, st))

            def _k_finish(self, stats, p, st, c_smooth, prob_all, loss_all, norms):
                a, b, n = BO.simmatch_shard_finish(stats, p, st, c_smooth)
                prob_all.copy_(a); loss_all.copy_(b); norms.copy_(n)

            def _k_grad(self, s, fk, fq, p, tt, st, norms, out):
                out.copy_(BO.simmatch_shard_grad(fk, fq, p, self.bank[s], self.labels[s], tt, st, norms))

            def _k_update(self, s, k, y, index_local):
                BO.update_bank(self.bank[s], self.labels[s], k, y, index_local)

        import stil_tta_b200.bank as bank_mod
        bank_mod.alloc_bank = lambda dim, k, dtype, device: torch.zeros(dim, k, dtype=dtype)   # CPU buffers for the test double
        g = torch.Generator().manual_seed(5)
        unit = torch.nn.functional.normalize
        bank_rows = unit(torch.randn(kb, d, generator=g))
        labels = torch.randint(0, c, (kb,), generator=g)
        fk_all = unit(bank_rows[torch.randint(0, kb, (world * rows,), generator=g)] + 0.3 * torch.randn(world * rows, d, generator=g))
        fq_all = unit(fk_all + 0.2 * torch.randn(world * rows, d, generator=g))
        p_all = torch.softmax(torch.randn(world * rows, c, generator=g) * 3, 1)
        sb = CpuSharded(d, kb, c, dtype=torch.float32, device="cpu")
        assert sb.world == world and sb.k_shard == kb // world
        sb.load(bank_rows, labels)
        loc = slice(rank * rows, (rank + 1) * rows)
        fq = fq_all[loc].clone().requires_grad_(True)
        prob_ku, loss_in = sb(fk_all[loc], fq, p_all[loc], 0.1, 0.1, 0.9)
        # _update_bank between forward and backward (the reference order): bank writes go to the owning shard
        idx = torch.arange(rank, kb, 7)[:8]
        newk = unit(torch.randn(len(idx), d, generator=torch.Generator().manual_seed(rank)))
        sb.update(newk, torch.full((len(idx),), 1, dtype=torch.int64), idx)
        (gq,) = torch.autograd.grad(loss_in.mean(), fq)
        fqr = fq_all.clone().requires_grad_(True)
        ref = O.simmatch_bank(fk_all, fqr, p_all, bank_rows, labels, 0.1, 0.1, 0.9)
        (g_ref,) = torch.autograd.grad(ref["loss_in"][loc].mean(), fqr)
        torch.testing.assert_close(prob_ku, ref["prob_ku"][loc], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(loss_in, ref["loss_in"][loc].detach(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(gq, g_ref[loc], rtol=1e-3, atol=1e-6)
        # every rank's update landed in the shard that owns the column
        full = torch.cat([t.clone() for t in [sb.bank[0]]], dim=1)
        gathered = [torch.zeros_like(full) for _ in range(world)]
        dist.all_gather(gathered, full)
        whole = torch.cat(gathered, dim=1)
        for r in range(world):
            idx_r = torch.arange(r, kb, 7)[:8]
            exp = unit(torch.randn(len(idx_r), d, generator=torch.Generator().manual_seed(r)))
            torch.testing.assert_close(whole[:, idx_r], exp.t())
    finally:
        dist.destroy_process_group()


def test_sharded_bank_schedule_matches_whole_bank_oracle():
    mp.spawn(_bank_worker, args=(2, _free_port(), 12, 256, 16, 5), nprocs=2, join=True)
