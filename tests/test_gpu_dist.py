"""2-GPU (NCCL) test of the data-parallel head: DistributedSTiLHead on two ranks == the single-process oracle on
the concatenated batch (global-batch InfoNCE, all-reduced prototype partials).  Skipped with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, use_graph, transport):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import stil_tta_b200 as S
        from stil_tta_b200 import synth
        from oracle import stil_head_oracle as O
        cfg = synth.dvm_config(batch)
        batches = [synth.make_batch(cfg, seed=2022, rank=r) for r in range(world)]
        head = S.DistributedSTiLHead(cfg, device=f"cuda:{rank}", use_graph=use_graph, transport=transport)
        head.load(batches[rank])
        for _ in range(4):          # every run accumulates again; p2p alternates its two regions
            head.run()
        torch.cuda.synchronize()
        # oracle: reference CLIPLoss on the concatenation of all ranks' rows
        a = [b["feat_i"].float().requires_grad_(True) for b in batches]
        bb = [b["feat_t"].float().requires_grad_(True) for b in batches]
        loss, _, _ = O.clip_loss_global(a, bb, cfg.temperature, cfg.lambda_0)
        ga, gb = torch.autograd.grad(loss, (a[rank], bb[rank]))
        got = float(head.out["losses"][0])
        assert abs(got - float(loss)) <= 1e-3 * float(loss), (got, float(loss))
        for name, ref in (("d_feat_i", ga), ("d_feat_t", gb)):
            err = float((head.out[name].float().cpu() - ref).abs().max() / ref.abs().max())
            assert err <= 1e-3, (name, err)      # fp32 gradients of bf16 embeddings (the head's default)
        # row-local outputs == single-rank oracle; prototype partials == sum over ranks
        os_ = [O.head_step(b, cfg, with_grads=False) for b in batches]
        for k in ("max_idx", "mask1", "case1", "case3"):
            assert torch.equal(head.out[k].cpu(), os_[rank][k]), k
        assert abs(float(head.out["losses"][1]) - float(os_[rank]["loss_pt"])) <= 1e-3 * float(os_[rank]["loss_pt"]) + 1e-6
        cs = sum(o["class_sum"] for o in os_)
        cc = sum(o["class_count"] for o in os_)
        assert float((head.out["class_sum"].cpu() - cs).abs().max()) <= 1e-4
        assert float((head.prototypes_sum.cpu() - 4 * cs).abs().max()) <= 4e-4
        assert float((head.prototypes_count_sum.cpu() - 4 * cc).abs().max()) <= 4e-5
        ok = True
    except BaseException:
        import traceback
        traceback.print_exc()
        ok = False
    # a graph that captured NCCL work pins the communicator: drop it, then leave without the process-group teardown
    try:
        head.release()
        torch.cuda.synchronize()
    except BaseException:
        pass
    os._exit(0 if ok else 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("use_graph,transport", [(False, "fused"), (True, "fused"), (False, "p2p"), (True, "p2p"), (True, "nccl")])
def test_distributed_head_two_ranks(use_graph, transport):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), 256, use_graph, transport), nprocs=2, join=True)


def _bank_worker(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = False
    try:
        import stil_tta_b200 as S
        from oracle import stil_head_oracle as O
        rows, kb, d, c = 96, 2048, 128, 10
        g = torch.Generator().manual_seed(3)
        unit = torch.nn.functional.normalize
        bank_rows = unit(torch.randn(kb, d, generator=g)).to(torch.bfloat16)
        labels = torch.randint(0, c, (kb,), generator=g)
        fk_all = unit(bank_rows.float()[torch.randint(0, kb, (world * rows,), generator=g)] + 0.3 * torch.randn(world * rows, d, generator=g)).to(torch.bfloat16)
        fq_all = unit(fk_all.float() + 0.2 * torch.randn(world * rows, d, generator=g)).to(torch.bfloat16)
        p_all = torch.softmax(torch.randn(world * rows, c, generator=g) * 3, 1)
        loc = slice(rank * rows, (rank + 1) * rows)
        sb = S.ShardedSimMatchBank(d, kb, c, dtype=torch.bfloat16, device=f"cuda:{rank}")
        assert sb.world == world
        sb.load(bank_rows, labels)
        fq = fq_all[loc].cuda().requires_grad_(True)
        prob_ku, loss_in = sb(fk_all[loc].cuda(), fq, p_all[loc].cuda(), 0.1, 0.1, 0.9)
        (gq,) = torch.autograd.grad(loss_in.mean(), fq)
        fqr = fq_all.float().requires_grad_(True)
        ref = O.simmatch_bank(fk_all.float(), fqr, p_all, bank_rows.float(), labels, 0.1, 0.1, 0.9)
        (g_ref,) = torch.autograd.grad(ref["loss_in"][loc].mean(), fqr)
        assert float((prob_ku.cpu() - ref["prob_ku"][loc]).abs().max()) <= 2e-5
        e1 = float((loss_in.cpu() - ref["loss_in"][loc].detach()).abs().max() / ref["loss_in"][loc].abs().max())
        e2 = float((gq.float().cpu() - g_ref[loc]).abs().max() / g_ref[loc].abs().max())
        assert e1 <= 1e-3 and e2 <= 4e-3, (e1, e2)       # gq is bf16 (input dtype): 2^-9 output rounding on top of 1e-3
        # the same sweep captured into one CUDA graph per rank (collectives included), replayed on the same inputs
        sg = S.ShardedSimMatchBank(d, kb, c, dtype=torch.bfloat16, device=f"cuda:{rank}", use_graph=True)
        sg.load(bank_rows, labels)
        for _ in range(2):
            fq2 = fq_all[loc].cuda().requires_grad_(True)
            prob_g, loss_g = sg(fk_all[loc].cuda(), fq2, p_all[loc].cuda(), 0.1, 0.1, 0.9)
            (gq2,) = torch.autograd.grad(loss_g.mean(), fq2)
            assert float((prob_g - prob_ku).abs().max()) <= 1e-6
            assert float((loss_g - loss_in).abs().max() / loss_in.abs().max()) <= 1e-5
            assert float((gq2.float() - gq.float()).abs().max() / gq.float().abs().max()) <= 1e-5
        ok = True
    except BaseException:
        import traceback
        traceback.print_exc()
    os._exit(0 if ok else 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_simmatch_bank_two_ranks():
    import torch.multiprocessing as mp
    mp.spawn(_bank_worker, args=(2, _free_port()), nprocs=2, join=True)
