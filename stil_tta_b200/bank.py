"""Memory-bank block of the SimMatch baseline (BASELINE config C5) — drop-in for the ``start_unlabel`` branch of
``SimMatchModel.forward``, ``models/MatchModel/simmatch_model.py:268-286``.

The bank keeps the reference layout ``[dim, K]`` (unit columns, ``:68-69``) and the ``labels [K]`` int64 buffer
(``:70``) so checkpoints round-trip; it is read in place by the tensor cores (as an MN-major operand for the logits,
as a K-major operand for the feature gradient).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._lib import check, dtype_code, ptr


def alloc_bank(dim: int, k_bank: int, dtype=torch.bfloat16, device="cuda") -> torch.Tensor:
    """A ``[dim, k_bank]`` bank whose rows are 16-byte aligned for any ``k_bank`` (view of a padded buffer)."""
    per16 = 8 if dtype == torch.bfloat16 else 4
    ld = (k_bank + per16 - 1) // per16 * per16
    return torch.zeros(dim, ld, dtype=dtype, device=device)[:, :k_bank]


class _SimMatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat_ku, feat_qu, prob_ku_orig, bank, labels, tt, st, c_smooth):
        dev = _lib.require_cuda(feat_ku, feat_qu, prob_ku_orig, bank, labels)
        _lib.ensure_device(dev)
        dt = bank.dtype
        if dt not in (torch.float32, torch.bfloat16):
            raise ValueError("bank must be float32 or bfloat16")
        fk = feat_ku.detach().to(dt).contiguous()
        fq = feat_qu.detach().to(dt).contiguous()
        p = prob_ku_orig.detach().to(torch.float32).contiguous()
        lab = labels.to(torch.int64).contiguous()
        if bank.dim() != 2 or bank.stride(1) != 1:
            raise ValueError("bank must be [dim, K] with unit stride along K (reference layout)")
        rows, d = fq.shape
        kb, c = bank.shape[1], p.shape[1]
        if bank.shape[0] != d or lab.numel() != kb:
            raise ValueError("bank / labels / feature shapes do not match")
        lib = _lib.load()
        code = dtype_code(fq)
        ws = torch.empty(lib.stil_simmatch_workspace_bytes(rows, kb, d, code), dtype=torch.uint8, device=dev)
        prob_ku = torch.empty(rows, c, dtype=torch.float32, device=dev)
        loss_in = torch.empty(rows, dtype=torch.float32, device=dev)
        gcode = _lib.STIL_F32            # dLoss/dLogits kept as a bf16 hi+lo pair, fp32 gradient
        with torch.cuda.device(dev):
            check(lib.stil_simmatch_fwd(ptr(fk), ptr(fq), code, rows, d, d, ptr(bank), bank.stride(0), ptr(lab), kb, ptr(p), c,
                                        float(tt), float(st), float(c_smooth), ptr(prob_ku), ptr(loss_in), gcode, ptr(ws),
                                        ws.numel(), _lib.stream_ptr(dev)))
        # d loss_in[i] / d feat_qu[i, :] is computed HERE, while the bank still holds what the forward saw: the reference
        # overwrites bank columns right after this block and before loss.backward() (simmatch_model.py:291; it protects
        # itself with bank.clone(), :237).  In-place writes through raw kernels do not move torch's version counter, so
        # reading the live bank in backward would silently use the new columns.
        jac = None
        if ctx.needs_input_grad[1]:
            jac = torch.empty(rows, d, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                check(lib.stil_simmatch_bwd(ptr(fq), code, rows, d, ptr(bank), bank.stride(0), kb, None, ptr(jac),
                                            _lib.STIL_F32, d, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.jac = jac
        ctx.in_dtype = feat_qu.dtype
        ctx.mark_non_differentiable(prob_ku)
        return prob_ku, loss_in

    @staticmethod
    def backward(ctx, _g_prob, g_loss):
        if ctx.jac is None:
            return (None,) * 8
        d_fq = g_loss.detach().to(torch.float32)[:, None] * ctx.jac
        return None, d_fq.to(ctx.in_dtype), None, None, None, None, None, None


def simmatch_bank(feat_ku: torch.Tensor, feat_qu: torch.Tensor, prob_ku_orig: torch.Tensor, bank: torch.Tensor,
                  labels: torch.Tensor, tt: float, st: float, c_smooth: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(prob_ku, loss_in)`` of ``simmatch_model.py:268-286``: teacher bank softmax re-weighted by the class
    probabilities of the bank labels, class-aggregated smoothing of the pseudo label, and the instance-similarity loss
    (per row; only ``feat_qu`` receives a gradient)."""
    return _SimMatchFn.apply(feat_ku, feat_qu, prob_ku_orig, bank, labels, tt, st, c_smooth)


# ----------------------------------------------------------------------------------------------------------------------
# Column-sharded bank over the GPUs of one node (BASELINE config C5: 65536 x 512 over 8 B200)
# ----------------------------------------------------------------------------------------------------------------------
class _ShardedSimMatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat_ku, feat_qu, prob_ku_orig, owner, tt, st, c_smooth):
        prob_ku, loss_in, jac = owner._run_sweep(feat_ku, feat_qu, prob_ku_orig, tt, st, c_smooth, ctx.needs_input_grad[1])
        ctx.jac = jac
        ctx.in_dtype = feat_qu.dtype
        ctx.mark_non_differentiable(prob_ku)
        return prob_ku, loss_in

    @staticmethod
    def backward(ctx, _g_prob, g_loss):
        if ctx.jac is None:
            return (None,) * 7
        return None, (g_loss.detach().to(torch.float32)[:, None] * ctx.jac).to(ctx.in_dtype), None, None, None, None, None


class ShardedSimMatchBank:
    """The SimMatch memory bank (``simmatch_model.py:68-70``) column-sharded over the ranks of ``group``: rank ``r`` holds
    ``bank[:, r*K/W : (r+1)*K/W]`` (reference layout ``[dim, K/W]``) and the matching ``labels`` slice.

    The reference instead REPLICATES the bank on every GPU and all-gathers the updates (``_update_bank``, ``:141-147``); at
    C5 sizes (65536 x 512) that is 64 MiB of bank swept by every GPU for its 448 rows.  Sharded, every rank sweeps
    ``K/W`` columns for the gathered rows of all ranks — the same tensor work per GPU, 1/W of the bank bytes per GPU, and
    the bank grows with the node.  One step (``__call__``, the block ``simmatch_model.py:268-286``):

      1. ONE all-gather of the packed rows ``[feat_ku | feat_qu | prob_ku_orig]`` over ranks (rows of all ranks, rank-major)
      2. ``stil_simmatch_shard_stats``: logits against the local shard, per-row statistics with a FIXED shift
         (unit vectors: ``e = exp((z - 1)/T)``), so that they are ADDITIVE over shards
      3. reduce-scatter(SUM) of the ``[W*rows, 3 + C]`` statistics: every rank gets the totals of its own rows
      4. ``stil_simmatch_shard_finish`` on those rows: ``prob_ku`` (``:280``), ``loss_in`` (``:286``) and the two normalisers
         per row, which are all-gathered (``[W*rows, 2]``) for the gradient pass
      5. (when ``feat_qu`` needs a gradient) ``stil_simmatch_shard_grad``: ``G = (S - T')/st`` on the shard's columns and
         the partial ``G · bank_shardᵀ``; reduce-scatter(SUM) gives every rank ``d loss_in / d feat_qu`` of its own rows.
         It is computed in the forward, while the bank still holds what the forward saw (the reference overwrites bank
         columns before ``loss.backward()``, ``:291``).

    ``emulate_shards=S`` (tests, one process): the bank is split into ``S`` shards swept one after the other on this GPU
    and the collectives become sums — the same kernels, offsets and additive statistics as ``S`` ranks.

    ``use_graph=True``: the whole sweep (gathers, kernels, collectives) is captured into ONE CUDA graph per input signature
    (shapes, dtype, temperatures, smoothing, gradient wanted) on first use and replayed afterwards — the eager sweep is ~15
    launches and four extension calls, and on a B200 the host needs longer to enqueue them than the GPU to run them.  The
    graph reads the bank and label buffers in place (``update`` / ``load`` stay visible); inputs are copied into the
    graph's static buffers and the three small results are cloned out, so callers see ordinary tensors.  Every rank must
    capture with the same signature sequence (the collectives are part of the graph).
    """

    def __init__(self, dim: int, k_bank: int, num_classes: int, dtype=torch.bfloat16, device="cuda", group=None,
                 emulate_shards: int = 1, use_graph: bool = False) -> None:
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.dev = torch.device(device)
        self._check_device()
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.nshards = emulate_shards if self.world == 1 else 1
        parts = self.world * self.nshards
        if k_bank % parts:
            raise ValueError(f"k_bank {k_bank} must be divisible by the number of shards {parts}")
        self.dim, self.k_bank, self.num_classes, self.dtype = dim, k_bank, num_classes, dtype
        self.k_shard = k_bank // parts
        # reference buffer names (simmatch_model.py:68-70); one (bank, labels) pair per local shard
        self.bank = [alloc_bank(dim, self.k_shard, dtype, self.dev) for _ in range(self.nshards)]
        self.labels = [torch.zeros(self.k_shard, dtype=torch.int64, device=self.dev) for _ in range(self.nshards)]
        self._ws = None
        self.use_graph = bool(use_graph)
        self._graphs = {}
        # this library's kernels per sweep and shard: logits, statistics, chunk reduce, G, dX (+ one finish per sweep); the
        # collectives, the packing cat / copies and the memset are not counted
        self.launches_per_step = self.nshards * 5 + 1

    def _check_device(self) -> None:
        if self.dev.type != "cuda":
            raise RuntimeError("ShardedSimMatchBank runs on a CUDA device (sm_100a) only; there is no CPU fallback")
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        _lib.ensure_device(self.dev)

    # ------------------------------------------------------------------------------------------
    def load_shard(self, bank_rows: torch.Tensor, labels: torch.Tensor, shard: int = 0) -> None:
        """Fill local shard ``shard`` from ``bank_rows [k_shard, dim]`` (row-major, unit rows) and its labels."""
        self._check_labels(labels)
        self.bank[shard].copy_(bank_rows.to(self.dev).t())
        self.labels[shard].copy_(labels.to(self.dev))

    def load(self, bank_rows: torch.Tensor, labels: torch.Tensor) -> None:
        """Fill every local shard from the WHOLE bank ``[k_bank, dim]`` (each rank keeps only its columns)."""
        for s in range(self.nshards):
            g = (self.rank * self.nshards + s) * self.k_shard
            self.load_shard(bank_rows[g:g + self.k_shard], labels[g:g + self.k_shard], s)

    def update(self, k: torch.Tensor, y: torch.Tensor, index: torch.Tensor) -> None:
        """``_update_bank`` (``simmatch_model.py:141-147``) with GLOBAL column indices: ``k``/``y``/``index`` are gathered
        over ranks like the reference does, then every rank writes the columns it owns."""
        self._check_labels(y)
        if self.world > 1:
            k, y, index = self._gather(k.contiguous()), self._gather(y.contiguous()), self._gather(index.contiguous())
        for s in range(self.nshards):
            g = (self.rank * self.nshards + s) * self.k_shard
            mine = (index >= g) & (index < g + self.k_shard)
            if bool(mine.any()):
                self._k_update(s, k[mine], y[mine], index[mine] - g)

    def _check_labels(self, y: torch.Tensor) -> None:
        """The sweep kernels index per-class accumulators with the stored labels (the reference's ``gather`` / ``scatter_add``
        raise on a bad class, ``simmatch_model.py:274-279``): refuse them where they enter the bank."""
        if y.numel() and (int(y.min()) < 0 or int(y.max()) >= self.num_classes):
            raise ValueError(f"bank labels must lie in [0, {self.num_classes})")

    def _k_update(self, s, k, y, index_local) -> None:
        from .bank_blocks import update_bank
        update_bank(self.bank[s], self.labels[s], k, y, index_local)

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    # ---- the three kernels of one sweep (tests/test_dist_gloo.py overrides them with the torch restatement of
    # oracle/bank_oracle.py to run the SCHEDULE below on CPU ranks over gloo)
    def _k_stats(self, s, fk, fq, p, tt, st, out):
        lib, bank = _lib.load(), self.bank[s]
        rows_all, d = fq.shape
        if fk.stride(0) != fq.stride(0) or fk.stride(1) != 1 or fq.stride(1) != 1:
            raise ValueError("feat_ku / feat_qu must share one row stride")
        check(lib.stil_simmatch_shard_stats(ptr(fk), ptr(fq), dtype_code(fq), rows_all, d, fq.stride(0), ptr(bank), bank.stride(0),
                                            ptr(self.labels[s]), self.k_shard, ptr(p), p.shape[1], float(tt), float(st), ptr(out),
                                            ptr(self._ws[s]), self._ws[s].numel(), _lib.stream_ptr(self.dev)))

    def _k_finish(self, stats, p, st, c_smooth, prob_all, loss_all, norms):
        check(_lib.load().stil_simmatch_shard_finish(ptr(stats), ptr(p), p.shape[0], p.shape[1], float(st), float(c_smooth),
                                                     ptr(prob_all), ptr(loss_all), ptr(norms), _lib.stream_ptr(self.dev)))

    def _k_grad(self, s, fk, fq, p, tt, st, norms, out):
        lib, bank = _lib.load(), self.bank[s]
        rows_all, d = fq.shape
        check(lib.stil_simmatch_shard_grad(dtype_code(fq), rows_all, d, ptr(bank), bank.stride(0), ptr(self.labels[s]),
                                           self.k_shard, ptr(p), p.shape[1], float(tt), float(st), ptr(norms), ptr(out), d,
                                           ptr(self._ws[s]), self._ws[s].numel(), _lib.stream_ptr(self.dev)))

    def _alloc_ws(self, rows_all, d, code):
        nbytes = _lib.load().stil_simmatch_workspace_bytes(rows_all, self.k_shard, d, code)
        if self._ws is None or self._ws[0].numel() < nbytes:
            self._ws = [torch.empty(nbytes, dtype=torch.uint8, device=self.dev) for _ in range(self.nshards)]

    def _sweep(self, feat_ku, feat_qu, prob_ku_orig, tt, st, c_smooth, need_grad):
        dev, W = self.dev, self.world
        fk = feat_ku.detach().to(self.dtype).contiguous()
        fq = feat_qu.detach().to(self.dtype).contiguous()
        p = prob_ku_orig.detach().to(torch.float32).contiguous()
        if W > 1:
            # ONE all-gather instead of three (each is ~12 us of latency at these sizes): a row of the packed buffer is
            # [feat_ku | feat_qu | prob_ku_orig | pad to 16 bytes]; the kernels read the features in place through the row
            # stride, the probabilities are copied out (the kernels want them contiguous)
            u8 = torch.uint8
            parts = [fk.view(u8), fq.view(u8), p.view(u8)]
            nbytes = sum(t.shape[1] for t in parts)
            if nbytes % 16:
                parts.append(torch.zeros(fk.shape[0], 16 - nbytes % 16, dtype=u8, device=fk.device))
            packed = self._gather(torch.cat(parts, dim=1))
            o1, o2 = parts[0].shape[1], parts[0].shape[1] + parts[1].shape[1]
            fk = packed[:, :o1].view(self.dtype)
            fq = packed[:, o1:o2].view(self.dtype)
            p = packed[:, o2:o2 + parts[2].shape[1]].view(torch.float32).contiguous()
        rows_all, d = fq.shape
        rows = rows_all // W
        c = p.shape[1]
        if d != self.dim or c != self.num_classes:
            raise ValueError("feature / probability shapes do not match the bank")
        self._alloc_ws(rows_all, d, dtype_code(fq))
        f32 = dict(dtype=torch.float32, device=dev)
        # one shard: the statistics kernels write every entry; several emulated shards: summed into a cleared buffer
        stats = torch.zeros(rows_all, 3 + c, **f32) if self.nshards > 1 else torch.empty(rows_all, 3 + c, **f32)
        part = torch.empty(rows_all, 3 + c, **f32) if self.nshards > 1 else stats
        with self._device_guard():
            for s in range(self.nshards):
                self._k_stats(s, fk, fq, p, tt, st, part)
                if self.nshards > 1:
                    stats += part
            lo = self.rank * rows
            if W > 1:
                # every rank finishes ITS rows: reduce-scatter of the statistics (instead of an all-reduce of all W*rows rows),
                # then only the two normalisers per row travel back to everybody for the gradient pass
                mine = torch.empty(rows, 3 + c, **f32)
                self.dist.reduce_scatter_tensor(mine, stats, group=self.group)
                prob_ku, loss_in, norms_mine = torch.empty(rows, c, **f32), torch.empty(rows, **f32), torch.empty(rows, 2, **f32)
                self._k_finish(mine, p[lo:lo + rows], st, c_smooth, prob_ku, loss_in, norms_mine)
                norms = self._gather(norms_mine) if need_grad else None
            else:
                prob_ku, loss_in, norms = torch.empty(rows_all, c, **f32), torch.empty(rows_all, **f32), torch.empty(rows_all, 2, **f32)
                self._k_finish(stats, p, st, c_smooth, prob_ku, loss_in, norms)
            jac = None
            if need_grad:
                jall = torch.zeros(rows_all, d, **f32) if self.nshards > 1 else None
                jp = torch.empty(rows_all, d, **f32)
                for s in range(self.nshards):
                    self._k_grad(s, fk, fq, p, tt, st, norms, jp)
                    if self.nshards > 1:
                        jall += jp
                jall = jp if jall is None else jall
                if W > 1:
                    jac = torch.empty(rows, d, **f32)
                    self.dist.reduce_scatter_tensor(jac, jall, group=self.group)
                else:
                    jac = jall
        return prob_ku, loss_in, jac

    def _device_guard(self):
        return torch.cuda.device(self.dev)

    # ------------------------------------------------------------------------------------------ CUDA graph
    def _run_sweep(self, feat_ku, feat_qu, prob_ku_orig, tt, st, c_smooth, need_grad):
        if not self.use_graph:
            return self._sweep(feat_ku, feat_qu, prob_ku_orig, tt, st, c_smooth, need_grad)
        key = (tuple(feat_ku.shape), feat_ku.dtype, feat_qu.dtype, tuple(prob_ku_orig.shape), prob_ku_orig.dtype,
               float(tt), float(st), float(c_smooth), bool(need_grad))
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = self._capture(feat_ku, feat_qu, prob_ku_orig, tt, st, c_smooth, need_grad)
        graph, s_in, s_out = ent
        with self._device_guard():
            s_in[0].copy_(feat_ku.detach())
            s_in[1].copy_(feat_qu.detach())
            s_in[2].copy_(prob_ku_orig.detach())
            graph.replay()
            return tuple(None if t is None else t.clone() for t in s_out)

    def _capture(self, feat_ku, feat_qu, prob_ku_orig, tt, st, c_smooth, need_grad):
        """Warm up on a side stream (workspace allocation, lazy kernel attributes, NCCL channels), then capture one sweep."""
        with self._device_guard():
            s_in = [feat_ku.detach().clone(), feat_qu.detach().clone(), prob_ku_orig.detach().clone()]
            cur = torch.cuda.current_stream(self.dev)
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._sweep(s_in[0], s_in[1], s_in[2], tt, st, c_smooth, need_grad)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                s_out = self._sweep(s_in[0], s_in[1], s_in[2], tt, st, c_smooth, need_grad)
        return graph, s_in, s_out

    def drop_graphs(self) -> None:
        """Forget the captured sweeps (after ``load`` replaced the bank tensors, or to release their memory)."""
        self._graphs.clear()

    def __call__(self, feat_ku, feat_qu, prob_ku_orig, tt: float, st: float, c_smooth: float):
        """``(prob_ku, loss_in)`` of ``simmatch_model.py:268-286`` for this rank's rows against the WHOLE (sharded) bank."""
        return _ShardedSimMatchFn.apply(feat_ku, feat_qu, prob_ku_orig, self, tt, st, c_smooth)
