"""Memory-bank head blocks of the CoMatch and MMatch baselines and the queue maintenance next
to a7-a9 — drop-ins for the tensor code inside ``CoMatchModel.forward`` (``models/MatchModel/comatch_model.py:268-321``),
``CoMatch.training_step`` (``models/MatchModel/CoMatch.py:92-110``) and ``MMatch.training_step``
(``models/SemiMultimodal/MMatch.py:215-235, 259``).

Queues keep the reference layouts (``queue [dim, K_q]``, ``probs [C, K_q]``, int64 ``[1]`` pointers) so checkpoints
round-trip; they are read in place by the tensor cores.  No CPU fallback.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch

from . import _lib
from ._lib import check, dtype_code, ptr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _queue_view(q: torch.Tensor, name: str) -> torch.Tensor:
    """The extension reads ``[dim, K_q]`` queues in place; rows must be 16-byte granular (K_q % 4 for fp32, % 8 for bf16)."""
    per16 = 8 if q.dtype == torch.bfloat16 else 4
    if q.dim() != 2 or q.stride(1) != 1 or q.stride(0) % per16 or q.data_ptr() % 16:
        raise ValueError(f"{name} must be [dim, K_q] with unit stride along K_q and 16-byte aligned rows "
                         f"(K_q a multiple of {per16}); got shape {tuple(q.shape)}, strides {q.stride()}")
    return q


class SmoothedLabels(NamedTuple):
    probs: torch.Tensor        # [rows, C]
    max_prob: torch.Tensor     # [rows]
    max_idx: torch.Tensor      # [rows] int64
    mask: torch.Tensor         # [rows] bool, max_prob >= th


@torch.no_grad()
def bank_smooth(probs: torch.Tensor, feat: Optional[torch.Tensor], queue_feat: Optional[torch.Tensor],
                queue_probs: Optional[torch.Tensor], temperature: float, c_keep: float, c_bank: float, th: float
                ) -> SmoothedLabels:
    """``c_keep*probs + c_bank * rownorm(exp(feat @ queue_feat / T)) @ queue_probs.T`` and its ``max / argmax / >= th``
    (``comatch_model.py:288-293`` + ``CoMatch.py:92-93``; ``MMatch.py:222-230``).  ``queue_feat=None`` is the epoch gate
    (``MMatch.py:221``, ``comatch_model.py:288``): the distribution passes through unchanged."""
    p = _f32c(probs)
    dev = _lib.require_cuda(p, feat, queue_feat, queue_probs)
    _lib.ensure_device(dev)
    rows, c = p.shape
    lib = _lib.load()
    out = torch.empty_like(p)
    max_prob = torch.empty(rows, dtype=torch.float32, device=dev)
    max_idx = torch.empty(rows, dtype=torch.int64, device=dev)
    mask = torch.empty(rows, dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        if queue_feat is None:
            check(lib.stil_bank_smooth(ptr(p), c, rows, c, None, 0, 0, 0, None, 0, None, 0, 0, 1.0, 1.0, 0.0, ptr(out), c,
                                       float(th), ptr(max_prob), ptr(max_idx), ptr(mask), None, 0, _lib.stream_ptr(dev)))
        else:
            q = _queue_view(queue_feat, "queue_feat")
            f = feat.detach().to(q.dtype).contiguous()
            qp = queue_probs
            if qp.dtype != torch.float32 or qp.stride(1) != 1:
                qp = _f32c(qp)
            d, kq = q.shape
            if f.shape != (rows, d) or qp.shape != (c, kq):
                raise ValueError("bank_smooth: feat / queue_feat / queue_probs shapes do not match")
            code = dtype_code(f)
            ws = _lib.workspace(dev, "bank_smooth", lib.stil_bank_smooth_workspace_bytes(rows, kq, d, c, code))
            check(lib.stil_bank_smooth(ptr(p), c, rows, c, ptr(f), code, d, d, ptr(q), q.stride(0), ptr(qp), qp.stride(0), kq,
                                       float(temperature), float(c_keep), float(c_bank), ptr(out), c, float(th),
                                       ptr(max_prob), ptr(max_idx), ptr(mask), ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
    return SmoothedLabels(out, max_prob, max_idx, mask)


def mmatch_pseudo_label(pseudo_label_orig: torch.Tensor, feat_m_u: torch.Tensor, embed_queue: torch.Tensor,
                        probs_queue: torch.Tensor, T: float, th1: float, current_epoch: int = 1) -> SmoothedLabels:
    """a9 — ``MMatch.py:221-230``: ``pseudo_label = 0.9*p + 0.1*A@probs_queue.T`` once ``current_epoch > 0``, then
    ``max_prob, max_idx`` and ``mask1 = max_prob >= th1``."""
    if current_epoch > 0:
        return bank_smooth(pseudo_label_orig, feat_m_u, embed_queue, probs_queue, T, 0.9, 0.1, th1)
    return bank_smooth(pseudo_label_orig, None, None, None, T, 1.0, 0.0, th1)


def comatch_smooth(probs: torch.Tensor, feature_u_w: torch.Tensor, queue_w: torch.Tensor, probs_xu: torch.Tensor,
                   temperature: float, alpha: float, thr: float, smooth: bool = True) -> SmoothedLabels:
    """a8 — ``comatch_model.py:288-293`` (``smooth`` = ``epoch > start_epoch``) and ``CoMatch.py:92-93``."""
    if smooth:
        return bank_smooth(probs, feature_u_w, queue_w, probs_xu, temperature, alpha, 1 - alpha, thr)
    return bank_smooth(probs, None, None, None, temperature, 1.0, 0.0, thr)


class _GraphsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, probs_u, feat_s0, feat_s1, queue_s, temperature):
        dev = _lib.require_cuda(probs, probs_u, feat_s0, feat_s1, queue_s)
        _lib.ensure_device(dev)
        q_s = _queue_view(queue_s, "queue_s")
        p = _f32c(probs)
        pu = probs_u if (probs_u.dtype == torch.float32 and probs_u.stride(1) == 1) else _f32c(probs_u)
        f0 = feat_s0.detach().to(q_s.dtype).contiguous()
        f1 = feat_s1.detach().to(q_s.dtype).contiguous()
        rows, c = p.shape
        d, kq = q_s.shape
        if f0.shape != (rows, d) or f1.shape != (rows, d) or pu.shape != (c, kq):
            raise ValueError("comatch_graphs: shapes do not match")
        lib = _lib.load()
        code = dtype_code(f0)
        ld = (rows + kq + 3) // 4 * 4
        Q = torch.empty(rows, ld, dtype=torch.float32, device=dev)[:, :rows + kq]
        sim = torch.empty(rows, ld, dtype=torch.float32, device=dev)[:, :rows + kq]
        ws = _lib.workspace(dev, "comatch_graphs", lib.stil_comatch_graphs_workspace_bytes(rows, kq, d, c, code))
        with torch.cuda.device(dev):
            check(lib.stil_comatch_graphs_fwd(ptr(p), c, rows, c, ptr(pu), pu.stride(0), ptr(f0), ptr(f1), code, d, d, ptr(q_s),
                                              q_s.stride(0), kq, float(temperature), ptr(Q), ptr(sim), ld, ptr(ws),
                                              ws.numel(), _lib.stream_ptr(dev)))
        # snapshot of queue_s: queue_enqueue overwrites it in place (raw kernel, torch's version counter does not move)
        # between this forward and loss.backward() — the reference passes self.queue_s.clone().detach() for the same
        # reason (comatch_model.py:310)
        ctx.save_for_backward(sim, f1, q_s.clone())
        ctx.meta = (float(temperature), feat_s0.dtype, ld)
        ctx.mark_non_differentiable(Q)
        return Q, sim

    @staticmethod
    def backward(ctx, _gq, g_sim):
        sim, f1, q_s = ctx.saved_tensors
        temperature, in_dtype, ld = ctx.meta
        dev = sim.device
        rows, d = f1.shape
        kq = q_s.shape[1]
        g = g_sim.detach().to(torch.float32)
        if g.stride(1) != 1 or g.stride(0) != ld:
            gp = torch.empty(rows, ld, dtype=torch.float32, device=dev)
            gp[:, :rows + kq] = g
            g = gp
        lib = _lib.load()
        code = dtype_code(f1)
        d_f0 = torch.empty(rows, d, dtype=torch.float32, device=dev)
        ws = _lib.workspace(dev, "comatch_sim_bwd", lib.stil_comatch_sim_bwd_workspace_bytes(rows, kq, d, code))
        with torch.cuda.device(dev):
            check(lib.stil_comatch_sim_bwd(ptr(g), ptr(sim), ld, rows, kq, ptr(f1), code, d, d, ptr(q_s), q_s.stride(0),
                                           temperature, ptr(d_f0), _lib.STIL_F32, d, ptr(ws), ws.numel(),
                                           _lib.stream_ptr(dev)))
        return None, None, d_f0.to(in_dtype), None, None, None


def comatch_graphs(probs: torch.Tensor, probs_u: torch.Tensor, features_u_s0: torch.Tensor, features_u_s1: torch.Tensor,
                   queue_s: torch.Tensor, temperature: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """a8 — ``(Q, sim)`` of ``comatch_model.py:298-312``: the pseudo-label graph ``[probs@probs.T (diag 1) | probs@probs_u]``
    (no gradient) and the embedding graph ``exp([f_s0@f_s1.T | f_s0@queue_s] / T)`` (gradient to ``features_u_s0``)."""
    return _GraphsFn.apply(probs, probs_u, features_u_s0, features_u_s1, queue_s, temperature)


class _ContrastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Q, sim, contrast_th):
        dev = _lib.require_cuda(Q, sim)
        _lib.ensure_device(dev)
        if Q.shape != sim.shape or Q.dtype != torch.float32 or sim.dtype != torch.float32:
            raise ValueError("graph_contrast_loss: Q and sim must be float32 of the same shape")
        if Q.stride(1) != 1 or sim.stride(1) != 1 or Q.stride(0) != sim.stride(0):
            Q, sim = Q.contiguous(), sim.detach().contiguous()
        rows, cols = Q.shape
        ld = Q.stride(0)
        lib = _lib.load()
        loss = torch.empty((), dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[1]
        d_sim = torch.empty(rows, ld, dtype=torch.float32, device=dev) if need else None
        ws = torch.empty(lib.stil_row_loss_workspace_bytes(rows), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.stil_graph_contrast_loss(ptr(Q), ptr(sim.detach()), ld, rows, cols, float(contrast_th), ptr(loss),
                                               ptr(d_sim), 1.0, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.d_sim = d_sim[:, :cols] if need else None
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d_sim
        return None, (d * g if d is not None else None), None


def graph_contrast_loss(Q: torch.Tensor, sim: torch.Tensor, contrast_th: float) -> torch.Tensor:
    """a8 consumer — ``loss_contrast`` of ``CoMatch.py:100-110`` (forward and d/d sim in one pass)."""
    return _ContrastFn.apply(Q, sim, contrast_th)


class _WeightedCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, mask):
        dev = _lib.require_cuda(logits, target, mask)
        _lib.ensure_device(dev)
        y = logits.detach()
        if y.dtype not in (torch.float32, torch.bfloat16):
            y = y.float()
        y = y.contiguous()
        rows, k = y.shape
        tp = tidx = None
        if target.dtype == torch.int64 and target.dim() == 1:
            tidx = target.contiguous()
        else:
            tp = _f32c(target)
        m = None if mask is None else (mask if mask.dtype in (torch.bool, torch.uint8) else mask != 0).contiguous()
        lib = _lib.load()
        loss = torch.empty((), dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[0]
        d_y = torch.empty(rows, k, dtype=torch.float32, device=dev) if need else None
        ws = torch.empty(lib.stil_row_loss_workspace_bytes(rows), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.stil_weighted_softce(ptr(y), dtype_code(y), k, ptr(tp), k, ptr(tidx), ptr(m), rows, k, ptr(loss),
                                           ptr(d_y), k, 1.0, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.d_y = d_y
        ctx.in_dtype = logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d_y
        return ((d * g).to(ctx.in_dtype) if d is not None else None), None, None


def masked_ce(logits: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """``mean_i mask_i * CE(logits_i, target_i)`` — the unlabelled loss of the single-head baselines:
    ``target`` [rows, K] probabilities (``SimMatch.py:91``, ``CoMatch.py:96-97``) or [rows] int64 class indices (the dense
    one-hot ``hard_label`` of ``MMatch.py:231-234``)."""
    return _WeightedCeFn.apply(logits, target, mask)


@torch.no_grad()
def queue_enqueue(queue_feat: torch.Tensor, queue_probs: torch.Tensor, queue_ptr: torch.Tensor, z: torch.Tensor,
                  t: torch.Tensor) -> None:
    """``_dequeue_and_enqueue`` of ``comatch_model.py:117-146`` / ``MMatch.py:102-117`` (single process; gather ``z`` and
    ``t`` over ranks first when distributed): in-place FIFO write, truncated at the wrap point; ``queue_ptr`` (int64 [1]
    on the device) is advanced on the device — no ``int(ptr)`` host sync."""
    dev = _lib.require_cuda(queue_feat, queue_probs, queue_ptr, z, t)
    _lib.ensure_device(dev)
    if queue_feat.stride(1) != 1 or queue_probs.stride(1) != 1 or queue_probs.dtype != torch.float32 \
            or queue_ptr.dtype != torch.int64:
        raise ValueError("queue_enqueue: queues must have unit stride along K_q, float32 probs and an int64 pointer")
    zc = z.detach()
    if zc.dtype not in (torch.float32, torch.bfloat16):
        zc = zc.float()
    zc = zc.contiguous()
    tc = _f32c(t)
    d, kq = queue_feat.shape
    c = queue_probs.shape[0]
    n = zc.shape[0]
    if zc.shape[1] != d or tc.shape != (n, c) or queue_probs.shape[1] != kq:
        raise ValueError("queue_enqueue: shapes do not match")
    with torch.cuda.device(dev):
        check(_lib.load().stil_queue_enqueue(ptr(queue_feat), dtype_code(queue_feat), queue_feat.stride(0), ptr(queue_probs),
                                             queue_probs.stride(0), kq, ptr(queue_ptr), ptr(zc), dtype_code(zc), d, n, d,
                                             ptr(tc), c, c, _lib.stream_ptr(dev)))


@torch.no_grad()
def update_bank(bank: torch.Tensor, labels: torch.Tensor, k: torch.Tensor, y: torch.Tensor, index: torch.Tensor) -> None:
    """``SimMatchModel._update_bank`` (``simmatch_model.py:141-147``, single process): ``bank[:, index] = k.T``,
    ``labels[index] = y``, in place."""
    dev = _lib.require_cuda(bank, labels, k, y, index)
    _lib.ensure_device(dev)
    if bank.stride(1) != 1 or labels.dtype != torch.int64:
        raise ValueError("update_bank: bank must have unit stride along K and int64 labels")
    kc = k.detach()
    if kc.dtype not in (torch.float32, torch.bfloat16):
        kc = kc.float()
    kc = kc.contiguous()
    n, d = kc.shape
    with torch.cuda.device(dev):
        check(_lib.load().stil_bank_update(ptr(bank), dtype_code(bank), bank.stride(0), ptr(labels), ptr(kc), dtype_code(kc), d,
                                           ptr(y.to(torch.int64).contiguous()), ptr(index.to(torch.int64).contiguous()), n, d,
                                           _lib.stream_ptr(dev)))


class HistAlignment:
    """CoMatch's distribution alignment (``comatch_model.py:271-285``): the reference keeps a Python list of the last
    128 batch means; here they live in a ``[128, C]`` device ring with a device-side counter (no host sync)."""

    def __init__(self, num_classes: int, device, hist_len: int = 128):
        self.hist = torch.zeros(hist_len, num_classes, dtype=torch.float32, device=device)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)

    @torch.no_grad()
    def __call__(self, probs: torch.Tensor, group=None) -> torch.Tensor:
        import torch.distributed as dist
        p = _f32c(probs)
        dev = _lib.require_cuda(p, self.hist)
        _lib.ensure_device(dev)
        rows, k = p.shape
        lib = _lib.load()
        mean = torch.empty(k, dtype=torch.float32, device=dev)
        scratch = torch.empty(k, dtype=torch.float32, device=dev)
        out = torch.empty_like(p)
        with torch.cuda.device(dev):
            check(lib.stil_da_batch_mean(ptr(p), k, rows, k, ptr(mean), _lib.stream_ptr(dev)))
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                dist.all_reduce(mean, group=group)                 # comatch_model.py:273-274
                mean /= dist.get_world_size(group)                 # :278
            check(lib.stil_da_apply_hist(ptr(p), k, rows, k, ptr(mean), ptr(self.hist), self.hist.shape[0], ptr(self.count),
                                         ptr(scratch), ptr(out), k, _lib.stream_ptr(dev)))
        return out
