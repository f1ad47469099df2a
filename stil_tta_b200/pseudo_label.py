"""CGPL consensus pseudo-labelling + PGLS prototype-guided smoothing for the unlabelled rows.

The reference has no named function for this: it is the inline ``torch.no_grad()`` block
``models/Disentangle/STiLModel.py:262-298``.  ``cgpl_pgls`` takes exactly the tensors that block
reads and returns exactly the tensors it binds.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from . import _lib
from ._lib import check, dtype_code, ptr


class PseudoLabels(NamedTuple):
    pseudo_label: torch.Tensor   # [B_u, K] f32   STiLModel.py:295
    prediction: Optional[torch.Tensor]  # [B_u, K] f32   :296 (None unless return_prediction)
    max_prob: torch.Tensor       # [B_u] f32      :297
    max_idx: torch.Tensor        # [B_u] int64    :297
    mask1: torch.Tensor          # [B_u] bool     :298
    case1: torch.Tensor          # [B_u] bool     :264
    case2_i: torch.Tensor        # :265
    case2_t: torch.Tensor        # :266
    case3: torch.Tensor          # :267
    top1: torch.Tensor           # [3, B_u] int64 :263 (m, i, t)
    teacher_logits: torch.Tensor  # [B_u, K] f32  :293 (raw feat·protoᵀ, before /T)


def prototype_logits(feat: torch.Tensor, prototypes: torch.Tensor) -> torch.Tensor:
    """``feat @ prototypes.t()`` in fp32 on the tensor cores (``STiLModel.py:293``, ``:350``)."""
    f = feat.detach()
    if f.dtype not in (torch.float32, torch.bfloat16):
        f = f.float()
    f = f.contiguous()
    protos = prototypes.detach().to(torch.float32).contiguous()
    dev = _lib.require_cuda(f, protos)
    _lib.ensure_device(dev)
    rows, d = f.shape
    k = protos.shape[0]
    ld = (k + 3) // 4 * 4
    out = torch.empty(rows, ld, dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = _lib.workspace(dev, "proto", lib.stil_proto_logits_workspace_bytes(rows, k, d, dtype_code(f)))
    with torch.cuda.device(dev):
        check(lib.stil_proto_logits(ptr(f), dtype_code(f), rows, d, d, ptr(protos), k, ptr(out), ld, ptr(ws),
                                    ws.numel(), _lib.stream_ptr(dev)))
    return out[:, :k]


@torch.no_grad()
def cgpl_pgls(y_m: torch.Tensor, y_i: torch.Tensor, y_t: torch.Tensor, feat_m_ue: torch.Tensor,
              prototypes: torch.Tensor, *, T: float, rate_pseudo: float, th1: float,
              return_prediction: bool = True, teacher_logits: Optional[torch.Tensor] = None,
              prediction_override: Optional[torch.Tensor] = None) -> PseudoLabels:
    """Fused CGPL (3x softmax/argmax, 4 agreement cases, case-averaged pseudo label) and PGLS (teacher
    prototype probabilities, smoothing mix, max/argmax, confidence mask).

    ``y_*`` are the TEACHER logits of the unlabelled rows [B_u, K]; ``feat_m_ue`` the teacher multimodal
    feature [B_u, P]; ``prototypes`` the (un-normalised) bank [K, P].  Index / mask outputs are bit-exact
    with torch semantics (first index among equal maxima, ``>=`` on fp32).

    ``prediction_override`` [B_u, K]: the ``prediction`` of ``STiLModel.py:276-277`` when ``DA == True`` —
    ``distribution_alignment(torch.softmax(y_hat_m_ue, dim=1))`` — which then replaces ``softmax(y_m)`` in the
    smoothing mix and the threshold (``:296-298``); cases and ``pseudo_label`` are unaffected (``:263-274, :295``).
    """
    ys = [y.detach() for y in (y_m, y_i, y_t)]
    if any(y.dtype not in (torch.float32, torch.bfloat16) for y in ys) or len({y.dtype for y in ys}) != 1:
        ys = [y.float() for y in ys]
    ys = [y.contiguous() for y in ys]
    dev = _lib.require_cuda(*ys, feat_m_ue, prototypes)
    _lib.ensure_device(dev)
    rows, k = ys[0].shape
    if teacher_logits is None:
        teacher_logits = prototype_logits(feat_m_ue, prototypes)
    tl = teacher_logits
    if tl.dtype != torch.float32 or tl.stride(1) != 1:
        tl = tl.float().contiguous()
    pin = prediction_override
    if pin is not None:
        pin = pin.detach()
        if pin.dtype != torch.float32 or pin.stride(-1) != 1:
            pin = pin.float().contiguous()
        if pin.shape != (rows, k):
            raise ValueError(f"prediction_override must be [{rows}, {k}], got {tuple(pin.shape)}")
        _lib.require_cuda(pin, ys[0])
    f32 = dict(dtype=torch.float32, device=dev)
    pl = torch.empty(rows, k, **f32)
    pred = torch.empty(rows, k, **f32) if return_prediction else None
    max_prob = torch.empty(rows, **f32)
    max_idx = torch.empty(rows, dtype=torch.int64, device=dev)
    flags = [torch.empty(rows, dtype=torch.bool, device=dev) for _ in range(5)]
    top1 = torch.empty(3, rows, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().stil_cgpl_pgls(ptr(ys[0]), ptr(ys[1]), ptr(ys[2]), dtype_code(ys[0]), k, ptr(tl),
                                         tl.stride(0) if rows > 0 else k, rows, k, float(T), float(rate_pseudo),
                                         float(th1), 1, ptr(pin), pin.stride(0) if pin is not None and rows > 0 else k,
                                         ptr(pl), k, ptr(pred), k, ptr(max_prob), ptr(max_idx),
                                         ptr(flags[0]), ptr(flags[1]), ptr(flags[2]), ptr(flags[3]), ptr(flags[4]),
                                         ptr(top1), None, None, _lib.stream_ptr(dev)))
    return PseudoLabels(pl, pred, max_prob, max_idx, flags[0], flags[1], flags[2], flags[3], flags[4], top1, tl)


@torch.no_grad()
def distribution_alignment(probs: torch.Tensor, DA_queue: torch.Tensor, DA_ptr: torch.Tensor, group=None
                           ) -> torch.Tensor:
    """Drop-in for ``STiLModel.distribution_alignment`` (``STiLModel.py:171-180``): the batch-mean class
    distribution (all-reduced over ranks when a process group is initialised) enters the ring buffer ``DA_queue``
    [256, K] at ``DA_ptr`` (int64 [1], advanced on the device — no host sync), ``probs`` are divided by the queue
    mean and row-renormalised.  ``DA_queue`` / ``DA_ptr`` are the caller's ``register_buffer``s, updated in place."""
    import torch.distributed as dist
    p = probs.detach().to(torch.float32).contiguous()
    dev = _lib.require_cuda(p, DA_queue, DA_ptr)
    _lib.ensure_device(dev)
    if DA_queue.dtype != torch.float32 or DA_ptr.dtype != torch.int64 or not DA_queue.is_contiguous():
        raise ValueError("DA_queue must be a contiguous float32 [len, K] buffer and DA_ptr an int64 [1] buffer")
    rows, k = p.shape
    lib = _lib.load()
    mean = torch.empty(k, dtype=torch.float32, device=dev)
    scratch = torch.empty(k, dtype=torch.float32, device=dev)
    out = torch.empty_like(p)
    with torch.cuda.device(dev):
        check(lib.stil_da_batch_mean(ptr(p), k, rows, k, ptr(mean), _lib.stream_ptr(dev)))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(mean, group=group)                     # STiLModel.py:174
            mean /= dist.get_world_size(group)                     # :176
        check(lib.stil_da_apply(ptr(p), k, rows, k, ptr(mean), ptr(DA_queue), DA_queue.shape[0], ptr(DA_ptr),
                                ptr(scratch), ptr(out), k, _lib.stream_ptr(dev)))
    return out
