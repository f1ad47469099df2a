"""Drop-in ``CLIPLoss`` / ``PrototypeLoss`` modules backed by the sm_100a kernels.

Same constructor and ``forward`` signatures, argument meaning, return tuples and error
behaviour as the reference ``utils/clip_loss.py:6-40`` and ``utils/prototype_loss.py:14-40``,
so ``STiLModel.__init__`` (``models/Disentangle/STiLModel.py:72-73``) can construct them and
``training_step`` (``:322``, ``:339``) / ``validation_step`` (``:435``) call them unchanged.
Each is a ``torch.autograd.Function`` over the C ABI (``include/stil_head.h``).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import check, dtype_code, ptr


def _prep_embed(t: torch.Tensor) -> torch.Tensor:
    """Embeddings go to the kernels as contiguous f32 or bf16 rows (other float types are upcast)."""
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()
    return t.contiguous()


class _InfoNCEFn(torch.autograd.Function):
    """loss = l0*CE(a b^T/T, arange) + (1-l0)*CE((a b^T/T)^T, arange) on L2-normalised rows."""

    @staticmethod
    def forward(ctx, out0, out1, temperature, lambda_0, want_logits, gather):
        a, b = _prep_embed(out0), _prep_embed(out1)
        if a.dtype != b.dtype:
            a, b = a.float(), b.float()
        dev = _lib.require_cuda(a, b)
        _lib.ensure_device(dev)
        if a.dim() != 2 or a.shape != b.shape:
            raise ValueError(f"CLIPLoss expects two [B, D] tensors of equal shape, got {tuple(a.shape)} and {tuple(b.shape)}")
        m, d = a.shape
        if gather is not None:
            a_all, b_all, off, n = gather.gather_rows(a), gather.gather_rows(b), gather.row_offset(m), gather.total_rows(m)
        else:
            a_all, b_all, off, n = a, b, 0, m
        lib = _lib.load()
        ws_bytes = lib.stil_infonce_workspace_bytes(m, n, d, dtype_code(a))
        ws = _lib.workspace(dev, "infonce", ws_bytes)
        loss_sum = torch.empty(1, dtype=torch.float32, device=dev)
        lse_row = torch.empty(m, dtype=torch.float32, device=dev)
        lse_col = torch.empty(m, dtype=torch.float32, device=dev)
        logits = torch.empty(m, n, dtype=torch.float32, device=dev) if want_logits else None
        with torch.cuda.device(dev):
            check(lib.stil_infonce_fwd(ptr(a), ptr(b), ptr(a_all), ptr(b_all), dtype_code(a), m, n, d, d, off,
                                       float(temperature), float(lambda_0), ptr(loss_sum), ptr(lse_row), ptr(lse_col),
                                       ptr(logits), n, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        if gather is not None:
            loss_sum = gather.all_reduce_sum(loss_sum)
            lse_row_all, lse_col_all = gather.gather_rows(lse_row), gather.gather_rows(lse_col)
        else:
            lse_row_all, lse_col_all = lse_row, lse_col
        ctx.save_for_backward(a, b, a_all, b_all, lse_row_all, lse_col_all)
        ctx.meta = (float(temperature), float(lambda_0), off, n, out0.dtype, out1.dtype)
        ctx.mark_non_differentiable(*([logits] if logits is not None else []))
        loss = loss_sum.reshape(())
        if logits is None:
            return loss, torch.empty(0, device=dev)
        return loss, logits

    @staticmethod
    def backward(ctx, grad_loss, _grad_logits):
        a, b, a_all, b_all, lse_row_all, lse_col_all = ctx.saved_tensors
        temperature, lambda_0, off, n, dt0, dt1 = ctx.meta
        dev = a.device
        m, d = a.shape
        lib = _lib.load()
        ws_bytes = lib.stil_infonce_workspace_bytes(m, n, d, dtype_code(a))
        ws = _lib.workspace(dev, "infonce", ws_bytes)
        g = grad_loss.detach().to(torch.float32).contiguous()
        # gradients are formed and written in fp32 whatever the embedding dtype (dLoss/dLogits travels as a bf16 hi+lo
        # pair); autograd's contract then rounds them ONCE to the dtype of the inputs
        d_a, d_b = torch.empty_like(a, dtype=torch.float32), torch.empty_like(b, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(lib.stil_infonce_bwd(ptr(a), ptr(b), ptr(a_all), ptr(b_all), dtype_code(a), m, n, d, d, off,
                                       temperature, lambda_0, ptr(lse_row_all), ptr(lse_col_all), ptr(g), ptr(d_a),
                                       ptr(d_b), dtype_code(d_a), d, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return d_a.to(dt0), d_b.to(dt1), None, None, None, None


class CLIPLoss(nn.Module):
    """Drop-in for ``utils/clip_loss.py:CLIPLoss`` (cross-modal InfoNCE, "ITC").

    ``forward(out0, out1, indices=None) -> (loss, logits, labels)`` like the reference (:27, :40).
    ``logits`` is returned detached (the reference only consumes it for validation top-k,
    ``STiLModel.py:437-438``); pass ``return_logits=False`` to skip materialising the B x B matrix.

    Extension (not in the reference): ``gather=GlobalBatch(group)`` all-gathers both
    embeddings over NCCL so the logits see the global batch; the returned loss is the global loss
    (identical on every rank) and the gradient is d(global loss)/d(local rows).
    """

    def __init__(self, temperature: float, lambda_0: float = 0.5, return_logits: bool = True, gather=None) -> None:
        super().__init__()
        self.temperature = temperature
        if lambda_0 > 1 or lambda_0 < 0:
            raise ValueError('lambda_0 must be a float between 0 and 1.')
        self.lambda_0 = lambda_0
        self.lambda_1 = 1 - lambda_0
        self.return_logits = return_logits
        self.gather = gather

    def forward(self, out0: torch.Tensor, out1: torch.Tensor, indices: List[int] = None) -> Tuple:
        loss, logits = _InfoNCEFn.apply(out0, out1, self.temperature, self.lambda_0, self.return_logits, self.gather)
        off = 0 if self.gather is None else self.gather.row_offset(len(out0))
        labels = torch.arange(len(out0), device=out0.device) + off
        return loss, (logits if self.return_logits else None), labels


def label_argmax(label: torch.Tensor, threshold: float):
    """(cls int32, conf bool, max_prob f32) of a dense soft label — prototype_loss.py:31-32 / STiLModel.py:204-205."""
    dev = _lib.require_cuda(label)
    _lib.ensure_device(dev)
    if label.dim() != 2:
        raise ValueError("label must be [B, K]")
    lab = label.detach()
    if lab.dtype != torch.float32:
        lab = lab.float()          # the reference's own smoke test passes integer labels (prototype_loss.py:43)
    lab = lab.contiguous()
    rows, k = lab.shape
    cls = torch.empty(rows, dtype=torch.int32, device=dev)
    conf = torch.empty(rows, dtype=torch.bool, device=dev)
    maxp = torch.empty(rows, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().stil_label_argmax(ptr(lab), k, rows, k, float(threshold), ptr(cls), ptr(conf), ptr(maxp),
                                            _lib.stream_ptr(dev)))
    return cls, conf, maxp


class _ProtoCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, prototypes, cls, conf, temperature):
        f = _prep_embed(feat)
        protos = prototypes.detach().to(torch.float32).contiguous()
        dev = _lib.require_cuda(f, protos, cls, conf)
        _lib.ensure_device(dev)
        rows, d = f.shape
        k = protos.shape[0]
        if protos.shape[1] != d:
            raise ValueError(f"prototypes [K, {protos.shape[1]}] do not match feat [B, {d}]")
        lib = _lib.load()
        ws = _lib.workspace(dev, "proto", lib.stil_proto_ce_workspace_bytes(rows, k, d, dtype_code(f)))
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        lse = torch.empty(rows, dtype=torch.float32, device=dev)
        w = torch.empty(rows, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.stil_proto_ce_fwd(ptr(f), dtype_code(f), rows, d, d, ptr(protos), k, ptr(cls), ptr(conf),
                                        float(temperature), ptr(loss), ptr(lse), ptr(w), ptr(ws), ws.numel(),
                                        _lib.stream_ptr(dev)))
        ctx.save_for_backward(f, protos, cls, lse, w)
        ctx.meta = (float(temperature), feat.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        f, protos, cls, lse, w = ctx.saved_tensors
        temperature, dt = ctx.meta
        dev = f.device
        rows, d = f.shape
        k = protos.shape[0]
        lib = _lib.load()
        ws = _lib.workspace(dev, "proto", lib.stil_proto_ce_workspace_bytes(rows, k, d, dtype_code(f)))
        g = grad_loss.detach().to(torch.float32).contiguous()
        d_f = torch.empty_like(f, dtype=torch.float32)      # fp32 gradient, rounded once to the input dtype below
        with torch.cuda.device(dev):
            check(lib.stil_proto_ce_bwd(ptr(f), dtype_code(f), rows, d, d, ptr(protos), k, ptr(cls), ptr(lse), ptr(w),
                                        temperature, ptr(g), ptr(d_f), dtype_code(d_f), d, ptr(ws), ws.numel(),
                                        _lib.stream_ptr(dev)))
        return d_f.to(dt), None, None, None, None


class PrototypeLoss(nn.Module):
    """Drop-in for ``utils/prototype_loss.py:PrototypeLoss``.

    ``forward(label, prototypes, feat) -> loss`` (:24, :40).  ``label`` is the dense [B, K] soft label
    (``pseudo_label_all``, ``STiLModel.py:321``); only its row max / argmax are used (:31-32), ``prototypes``
    and ``label`` receive no gradient, ``feat`` does.
    """

    def __init__(self, temperature, threshold) -> None:
        super().__init__()
        self.temperature = temperature
        self.threshold = threshold

    def forward(self, label: torch.Tensor, prototypes: torch.Tensor, feat: torch.Tensor):
        cls, conf, _ = label_argmax(label, self.threshold)
        return _ProtoCEFn.apply(feat, prototypes, cls, conf, self.temperature)

    def forward_hard(self, cls: torch.Tensor, conf: torch.Tensor, prototypes: torch.Tensor, feat: torch.Tensor):
        """Same loss from pre-reduced (class index int32, confident bool) rows — what the fused head uses."""
        return _ProtoCEFn.apply(feat, prototypes, cls.to(torch.int32), conf.to(torch.bool), self.temperature)


class _MaskedSoftCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_m, y_i, y_t, pseudo_label, mask1, case1, case2_i, case2_t, case3, mask_random):
        ys = [y.contiguous() if y.dtype in (torch.float32, torch.bfloat16) else y.float().contiguous()
              for y in (y_m, y_i, y_t)]
        if len({y.dtype for y in ys}) != 1:
            ys = [y.float() for y in ys]
        pl = pseudo_label.detach().to(torch.float32).contiguous()
        flags = [t.to(torch.bool).contiguous() for t in (mask1, case1, case2_i, case2_t, case3, mask_random)]
        dev = _lib.require_cuda(*ys, pl, *flags)
        _lib.ensure_device(dev)
        rows, k = ys[0].shape
        lib = _lib.load()
        ws = _lib.workspace(dev, "softce", lib.stil_masked_softce_workspace_bytes(rows))
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        grads = [torch.empty(rows, k, dtype=torch.float32, device=dev) for _ in range(3)]
        with torch.cuda.device(dev):
            check(lib.stil_masked_softce(ptr(ys[0]), ptr(ys[1]), ptr(ys[2]), dtype_code(ys[0]), k, ptr(pl), k,
                                         *[ptr(f) for f in flags], rows, k, ptr(losses), ptr(grads[0]), ptr(grads[1]),
                                         ptr(grads[2]), k, 1.0, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.save_for_backward(*grads)
        ctx.dts = (y_m.dtype, y_i.dtype, y_t.dtype)
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_m, g_i, g_t):
        grads = ctx.saved_tensors
        outs = [(g * u).to(dt) for g, u, dt in zip((g_m, g_i, g_t), grads, ctx.dts)]
        return (*outs, None, None, None, None, None, None, None)


def masked_soft_ce(y_m_u, y_i_u, y_t_u, pseudo_label, mask1, case1, case2_i, case2_t, case3, mask_random):
    """The three masked soft-target CE losses of ``STiLModel.py:301-303`` on the unlabelled student logits:
    ``(F.cross_entropy(y, pseudo_label, 'none') * mask1 * case_weight).mean()`` for m / i / t."""
    return _MaskedSoftCEFn.apply(y_m_u, y_i_u, y_t_u, pseudo_label, mask1, case1, case2_i, case2_t, case3, mask_random)
