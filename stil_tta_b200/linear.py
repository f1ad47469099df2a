"""The Linear layers either side of the head on the sm_100a tensor cores.

``projector_imaging`` / ``projector_tabular`` are ``nn.Linear(multimodal_embedding_dim, projection_dim)`` followed by
``F.normalize`` in ``project_3features`` (``models/Disentangle/STiLModel.py:56-63, 182-192``); the three classifiers are
``nn.Linear(hidden*3 | hidden*2, num_classes)`` (``models/Disentangle/utils/STiLModel_backbone.py:66-68, 153-155``).
``Linear`` keeps ``nn.Linear``'s parameter names, shapes and initialisation (``weight [out, in]``, ``bias [out]``), so state
dicts round-trip; ``normalize=True`` fuses ``F.normalize`` into the GEMM epilogue.  Forward and backward (d_x, d_weight,
d_bias) run through ``stil_linear_fwd`` / ``stil_linear_bwd`` (include/stil_head.h): fp32-accurate products from bf16
tensor-core passes, no CPU / eager fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import check, dtype_code, ptr


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, normalize):
        xc = x.detach()
        if xc.dtype not in (torch.float32, torch.bfloat16):
            xc = xc.float()
        xc = xc.contiguous()
        w = weight.detach().to(torch.float32).contiguous()
        b = None if bias is None else bias.detach().to(torch.float32).contiguous()
        dev = _lib.require_cuda(xc, w, b)
        _lib.ensure_device(dev)
        if xc.dim() != 2 or w.dim() != 2 or xc.shape[1] != w.shape[1]:
            raise ValueError(f"Linear expects x [B, {w.shape[1]}], got {tuple(xc.shape)}")
        rows, din = xc.shape
        dout = w.shape[0]
        lib = _lib.load()
        code = dtype_code(xc)
        ws = _lib.workspace(dev, "linear", lib.stil_linear_workspace_bytes(rows, din, dout, code))
        y = torch.empty(rows, dout, dtype=torch.float32, device=dev)
        y_raw = torch.empty(rows, dout, dtype=torch.float32, device=dev) if normalize else None
        inv = torch.empty(rows, dtype=torch.float32, device=dev) if normalize else None
        with torch.cuda.device(dev):
            check(lib.stil_linear_fwd(ptr(xc), code, rows, din, din, ptr(w), ptr(b), dout, int(bool(normalize)), ptr(y), dout,
                                      ptr(y_raw), ptr(inv), ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.save_for_backward(xc, w, y_raw, inv)
        ctx.meta = (x.dtype, bias is not None, weight.dtype, None if bias is None else bias.dtype)
        return y

    @staticmethod
    def backward(ctx, d_y):
        xc, w, y_raw, inv = ctx.saved_tensors
        x_dtype, has_bias, w_dtype, b_dtype = ctx.meta
        dev = xc.device
        rows, din = xc.shape
        dout = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        g = d_y.detach().to(torch.float32).contiguous()
        lib = _lib.load()
        code = dtype_code(xc)
        ws = _lib.workspace(dev, "linear", lib.stil_linear_workspace_bytes(rows, din, dout, code))
        d_x = torch.empty(rows, din, dtype=torch.float32, device=dev) if need_x else None
        d_w = torch.empty(dout, din, dtype=torch.float32, device=dev) if need_w else None
        d_b = torch.empty(dout, dtype=torch.float32, device=dev) if need_b else None
        if need_x or need_w or need_b:
            with torch.cuda.device(dev):
                check(lib.stil_linear_bwd(ptr(xc), code, rows, din, din, ptr(w), dout, ptr(y_raw), ptr(inv), ptr(g), dout, ptr(d_x),
                                          _lib.STIL_F32, din, ptr(d_w), ptr(d_b), ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return (None if d_x is None else d_x.to(x_dtype), None if d_w is None else d_w.to(w_dtype),
                None if d_b is None else d_b.to(b_dtype), None)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, normalize: bool = False) -> torch.Tensor:
    """``F.linear(x, weight, bias)`` (fp32 output), followed by ``F.normalize(., dim=1)`` when ``normalize``."""
    return _LinearFn.apply(x, weight, bias, normalize)


class Linear(nn.Module):
    """Drop-in for the ``nn.Linear`` projectors / classifiers around the head; ``normalize=True`` = ``Linear`` + ``F.normalize``
    (``project_3features``, ``STiLModel.py:182-192``; needs ``out_features <= 128``)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, normalize: bool = False, device=None) -> None:
        super().__init__()
        if normalize and out_features > 128:
            raise ValueError("the fused F.normalize needs out_features <= 128")
        self.in_features, self.out_features, self.normalize = in_features, out_features, normalize
        self.weight = nn.Parameter(torch.empty(out_features, in_features, device=device))
        self.bias = nn.Parameter(torch.empty(out_features, device=device)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:          # nn.Linear's own initialisation
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features) if self.in_features > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return linear(x, self.weight, self.bias, self.normalize)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, normalize={self.normalize}"
