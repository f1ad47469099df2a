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
