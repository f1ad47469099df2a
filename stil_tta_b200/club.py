"""CLUBMean mutual-information bound — drop-in for ``models/Disentangle/utils/club.py`` ``CLUBMean``
(constructor ``:88-103``, ``forward`` ``:107-121``, ``learning_loss`` ``:125-130``; used ``STiLModel.py:67-68, 327-330``).

``p_mu`` stays a torch module (its weights are the reference's parameters, so checkpoints load unchanged); what is
replaced is the tensor code after it.  The reference materialises a ``B x B x D`` broadcast (512 MB at B = D = 512); the
bound collapses algebraically to column sums, ``sum_i mu_i.y_i / B - (sum_i mu_i).(sum_j y_j) / B^2``, computed by two small
kernels in O(B D).  No CPU fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from ._lib import check, dtype_code, ptr


class _ClubFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, y, want_bound, want_est):
        dev = _lib.require_cuda(mu, y)
        _lib.ensure_device(dev)
        m, v = mu.detach(), y.detach()
        if m.dtype not in (torch.float32, torch.bfloat16) or v.dtype != m.dtype:
            m, v = m.float(), v.float()
        m, v = m.contiguous(), v.contiguous()
        if m.shape != v.shape or m.dim() != 2:
            raise ValueError("CLUBMean: mu and y_samples must be [B, D] of the same shape")
        rows, dim = m.shape
        stats = torch.empty(4 * dim, dtype=torch.float32, device=dev)
        out = torch.zeros(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().stil_club_fwd(ptr(m), ptr(v), dtype_code(m), rows, dim, dim, ptr(stats),
                                            ptr(out[0:1]) if want_bound else None, ptr(out[1:2]) if want_est else None,
                                            _lib.stream_ptr(dev)))
        ctx.save_for_backward(m, v, stats)
        ctx.meta = (mu.dtype, y.dtype, want_bound, want_est)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_bound, g_est):
        m, v, stats = ctx.saved_tensors
        mu_dtype, y_dtype, want_bound, want_est = ctx.meta
        dev = m.device
        rows, dim = m.shape
        gb = g_bound.detach().float().reshape(1).contiguous() if want_bound else None
        ge = g_est.detach().float().reshape(1).contiguous() if want_est else None
        d_mu = torch.empty(rows, dim, dtype=torch.float32, device=dev)
        d_y = torch.empty(rows, dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().stil_club_bwd(ptr(m), ptr(v), dtype_code(m), rows, dim, dim, ptr(stats), ptr(gb), ptr(ge),
                                            ptr(d_mu), ptr(d_y), dim, _lib.stream_ptr(dev)))
        return d_mu.to(mu_dtype), d_y.to(y_dtype), None, None


def club_bound(mu: torch.Tensor, y_samples: torch.Tensor) -> torch.Tensor:
    """``CLUBMean.forward`` from ``mu = p_mu(x_samples)`` on (``club.py:107-121``)."""
    return _ClubFn.apply(mu, y_samples, True, False)[0]


def club_learning_loss(mu: torch.Tensor, y_samples: torch.Tensor) -> torch.Tensor:
    """``CLUBMean.learning_loss`` from ``mu`` on (``club.py:125-130``)."""
    return _ClubFn.apply(mu, y_samples, False, True)[1]


def club_both(mu: torch.Tensor, y_samples: torch.Tensor):
    """``(forward, learning_loss)`` of the same pair in one pass (the trainer calls both, ``STiLModel.py:327-330``)."""
    return _ClubFn.apply(mu, y_samples, True, True)


class CLUBMean(nn.Module):
    """Same constructor, parameters (``p_mu``) and methods as the reference ``CLUBMean``."""

    def __init__(self, x_dim, y_dim, hidden_size=512):
        super().__init__()
        if hidden_size is None:
            self.p_mu = nn.Linear(x_dim, y_dim)
        else:
            self.p_mu = nn.Sequential(nn.Linear(x_dim, int(hidden_size)), nn.ReLU(), nn.Linear(int(hidden_size), y_dim))

    def get_mu_logvar(self, x_samples):
        return self.p_mu(x_samples), 0

    def forward(self, x_samples, y_samples):
        return club_bound(self.p_mu(x_samples), y_samples)

    def loglikeli(self, x_samples, y_samples):
        return -club_learning_loss(self.p_mu(x_samples), y_samples)

    def learning_loss(self, x_samples, y_samples):
        return club_learning_loss(self.p_mu(x_samples), y_samples)
