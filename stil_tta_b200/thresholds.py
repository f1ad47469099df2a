"""Pseudo-label thresholds of the remaining baselines, on the device (``csrc/threshold_kernels.cu``):

* ``FreeMatchThreshold`` — the self-adaptive threshold state and ``masking`` of ``FreeMatchModel``
  (``models/MatchModel/FreeMatchFolder/freematch_model.py:49-53`` state, ``:128-144`` ``update``, ``:146-165`` ``masking``);
* ``entropy_loss`` — the fairness loss of ``FreeMatchFolder/freematch_utils.py:17-45`` with its gradient;
* ``cotraining_pseudo_labels`` — ``models/SemiMultimodal/CoTraining.py:141-146`` (the two unsupervised losses of ``:148-149``
  are ``masked_ce(y_hat_i[B_l:], pseudo_label_t, mask_t)`` and ``masked_ce(y_hat_t[B_l:], pseudo_label_i, mask_i)``).

CUDA only: a CPU tensor raises (there is no fallback path).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return (t if t.dtype == torch.float32 else t.float()).contiguous()


def _ws(dev: torch.device, rows: int, c: int) -> torch.Tensor:
    return _lib.workspace(dev, "thresholds", _lib.load().stil_threshold_workspace_bytes(rows, c))


class FreeMatchThreshold:
    """Device-resident ``time_p`` / ``p_model`` / ``label_hist`` (the reference keeps them as plain tensor attributes and moves
    them to the GPU on first use, ``:148-153``) and ``masking`` = ``update`` + mask of one unlabelled batch.

    ``group``: when given (and the process group is initialised), the batch statistics of all ranks are summed before the
    update — what ``concat_all_gather`` achieves in the reference (``:129-130``) at a fraction of the traffic."""

    def __init__(self, num_classes: int, momentum: float = 0.999, clip_thresh: float = 0.0, device="cuda", group=None) -> None:
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("FreeMatchThreshold runs on a CUDA device (sm_100a) only; there is no CPU fallback")
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        _lib.ensure_device(self.dev)
        self.num_classes, self.m, self.clip_thresh, self.group = num_classes, float(momentum), float(clip_thresh), group
        self.p_model = torch.full((num_classes,), 1.0 / num_classes, dtype=torch.float32, device=self.dev)       # :51
        self.label_hist = torch.full((num_classes,), 1.0 / num_classes, dtype=torch.float32, device=self.dev)    # :52
        self.time_p = self.p_model.mean().reshape(1).clone()                                                      # :53
        self.max_probs: Optional[torch.Tensor] = None
        self.max_idx: Optional[torch.Tensor] = None

    @torch.no_grad()
    def masking(self, logits_x_ulb: torch.Tensor, softmax_x_ulb: bool = True) -> torch.Tensor:
        """``mask`` [rows] float 0/1 (``:165``); ``self.max_probs`` / ``self.max_idx`` hold ``probs.max(dim=-1)`` (``:161``)."""
        _lib.require_cuda(logits_x_ulb)
        lib = _lib.load()
        x = _f32c(logits_x_ulb)
        rows, c = x.shape
        if c != self.num_classes:
            raise ValueError("class count does not match the threshold state")
        f32 = dict(dtype=torch.float32, device=self.dev)
        stats = torch.empty(2 * c + 2, **f32)
        self.max_probs = torch.empty(rows, **f32)
        self.max_idx = torch.empty(rows, dtype=torch.int64, device=self.dev)
        mask = torch.empty(rows, **f32)
        ws = _ws(self.dev, rows, c)
        s = _lib.stream_ptr(self.dev)
        with torch.cuda.device(self.dev):
            check(lib.stil_freematch_stats(ptr(x), x.stride(0), rows, c, 1 if softmax_x_ulb else 0, ptr(stats), ptr(self.max_probs),
                                           ptr(self.max_idx), None, 0, ptr(ws), ws.numel(), s))
            dist = torch.distributed
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(stats, group=self.group)
            check(lib.stil_freematch_update_mask(ptr(stats), rows, c, self.m, self.clip_thresh, ptr(self.time_p), ptr(self.p_model),
                                                 ptr(self.label_hist), ptr(self.max_probs), ptr(self.max_idx), ptr(mask), ptr(ws),
                                                 ws.numel(), s))
        return mask


class _EntropyLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask, logits_s, prob_model, label_hist):
        dev = _lib.require_cuda(mask, logits_s, prob_model, label_hist)
        _lib.ensure_device(dev)
        lib = _lib.load()
        x, mk = _f32c(logits_s), _f32c(mask)
        rows, c = x.shape
        out = torch.empty(2, dtype=torch.float32, device=dev)
        # the workspace carries probabilities and the gradient vector to the backward: private to this call
        ws = torch.empty(lib.stil_threshold_workspace_bytes(rows, c), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.stil_freematch_entropy_fwd(ptr(mk), ptr(x), x.stride(0), rows, c, ptr(_f32c(prob_model)), ptr(_f32c(label_hist)),
                                                 ptr(out[0:1]), ptr(out[1:2]), ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.ws, ctx.shape, ctx.in_dtype = ws, (rows, c), logits_s.dtype
        loss, hist_mean = out[0].clone(), out[1].clone()
        ctx.mark_non_differentiable(hist_mean)
        return loss, hist_mean

    @staticmethod
    def backward(ctx, g_loss, _g_hist):
        rows, c = ctx.shape
        dev = ctx.ws.device
        g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
        d = torch.empty(rows, c, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().stil_freematch_entropy_bwd(rows, c, ptr(g), ptr(d), c, ptr(ctx.ws), ctx.ws.numel(), _lib.stream_ptr(dev)))
        return None, d.to(ctx.in_dtype), None, None


def entropy_loss(mask: torch.Tensor, logits_s: torch.Tensor, prob_model: torch.Tensor, label_hist: torch.Tensor
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``entropy_loss`` of ``freematch_utils.py:17-45``: ``(loss, hist_s.mean())``; only ``logits_s`` gets a gradient.  Like the
    reference it must not be called with an all-zero mask (``freematch_model.py:196-199`` guards it)."""
    return _EntropyLossFn.apply(mask, logits_s, prob_model, label_hist)


@torch.no_grad()
def threshold_rows(logits: torch.Tensor, threshold: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``probs = softmax(logits)``, ``max_probs, max_idx = probs.max(1)``, ``mask = max_probs >= threshold`` (float 0/1)."""
    dev = _lib.require_cuda(logits)
    _lib.ensure_device(dev)
    x = _f32c(logits)
    rows, c = x.shape
    probs = torch.empty(rows, c, dtype=torch.float32, device=dev)
    max_p = torch.empty(rows, dtype=torch.float32, device=dev)
    max_i = torch.empty(rows, dtype=torch.int64, device=dev)
    mask = torch.empty(rows, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().stil_threshold_rows(ptr(x), x.stride(0), rows, c, float(threshold), ptr(probs), c, ptr(max_p), ptr(max_i),
                                              ptr(mask), _lib.stream_ptr(dev)))
    return probs, max_p, max_i, mask


@torch.no_grad()
def cotraining_pseudo_labels(y_hat_i_e: torch.Tensor, y_hat_t_e: torch.Tensor, threshold: float):
    """``CoTraining.py:141-146`` on the unlabelled rows of the two teacher heads:
    ``(pseudo_label_i, pseudo_label_t, mask_i, mask_t)`` — each modality's mask gates the OTHER modality's loss (``:148-149``)."""
    pl_i, _, _, mask_i = threshold_rows(y_hat_i_e, threshold)
    pl_t, _, _, mask_t = threshold_rows(y_hat_t_e, threshold)
    return pl_i, pl_t, mask_i, mask_t
