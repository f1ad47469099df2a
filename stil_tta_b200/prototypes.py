"""Prototype bank of PGLS: per-class partial sums, running accumulators and the epoch-end replace.

Mirrors ``STiLModel.cal_prototypes`` (``models/Disentangle/STiLModel.py:199-214``),
``cal_prototypes_separate`` (``:216-226``), the buffers ``prototypes`` / ``prototypes_sum`` /
``prototypes_count_sum`` (``:94-96``), the per-step accumulate (``:374-381``) and the epoch-end
finalise (``:408-415``).  NB the reference does NOT EMA the prototypes (``prototype_momentum`` is a
dead hparam): it replaces them with ``sum / count`` once per epoch — that is the behaviour here.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import check, dtype_code, ptr
from .losses import label_argmax


def _accumulate(feat, cls, conf, b_l, repeat_ratio, k, psum=None, pcount=None):
    f = feat.detach()
    if f.dtype not in (torch.float32, torch.bfloat16):
        f = f.float()
    f = f.contiguous()
    dev = _lib.require_cuda(f, cls, conf, psum, pcount)
    _lib.ensure_device(dev)
    rows, d = f.shape
    class_sum = torch.empty(k, d, dtype=torch.float32, device=dev)
    class_count = torch.empty(k, 1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().stil_proto_accumulate(ptr(f), dtype_code(f), rows, d, d, ptr(cls), ptr(conf), int(b_l),
                                                float(repeat_ratio), k, ptr(class_sum), ptr(class_count), ptr(psum),
                                                ptr(pcount), _lib.stream_ptr(dev)))
    return class_sum, class_count


@torch.no_grad()
def cal_prototypes(label: torch.Tensor, feat: torch.Tensor, th1: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(class_sum [K,P], class_count [K,1])`` over rows whose ``label.max(1) >= th1`` (STiLModel.py:199-214)."""
    cls, conf, _ = label_argmax(label, th1)
    return _accumulate(feat, cls, conf, 0, 1.0, label.shape[1])


@torch.no_grad()
def cal_prototypes_separate(label: torch.Tensor, feat: torch.Tensor, B_l: int, th1: float, repeat_ratio: float
                            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Labelled rows ``[:B_l]`` weigh ``1/repeat_ratio`` (STiLModel.py:216-226)."""
    cls, conf, _ = label_argmax(label, th1)
    return _accumulate(feat, cls, conf, B_l, repeat_ratio, label.shape[1])


class PrototypeBank(nn.Module):
    """Holds the three reference buffers under their reference names so checkpoints round-trip
    (``STiLModel.py:94-96``); methods are the reference's call sites."""

    def __init__(self, num_classes: int, projection_dim: int, th1: float, repeat_ratio: float = 1.0,
                 process_group=None) -> None:
        super().__init__()
        self.th1, self.repeat_ratio, self.process_group = th1, repeat_ratio, process_group
        self.register_buffer("prototypes", torch.zeros(num_classes, projection_dim))
        self.register_buffer("prototypes_sum", torch.zeros(num_classes, projection_dim))
        self.register_buffer("prototypes_count_sum", torch.zeros(num_classes, 1))
        self.register_buffer("empty_classes", torch.zeros(1, dtype=torch.int32), persistent=False)

    def cal_prototypes(self, label, feat):
        return cal_prototypes(label, feat, self.th1)

    def cal_prototypes_separate(self, label, feat, B_l):
        return cal_prototypes_separate(label, feat, B_l, self.th1, self.repeat_ratio)

    def _distributed(self) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1

    @torch.no_grad()
    def update(self, label: torch.Tensor, feat: torch.Tensor, B_l: int, cls=None, conf=None):
        """One training step's ``STiLModel.py:374-381``: partial sums (+ all-reduce when distributed) added into
        the running buffers.  Returns ``(class_sum, class_count)`` (after the all-reduce)."""
        if cls is None:
            cls, conf, _ = label_argmax(label, self.th1)
        k = self.prototypes.shape[0]
        if not self._distributed():
            return _accumulate(feat, cls, conf, B_l, self.repeat_ratio, k, self.prototypes_sum,
                               self.prototypes_count_sum)
        from .distributed import all_reduce_prototype_partials
        cs, cc = _accumulate(feat, cls, conf, B_l, self.repeat_ratio, k)
        cs, cc = all_reduce_prototype_partials(cs, cc, self.process_group)     # :377-379, one packed collective
        dev = cs.device
        with torch.cuda.device(dev):
            check(_lib.load().stil_proto_add(ptr(cs), ptr(cc), k, cs.shape[1], ptr(self.prototypes_sum),
                                             ptr(self.prototypes_count_sum), _lib.stream_ptr(dev)))
        return cs, cc

    @torch.no_grad()
    def finalize(self, check_empty: bool = False) -> torch.Tensor:
        """Epoch end (``STiLModel.py:408-415``): ``prototypes = sum / count``, zero the accumulators.  The
        reference host-asserts that every class was seen (:411-412); here the count of empty classes is left
        in the device flag ``empty_classes`` and only read (one sync) when ``check_empty``."""
        dev = _lib.require_cuda(self.prototypes)
        k, d = self.prototypes.shape
        with torch.cuda.device(dev):
            check(_lib.load().stil_proto_finalize(ptr(self.prototypes), ptr(self.prototypes_sum),
                                                  ptr(self.prototypes_count_sum), k, d, ptr(self.empty_classes),
                                                  _lib.stream_ptr(dev)))
        if check_empty:
            assert int(self.empty_classes) == 0, "a class received no confident sample this epoch"
        return self.empty_classes
