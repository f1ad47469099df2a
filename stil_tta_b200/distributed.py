"""Data-parallel plumbing of the head over ``torch.distributed`` (NCCL on NVLink; gloo in CPU tests).

Two exchange steps exist on the path:

* the prototype partials ``class_sum [K,P]`` / ``class_count [K,1]`` are summed over ranks before they
  are accumulated (``STiLModel.py:377-379`` issues two all-reduces; here ONE on a packed ``[K, P+1]``);
* (extension) global-batch InfoNCE: both embeddings are all-gathered so each rank scores its rows against
  every column; row- and column-LSEs of the local rows are computed locally and only two ``[m]`` vectors are
  gathered for the backward — no reduce-scatter of gradients is needed because each rank computes both the
  ``a``-side and the ``b``-side gradient of its own rows.
Everything else in the head is row-independent and needs no communication.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def all_reduce_prototype_partials(class_sum: torch.Tensor, class_count: torch.Tensor, group=None
                                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """SUM over ranks of both partials with a single collective on a packed [K, P+1] buffer."""
    k, p = class_sum.shape
    packed = torch.empty(k, p + 1, dtype=class_sum.dtype, device=class_sum.device)
    packed[:, :p] = class_sum
    packed[:, p:] = class_count.reshape(k, 1)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[:, :p].contiguous(), packed[:, p:].contiguous()


class GlobalBatch:
    """Row bookkeeping + collectives for the global-batch InfoNCE (equal rows per rank)."""

    def __init__(self, group=None) -> None:
        self.group = group

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    def total_rows(self, local_rows: int) -> int:
        return local_rows * self.world_size

    def row_offset(self, local_rows: int) -> int:
        return local_rows * self.rank

    def gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """[m, ...] per rank -> [W*m, ...], rank-major (no gradient: the backward is analytic)."""
        if self.world_size == 1:
            return t
        t = t.detach().contiguous()
        out = torch.empty((self.world_size * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def sharded_infonce(a_loc: torch.Tensor, b_loc: torch.Tensor, temperature: float, lambda_0: float, gb: GlobalBatch,
                    fwd_local, bwd_local):
    """The communication schedule of the global-batch InfoNCE with the per-rank compute injected
    (``fwd_local`` / ``bwd_local``), so that the schedule itself can be exercised on CPU with gloo.

    fwd_local(a_all, b_all, off, m) -> (loss_sum_local, lse_row_local, lse_col_local)
    bwd_local(a_all, b_all, off, m, lse_row_all, lse_col_all) -> (d_a_local, d_b_local)
    Returns (global_loss, d_a_local, d_b_local)."""
    m = a_loc.shape[0]
    off = gb.row_offset(m)
    a_all, b_all = gb.gather_rows(a_loc), gb.gather_rows(b_loc)
    loss_sum, lse_row, lse_col = fwd_local(a_all, b_all, off, m)
    loss = gb.all_reduce_sum(loss_sum.clone())
    lse_row_all, lse_col_all = gb.gather_rows(lse_row), gb.gather_rows(lse_col)
    d_a, d_b = bwd_local(a_all, b_all, off, m, lse_row_all, lse_col_all)
    return loss, d_a, d_b
