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


class P2PBuffer:
    """One zeroed device buffer per rank, mapped into every peer of the node with CUDA IPC, plus the
    ``stil_p2p_exchange`` all-gather over it (remote NVLink stores + arrival flags; see csrc/p2p.cu).

    Layout: ``[flags 512 B | control 768 B | pad to 4096 | user bytes]``.  ``view(offset, shape, dtype)`` gives a
    torch tensor over the local buffer; ``exchange(channel, [(src_tensor, dst_offset_bytes), ...])`` stores the
    segments at the same offsets of every rank's buffer."""

    HEADER = 4096

    def __init__(self, user_bytes: int, device: torch.device, group=None) -> None:
        import ctypes as C
        from . import _lib
        self._C, self._lib, self.group, self.dev = C, _lib, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("P2PBuffer supports up to 8 ranks (one NVSwitch node)")
        self.nbytes = self.HEADER + (int(user_bytes) + 255) // 256 * 256
        lib = _lib.load()
        with torch.cuda.device(device):
            p = C.c_void_p()
            _lib.check(lib.stil_p2p_alloc(self.nbytes, C.byref(p)))
            self.local = p.value
            h = C.create_string_buffer(64)
            _lib.check(lib.stil_p2p_export(self.local, h))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(h.raw), group=group)
            self.peers = []
            for r, hb in enumerate(handles):
                if r == self.rank:
                    self.peers.append(self.local)
                else:
                    q = C.c_void_p()
                    _lib.check(lib.stil_p2p_import(hb, C.byref(q)))
                    self.peers.append(q.value)
        self._bases = (C.c_void_p * self.world)(*self.peers)
        dist.barrier(group=group)

    def view(self, offset: int, shape, dtype) -> torch.Tensor:
        """Tensor over ``[HEADER + offset, ...)`` of the LOCAL buffer."""
        n = 1
        for s in shape:
            n *= int(s)
        esz = torch.tensor([], dtype=dtype).element_size()
        tstr = {torch.float32: "<f4", torch.bfloat16: "<u2", torch.uint8: "|u1", torch.int64: "<i8"}[dtype]

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (n,), "typestr": tstr, "data": (self.local + self.HEADER + offset, False),
                                        "version": 3}
        t = torch.as_tensor(raw, device=self.dev)
        if dtype == torch.bfloat16:
            t = t.view(torch.bfloat16)
        assert t.numel() * esz == n * esz
        return t.view(*shape)

    def exchange(self, channel: int, segments, stream_ptr: int) -> None:
        C, lib = self._C, self._lib.load()
        n = len(segments)
        src = (C.c_void_p * n)(*[t.data_ptr() for t, _ in segments])
        nb = (C.c_int64 * n)(*[t.numel() * t.element_size() for t, _ in segments])
        off = (C.c_int64 * n)(*[self.HEADER + o for _, o in segments])
        self._lib.check(lib.stil_p2p_exchange(self._bases, self.world, self.rank, 0, 512, channel, n, src, nb, off,
                                              stream_ptr))

    def close(self) -> None:
        if self.local is None:
            return
        lib = self._lib.load()
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.dev):
            for r, p in enumerate(self.peers):
                if r != self.rank:
                    lib.stil_p2p_close(p)
            dist.barrier(group=self.group)
            lib.stil_p2p_free(self.local)
        self.local = None
