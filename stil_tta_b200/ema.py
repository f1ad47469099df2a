"""EMA teacher update — drop-in for ``STiLModel.momentum_update_ema`` (``models/Disentangle/STiLModel.py:154-168``).

The reference walks both state dicts in Python every step and issues ``mul_`` / ``add_`` (or ``copy_`` for
``num_batches_tracked``) per tensor: hundreds of tiny launches.  ``EmaTeacher`` builds a device-resident table of
``(ema, main, numel, dtype, kind)`` entries once and runs ONE kernel per step (``stil_ema_update``); fp32 results are
bit-identical to the reference (each product and the sum are rounded like the eager ops).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Tuple

import torch

from . import _lib
from ._lib import EmaEntry, check

CHUNK = 32768      # elements (bytes for copies) per block


class EmaTeacher:
    def __init__(self, pairs: Iterable[Tuple[torch.Tensor, torch.Tensor, bool]], momentum: float) -> None:
        """``pairs``: ``(ema_tensor, main_tensor, copy)`` — ``copy=True`` entries are copied verbatim (``:163-164``), the others
        follow ``ema.mul_(m).add_((1 - m) * main)`` (``:165-166`` / ``:168``).  Tensors must be contiguous CUDA tensors; the
        table keeps their addresses, so they must stay alive and in place (parameters and buffers of a module do)."""
        self.momentum = float(momentum)
        entries: List[EmaEntry] = []
        chunk_entry, chunk_start = [], []
        self._keep = []
        dev = None
        for ema, main, copy in pairs:
            dev = _lib.require_cuda(ema, main)
            if ema.shape != main.shape or ema.dtype != main.dtype:
                raise ValueError("state_dict shapes are different!")          # the reference's assert, :161
            if not (ema.is_contiguous() and main.is_contiguous()):
                raise ValueError("EmaTeacher needs contiguous tensors")
            if ema.numel() == 0:
                continue
            e = EmaEntry()
            e.ema, e.main = ema.data_ptr(), main.data_ptr()
            if copy or ema.dtype not in (torch.float32, torch.bfloat16):
                if not copy:
                    raise ValueError(f"EMA of dtype {ema.dtype} is not supported (float32 / bfloat16)")
                e.kind, e.dtype, e.numel = 1, 0, ema.numel() * ema.element_size()
            else:
                e.kind, e.dtype, e.numel = 0, _lib.dtype_code(ema), ema.numel()
            for start in range(0, e.numel, CHUNK):
                chunk_entry.append(len(entries))
                chunk_start.append(start)
            entries.append(e)
            self._keep.append((ema, main))
        self.n_entries, self.n_chunks = len(entries), len(chunk_entry)
        self.dev = dev
        if self.n_entries:
            _lib.ensure_device(dev)
            raw = (EmaEntry * self.n_entries)(*entries)
            host = torch.frombuffer(bytearray(bytes(raw)), dtype=torch.uint8)
            self._table = host.to(dev)
            self._chunk_entry = torch.tensor(chunk_entry, dtype=torch.int32, device=dev)
            self._chunk_start = torch.tensor(chunk_start, dtype=torch.int64, device=dev)

    @classmethod
    def from_modules(cls, model: torch.nn.Module, ema: torch.nn.Module, momentum: float, eman: bool = True) -> "EmaTeacher":
        """The two branches of ``momentum_update_ema``: ``eman`` — every state-dict entry, ``num_batches_tracked`` copied
        (``:156-166``); otherwise parameters only (``:167-168``)."""
        pairs = []
        if eman:
            sm, se = model.state_dict(), ema.state_dict()
            for (k_main, v_main), (k_ema, v_ema) in zip(sm.items(), se.items()):
                if k_main != k_ema:
                    raise ValueError("state_dict names are different!")       # the reference's assert, :160
                pairs.append((v_ema, v_main, "num_batches_tracked" in k_ema))
        else:
            for p_q, p_k in zip(model.parameters(), ema.parameters()):
                pairs.append((p_k.data, p_q.data, False))
        return cls(pairs, momentum)

    @torch.no_grad()
    def step(self) -> None:
        """One ``momentum_update_ema()`` — a single kernel launch on the current stream (CUDA-graph capturable)."""
        if not self.n_chunks:
            return
        with torch.cuda.device(self.dev):
            check(_lib.load().stil_ema_update(self._table.data_ptr(), self.n_entries, self._chunk_entry.data_ptr(),
                                              self._chunk_start.data_ptr(), self.n_chunks, CHUNK, self.momentum,
                                              _lib.stream_ptr(self.dev)))


def momentum_update_ema(model: torch.nn.Module, ema: torch.nn.Module, momentum: float, eman: bool = True, _cache={}) -> None:
    """Function form with the signature of the reference method's state (``self.model``, ``self.ema``, ``self.momentum``,
    ``self.eman``); the table is built on the first call for a given pair of modules."""
    key = (id(model), id(ema), float(momentum), bool(eman))
    upd = _cache.get(key)
    if upd is None:
        upd = _cache[key] = EmaTeacher.from_modules(model, ema, momentum, eman)
    upd.step()
