"""``STiLHead`` — the whole per-batch semi-supervised head as one enqueued step.

One call of :meth:`STiLHead.run` does what ``STiLModel.training_step`` does between the encoders and
``loss.backward()`` (``models/Disentangle/STiLModel.py:262-303, 317-322, 339, 374-381``): CGPL, PGLS,
the ITC (InfoNCE) and PT (prototype) losses with their gradients w.r.t. the embeddings, the masked
soft-target CE of the three student heads with its gradients, and the prototype partial sums added into
the running buffers — ~10 kernel launches through ``stil_head_step`` (include/stil_head.h), captured in
a CUDA graph.  All buffers are static (allocated once), so a replay has no host-side work.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import HeadStepArgs, check
from .synth import HeadConfig

_IN_EMBED = ("feat_i", "feat_t", "feat_m", "feat_m_e")
_IN_TEACHER = ("y_m_ue", "y_i_ue", "y_t_ue")
_IN_STUDENT = ("y_m", "y_i", "y_t")


class STiLHead:
    def __init__(self, cfg: HeadConfig, device="cuda", student_ce: bool = True, use_graph: bool = True,
                 rate_uce: float = 1.0, logit_dtype: torch.dtype = torch.float32,
                 grad_dtype: torch.dtype = torch.float32, da: bool = False, da_len: int = 256) -> None:
        """``grad_dtype``: dtype of ``d_feat_i/t/m``.  The default fp32 keeps dLoss/dLogits as a bf16 hi+lo pair and
        writes fp32 gradients, so bf16 *embeddings* (C2) still give gradients within 1e-3 of the fp32 reference;
        ``torch.bfloat16`` rounds both (2^-9) like autograd returning a bf16 ``.grad``.
        ``da``: ``hparams.DA == True`` (``STiLModel.py:276-277``) — ``prediction`` is the distribution-aligned
        ``softmax(y_m_ue)``; the ring buffer lives in ``DA_queue`` / ``DA_ptr`` (reference names, ``:98-100``)."""
        self.cfg = cfg
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("STiLHead runs on a CUDA device (sm_100a) only")
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        _lib.ensure_device(self.dev)
        self.student_ce, self.use_graph, self.rate_uce = student_ce, use_graph, rate_uce
        B, B_l, B_u, K, P = cfg.batch, cfg.b_l, cfg.b_u, cfg.num_classes, cfg.proj_dim
        edt = torch.bfloat16 if cfg.embed_dtype == "bf16" else torch.float32
        dev = self.dev
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=dev)
        # every per-batch input lives in ONE device allocation (and one pinned host mirror, see pin()), so that a
        # batch arrives with a single host->device copy
        spec = [(k, (B, P), edt) for k in _IN_EMBED] + [(k, (B_u, K), logit_dtype) for k in _IN_TEACHER]
        if student_ce:
            spec += [(k, (B, K), logit_dtype) for k in _IN_STUDENT]
        spec += [("y_l", (B_l,), torch.int64), ("mask_random", (B_u,), torch.bool)]
        self._layout, off = [], 0
        for name, shape, dtype in spec:
            nbytes = int(torch.tensor([], dtype=dtype).element_size()) * int(torch.Size(shape).numel())
            self._layout.append((name, shape, dtype, off, nbytes))
            off += (nbytes + 255) // 256 * 256
        self._packed_in = torch.zeros(max(off, 256), dtype=torch.uint8, device=dev)
        self.inp: Dict[str, torch.Tensor] = self._views(self._packed_in)
        # state buffers carry the reference names (STiLModel.py:94-96)
        self.prototypes = z(K, P)
        self.prototypes_sum = z(K, P)
        self.prototypes_count_sum = z(K, 1)
        # ... and so does every per-batch output (one device allocation, one pinned host mirror): a caller that wants the
        # gradients, pseudo labels and masks on the host gets them with a single device->host copy (step_host_full)
        ospec = [("losses", (5,), torch.float32),                      # itc, pt, m_u, i_u, t_u
                 ("d_feat_i", (B, P), grad_dtype), ("d_feat_t", (B, P), grad_dtype), ("d_feat_m", (B, P), grad_dtype),
                 ("pseudo_label", (B_u, K), torch.float32), ("max_prob", (B_u,), torch.float32),
                 ("max_idx", (B_u,), torch.int64)]
        ospec += [(k, (B_u,), torch.bool) for k in ("mask1", "case1", "case2_i", "case2_t", "case3")]
        if student_ce:
            ospec += [(k, (B, K), torch.float32) for k in ("d_y_m", "d_y_i", "d_y_t")]
        self._out_layout, off = [], 0
        for name, shape, dtype in ospec:
            nbytes = int(torch.tensor([], dtype=dtype).element_size()) * int(torch.Size(shape).numel())
            self._out_layout.append((name, shape, dtype, off, nbytes))
            off += (nbytes + 255) // 256 * 256
        self._packed_out = torch.zeros(max(off, 256), dtype=torch.uint8, device=dev)
        self.out: Dict[str, torch.Tensor] = self._views(self._packed_out, self._out_layout)
        self.out.update({"class_sum": z(K, P), "class_count": z(K, 1)})
        self.da = bool(da)
        if self.da:
            self.DA_queue, self.DA_ptr = z(da_len, K), z(1, dtype=torch.int64)
            self._da = {"probs": z(B_u, K), "mean": z(K), "qmean": z(K), "aligned": z(B_u, K)}
        lib = _lib.load()
        code = _lib.STIL_BF16 if edt == torch.bfloat16 else _lib.STIL_F32
        self._ws = torch.zeros(lib.stil_head_step_workspace_bytes(B, B_l, K, P, code), dtype=torch.uint8, device=dev)
        self._proto_version = None       # torch version counter of `prototypes` at the last operand conversion
        self._args = self._make_args(code, _lib.STIL_BF16 if logit_dtype == torch.bfloat16 else _lib.STIL_F32)
        self.launches_per_step = lib.stil_head_step_launches(C.byref(self._args))
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._pinned: Optional[Dict[str, torch.Tensor]] = None
        self._losses_host = torch.zeros(5, dtype=torch.float32).pin_memory()
        self._out_host: Optional[torch.Tensor] = None
        self.h2d_bytes = self._packed_in.numel()
        self.d2h_bytes = self._losses_host.numel() * 4
        self.d2h_bytes_full = self._packed_out.numel()

    # ------------------------------------------------------------------------------------------
    def _views(self, buf: torch.Tensor, layout=None) -> Dict[str, torch.Tensor]:
        """Typed views of a packed input / output buffer (device or pinned host) following self._layout."""
        out = {}
        for name, shape, dtype, off, nbytes in (self._layout if layout is None else layout):
            out[name] = buf[off:off + nbytes].view(dtype).view(shape) if nbytes else torch.empty(shape, dtype=dtype,
                                                                                                 device=buf.device)
        return out

    def _make_args(self, embed_code: int, logit_code: int) -> HeadStepArgs:
        c, i, o = self.cfg, self.inp, self.out
        p = lambda t: t.data_ptr()
        a = HeadStepArgs()
        a.batch, a.b_l, a.k, a.dim = c.batch, c.b_l, c.num_classes, c.proj_dim
        a.embed_dtype, a.logit_dtype, a.grad_dtype = embed_code, logit_code, _lib.dtype_code(o["d_feat_i"])
        a.temperature, a.lambda0, a.th1 = c.temperature, c.lambda_0, c.th1
        a.rate_pseudo, a.repeat_ratio = c.rate_pseudo, c.repeat_ratio
        a.past_start_epoch = int(c.past_start_epoch)
        for k in _IN_EMBED + _IN_TEACHER:
            setattr(a, k, p(i[k]))
        if self.student_ce:
            for k in _IN_STUDENT:
                setattr(a, k, p(i[k]))
            a.d_y_m, a.d_y_i, a.d_y_t = p(o["d_y_m"]), p(o["d_y_i"]), p(o["d_y_t"])
        a.y_l, a.prototypes, a.mask_random = p(i["y_l"]), p(self.prototypes), p(i["mask_random"])
        a.losses = p(o["losses"])
        a.d_feat_i, a.d_feat_t, a.d_feat_m = p(o["d_feat_i"]), p(o["d_feat_t"]), p(o["d_feat_m"])
        a.pseudo_label, a.max_prob, a.max_idx = p(o["pseudo_label"]), p(o["max_prob"]), p(o["max_idx"])
        for k in ("mask1", "case1", "case2_i", "case2_t", "case3", "class_sum", "class_count"):
            setattr(a, k, p(o[k]))
        a.prototypes_sum, a.prototypes_count_sum = p(self.prototypes_sum), p(self.prototypes_count_sum)
        a.rate_uce_scale = self.rate_uce
        a.workspace, a.workspace_bytes = p(self._ws), self._ws.numel()
        if self.da:
            a.prediction_in = p(self._da["aligned"])
        return a

    def set_hparams(self, **kw) -> None:
        """Change step hyper-parameters after construction — ``past_start_epoch`` (the reference flips the gate of
        ``STiLModel.py:317-320`` when ``current_epoch > start_epoch``), ``th1``, ``rate_pseudo``, ``temperature``,
        ``lambda_0``, ``repeat_ratio``, ``rate_uce``.  They are kernel parameters baked into the captured CUDA graph,
        so the graph is dropped and re-captured on the next ``run()`` (in the data-parallel head every rank must make
        the same call before its next step: capture is collective)."""
        names = {"past_start_epoch", "th1", "rate_pseudo", "temperature", "lambda_0", "repeat_ratio"}
        for k, v in kw.items():
            if k == "rate_uce":
                self.rate_uce = float(v)
            elif k in names:
                setattr(self.cfg, k, v)
            else:
                raise ValueError(f"unknown head hyper-parameter {k!r}")
        a, c = self._args, self.cfg
        a.temperature, a.lambda0, a.th1 = c.temperature, c.lambda_0, c.th1
        a.rate_pseudo, a.repeat_ratio = c.rate_pseudo, c.repeat_ratio
        a.past_start_epoch = int(c.past_start_epoch)
        a.rate_uce_scale = self.rate_uce
        self._drop_graphs()

    def _drop_graphs(self) -> None:
        self._graph = None

    def load(self, batch: Dict[str, torch.Tensor]) -> None:
        """Copy one batch (CPU or CUDA tensors, keys as in synth.make_batch) into the static input buffers."""
        for k, dst in self.inp.items():
            dst.copy_(batch[k].to(dst.dtype) if batch[k].dtype != dst.dtype else batch[k], non_blocking=True)
        if "prototypes" in batch:
            self.prototypes.copy_(batch["prototypes"], non_blocking=True)

    def _ensure_prototypes(self) -> None:
        """The prototypes change once per epoch (STiLModel.py:408-415), their tensor-core operand form is cached in
        the workspace and refreshed only when the tensor's version counter moved (any in-place torch write) or
        finalize_prototypes() ran."""
        v = self.prototypes._version
        if v != self._proto_version:
            self._args.stream = torch.cuda.current_stream(self.dev).cuda_stream
            check(_lib.load().stil_head_prepare_prototypes(C.byref(self._args)))
            self._args.prototypes_prepared = 1
            self._proto_version = v

    def _all_reduce_mean(self, mean: torch.Tensor) -> None:
        """``torch.distributed.all_reduce(probs_bt_mean)`` / world size (``STiLModel.py:174-176``); one rank: no-op."""

    def _enqueue_da(self, stream: int) -> None:
        """``prediction = distribution_alignment(torch.softmax(y_hat_m_ue, dim=1))`` (``STiLModel.py:276-277``) into the
        buffer ``stil_head_step`` reads as ``prediction_in``."""
        lib, d, p = _lib.load(), self._da, (lambda t: t.data_ptr())
        B_u, K = self.cfg.b_u, self.cfg.num_classes
        y = self.inp["y_m_ue"]
        check(lib.stil_softmax_rows(p(y), _lib.dtype_code(y), K, B_u, K, p(d["probs"]), K, stream))
        check(lib.stil_da_batch_mean(p(d["probs"]), K, B_u, K, p(d["mean"]), stream))
        self._all_reduce_mean(d["mean"])
        check(lib.stil_da_apply(p(d["probs"]), K, B_u, K, p(d["mean"]), p(self.DA_queue), self.DA_queue.shape[0],
                                p(self.DA_ptr), p(d["qmean"]), p(d["aligned"]), K, stream))

    def _enqueue(self) -> None:
        self._args.stream = torch.cuda.current_stream(self.dev).cuda_stream
        if self.da:
            self._enqueue_da(self._args.stream)
        check(_lib.load().stil_head_step(C.byref(self._args)))

    def capture(self) -> None:
        """Warm up once, then record the step into a CUDA graph (kernel params are baked in)."""
        with torch.cuda.device(self.dev):
            self._ensure_prototypes()
            # the warm-up run must not leave a trace in the running accumulators (STiLModel.py:380-381) or the DA ring
            keep = [t.clone() for t in self._state_tensors()]
            s = torch.cuda.Stream(self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                self._enqueue()
            torch.cuda.current_stream(self.dev).wait_stream(s)
            for t, k in zip(self._state_tensors(), keep):
                t.copy_(k)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self._graph = g

    def _state_tensors(self):
        st = [self.prototypes_sum, self.prototypes_count_sum]
        if self.da:
            st += [self.DA_queue, self.DA_ptr]
        return st

    def timed_run(self, names=False):
        """One un-captured step with CUDA events recorded (on the launching streams) around the launches of the two
        critical chains; returns the per-launch durations in milliseconds (bench.py's live kernel timing)."""
        pairs = [("gemm_tc05_kernel<STATS>[proto, teacher]", 1, 5), ("cgpl_pgls_kernel", 5, 6),
                 ("gemm_tc05_kernel<GRAD>[proto]", 6, 7), ("gemm_tc05_kernel<STORE>[proto]", 7, 8),
                 ("prep_kernel[infonce]", 9, 10), ("gemm_tc05_kernel<STATS>[infonce x2]", 10, 2),
                 ("gemm_tc05_kernel<GRAD>[infonce x2]", 2, 3), ("gemm_tc05_kernel<STORE>[infonce x2]", 3, 4)]
        if self.cfg.embed_dtype != "bf16":
            pairs.insert(0, ("prep_kernel[fp32 split]", 0, 1))
        n = 11
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
            for e in evs:
                e.record(st)            # materialise the cudaEvent_t handles
            arr = (C.c_void_p * n)(*[e.cuda_event for e in evs])
            self._args.timing_events = C.cast(arr, C.c_void_p)
            self._args.n_timing_events = n
            self._ensure_prototypes()
            try:
                # keep the GPU busy while the host enqueues the whole step, so the event intervals are GPU-side
                # durations rather than host launch gaps
                torch.cuda._sleep(400_000)
                self._enqueue()
            finally:
                self._args.timing_events = None
                self._args.n_timing_events = 0
            torch.cuda.synchronize(self.dev)
            ms = [evs[i].elapsed_time(evs[j]) for _, i, j in pairs]
        return list(zip([p[0] for p in pairs], ms)) if names else ms

    def run(self) -> None:
        """Enqueue one head step on the current stream (graph replay when captured)."""
        with torch.cuda.device(self.dev):
            self._ensure_prototypes()
            if self.use_graph:
                if self._graph is None:
                    self.capture()
                self._graph.replay()
            else:
                self._enqueue()

    # ------------------------------------------------------------------------------------------ end to end
    def pin(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Stage a host batch in ONE pinned buffer laid out like the device inputs (a data loader's job, done
        once per batch outside the step)."""
        buf = torch.zeros(self._packed_in.numel(), dtype=torch.uint8).pin_memory()
        for k, v in self._views(buf).items():
            v.copy_(batch[k].to(v.dtype))
        return buf

    def copy_in(self, pinned: torch.Tensor) -> None:
        """Host -> device copy of one packed batch on the current stream."""
        self._packed_in.copy_(pinned, non_blocking=True)

    def step_host(self, pinned: torch.Tensor) -> torch.Tensor:
        """One end-to-end step from HOST buffers: H2D of every input (one copy), the head, D2H of the five losses.
        Returns the pinned host tensor of losses (valid after the stream synchronises)."""
        with torch.cuda.device(self.dev):
            self.copy_in(pinned)
            self.run()
            self._losses_host.copy_(self.out["losses"], non_blocking=True)
        return self._losses_host

    def copy_out_full(self) -> Dict[str, torch.Tensor]:
        """Device -> host copy (current stream, ONE transfer) of every per-batch output — losses, the three embedding
        gradients, the logit gradients, pseudo labels, max_prob / max_idx and the five masks — into a pinned host mirror;
        returns typed host views (valid once the stream has synchronised)."""
        if self._out_host is None:
            self._out_host = torch.zeros(self._packed_out.numel(), dtype=torch.uint8).pin_memory()
            self._out_host_views = self._views(self._out_host, self._out_layout)
        self._out_host.copy_(self._packed_out, non_blocking=True)
        return self._out_host_views

    def step_host_full(self, pinned: torch.Tensor) -> Dict[str, torch.Tensor]:
        """step_host, but EVERYTHING the step produces per batch goes back to the host (one D2H copy)."""
        with torch.cuda.device(self.dev):
            self.copy_in(pinned)
            self.run()
            return self.copy_out_full()

    def finalize_prototypes(self) -> torch.Tensor:
        """Epoch end (STiLModel.py:408-415)."""
        empty = torch.zeros(1, dtype=torch.int32, device=self.dev)
        k, d = self.prototypes.shape
        with torch.cuda.device(self.dev):
            check(_lib.load().stil_proto_finalize(self.prototypes.data_ptr(), self.prototypes_sum.data_ptr(),
                                                  self.prototypes_count_sum.data_ptr(), k, d, empty.data_ptr(),
                                                  torch.cuda.current_stream(self.dev).cuda_stream))
        self._proto_version = None      # written by the kernel, not through torch: force the operand refresh
        return empty


class DistributedSTiLHead(STiLHead):
    """Data-parallel head over one node (one process per GPU, ``torch.distributed``; NVLink).

    Everything row-local runs in ``stil_head_step`` with ``skip_infonce``; the two coupled pieces follow
    SURVEY §8e: (i) InfoNCE on the GLOBAL batch — ``[feat_i | feat_t]`` rows all-gathered (one exchange, leading
    dimension 2P), local rows scored against all columns, only the loss partial and the two LSE vectors exchanged
    before the backward (oracle: reference ``CLIPLoss`` on the concatenated batch) — no gradient reduce-scatter;
    (ii) the prototype partial sums of all ranks are summed before they are accumulated (``STiLModel.py:377-379``).

    ``transport="fused"`` (default for bf16 embeddings): compute and exchange are fused over CUDA-IPC peer memory — one
    kernel packs ``[feat_i | feat_t]``, computes the inverse norms and stores both into every rank's buffer; the
    statistics GEMM consumes the gathered rows tile by tile as their owner's arrival flag lands (local tiles do not wait);
    one kernel merges the statistics into row LSEs and stores them into every rank's buffer; the gradient GEMM's epilogue
    waits only for the owners of the column LSEs it reads; the class partials are pushed and waited for at the very end
    of the row-local chain.  No kernel sits waiting for a peer before useful work.
    ``transport="p2p"``: the three exchanges are blocking ``stil_p2p_exchange`` all-gather kernels (remote NVLink stores
    + flags).  Both peer-memory transports alternate destination regions between two halves on successive steps, so two
    CUDA graphs (even / odd) are captured and replayed in turn.
    ``transport="nccl"``: one NCCL all-gather and two all-reduces, captured in one graph.
    ``losses[0]`` is the global InfoNCE loss (same on every rank); ``d_feat_i/t`` its gradients w.r.t. the local rows.
    """

    def __init__(self, cfg: HeadConfig, device="cuda", group=None, transport: str = "fused", **kw) -> None:
        super().__init__(cfg, device=device, **kw)
        import torch.distributed as dist
        from .distributed import GlobalBatch
        self.dist, self.group, self.gb = dist, group, GlobalBatch(group)
        self.world, self.rank = self.gb.world_size, self.gb.rank
        if self.world > 1:
            # every rank must launch the same shapes: the arrival counters / LL tags of the peer-memory transports count
            # pushes per step, so a rank with a different batch (a ragged last batch) would make its peers wait for ever
            sig = torch.tensor([cfg.batch, cfg.b_l, cfg.num_classes, cfg.proj_dim, int(self.inp["feat_i"].dtype == torch.bfloat16),
                                int(self.student_ce), int(self.da)], dtype=torch.int64, device=self.dev)
            allsig = [torch.zeros_like(sig) for _ in range(self.world)]
            dist.all_gather(allsig, sig, group=group)
            if any(not torch.equal(s_, sig) for s_ in allsig):
                raise ValueError("DistributedSTiLHead: every rank must use the same batch / class / dimension / dtype "
                                 f"configuration (got {[s_.tolist() for s_ in allsig]}); pad or drop a ragged last batch")
        if transport == "fused" and self.inp["feat_i"].dtype != torch.bfloat16:
            transport = "p2p"           # the fused schedule reads bf16 rows in place; fp32 needs the operand split pass
        self.transport = transport if self.world > 1 else "nccl"
        B, K, P = cfg.batch, cfg.num_classes, cfg.proj_dim
        n, W = B * self.world, self.world
        dev, edt = self.dev, self.inp["feat_i"].dtype
        esz = 2 if edt == torch.bfloat16 else 4
        self._slot = (K * P + K + 3) // 4 * 4                       # floats per rank: [class_sum | class_count]
        self._ab_loc = torch.empty(B, 2 * P, dtype=edt, device=dev)  # [feat_i | feat_t], leading dimension 2P
        self._cls_loc = torch.zeros(self._slot, dtype=torch.float32, device=dev)
        self._nce_loc = torch.zeros(4 + 2 * B, dtype=torch.float32, device=dev)   # [loss partial (4) | lse_row | lse_col]
        self._nce_stream = None
        self._graphs = [None, None]
        self._parity = 0
        lib = _lib.load()
        code = _lib.dtype_code(self.inp["feat_i"])
        self._nce_ws = torch.zeros(lib.stil_infonce_workspace_bytes(B, n, P, code), dtype=torch.uint8, device=dev)
        self._loss_stream = None
        a = self._args
        a.skip_infonce = 1
        a.class_sum = self._cls_loc.data_ptr()
        a.class_count = self._cls_loc[K * P:].data_ptr()
        a.prototypes_sum, a.prototypes_count_sum = None, None
        self._sum_out = torch.zeros(K, P, dtype=torch.float32, device=dev)
        self._cnt_out = torch.zeros(K, 1, dtype=torch.float32, device=dev)
        self.out["class_sum"], self.out["class_count"] = self._sum_out, self._cnt_out
        if self.transport in ("p2p", "fused"):
            from .distributed import P2PBuffer
            r256 = lambda x: (x + 255) // 256 * 256
            self._o_ab = 0
            self._o_loss = self._o_ab + r256(n * 2 * P * esz)
            self._o_lr = self._o_loss + r256(W * 16)
            self._o_lc = self._o_lr + r256(n * 8)       # 8 bytes per entry: the fused transport stores LL words
            self._o_cls = self._o_lc + r256(n * 8)
            self._o_ra = self._o_cls + r256(W * self._slot * 4)
            self._o_rb = self._o_ra + r256(n * 4)
            self._half = self._o_rb + r256(n * 4)
            self._p2p = P2PBuffer(2 * self._half, dev, group)
            ch = _lib.P2PChannel()
            for i, b in enumerate(self._p2p.peers):
                ch.bases[i] = b
            ch.world, ch.rank, ch.flags_offset, ch.ctrl_offset, ch.channel = W, self.rank, 0, 512, 2
            self._push_channel = ch
            v = self._p2p.view
            self._r = []
            for h in (0, 1):
                o = h * self._half
                self._r.append(dict(
                    off=o, ab=v(o + self._o_ab, (n, 2 * P), edt), loss=v(o + self._o_loss, (W, 4), torch.float32),
                    lse_row=v(o + self._o_lr, (n,), torch.float32), lse_col=v(o + self._o_lc, (n,), torch.float32),
                    cls=v(o + self._o_cls, (W, self._slot), torch.float32),
                    ra=v(o + self._o_ra, (n,), torch.float32), rb=v(o + self._o_rb, (n,), torch.float32)))
        else:
            # NCCL: [loss partial (4) | LSE slots [2, W, B]] all-reduced for the InfoNCE chain (every rank fills only
            # its own LSE slot, so SUM is the gather) and [class_sum | class_count] all-reduced for the bank
            self._ab_all = torch.empty(n, 2 * P, dtype=edt, device=dev)
            self._packed = torch.zeros(4 + 2 * n, dtype=torch.float32, device=dev)
            self._lse = self._packed[4:].view(2, W, B)
        # + cat, exchanges / collectives, infonce fwd (prep, gemm, finish), bwd (prep, gemm, gemm), add, loss
        if self.transport == "fused":
            # row-local step (+ the class-partial push inside it), push_embeddings, stats GEMM, push_lse, loss finish,
            # grad GEMM, dX GEMM (+ slice reduction for a long split contraction), waiting proto_add
            self.launches_per_step = lib.stil_head_step_launches(C.byref(a)) + 1 + 6 + (1 if (n >= 2048 or P > 128) else 0) + 1
        else:
            self.launches_per_step = lib.stil_head_step_launches(C.byref(a)) + 1 + 3 + 3 + 3 + 2

    # ------------------------------------------------------------------------------------------
    def _drop_graphs(self) -> None:
        self._graph = None
        self._graphs = [None, None]

    def _all_reduce_mean(self, mean: torch.Tensor) -> None:
        if self.world > 1:
            self.dist.all_reduce(mean, group=self.group)
            mean.div_(self.world)

    def capture(self) -> None:
        """Record kernels AND exchanges of one step into CUDA graphs (every rank must call this and later replay
        in lockstep).  The p2p transport alternates destination halves, hence one graph per parity."""
        with torch.cuda.device(self.dev):
            self._ensure_prototypes()
            keep = [t.clone() for t in self._state_tensors()]
            s = torch.cuda.Stream(self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for i in range(2):
                    self._run_eager(i & 1)
            torch.cuda.current_stream(self.dev).wait_stream(s)
            torch.cuda.synchronize(self.dev)
            self.dist.barrier(group=self.group)
            for parity in ((0, 1) if self.transport in ("p2p", "fused") else (0,)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_eager(parity)
                self._graphs[parity] = g
            for t, k in zip(self._state_tensors(), keep):
                t.copy_(k)
            torch.cuda.synchronize(self.dev)
            self.dist.barrier(group=self.group)
            self._graph = self._graphs[0]

    def run(self) -> None:
        with torch.cuda.device(self.dev):
            self._ensure_prototypes()
            parity = self._parity if self.transport in ("p2p", "fused") else 0
            self._parity ^= 1
            if self.use_graph:
                if self._graph is None:
                    self.capture()
                self._graphs[parity].replay()
            else:
                self._run_eager(parity)

    def _run_fused(self, parity: int) -> None:
        """One step of the fused compute + exchange schedule (see the class docstring)."""
        cfg, lib = self.cfg, _lib.load()
        B, K, P, W = cfg.batch, cfg.num_classes, cfg.proj_dim, self.world
        n, off, r = B * W, B * self.rank, self.rank
        code = _lib.dtype_code(self.inp["feat_i"])
        p = lambda t: t.data_ptr()
        pb, R = self._p2p, self._r[parity]
        H = pb.HEADER + R["off"]                      # byte offset of this parity's region inside every buffer
        flags = lambda ch: pb.local + ch * 64          # u64 flags[channel][8] at offset 0
        seq = lambda ch: pb.local + 512 + ch * 8       # u64 seq[channel] at offset 512
        chan = lambda ch: (pb._bases, W, r, 0, 512, ch)
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            if self._nce_stream is None:
                self._nce_stream = torch.cuda.Stream(self.dev)
                self._loss_stream = torch.cuda.Stream(self.dev)
            sb, sl = self._nce_stream, self._loss_stream
            sb.wait_stream(cur)
            a_all, b_all = p(R["ab"]), p(R["ab"]) + P * 2
            with torch.cuda.stream(sb):
                check(lib.stil_p2p_push_embeddings(*chan(0), p(self.inp["feat_i"]), p(self.inp["feat_t"]), code, B, P, off,
                                                   H + self._o_ab, H + self._o_ra, H + self._o_rb, sb.cuda_stream))
                check(lib.stil_infonce_stats_gathered(a_all, b_all, p(R["ra"]), p(R["rb"]), code, B, n, P, 2 * P, off,
                                                      cfg.temperature, flags(0), seq(0), B, p(self._nce_ws),
                                                      self._nce_ws.numel(), sb.cuda_stream))
                sl.wait_stream(sb)
                # row LSEs as LL words {value, step tag}: no flag, no fence; the tag is channel 0's arrival target
                check(lib.stil_p2p_push_lse(*chan(0), p(self._nce_ws), B, n, P, code, off, H + self._o_lr, H + self._o_lc,
                                            sb.cuda_stream))
                check(lib.stil_infonce_bwd_gathered(a_all, b_all, p(R["ra"]), p(R["rb"]), code, B, n, P, 2 * P, off,
                                                    cfg.temperature, cfg.lambda_0, p(R["lse_row"]), p(R["lse_col"]),
                                                    seq(0), None, p(self.out["d_feat_i"]),
                                                    p(self.out["d_feat_t"]), _lib.dtype_code(self.out["d_feat_i"]), P,
                                                    p(self._nce_ws), self._nce_ws.numel(), sb.cuda_stream))
            with torch.cuda.stream(sl):
                # loss partial of the local rows from the statistics, beside the chain: it leaves for every rank as one
                # LL word (value + step tag in a single 8-byte store)
                check(lib.stil_infonce_loss_gathered(a_all, b_all, p(R["ra"]), p(R["rb"]), code, B, n, P, 2 * P, off,
                                                     cfg.temperature, cfg.lambda_0, p(self._nce_loc), p(self._nce_loc[4:]),
                                                     p(self._nce_loc[4 + B:]), pb._bases, W, r, H + self._o_loss, seq(0),
                                                     p(self._nce_ws), self._nce_ws.numel(), sl.cuda_stream))
            # everything row-local (current stream); the class partials are pushed from inside the step, mid-way, as soon
            # as proto_accumulate has produced them
            a = self._args
            a.partials_push = C.pointer(self._push_channel)
            a.partials_dst_offset = H + self._o_cls + r * self._slot * 4
            self._enqueue()
            cur.wait_stream(sl)
            # last kernel of the step: waits for every rank's partials (arrival counters) and loss word, adds in rank order
            check(lib.stil_proto_add_gathered_wait(p(R["cls"]), W, self._slot, K, P, p(self._sum_out), p(self._cnt_out),
                                                   p(self.prototypes_sum), p(self.prototypes_count_sum), flags(2), seq(2),
                                                   pb.local + H + self._o_loss, seq(0), p(self.out["losses"]),
                                                   cur.cuda_stream))
            cur.wait_stream(sb)

    def _run_eager(self, parity: int = 0) -> None:
        if self.transport == "fused":
            return self._run_fused(parity)
        cfg, dist, lib = self.cfg, self.dist, _lib.load()
        B, K, P, W = cfg.batch, cfg.num_classes, cfg.proj_dim, self.world
        n, off, r = B * W, B * self.rank, self.rank
        code = _lib.dtype_code(self.inp["feat_i"])
        p = lambda t: t.data_ptr()
        esz = self._ab_loc.element_size()
        p2p = self.transport == "p2p"
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            if self._nce_stream is None:
                self._nce_stream = torch.cuda.Stream(self.dev)
            sb = self._nce_stream
            torch.cat((self.inp["feat_i"], self.inp["feat_t"]), dim=1, out=self._ab_loc)
            sb.wait_stream(cur)
            # chain B (own stream): gather -> global InfoNCE forward -> exchange(loss, LSE) -> backward
            with torch.cuda.stream(sb):
                if p2p:
                    R = self._r[parity]
                    self._p2p.exchange(0, [(self._ab_loc, R["off"] + self._o_ab + off * 2 * P * esz)], sb.cuda_stream)
                    ab_all, loss_dst = R["ab"], self._nce_loc
                    lse_row_loc, lse_col_loc = self._nce_loc[4:4 + B], self._nce_loc[4 + B:]
                else:
                    dist.all_gather_into_tensor(self._ab_all, self._ab_loc, group=self.group)
                    self._lse.zero_()
                    ab_all, loss_dst = self._ab_all, self._packed
                    lse_row_loc, lse_col_loc = self._lse[0, r], self._lse[1, r]
                a_all, b_all = p(ab_all), p(ab_all) + P * esz
                a_loc, b_loc = a_all + off * 2 * P * esz, b_all + off * 2 * P * esz
                check(lib.stil_infonce_fwd(a_loc, b_loc, a_all, b_all, code, B, n, P, 2 * P, off, cfg.temperature,
                                           cfg.lambda_0, p(loss_dst), p(lse_row_loc), p(lse_col_loc), None, 0,
                                           p(self._nce_ws), self._nce_ws.numel(), sb.cuda_stream))
                if p2p:
                    self._p2p.exchange(1, [(self._nce_loc[:4], R["off"] + self._o_loss + r * 16),
                                           (lse_row_loc, R["off"] + self._o_lr + off * 4),
                                           (lse_col_loc, R["off"] + self._o_lc + off * 4)], sb.cuda_stream)
                    lse_row_all, lse_col_all = R["lse_row"], R["lse_col"]
                    torch.sum(R["loss"][:, 0], dim=0, keepdim=True, out=self.out["losses"][0:1])
                else:
                    dist.all_reduce(self._packed, op=dist.ReduceOp.SUM, group=self.group)
                    lse_row_all, lse_col_all = self._lse[0].reshape(-1), self._lse[1].reshape(-1)
                    self.out["losses"][0:1].copy_(self._packed[0:1], non_blocking=True)
                # the forward's inverse norms are still in the workspace: no second preparation pass
                check(lib.stil_infonce_bwd_after_fwd(a_all, b_all, code, B, n, P, 2 * P, off, cfg.temperature,
                                                     cfg.lambda_0, p(lse_row_all), p(lse_col_all), None,
                                                     p(self.out["d_feat_i"]), p(self.out["d_feat_t"]),
                                                     _lib.dtype_code(self.out["d_feat_i"]), P, p(self._nce_ws),
                                                     self._nce_ws.numel(), sb.cuda_stream))
            # chain A (current stream): everything row-local, then the prototype partials of all ranks
            self._enqueue()
            if p2p:
                self._p2p.exchange(2, [(self._cls_loc, R["off"] + self._o_cls + r * self._slot * 4)], cur.cuda_stream)
                parts, nparts = R["cls"], W
            else:
                dist.all_reduce(self._cls_loc, op=dist.ReduceOp.SUM, group=self.group)
                parts, nparts = self._cls_loc, 1
            check(lib.stil_proto_add_gathered(p(parts), nparts, self._slot, K, P, p(self._sum_out), p(self._cnt_out),
                                              p(self.prototypes_sum), p(self.prototypes_count_sum), cur.cuda_stream))
            cur.wait_stream(sb)

    def release(self) -> None:
        """Drop the captured graphs (they pin the communicator / peer mappings) and unmap the peer memory; call on
        every rank before the process group is destroyed."""
        self._graph = None
        self._graphs = [None, None]
        if getattr(self, "_p2p", None) is not None:
            self._r = None
            self._p2p.close()
            self._p2p = None
