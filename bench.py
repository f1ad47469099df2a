#!/usr/bin/env python
"""bench.py — STiL head-step throughput on B200 (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2|C3|C5]

One "step" = one pass of the whole per-batch head (CGPL, PGLS, InfoNCE fwd+bwd, prototype loss fwd+bwd,
masked soft-target CE fwd+bwd, prototype partial sums) over one synthetic DVM-shaped batch (C2: B=512 = 64
labelled + 448 unlabelled, K=286, P=128, bf16 embeddings).  `value` = samples/s with inputs resident in HBM
(CUDA-graph replay), `e2e` = the same through STiLHead.step_host with pinned HOST buffers (H2D of every
input and D2H of the losses inside the timed region), `e2e_full` = with EVERY per-batch output (gradients,
pseudo labels, masks) copied back as well.  Timed steps rotate over enough distinct batches that the touched
working set exceeds the 126 MB L2.  Before anything is timed the step is checked against the oracle
(`parity_check`; on N > 1 ranks against the single-process oracle on the concatenated batch) and a mismatch
exits non-zero.  Rank 0 prints ONE JSON line.

  --config C5: the SimMatch memory-bank sweep (bank 65536 x 512 bf16, 286 classes, 448 unlabelled rows per GPU),
  column-sharded over the N GPUs: see run_bank_arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

METRIC = "stil_head_step_samples_per_sec"
UNIT = "samples/s"
L2_BYTES = 126 * 2 ** 20


PREWARM_STEPS = 6000   # untimed steps before the W warm-up steps of every timed region (~0.2-0.4 s of GPU work)


def peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_tflops_burst": d["bf16_tflops"], "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "src": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        f = REPO / "profiles" / name
        if f.exists():
            d = json.loads(f.read_text()).get(kernel)
            if d:
                return d["dram_bytes_per_launch"]
    return None


def get_cfg(name: str):
    from stil_tta_b200 import synth
    return synth.CONFIGS[name]()


def workload_name(cfg, name):
    return (f"{name} {cfg.name} head step B={cfg.batch} ({cfg.b_l}l+{cfg.b_u}u) K={cfg.num_classes} "
            f"P={cfg.proj_dim} {cfg.embed_dtype} embeddings")


def make_config(workload: str, per_gpu_batch: int, world: int, l2: str, cuda_graph, collectives: str, infonce: str) -> dict:
    """The `config` object of the JSON line — the SAME key set on both arms (the driver compares them)."""
    return {"workload": workload, "per_gpu_batch": per_gpu_batch, "parallelism": f"dp{world}", "l2": l2,
            "cuda_graph": cuda_graph, "prewarm_steps": PREWARM_STEPS, "collectives": collectives, "infonce": infonce}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_steps(cfg, steps: int, warmup: int, budget_s: float = 30.0):
    """The reference head on the host cores: the oracle port (torch CPU, fp32, fwd+bwd), every op of
    STiLModel.training_step lines 262-303, 317-322, 339, 374-381.  Returns (ms_per_step, steps_done, threads)."""
    from oracle import stil_head_oracle as O
    from stil_tta_b200 import synth
    ncpu = len(os.sched_getaffinity(0))
    batches = [synth.make_batch(cfg, seed=100 + i) for i in range(4)]
    state = {"prototypes_sum": torch.zeros(cfg.num_classes, cfg.proj_dim),
             "prototypes_count_sum": torch.zeros(cfg.num_classes, 1)}
    # give the reference its best thread count (oversubscribed intra-op threads can be slower than one): 20 timed steps
    # after 3 warm-up steps per candidate, median step time
    best = (float("inf"), 1)
    for n in sorted({1, 2, 4, 8, 16, 32, ncpu}):
        if n > ncpu:
            continue
        torch.set_num_threads(n)
        for i in range(3):
            O.head_step(batches[i % 4], cfg, state=state)
        ts = []
        for i in range(20):
            t0 = time.perf_counter()
            O.head_step(batches[i % 4], cfg, state=state)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        best = min(best, (ts[len(ts) // 2], n))
    n = best[1]
    torch.set_num_threads(n)
    for i in range(warmup):
        O.head_step(batches[i % 4], cfg, state=state)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        O.head_step(batches[i % 4], cfg, state=state)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dt / done * 1e3, done, n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == "C5":
        return run_bank_reference_arm(args)
    cfg = get_cfg(args.config)
    ms, done, n = cpu_reference_steps(cfg, args.steps, args.warmup, budget_s=120.0)
    value = cfg.batch / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(workload_name(cfg, args.config), cfg.batch, args.gpus, "n/a (host)", False,
                              "n/a (one host process)", f"local batch {cfg.batch} (reference CLIPLoss)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n, "kind": "port",
                         "sample": f"{done} head steps (fwd+bwd) of the oracle port of STiLModel.training_step's head, "
                                   f"torch {torch.__version__} CPU fp32, best of 1..{len(os.sched_getaffinity(0))} threads = {n}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def timed_region(fn_step, steps, warmup, dist_on, dev, prewarm=PREWARM_STEPS):
    """warm-up, barrier+sync, K steps between CUDA events on the current stream, barrier+sync; ms total."""
    import torch.distributed as dist
    # untimed pre-warm before the W warm-up steps: a fixed NUMBER of steps (the same on every rank — the data-parallel
    # step is collective) worth ~0.3 s, so that a short timed region (small K) is not measured on clocks still ramping up
    # from the idle state the clock sampler's start-up sleep leaves the GPU in
    pre = prewarm
    for i in range(pre):
        fn_step(i)
    for i in range(warmup):
        fn_step(pre + i)
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn_step(pre + warmup + i)
    e1.record()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def time_graph(launch, dev, reps=20, replays=5):
    """Seconds per launch: a CUDA graph of `reps` back-to-back launches (rotating buffers), timed over `replays` replays."""
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        for i in range(3):
            launch(i)
    torch.cuda.synchronize(dev)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(reps):
            launch(i)
    gr.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        gr.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / (replays * reps) * 1e-3


def row_kernel_rooflines(rows, K, dev, pk, nbuf, reps=20):
    """Average launch duration of the two row kernels (cgpl_pgls, masked_softce) on `rows` x `K` logits: CUDA events around a
    CUDA graph of `reps` back-to-back launches on `nbuf` rotating buffer sets (working set > L2, or said otherwise)."""
    from stil_tta_b200 import _lib
    from stil_tta_b200._lib import ptr
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(1)
    out = []
    ys = [[(torch.randn(rows, K, generator=g) * 3).to(dev) for _ in range(3)] for _ in range(nbuf)]
    tl = [torch.randn(rows, K, generator=g).to(dev) for _ in range(nbuf)]
    pl = [torch.empty(rows, K, device=dev) for _ in range(nbuf)]
    mp = torch.empty(rows, device=dev); mi = torch.empty(rows, dtype=torch.int64, device=dev)
    fl = [torch.empty(rows, dtype=torch.bool, device=dev) for _ in range(5)]
    cls = torch.empty(rows, dtype=torch.int32, device=dev); conf = torch.empty(rows, dtype=torch.bool, device=dev)

    def launch_cgpl(i):
        j = i % nbuf
        _lib.check(lib.stil_cgpl_pgls(ptr(ys[j][0]), ptr(ys[j][1]), ptr(ys[j][2]), 0, K, ptr(tl[j]), K, rows, K,
                                      0.1, 0.9, 0.9, 1, None, 0, ptr(pl[j]), K, None, 0, ptr(mp),
                                      ptr(mi), ptr(fl[0]), ptr(fl[1]), ptr(fl[2]), ptr(fl[3]), ptr(fl[4]), None,
                                      ptr(cls), ptr(conf), _lib.stream_ptr(dev)))

    t = time_graph(launch_cgpl, dev, reps)
    bytes_alg = 4 * rows * K * 4 + rows * K * 4 + rows * (4 + 8 + 5 + 4 + 1)
    wset = nbuf * 5 * rows * K * 4
    out.append({"kernel": "cgpl_pgls_kernel", "rows": rows, "k": K, "bound": "hbm", "us_per_launch": t * 1e6,
                "alg_bytes_per_launch": bytes_alg, "achieved": bytes_alg / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": bytes_alg / t / 1e9 / pk["hbm_gbs"], "working_set_MiB": wset / 2 ** 20})
    # masked soft-target CE fwd + gradient: reads 3 logit rows + pseudo label, writes 3 gradients
    for f in fl:
        f.fill_(True)
    mr = (torch.rand(rows, generator=g) >= 0.5).to(dev)
    losses = torch.empty(3, device=dev)
    gr_ = [[torch.empty(rows, K, device=dev) for _ in range(3)] for _ in range(min(nbuf, 4))]
    ws = torch.zeros(lib.stil_masked_softce_workspace_bytes(rows), dtype=torch.uint8, device=dev)
    for j in range(nbuf):
        pl[j].copy_(torch.softmax(ys[j][0], 1))

    def launch_ce(i):
        j, jg = i % nbuf, i % len(gr_)
        _lib.check(lib.stil_masked_softce(ptr(ys[j][0]), ptr(ys[j][1]), ptr(ys[j][2]), 0, K, ptr(pl[j]), K, ptr(fl[0]),
                                          ptr(fl[1]), ptr(fl[2]), ptr(fl[3]), ptr(fl[4]), ptr(mr), rows, K, ptr(losses),
                                          ptr(gr_[jg][0]), ptr(gr_[jg][1]), ptr(gr_[jg][2]), K, 1.0, ptr(ws), ws.numel(),
                                          _lib.stream_ptr(dev)))

    t = time_graph(launch_ce, dev, reps)
    bytes_alg = 7 * rows * K * 4 + rows * 6
    out.append({"kernel": "masked_softce_kernel", "rows": rows, "k": K, "bound": "hbm", "us_per_launch": t * 1e6,
                "alg_bytes_per_launch": bytes_alg, "achieved": bytes_alg / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": bytes_alg / t / 1e9 / pk["hbm_gbs"], "working_set_MiB": (nbuf * 4 + len(gr_) * 3) * rows * K * 4 / 2 ** 20})
    return out


def time_fn(fn, dev, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters * 1e-3


def clip_loss_sweep(dev, pk, shapes=((4096, 128), (4096, 512), (4096, 2048))):
    """CLIPLoss fwd+bwd on shapes where the tensor pipe is the bound ('also stress D in {512, 2048}').
    Whole-op entries: algorithmic flops 6 n^2 d / CUDA-event time of the forward+backward (all kernels).  Per-kernel
    entries: device durations of the library's launches from the profiler's CUPTI records."""
    import stil_tta_b200 as S
    from torch.profiler import ProfilerActivity, profile
    g = torch.Generator().manual_seed(1)
    res = []
    for n, d in shapes:
        a = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
        b = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
        crit = S.CLIPLoss(0.1, 0.5, return_logits=False)

        def fwd_bwd():
            a.grad = b.grad = None
            loss, _, _ = crit(a, b)
            loss.backward()
        t = time_fn(fwd_bwd, dev, iters=8, warm=3)
        fl = 6.0 * n * n * d
        res.append({"op": "CLIPLoss fwd+bwd", "n": n, "d": d, "dtype": "bf16", "us": t * 1e6, "alg_flops": fl,
                    "achieved": fl / t / 1e12, "unit": "TFLOP/s", "frac_sustained": fl / t / 1e12 / pk["bf16_tflops"],
                    "frac_burst": fl / t / 1e12 / pk["bf16_tflops_burst"], "basis": "algorithmic 6 n^2 d"})
        try:
            # per-kernel durations WITHOUT programmatic dependent launch: with PDL a dependent kernel's record includes its
            # wait for the predecessor, so the records would overlap and over-count
            from stil_tta_b200 import _lib
            _lib.load().stil_debug_pdl(0)
            try:
                fwd_bwd()
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for _ in range(3):
                        fwd_bwd()
                    torch.cuda.synchronize(dev)
            finally:
                _lib.load().stil_debug_pdl(1)
            per = {}
            for e in prof.events():
                if "stil::" in e.name:
                    per.setdefault(e.name.split("stil::")[-1].split("(")[0][:80], []).append(e.time_range.end - e.time_range.start)
            for name, v in sorted(per.items()):
                v.sort()
                res.append({"kernel": name, "n": n, "d": d, "us_per_launch": v[len(v) // 2], "launches_per_fwd_bwd": len(v) // 3})
        except Exception as ex:       # the profiler is evidence, never a reason to lose the bench line
            res.append({"profiler_error": repr(ex)[:160], "n": n, "d": d})
        del a, b
    return res


def gemm_ingraph_times(cfg, dev, batch, reps=6):
    """In-graph duration of every gemm_tc05_kernel launch of ONE captured head step: the CTAs' own %globaltimer stamps
    (stil_debug_trace; the buffer pointer travels in the kernel parameters, so it is installed before the capture).
    Per launch: mode, ctas, start = first CTA start, ready = first `griddepcontrol.wait` passed (the stream predecessor
    is complete), end = last CTA end, in microseconds relative to the step's first GEMM start; median over replays."""
    import stil_tta_b200 as S
    from stil_tta_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(64, 64, 8, dtype=torch.int64, device=dev)
    _lib.check(lib.stil_debug_trace(buf.data_ptr()))
    samples = []
    try:
        head = S.STiLHead(cfg, device=dev, use_graph=True)
        head.load(batch)
        head.capture()
        for _ in range(3):
            head.run()
        torch.cuda.synchronize(dev)
        for _ in range(reps):
            buf.zero_()
            head.run()
            torch.cuda.synchronize(dev)
            t = buf.cpu()
            used = t[:, :, 0] > 0
            ids = [i for i in range(64) if used[i].any()]
            if not ids:
                continue
            t0 = min(int(t[i][used[i]][:, 0].min()) for i in ids)
            one = []
            for i in ids:
                r = t[i][used[i]]
                ready = r[:, 1][r[:, 1] > 0]
                one.append((int(r[0, 7]), int(used[i].sum()), (int(r[:, 0].min()) - t0) / 1e3,
                            ((int(ready.min()) if ready.numel() else int(r[:, 0].min())) - t0) / 1e3,
                            (int(r[:, 6].max()) - t0) / 1e3))
            one.sort(key=lambda x: x[2])
            samples.append(one)
    finally:
        _lib.check(lib.stil_debug_trace(None))
    if not samples:
        return []
    names = {0: "STATS", 1: "STORE", 2: "GRAD", 3: "BWD"}
    out = []
    full = [s for s in samples if len(s) == len(samples[0])]
    for j in range(len(samples[0])):
        col = [s[j] for s in full]
        med = lambda k: sorted(c[k] for c in col)[len(col) // 2]
        out.append({"mode": names.get(col[0][0], "?"), "ctas": col[0][1], "start_us": round(med(2), 2),
                    "ready_us": round(med(3), 2), "end_us": round(med(4), 2),
                    "exclusive_us": round(med(4) - max(med(2), med(3)), 2), "lifetime_us": round(med(4) - med(2), 2)})
    return out


def parity_check(head, cfg, rank, world, dev, dist_on):
    """Untimed pre-check of the benchmarked path against the oracle (CPU, fp32) on one seeded batch per rank: N = 1: the
    whole step; N > 1: DistributedSTiLHead against the reference CLIPLoss on the CONCATENATED batch (oracle
    clip_loss_global) plus the sum of all ranks' prototype partials, like tests/test_gpu_dist.py.  Returns the measured
    errors; `ok` is the AND over ranks.  The caller exits non-zero when it is false."""
    import torch.distributed as dist
    from oracle import stil_head_oracle as O
    from stil_tta_b200 import synth
    torch.set_num_threads(min(8, max(1, len(os.sched_getaffinity(0)) // max(world, 1))))
    batches = [synth.make_batch(cfg, seed=4242, rank=r) for r in range(world)]
    keep = [t.clone() for t in (head.prototypes_sum, head.prototypes_count_sum)]
    head.prototypes_sum.zero_()
    head.prototypes_count_sum.zero_()
    head.load(batches[rank])
    head.run()
    torch.cuda.synchronize(dev)
    out = {k: v.detach().float().cpu() if v.is_floating_point() else v.cpu() for k, v in head.out.items()}
    psum = head.prototypes_sum.cpu()
    os_ = [O.head_step(b, cfg, with_grads=(world == 1)) for b in batches]
    o = os_[rank]
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    res = {}
    if world == 1:
        loss_ref, ga, gb = o["loss_itc"], o["d_feat_i"], o["d_feat_t"]
        res["d_feat_m_rel"] = rel(out["d_feat_m"], o["d_feat_m"])
    else:
        a = [b["feat_i"].float().requires_grad_(True) for b in batches]
        bb = [b["feat_t"].float().requires_grad_(True) for b in batches]
        loss_ref, _, _ = O.clip_loss_global(a, bb, cfg.temperature, cfg.lambda_0)
        ga, gb = torch.autograd.grad(loss_ref, (a[rank], bb[rank]))
        loss_ref = loss_ref.detach()
    res["loss_rel"] = abs(float(out["losses"][0]) - float(loss_ref)) / abs(float(loss_ref))
    res["loss_pt_rel"] = abs(float(out["losses"][1]) - float(o["loss_pt"])) / max(abs(float(o["loss_pt"])), 1e-6)
    res["grad_rel"] = max(rel(out["d_feat_i"], ga), rel(out["d_feat_t"], gb))
    amb = O.ambiguous_rows(batches[rank], cfg)
    dec_ok = all(torch.equal(out[k][~amb], o[k][~amb]) for k in ("max_idx", "mask1", "case1", "case2_i", "case2_t", "case3"))
    res["decisions_bit_exact"] = bool(dec_ok)
    res["ambiguous_rows"] = int(amb.sum())
    cs = sum(x["class_sum"] for x in os_)
    res["class_sum_abs"] = float((out["class_sum"] - cs).abs().max())
    res["prototypes_sum_abs"] = float((psum - cs).abs().max())
    ok = (res["loss_rel"] <= 1e-3 and res["loss_pt_rel"] <= 1e-3 and res["grad_rel"] <= 1e-3 and dec_ok and
          res["class_sum_abs"] <= 1e-4 and res["prototypes_sum_abs"] <= 1e-4 and res.get("d_feat_m_rel", 0.0) <= 1e-3)
    if dist_on:
        t = torch.tensor([1.0 if ok else 0.0, -res["loss_rel"], -res["grad_rel"], -res["class_sum_abs"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t[0].item() > 0.5)
        res["loss_rel"], res["grad_rel"], res["class_sum_abs"] = -float(t[1]), -float(t[2]), -float(t[3])
        res["over"] = f"max over {world} ranks"
    res["ok"] = ok
    res["tolerance"] = "losses, gradients 1e-3 relative (max|diff|/max|ref|); decisions bit-exact on non-ambiguous rows"
    res["oracle"] = "oracle/stil_head_oracle.py on CPU fp32" + (" (clip_loss_global on the concatenated batch)" if world > 1 else "")
    head.prototypes_sum.copy_(keep[0])
    head.prototypes_count_sum.copy_(keep[1])
    torch.cuda.synchronize(dev)
    return res


def init_dist(dev):
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # keep NCCL's version banner (printed at NCCL_DEBUG=VERSION/WARN) off stdout: rank 0 prints ONE JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ["NCCL_DEBUG_FILE"] = os.devnull
    # ... and whatever else the communicator set-up prints goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def run_gpu_arm(args):
    import torch.distributed as dist
    import stil_tta_b200 as S
    from stil_tta_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the STiL head has no CPU fallback (use --impl reference "
                         "for the host-CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        init_dist(dev)
    if args.config == "C5":
        return run_bank_arm(args, rank, world, dev, dist_on)
    cfg = get_cfg(args.config)
    pk = peaks()

    # Two heads (double buffering for the end-to-end copy overlap) with static buffers and one captured graph each.
    # Fresh inputs every step come from a pool of distinct batches in HBM that is larger than L2: the resident
    # measurement streams step i's inputs from pool[i % npool] into the head's input buffer with one device-to-device
    # copy INSIDE the timed region (like the encoders writing their outputs right before the head runs), so no step
    # ever re-reads cached inputs.  (Rotating over dozens of separately captured graphs instead would measure the
    # driver's graph-instance switching, not the step: 38 us/step with 4 instances, 48 with 15, 64 with 40.)
    Head = (lambda c, device: S.DistributedSTiLHead(c, device=device, use_graph=not args.no_graph, transport=args.transport)) if dist_on else \
           (lambda c, device: S.STiLHead(c, device=device))
    heads = [Head(cfg, device=dev) for _ in range(2)]
    nheads = len(heads)
    in_bytes = heads[0].h2d_bytes
    npool = args.nbuf if args.nbuf else max(4, int(1.25 * L2_BYTES / in_bytes) + 1)
    host_batches = [synth.make_batch(cfg, seed=2022 + i, rank=rank) for i in range(8)]
    for i, h in enumerate(heads):
        h.load(host_batches[i])
        if not (dist_on and args.no_graph):
            h.capture()
    # ---- parity of the benchmarked path (same head objects, same captured graphs) before anything is timed
    parity = parity_check(heads[0], cfg, rank, world, dev, dist_on)
    if not parity["ok"]:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": "parity check failed", "parity_check": parity, "n_gpus": world}),
                  flush=True)
        sys.stdout.flush()
        os._exit(3)
    heads[0].load(host_batches[0])
    pinned = [heads[0].pin(b) for b in host_batches]
    pool = []
    for i in range(npool):
        t = torch.empty_like(heads[0]._packed_in)
        t.copy_(pinned[i % len(pinned)], non_blocking=True)
        pool.append(t)
    for r in range(2):
        for h in heads:
            h.run()
    torch.cuda.synchronize(dev)

    # All measurements use the same two-deep pipeline: the batch of step i+1 is copied into the other head's input
    # buffer on a copy stream while step i computes (a data loader / encoder running one step ahead); each step waits
    # for its own batch, and every byte moved is inside the timed region.
    #   resident: source = the >L2 pool in HBM (device-to-device)                                   -> `value`
    #   e2e     : source = pinned HOST memory (3.8 MB H2D), plus D2H of the five losses              -> `e2e`
    #   e2e_full: same, plus D2H of EVERY per-batch output (gradients, pseudo labels, masks) on a second copy stream
    copy_stream = torch.cuda.Stream(dev)
    out_stream = torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(nheads)]
    ev_done = [torch.cuda.Event() for _ in range(nheads)]
    ev_out = [torch.cuda.Event() for _ in range(nheads)]

    def make_step(source, read_back):
        for e in ev_done + ev_out:
            e.record()

        def step(i):
            j = i % nheads
            h = heads[j]
            cur = torch.cuda.current_stream(dev)
            copy_stream.wait_event(ev_done[j])          # the previous user of this input buffer has finished
            if i < 2:
                copy_stream.wait_stream(cur)
            with torch.cuda.stream(copy_stream):
                h.copy_in(source(i))
                ev_in[j].record(copy_stream)
            cur.wait_event(ev_in[j])
            if read_back == 2:
                cur.wait_event(ev_out[j])               # this head's previous outputs have left for the host
            h.run()
            ev_done[j].record(cur)
            if read_back == 1:
                h._losses_host.copy_(h.out["losses"], non_blocking=True)
            elif read_back == 2:
                out_stream.wait_event(ev_done[j])
                with torch.cuda.stream(out_stream):
                    h.copy_out_full()
                    ev_out[j].record(out_stream)
        return step

    step_resident = make_step(lambda i: pool[i % npool], 0)
    step_e2e = make_step(lambda i: pinned[i % len(pinned)], 1)
    step_e2e_full = make_step(lambda i: pinned[i % len(pinned)], 2)

    with ClockSampler(local) as cs:
        time.sleep(0.3)                      # let nvidia-smi start polling before the load begins
        ms = timed_region(step_resident, args.steps, args.warmup, dist_on, dev)
        ms_e2e = timed_region(step_e2e, args.steps, args.warmup, dist_on, dev)
        ms_full = timed_region(step_e2e_full, args.steps, args.warmup, dist_on, dev, prewarm=PREWARM_STEPS // 4)
        torch.cuda.current_stream(dev).wait_stream(out_stream)
        torch.cuda.synchronize(dev)
    clocks = cs.summary()
    ms_step, ms_step_e2e, ms_step_full = ms / args.steps, ms_e2e / args.steps, ms_full / args.steps
    value = cfg.batch * world / (ms_step * 1e-3)
    e2e = cfg.batch * world / (ms_step_e2e * 1e-3)

    if dist_on:
        coll = ("peer-memory exchange kernels over NVLink (CUDA IPC; remote stores + flags), captured in the CUDA graphs: "
                "gather([feat_i|feat_t]), gather(loss, LSE), gather(class partials); InfoNCE chain on its own stream"
                if args.transport == "p2p" else
                "fused compute + exchange over NVLink peer memory (CUDA IPC; remote stores + arrival flags), captured in the CUDA "
                "graphs: one kernel packs [feat_i|feat_t], normalises and stores the rows into every rank's buffer; the statistics "
                "GEMM consumes each peer's rows as its flag lands; one kernel merges and stores the row LSEs; the gradient GEMM's "
                "epilogue waits for the column LSEs it reads; class partials pushed at the end of the row-local chain"
                if args.transport == "fused" else
                "NCCL, captured in the CUDA graph: all_gather([feat_i|feat_t]), all_reduce(loss, LSE slots), "
                "all_reduce(class_sum|class_count); InfoNCE chain on its own stream")
        nce = f"global batch {cfg.batch * world} (all-gathered)"
    else:
        coll, nce = "none (one GPU)", f"local batch {cfg.batch}"
    l2 = (f"inputs larger than L2: every step's {in_bytes / 2**20:.2f} MiB of inputs are streamed (D2D on a copy "
          f"stream, inside the timed region, one step ahead) from a pool of {npool} distinct batches = "
          f"{npool * in_bytes / 2**20:.0f} MiB in HBM"
          if npool * in_bytes > L2_BYTES else f"EXPERIMENT: pool of {npool} batches fits in L2")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "prewarm_steps": PREWARM_STEPS,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("bf16 operands / f32 accumulate / f32 gradients" if cfg.embed_dtype == "bf16"
                  else "f32 (3xbf16 split) / f32 accumulate"),
        "data": "synthetic",
        "config": make_config(workload_name(cfg, args.config), cfg.batch, world, l2,
                              True if not dist_on else not args.no_graph, coll, nce),
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_step_e2e, "h2d_bytes_per_step": heads[0].h2d_bytes,
                "d2h_bytes_per_step": heads[0].d2h_bytes},
        "e2e_full": {"value": cfg.batch * world / (ms_step_full * 1e-3), "unit": UNIT, "ms_per_step": ms_step_full,
                     "h2d_bytes_per_step": heads[0].h2d_bytes, "d2h_bytes_per_step": heads[0].d2h_bytes_full,
                     "what": "losses, d_feat_i/t/m, d_y_m/i/t, pseudo_label, max_prob, max_idx and the five masks copied back "
                             "to pinned host memory every step (one D2H on a second copy stream)"},
        "gpu_launches": heads[0].launches_per_step * args.steps,   # + one D2D input copy per step (not ours)
        "parity_check": parity,
        "clocks": clocks,
    }
    if rank == 0 and dist_on:
        print(json.dumps(line), flush=True)
    if rank == 0 and not dist_on:
        B, B_u, K, P = cfg.batch, cfg.b_u, cfg.num_classes, cfg.proj_dim
        # ---- dominant kernel: gemm_tc05_kernel.  In-graph durations from the CTAs' own %globaltimer stamps of a captured
        # step (PDL overlap included): exclusive = last CTA end - max(first CTA start, first wait passed), so the launches of
        # one chain add up to no more than the step.
        gl = gemm_ingraph_times(cfg, dev, host_batches[0])
        # algorithmic GEMM flops of one head step: a1 6*B^2*P, a3 2*B_u*K*P, a4 4*B*K*P — no
        # recompute, no padding, no split-precision passes — spread over the kernel's launches per step
        flops_step = 6.0 * B * B * P + 2.0 * B_u * K * P + 4.0 * B * K * P
        if gl:
            n_l = len(gl)
            t_avg = sum(x["exclusive_us"] for x in gl) / n_l * 1e-6
            ach = flops_step / n_l / t_avg / 1e12
            line["roofline"] = {"kernel": "gemm_tc05_kernel", "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"],
                                "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                                "traffic": ncu_traffic("gemm_tc05_kernel") if args.config == "C2" else None,
                                "peak_source": pk["src"] + " bf16 sustained", "launches_per_step": n_l,
                                "us_per_launch": t_avg * 1e6, "alg_flops_per_launch": flops_step / n_l,
                                "timing": "in-graph: %globaltimer stamps of every CTA of a captured step (stil_debug_trace), "
                                          "median of 6 replays; per launch last CTA end - max(first CTA start, predecessor "
                                          "complete)", "launches": gl,
                                "sum_exclusive_us": round(sum(x["exclusive_us"] for x in gl), 2),
                                "step_us": round(ms_step * 1e3, 2)}
        # un-captured per-launch durations (CUDA events around each launch; no PDL overlap) kept for comparison
        per = {}
        for i in range(3 + 20):
            heads[i % nheads]._packed_in.copy_(pool[i % npool], non_blocking=True)
            tr = heads[i % nheads].timed_run(names=True)
            if i >= 3:
                for name, ms1 in tr:
                    per.setdefault(name, []).append(ms1)
        line["kernel_us_uncaptured"] = {k: round(sum(v) / len(v) * 1e3, 2) for k, v in per.items()}
        # ---- HBM-bound row kernel at the benchmarked shape, then the large-shape sweeps (same run, same clocks record)
        rl = row_kernel_rooflines(B_u, K, dev, pk, nbuf=48)
        top = rl[0]
        line["roofline_hbm_kernel"] = {"kernel": "cgpl_pgls_kernel", "bound": "hbm", "achieved": top["achieved"],
                                       "peak": top["peak"], "unit": "GB/s", "frac": top["frac"],
                                       "us_per_launch": top["us_per_launch"], "alg_bytes_per_launch": top["alg_bytes_per_launch"],
                                       "traffic": ncu_traffic("cgpl_pgls_kernel") if args.config == "C2" else None}
        if not args.no_sweep:
            with ClockSampler(local) as cs2:
                sweep = clip_loss_sweep(dev, pk)
                sweep += row_kernel_rooflines(1 << 18, 286, dev, pk, nbuf=2, reps=4)
                sweep += row_kernel_rooflines(1 << 20, 2, dev, pk, nbuf=8, reps=8)
            line["roofline_sweep"] = sweep
            line["roofline_sweep_clocks"] = cs2.summary()
        if world == 1 and not args.no_cpu_baseline:
            ms_cpu, done, n = cpu_reference_steps(cfg, 400, 3, budget_s=15.0)
            line["cpu_baseline"] = {"value": cfg.batch / (ms_cpu * 1e-3), "unit": UNIT, "cores": n, "kind": "port",
                                    "sample": f"{done} head steps (fwd+bwd) of the oracle port on the host, "
                                              f"{ms_cpu:.2f} ms/step, best thread count of 1..{len(os.sched_getaffinity(0))} = {n}"}
        if args.torch_gpu_baseline:
            # "beat the library" bar: the same oracle port, eager PyTorch on THIS GPU (fp32, fwd+bwd); the
            # oracle is the thing measured here only as a baseline, like the cpu_baseline leg
            try:
                from oracle import stil_head_oracle as O
                gb = [{k: v.to(dev) for k, v in synth.make_batch(cfg, seed=100 + i).items()} for i in range(4)]
                st = {"prototypes_sum": torch.zeros(cfg.num_classes, cfg.proj_dim, device=dev),
                      "prototypes_count_sum": torch.zeros(cfg.num_classes, 1, device=dev)}
                for i in range(5):
                    O.head_step(gb[i % 4], cfg, state=st)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                nstep = 100
                e0.record()
                for i in range(nstep):
                    O.head_step(gb[i % 4], cfg, state=st)
                e1.record()
                torch.cuda.synchronize(dev)
                ms_t = e0.elapsed_time(e1) / nstep
                line["torch_gpu_baseline"] = {"value": cfg.batch / (ms_t * 1e-3), "unit": UNIT, "ms_per_step": ms_t,
                                              "kind": "oracle port, eager torch CUDA fp32 on the same B200 (launch-bound)",
                                              "sample": f"{nstep} head steps (fwd+bwd)"}
            except Exception as ex:   # a baseline must never take the bench line down
                line["torch_gpu_baseline"] = {"error": repr(ex)[:200]}
        print(json.dumps(line), flush=True)
    if dist_on:
        # graphs that captured NCCL work must die before the communicator; then leave without the (slow, and with
        # captured collectives occasionally hanging) process-group teardown
        torch.cuda.synchronize(dev)
        dist.barrier()
        for h in heads:
            h.release()
        del heads
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------ C5: SimMatch bank sweep
BANK_METRIC = "simmatch_bank_sweep_samples_per_sec"


def bank_workload(world):
    return (f"C5 SimMatch memory-bank block (simmatch_model.py:268-286): bank 65536 x 512 bf16, 286 classes, 448 unlabelled "
            f"rows per GPU, fwd+bwd, bank column-sharded over {world} GPU(s)")


def run_bank_reference_arm(args):
    """The oracle port of the SimMatch bank block on the host cores (one rank's 448 rows against the whole bank)."""
    from oracle import stil_head_oracle as O
    from stil_tta_b200 import synth
    rows, kb, d, c = 448, 65536, 512, 286
    bk = synth.make_bank(kb, d, c, dtype=torch.float32)
    g = torch.Generator().manual_seed(3)
    unit = torch.nn.functional.normalize
    fk = unit(torch.randn(rows, d, generator=g))
    p = torch.softmax(torch.randn(rows, c, generator=g) * 3, 1)
    n = len(os.sched_getaffinity(0))
    torch.set_num_threads(n)

    def step():
        fq = unit(fk + 0.2 * torch.randn(rows, d, generator=g)).requires_grad_(True)
        out = O.simmatch_bank(fk, fq, p, bk["bank"], bk["labels"], 0.1, 0.1, 0.9)
        out["loss_in"].mean().backward()
    for _ in range(min(args.warmup, 2)):
        step()
    t0, done = time.perf_counter(), 0
    for _ in range(args.steps):
        step()
        done += 1
        if time.perf_counter() - t0 > 60:
            break
    ms = (time.perf_counter() - t0) / done * 1e3
    value = rows / (ms * 1e-3)
    line = {"impl": "reference", "metric": BANK_METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": make_config(bank_workload(args.gpus), rows, args.gpus, "n/a (host)", False, "n/a (one host process)", "n/a"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n, "kind": "port",
                             "sample": f"{done} fwd+bwd sweeps of the oracle port (448 rows x 65536 x 512, fp32, torch CPU, {n} threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_bank_arm(args, rank, world, dev, dist_on):
    """C5 (BASELINE config 5): every rank holds a column shard [512, 65536 / N] of the bank and 448 query rows; the queries
    of all ranks are all-gathered, every rank sweeps its shard for all N*448 rows, the per-row statistics are combined
    across ranks and every rank finishes its own rows (stil_tta_b200.ShardedSimMatchBank)."""
    import torch.distributed as dist
    import stil_tta_b200 as S
    from stil_tta_b200 import synth
    pk = peaks()
    rows, kb, d, c = 448, 65536, 512, 286
    bk = synth.make_bank(kb, d, c)
    unit = torch.nn.functional.normalize
    shard = kb // world
    sb = S.ShardedSimMatchBank(d, kb, c, dtype=torch.bfloat16, device=dev, use_graph=not args.bank_eager)
    sb.load_shard(bk["bank"][rank * shard:(rank + 1) * shard], bk["labels"][rank * shard:(rank + 1) * shard])
    nb = 8
    gen = torch.Generator().manual_seed(100 + rank)
    fks = [unit(bk["bank"].float()[torch.randint(0, kb, (rows,), generator=gen)] + 0.3 * torch.randn(rows, d, generator=gen)).to(torch.bfloat16)
           for _ in range(nb)]
    fqs = [unit(f.float() + 0.2 * torch.randn(rows, d, generator=gen)).to(torch.bfloat16) for f in fks]
    ps = [torch.softmax(torch.randn(rows, c, generator=gen) * 3, 1) for _ in range(nb)]
    # ---- parity before timing: every rank's rows against the oracle on the WHOLE bank
    from oracle import stil_head_oracle as O
    torch.set_num_threads(min(16, max(1, len(os.sched_getaffinity(0)) // world)))
    fqr = fqs[0].float().requires_grad_(True)
    ref = O.simmatch_bank(fks[0].float(), fqr, ps[0], bk["bank"].float(), bk["labels"], 0.1, 0.1, 0.9)
    (g_ref,) = torch.autograd.grad(ref["loss_in"].mean(), fqr)
    fqc = fqs[0].to(dev).requires_grad_(True)
    prob_ku, loss_in = sb(fks[0].to(dev), fqc, ps[0].to(dev), 0.1, 0.1, 0.9)
    (g_c,) = torch.autograd.grad(loss_in.mean(), fqc)
    torch.cuda.synchronize(dev)
    def rel(a, b):
        """max |a - b| / max |b|; a bf16-typed result (autograd hands the bf16 leaf a bf16 gradient) is allowed its own output
        rounding, 2^-8 |b| element-wise, like tests/conftest.py:assert_rel"""
        a_, b_ = a.detach().float().cpu(), b.detach().float()
        slack = b_.abs() * 2.0 ** -8 if a.dtype == torch.bfloat16 else torch.zeros_like(b_)
        return float(((a_ - b_).abs() - slack).clamp_min(0).max() / b_.abs().max().clamp_min(1e-30))
    parity = {"prob_ku_abs": float((prob_ku.cpu() - ref["prob_ku"]).abs().max()), "loss_in_rel": rel(loss_in, ref["loss_in"].detach()),
              "grad_rel": rel(g_c, g_ref)}
    ok = parity["prob_ku_abs"] <= 2e-5 and parity["loss_in_rel"] <= 1e-3 and parity["grad_rel"] <= 1e-3
    if dist_on:
        t = torch.tensor([1.0 if ok else 0.0, -parity["prob_ku_abs"], -parity["loss_in_rel"], -parity["grad_rel"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t[0].item() > 0.5)
        parity = {"prob_ku_abs": -float(t[1]), "loss_in_rel": -float(t[2]), "grad_rel": -float(t[3]), "over": f"max over {world} ranks"}
    parity["ok"] = ok
    parity["oracle"] = "oracle simmatch_bank (CPU fp32) on the whole 65536-column bank"
    if not ok:
        if rank == 0:
            print(json.dumps({"metric": BANK_METRIC, "error": "parity check failed", "parity_check": parity, "n_gpus": world}), flush=True)
        sys.stdout.flush()
        os._exit(3)
    dfk = [t.to(dev) for t in fks]
    dfq = [t.to(dev).requires_grad_(True) for t in fqs]
    dps = [t.to(dev) for t in ps]
    pin = [(a.pin_memory(), b.pin_memory(), p_.pin_memory()) for a, b, p_ in zip(fks, fqs, ps)]
    host_out = torch.zeros(rows, d, dtype=torch.bfloat16).pin_memory()

    def step(i):
        j = i % nb
        dfq[j].grad = None
        prob_ku, loss_in = sb(dfk[j], dfq[j], dps[j], 0.1, 0.1, 0.9)
        loss_in.mean().backward()

    def step_e2e(i):
        j = i % nb
        a, b, p_ = (t.to(dev, non_blocking=True) for t in pin[j])
        b.requires_grad_(True)
        prob_ku, loss_in = sb(a, b, p_, 0.1, 0.1, 0.9)
        loss_in.mean().backward()
        host_out.copy_(b.grad, non_blocking=True)

    steps = min(args.steps, 200)
    with ClockSampler(dev.index) as cs:
        time.sleep(0.3)
        ms = timed_region(step, steps, args.warmup, dist_on, dev, prewarm=20)
        ms_e2e = timed_region(step_e2e, steps, args.warmup, dist_on, dev, prewarm=5)
    ms_step, ms_step_e2e = ms / steps, ms_e2e / steps
    value = rows * world / (ms_step * 1e-3)
    fl = 6.0 * rows * world * kb * d          # algorithmic: teacher + student logits + dX, all ranks' rows against the whole bank
    line = {"metric": BANK_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 operands / f32 accumulate / f32 gradients", "data": "synthetic",
            "config": make_config(bank_workload(world), rows, world,
                                  "the bank shard (64 MiB / N) and the logits tiles stream from HBM every sweep; inputs rotate over 8 batches",
                                  not args.bank_eager,
                                  "NCCL all_gather(queries), all_reduce(per-row statistics), reduce_scatter(dX partials)" if dist_on else "none (one GPU)",
                                  "n/a"),
            "e2e": {"value": rows * world / (ms_step_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_step_e2e,
                    "h2d_bytes_per_step": 2 * rows * d * 2 + rows * c * 4, "d2h_bytes_per_step": rows * d * 2},
            "gpu_launches": sb.launches_per_step * steps, "parity_check": parity, "clocks": cs.summary(),
            "roofline": {"kernel": "simmatch bank sweep (all kernels of fwd+bwd)", "bound": "tensor", "achieved": fl / world / (ms_step * 1e-3) / 1e12,
                         "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": fl / world / (ms_step * 1e-3) / 1e12 / pk["bf16_tflops"],
                         "traffic": None, "basis": "algorithmic 6 * rows * K_b * D per rank-step (SURVEY §8d) / whole fwd+bwd time"}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist_on:
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-CPU leg (profiling runs)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the large-shape roofline sweeps (profiling runs)")
    ap.add_argument("--torch-gpu-baseline", action="store_true",
                    help="also time the oracle port as eager PyTorch on the GPU (N=1; reported beside cpu_baseline)")
    ap.add_argument("--transport", default="fused", choices=["fused", "p2p", "nccl"],
                    help="N>1 exchange: fused peer-memory schedule, blocking peer-memory all-gathers, or NCCL")
    ap.add_argument("--nbuf", type=int, default=0, help="experiments: override the number of rotating batches")
    ap.add_argument("--bank-eager", action="store_true", help="--config C5: eager launches instead of one CUDA graph per sweep")
    ap.add_argument("--no-graph", action="store_true", help="N>1 only: do not capture kernels+NCCL in a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
