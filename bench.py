#!/usr/bin/env python
"""bench.py — STiL head-step throughput on B200 (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2|C3]

One "step" = one pass of the whole per-batch head (CGPL, PGLS, InfoNCE fwd+bwd, prototype loss fwd+bwd,
masked soft-target CE fwd+bwd, prototype partial sums) over one synthetic DVM-shaped batch (C2: B=512 = 64
labelled + 448 unlabelled, K=286, P=128, bf16 embeddings).  `value` = samples/s with inputs resident in HBM
(CUDA-graph replay), `e2e` = the same through STiLHead.step_host with pinned HOST buffers (H2D of every
input and D2H of the losses inside the timed region).  Timed steps rotate over enough distinct batches that
the touched working set exceeds the 126 MB L2.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
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
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/), or None."""
    f = REPO / "profiles" / "r1_traffic.json"
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
    # give the reference its best thread count (oversubscribed intra-op threads can be slower than one)
    best = (float("inf"), 1)
    for n in sorted({1, 2, 4, 8, 16, 32, ncpu}):
        if n > ncpu:
            continue
        torch.set_num_threads(n)
        O.head_step(batches[0], cfg, state=state)
        t0 = time.perf_counter()
        for i in range(3):
            O.head_step(batches[i % 4], cfg, state=state)
        best = min(best, ((time.perf_counter() - t0) / 3, n))
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
    cfg = get_cfg(args.config)
    ms, done, n = cpu_reference_steps(cfg, args.steps, args.warmup, budget_s=120.0)
    value = cfg.batch / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, args.config), "l2": "n/a (host)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n, "kind": "port",
                         "sample": f"{done} head steps (fwd+bwd) of the oracle port of STiLModel.training_step's head, "
                                   f"torch {torch.__version__} CPU fp32, best of 1..{len(os.sched_getaffinity(0))} threads = {n}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def timed_region(fn_step, steps, warmup, dist_on, dev):
    """warm-up, barrier+sync, K steps between CUDA events on the current stream, barrier+sync; ms total."""
    import torch.distributed as dist
    # untimed pre-warm before the W warm-up steps: a fixed NUMBER of steps (the same on every rank — the data-parallel
    # step is collective) worth ~0.3 s, so that a short timed region (small K) is not measured on clocks still ramping up
    # from the idle state the clock sampler's start-up sleep leaves the GPU in
    pre = PREWARM_STEPS
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


def kernel_rooflines(cfg, dev, pk, reps=20, nbuf=48):
    """Average launch duration of the row kernels, measured live with CUDA events around a CUDA graph of
    `reps` back-to-back launches on rotating buffers (so the working set exceeds L2)."""
    import stil_tta_b200 as S
    from stil_tta_b200 import _lib
    from stil_tta_b200._lib import ptr
    lib = _lib.load()
    B_u, K = cfg.b_u, cfg.num_classes
    g = torch.Generator(device="cpu").manual_seed(1)
    res = {}
    # ---- cgpl_pgls: reads 3 logit rows + teacher logits (f32), writes pseudo_label + per-row outputs
    ys = [[torch.randn(B_u, K, generator=g).to(dev) * 3 for _ in range(3)] for _ in range(nbuf)]
    tl = [torch.randn(B_u, K, generator=g).to(dev) for _ in range(nbuf)]
    pl = [torch.empty(B_u, K, device=dev) for _ in range(nbuf)]
    mp = torch.empty(B_u, device=dev); mi = torch.empty(B_u, dtype=torch.int64, device=dev)
    fl = [torch.empty(B_u, dtype=torch.bool, device=dev) for _ in range(5)]
    cls = torch.empty(B_u, dtype=torch.int32, device=dev); conf = torch.empty(B_u, dtype=torch.bool, device=dev)

    def launch_cgpl(i):
        j = i % nbuf
        _lib.check(lib.stil_cgpl_pgls(ptr(ys[j][0]), ptr(ys[j][1]), ptr(ys[j][2]), 0, K, ptr(tl[j]), K, B_u, K,
                                      cfg.temperature, cfg.rate_pseudo, cfg.th1, 1, ptr(pl[j]), K, None, 0, ptr(mp),
                                      ptr(mi), ptr(fl[0]), ptr(fl[1]), ptr(fl[2]), ptr(fl[3]), ptr(fl[4]), None,
                                      ptr(cls), ptr(conf), _lib.stream_ptr(dev)))

    def time_graph(launch):
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
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / (5 * reps) * 1e-3   # seconds per launch

    t = time_graph(launch_cgpl)
    bytes_alg = 4 * B_u * K * 4 + B_u * K * 4 + B_u * (4 + 8 + 5 + 4 + 1)
    res["cgpl_pgls_kernel"] = {"bound": "hbm", "seconds": t, "alg_bytes": bytes_alg,
                               "achieved": bytes_alg / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
    return res


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

    # Both measurements use the same two-deep pipeline: the batch of step i+1 is copied into the other head's input
    # buffer on a copy stream while step i computes (a data loader / encoder running one step ahead); each step waits
    # for its own batch, and every byte moved is inside the timed region.
    #   resident: source = the >L2 pool in HBM (device-to-device)          -> `value`
    #   e2e     : source = pinned HOST memory (3.8 MB H2D), plus D2H of the five losses -> `e2e`
    copy_stream = torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(nheads)]
    ev_done = [torch.cuda.Event() for _ in range(nheads)]

    def make_step(source, read_back):
        for e in ev_done:
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
            h.run()
            if read_back:
                h._losses_host.copy_(h.out["losses"], non_blocking=True)
            ev_done[j].record(cur)
        return step

    step_resident = make_step(lambda i: pool[i % npool], False)
    step_e2e = make_step(lambda i: pinned[i % len(pinned)], True)

    with ClockSampler(local) as cs:
        time.sleep(0.3)                      # let nvidia-smi start polling before the load begins
        ms = timed_region(step_resident, args.steps, args.warmup, dist_on, dev)
        ms_e2e = timed_region(step_e2e, args.steps, args.warmup, dist_on, dev)
    clocks = cs.summary()
    ms_step, ms_step_e2e = ms / args.steps, ms_e2e / args.steps
    value = cfg.batch * world / (ms_step * 1e-3)
    e2e = cfg.batch * world / (ms_step_e2e * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 operands / f32 accumulate" if cfg.embed_dtype == "bf16" else "f32 (3xbf16 split) / f32 accumulate",
        "data": "synthetic",
        "config": {"workload": workload_name(cfg, args.config), "per_gpu_batch": cfg.batch,
                   "l2": (f"inputs larger than L2: every step's {in_bytes / 2**20:.2f} MiB of inputs are streamed (D2D on a copy "
                          f"stream, inside the timed region, one step ahead) from a pool of {npool} distinct batches = "
                          f"{npool * in_bytes / 2**20:.0f} MiB in HBM"
                          if npool * in_bytes > L2_BYTES else f"EXPERIMENT: pool of {npool} batches fits in L2"),
                   "parallelism": f"dp{world}", "cuda_graph": True,
                   "prewarm": f"{PREWARM_STEPS} untimed steps before the W warm-up steps of each timed region (clock ramp)"},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_step_e2e, "h2d_bytes_per_step": heads[0].h2d_bytes,
                "d2h_bytes_per_step": heads[0].d2h_bytes},
        "gpu_launches": heads[0].launches_per_step * args.steps,   # + one D2D input copy per step (not ours)
        "clocks": clocks,
    }
    # live per-launch durations of the main-chain kernels: CUDA events recorded on the launching stream around
    # every launch of an un-captured step (stil_head_step's instrumentation hook), rotating over the batches
    per = {}
    reps = 0 if dist_on else max(20, min(args.steps, 100))
    for i in range((3 + reps) if reps else 0):
        heads[i % nheads]._packed_in.copy_(pool[i % npool], non_blocking=True)
        tr = heads[i % nheads].timed_run(names=True)
        if i >= 3:
            for name, ms1 in tr:
                per.setdefault(name, []).append(ms1)
    kern = {k: sum(v) / len(v) * 1e3 for k, v in per.items()}      # us per launch
    if rank == 0 and dist_on:
        line["config"]["cuda_graph"] = not args.no_graph
        line["config"]["collectives"] = (
            "peer-memory exchange kernels over NVLink (CUDA IPC; remote stores + flags), captured in the CUDA graphs: "
            "gather([feat_i|feat_t]), gather(loss, LSE), gather(class partials); InfoNCE chain on its own stream"
            if args.transport == "p2p" else
            "fused compute + exchange over NVLink peer memory (CUDA IPC; remote stores + arrival flags), captured in the CUDA "
            "graphs: one kernel packs [feat_i|feat_t], normalises and stores the rows into every rank's buffer; the statistics "
            "GEMM consumes each peer's rows as its flag lands; one kernel merges and stores the row LSEs; the gradient GEMM's "
            "epilogue waits for the column LSEs it reads; class partials pushed at the end of the row-local chain"
            if args.transport == "fused" else
            "NCCL, captured in the CUDA graph: all_gather([feat_i|feat_t]), all_reduce(loss, LSE slots), "
            "all_reduce(class_sum|class_count); InfoNCE chain on its own stream")
        line["config"]["infonce"] = f"global batch {cfg.batch * world} (all-gathered)"
        print(json.dumps(line), flush=True)
    if rank == 0 and not dist_on:
        B, B_u, K, P = cfg.batch, cfg.b_u, cfg.num_classes, cfg.proj_dim
        gemm_us = [v for k, v in kern.items() if k.startswith("gemm_tc05_kernel")]
        # algorithmic GEMM flops of one head step: a1 6*B^2*P, a3 2*B_u*K*P, a4 4*B*K*P — no
        # recompute, no padding, no split-precision passes — spread over the kernel's launches per step
        flops_step = 6.0 * B * B * P + 2.0 * B_u * K * P + 4.0 * B * K * P
        n_l = len(gemm_us)
        t_avg = sum(gemm_us) / n_l * 1e-6
        ach = flops_step / n_l / t_avg / 1e12
        line["roofline"] = {"kernel": "gemm_tc05_kernel", "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"],
                            "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                            "traffic": ncu_traffic("gemm_tc05_kernel") if args.config == "C2" else None,
                            "peak_source": pk["src"] + " bf16 sustained", "launches_per_step": n_l,
                            "us_per_launch": t_avg * 1e6, "alg_flops_per_launch": flops_step / n_l,
                            "share_of_timed_launches": sum(gemm_us) / sum(kern.values())}
        rl = kernel_rooflines(cfg, dev, pk)
        top = rl["cgpl_pgls_kernel"]
        line["roofline_hbm_kernel"] = {"kernel": "cgpl_pgls_kernel", "bound": "hbm", "achieved": top["achieved"],
                                       "peak": top["peak"], "unit": "GB/s", "frac": top["achieved"] / top["peak"],
                                       "us_per_launch": top["seconds"] * 1e6, "alg_bytes_per_launch": top["alg_bytes"],
                                       "traffic": ncu_traffic("cgpl_pgls_kernel") if args.config == "C2" else None}
        line["kernel_us"] = {k: round(v, 2) for k, v in kern.items()}
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-CPU leg (profiling runs)")
    ap.add_argument("--torch-gpu-baseline", action="store_true",
                    help="also time the oracle port as eager PyTorch on the GPU (N=1; reported beside cpu_baseline)")
    ap.add_argument("--transport", default="fused", choices=["fused", "p2p", "nccl"],
                    help="N>1 exchange: fused peer-memory schedule, blocking peer-memory all-gathers, or NCCL")
    ap.add_argument("--nbuf", type=int, default=0, help="experiments: override the number of rotating batches")
    ap.add_argument("--no-graph", action="store_true", help="N>1 only: do not capture kernels+NCCL in a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
