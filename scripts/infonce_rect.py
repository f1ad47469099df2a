"""Single-GPU emulation of ONE rank's global-batch InfoNCE work (rows m = 512 local, columns n = 512*W gathered):
stil_infonce_fwd + stil_infonce_bwd on a [n, 2P] gathered buffer, captured in a CUDA graph.  No exchange kernels —
this isolates the compute that grows with the world size."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from stil_tta_b200 import _lib  # noqa: E402
from stil_tta_b200._lib import check  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
m, P = 512, 128
worlds = [int(x) for x in sys.argv[1:] if x.isdigit()] or [1, 2, 4, 8]
for W in worlds:
    n = m * W
    g = torch.Generator().manual_seed(W)
    ab = torch.nn.functional.normalize(torch.randn(n, 2, P, generator=g), dim=2).reshape(n, 2 * P).to(torch.bfloat16).to(dev)
    off = m * (W - 1)
    esz = 2
    a_all, b_all = ab.data_ptr(), ab.data_ptr() + P * esz
    a_loc, b_loc = a_all + off * 2 * P * esz, b_all + off * 2 * P * esz
    ws = torch.empty(lib.stil_infonce_workspace_bytes(m, n, P, 1), dtype=torch.uint8, device=dev)
    loss = torch.zeros(4, device=dev)
    lse = torch.zeros(2, n, device=dev)
    import os
    gf32 = os.environ.get("GRAD_F32", "1") == "1"          # the head's default: fp32 gradients (G as bf16 hi+lo)
    d_a = torch.empty(m, P, dtype=torch.float32 if gf32 else torch.bfloat16, device=dev)
    d_b = torch.empty_like(d_a)
    s = torch.cuda.Stream()

    def step(st):
        check(lib.stil_infonce_fwd(a_loc, b_loc, a_all, b_all, 1, m, n, P, 2 * P, off, 0.1, 0.5, loss.data_ptr(),
                                   lse[0, off:].data_ptr(), lse[1, off:].data_ptr(), None, 0, ws.data_ptr(), ws.numel(), st))
        check(lib.stil_infonce_bwd_after_fwd(a_all, b_all, 1, m, n, P, 2 * P, off, 0.1, 0.5, lse[0].data_ptr(),
                                             lse[1].data_ptr(), None, d_a.data_ptr(), d_b.data_ptr(), 0 if gf32 else 1, P, ws.data_ptr(),
                                             ws.numel(), st))

    with torch.cuda.stream(s):
        for _ in range(3):
            step(s.cuda_stream)
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            step(s.cuda_stream)
        for _ in range(10):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 500
        e0.record(s)
        for _ in range(reps):
            gr.replay()
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    flops = 6.0 * m * n * P * 2 / 2   # 6*G^2*D/W per rank = 6*m*n*P
    print(f"W={W} n={n}: fwd+bwd {us:7.2f} us/step  ({6.0 * m * n * P / us / 1e6:6.1f} TFLOP/s algorithmic)", flush=True)

if "--profile" in sys.argv or True:
    # per-kernel durations of the last configuration (eager launches, torch profiler / CUPTI)
    try:
        from torch.profiler import ProfilerActivity, profile
        with torch.cuda.stream(s):
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(5):
                    gr.replay()
                s.synchronize()
        evs = sorted([e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()],
                     key=lambda e: e.time_range.start)
        t0 = evs[-len(evs) // 5].time_range.start if evs else 0
        for e in evs[-len(evs) // 5:]:
            print(f"   +{e.time_range.start - t0:8.1f} us  {e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
    except Exception as ex:   # profiler unavailable on the box
        print("profiler unavailable:", ex)
