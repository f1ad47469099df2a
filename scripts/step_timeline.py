"""Kernel timeline (torch profiler / CUPTI) of one single-GPU head step (CUDA graph replay)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import synth  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]()
head = S.STiLHead(cfg, device="cuda")
head.load(synth.make_batch(cfg, seed=2022))
head.capture()
for _ in range(10):
    head.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(1000):
    head.run()
e1.record()
torch.cuda.synchronize()
print(f"{cfg.name} B={cfg.batch}: {e0.elapsed_time(e1)} us/step (L2-resident inputs, back-to-back graph replays)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(6):
        head.run()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if "cuda" in str(e.device_type).lower()], key=lambda e: e.time_range.start)
n = len(evs) // 6
last = evs[-n:]
t0 = last[0].time_range.start
prev_end = evs[-n - 1].time_range.end if len(evs) > n else t0
print(f"   previous step's last kernel ended {t0 - prev_end:.1f} us before this step's first kernel started")
for e in last:
    print(f"   +{e.time_range.start - t0:7.1f} us  {e.time_range.end - e.time_range.start:6.1f} us  {e.name[:100]}")
