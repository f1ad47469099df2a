"""Kernel durations (torch profiler) of one CLIPLoss fwd+bwd at a large shape: python scripts/clip_profile.py [n] [d]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
g = torch.Generator().manual_seed(0)
a = torch.randn(n, d, generator=g).to(torch.bfloat16).cuda().requires_grad_(True)
b = torch.randn(n, d, generator=g).to(torch.bfloat16).cuda().requires_grad_(True)
crit = S.CLIPLoss(0.1, 0.5)
for _ in range(3):
    loss, _, _ = crit(a, b)
    loss.backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    loss, _, _ = crit(a, b)
    loss.backward()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if "cuda" in str(e.device_type).lower()], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"   +{e.time_range.start - t0:8.1f} us  {e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
