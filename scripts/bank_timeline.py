"""Kernel timeline (torch profiler / CUPTI) of one C5 SimMatch bank sweep (448 rows per rank x 65536 x 512, 286 classes),
single GPU or under torchrun (rank 0 prints).  Programmatic dependent launch is switched off for the profiled pass so
kernel records do not overlap."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import _lib, synth  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev = torch.device("cuda", torch.cuda.current_device())
rows, kb, d, c = 448, 65536, 512, 286
g = torch.Generator().manual_seed(rank)
bk = synth.make_bank(kb, d, c)
sb = S.ShardedSimMatchBank(d, kb, c, dtype=torch.bfloat16, device=dev)
sb.load(bk["bank"], bk["labels"])
unit = torch.nn.functional.normalize
fk = unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16).to(dev)
fq = unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
p = torch.softmax(torch.randn(rows, c, generator=g) * 3, 1).to(dev)


def step():
    fq.grad = None
    sb(fk, fq, p, 0.1, 0.1, 0.9)[1].mean().backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/step (fwd+bwd, eager)")
_lib.load().stil_debug_pdl(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
_lib.load().stil_debug_pdl(1)
if rank == 0:
    evs = sorted([e for e in prof.events() if "cuda" in str(e.device_type).lower()], key=lambda e: e.time_range.start)
    n = len(evs) // 3
    last = evs[-n:]
    t0 = last[0].time_range.start
    for e in last:
        print(f"   +{e.time_range.start - t0:7.1f} us  {e.time_range.end - e.time_range.start:6.1f} us  {e.name[:110]}")
if world > 1:
    dist.destroy_process_group()
