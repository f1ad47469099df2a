"""Kernel timeline (torch profiler / CUPTI) of one data-parallel head step on rank 0.
torchrun --nproc-per-node N scripts/dist_timeline.py [p2p|nccl]"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
transport = sys.argv[1] if len(sys.argv) > 1 else "fused"
cfg = synth.CONFIGS["C2"]()
head = S.DistributedSTiLHead(cfg, device=dev, use_graph=True, transport=transport)
head.load(synth.make_batch(cfg, seed=2022, rank=rank))
head.capture()
for _ in range(10):
    head.run()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(500):
    head.run()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"world {world} transport {transport}: {e0.elapsed_time(e1) / 500 * 1e3:.1f} us/step", flush=True)
dist.barrier()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(6):
        head.run()
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([e for e in prof.events() if "cuda" in str(e.device_type).lower()], key=lambda e: e.time_range.start)
    n = len(evs) // 6
    last = evs[-n:]
    t0 = last[0].time_range.start
    for e in last:
        print(f"   +{e.time_range.start - t0:7.1f} us  {e.time_range.end - e.time_range.start:6.1f} us  {e.name[:100]}")
torch.cuda.synchronize()
dist.barrier()
head.release()
os._exit(0)
