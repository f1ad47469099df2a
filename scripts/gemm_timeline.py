"""Per-CTA phase timeline of the tcgen05 GEMM launches of one head step (uses stil_debug_trace)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import _lib, synth  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]()
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
# the trace pointer travels in the kernel parameters, so install it before a graph is captured
buf = torch.zeros(64, 64, 8, dtype=torch.int64, device="cuda")
_lib.check(_lib.load().stil_debug_trace(buf.data_ptr()))
head = S.STiLHead(cfg, device="cuda", use_graph=graph)
head.load(synth.make_batch(cfg, seed=2022))
for _ in range(4):
    head.run()
torch.cuda.synchronize()
buf.zero_()
torch.cuda.synchronize()
for _ in range(2):
    head.run()
torch.cuda.synchronize()
_lib.check(_lib.load().stil_debug_trace(None))
t = buf.cpu()
used = (t[:, :, 0] > 0)
names = {0: "STATS", 1: "STORE", 2: "GRAD", 3: "BWD (slots: start, wait passed, row/col values, first G, last G, dX ready, end)"}
slots = ["start", "prologue", "tma_done", "mma_done", "epi_ready", "acc_ready", "end"]   # BWD: see names[3]
launch_ids = [i for i in range(64) if used[i].any()]
t0_all = min(int(t[i][used[i]][:, 0].min()) for i in launch_ids)
for i in launch_ids:
    rows = t[i][used[i]]
    base = int(rows[:, 0].min())
    print(f"launch slot {i}: mode={names.get(int(rows[0, 7]), '?')} ctas={rows.shape[0]} "
          f"first start @ {(base - t0_all) / 1e3:8.2f} us; last end +{(int(rows[:, 6].max()) - base) / 1e3:6.2f} us")
    for c in list(range(min(3, rows.shape[0]))) + ([rows.shape[0] - 1] if rows.shape[0] > 3 else []):
        r = rows[c]
        print("    cta", c, " ".join(f"{n}=+{(int(r[k]) - base) / 1e3:6.2f}" for k, n in enumerate(slots) if int(r[k]) > 0))
    d = (rows[:, 6] - rows[:, 0]).float() / 1e3
    print(f"    CTA lifetime us: mean {d.mean():.2f} max {d.max():.2f};  prologue {((rows[:,1]-rows[:,0]).float()/1e3).mean():.2f}"
          f"  acc_ready-prologue {((rows[:,5]-rows[:,1]).float()/1e3).mean():.2f}  epilogue {((rows[:,6]-rows[:,5]).float()/1e3).mean():.2f}")
