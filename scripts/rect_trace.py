"""Per-CTA phase stamps (stil_debug_trace) of the GEMM launches of ONE rank's global-batch InfoNCE backward on one GPU
(rows m = 512, columns n = 512 * W).  python scripts/rect_trace.py W   [env STIL_DX_CLUSTER=0|2|4|8]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from stil_tta_b200 import _lib  # noqa: E402
from stil_tta_b200._lib import check  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
m, P = 512, 128
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = m * W
g = torch.Generator().manual_seed(W)
ab = torch.nn.functional.normalize(torch.randn(n, 2, P, generator=g), dim=2).reshape(n, 2 * P).to(torch.bfloat16).to(dev)
off = m * (W - 1)
a_all, b_all = ab.data_ptr(), ab.data_ptr() + P * 2
a_loc, b_loc = a_all + off * 2 * P * 2, b_all + off * 2 * P * 2
ws = torch.empty(lib.stil_infonce_workspace_bytes(m, n, P, 1), dtype=torch.uint8, device=dev)
loss = torch.zeros(4, device=dev)
lse = torch.zeros(2, n, device=dev)
d_a = torch.empty(m, P, dtype=torch.float32, device=dev)
d_b = torch.empty_like(d_a)
buf = torch.zeros(64, 64, 8, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream


def step():
    check(lib.stil_infonce_fwd(a_loc, b_loc, a_all, b_all, 1, m, n, P, 2 * P, off, 0.1, 0.5, loss.data_ptr(),
                               lse[0, off:].data_ptr(), lse[1, off:].data_ptr(), None, 0, ws.data_ptr(), ws.numel(), st))
    check(lib.stil_infonce_bwd_after_fwd(a_all, b_all, 1, m, n, P, 2 * P, off, 0.1, 0.5, lse[0].data_ptr(), lse[1].data_ptr(),
                                         None, d_a.data_ptr(), d_b.data_ptr(), 0, P, ws.data_ptr(), ws.numel(), st))


for _ in range(3):
    step()
torch.cuda.synchronize()
check(lib.stil_debug_trace(buf.data_ptr()))
buf.zero_()
step()
torch.cuda.synchronize()
check(lib.stil_debug_trace(None))
t = buf.cpu()
used = t[:, :, 0] > 0
names = {0: "STATS", 1: "STORE", 2: "GRAD"}
slots = ["start", "wait_passed", "tma_done", "mma_done", "epi_ready", "acc_ready", "end"]
ids = [i for i in range(64) if used[i].any()]
t0 = min(int(t[i][used[i]][:, 0].min()) for i in ids)
for i in ids:
    rows = t[i][used[i]]
    base = int(rows[:, 0].min())
    print(f"launch {i}: {names.get(int(rows[0, 7]), '?')} ctas(traced)={rows.shape[0]} start @{(base - t0) / 1e3:7.2f} us, last end +{(int(rows[:, 6].max()) - base) / 1e3:6.2f} us")
    for c in sorted(set(list(range(min(9, rows.shape[0]))) + [rows.shape[0] - 1])):
        r = rows[c]
        print("    cta", c, " ".join(f"{nm}=+{(int(r[k]) - base) / 1e3:6.2f}" for k, nm in enumerate(slots) if int(r[k]) > 0))
