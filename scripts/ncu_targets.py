"""Small, fixed workloads for `ncu` captures of the hot kernels on shapes where their rooflines mean something
(profiles/README.md lists the exact command lines).

  python scripts/ncu_targets.py rows     # cgpl_pgls / masked_softce on 2^18 x 286 logits (HBM roofline)
  python scripts/ncu_targets.py rows2    # the same kernels on 2^20 x 2 (cardiac-shaped)
  python scripts/ncu_targets.py clip     # CLIPLoss forward + backward, 4096 x 2048 bf16 (tensor roofline)
  python scripts/ncu_targets.py step     # one C2 head step, un-captured
  python scripts/ncu_targets.py bank     # SimMatch bank block at the C5 shape (448 x 65536 x 512)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import _lib, synth  # noqa: E402
from stil_tta_b200._lib import ptr  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
what = sys.argv[1] if len(sys.argv) > 1 else "rows"
g = torch.Generator().manual_seed(0)

if what in ("rows", "rows2"):
    rows, k = ((1 << 18), 286) if what == "rows" else ((1 << 20), 2)
    ys = [(torch.randn(rows, k, generator=g) * 3).to(dev) for _ in range(3)]
    tl = torch.randn(rows, k, generator=g).to(dev)
    pl = torch.empty(rows, k, device=dev)
    mp = torch.empty(rows, device=dev); mi = torch.empty(rows, dtype=torch.int64, device=dev)
    fl = [torch.empty(rows, dtype=torch.bool, device=dev) for _ in range(5)]
    cls = torch.empty(rows, dtype=torch.int32, device=dev); conf = torch.empty(rows, dtype=torch.bool, device=dev)
    mr = (torch.rand(rows, generator=g) >= 0.5).to(dev)
    losses = torch.empty(3, device=dev)
    gr = [torch.empty(rows, k, device=dev) for _ in range(3)]
    ws = torch.zeros(lib.stil_masked_softce_workspace_bytes(rows), dtype=torch.uint8, device=dev)
    for _ in range(2):
        _lib.check(lib.stil_cgpl_pgls(ptr(ys[0]), ptr(ys[1]), ptr(ys[2]), 0, k, ptr(tl), k, rows, k, 0.1, 0.9, 0.9, 1, None, 0,
                                      ptr(pl), k, None, 0, ptr(mp), ptr(mi), ptr(fl[0]), ptr(fl[1]), ptr(fl[2]), ptr(fl[3]),
                                      ptr(fl[4]), None, ptr(cls), ptr(conf), _lib.stream_ptr(dev)))
        _lib.check(lib.stil_masked_softce(ptr(ys[0]), ptr(ys[1]), ptr(ys[2]), 0, k, ptr(pl), k, ptr(fl[0]), ptr(fl[1]),
                                          ptr(fl[2]), ptr(fl[3]), ptr(fl[4]), ptr(mr), rows, k, ptr(losses), ptr(gr[0]),
                                          ptr(gr[1]), ptr(gr[2]), k, 1.0, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
elif what == "clip":
    n, d = 4096, 2048
    a = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    b = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    crit = S.CLIPLoss(0.1, 0.5, return_logits=False)
    for _ in range(2):
        a.grad = b.grad = None
        crit(a, b)[0].backward()
elif what == "step":
    cfg = synth.CONFIGS["C2"]()
    head = S.STiLHead(cfg, device="cuda", use_graph=False)
    head.load(synth.make_batch(cfg, seed=2022))
    for _ in range(3):
        head.run()
elif what == "bank":
    rows, kb, d, c = 448, 65536, 512, 286
    bk = synth.make_bank(kb, d, c)
    sb = S.ShardedSimMatchBank(d, kb, c, dtype=torch.bfloat16, device="cuda")
    sb.load(bk["bank"], bk["labels"])
    unit = torch.nn.functional.normalize
    fk = unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16).to(dev)
    fq = unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
    p = torch.softmax(torch.randn(rows, c, generator=g) * 3, 1).to(dev)
    for _ in range(2):
        fq.grad = None
        sb(fk, fq, p, 0.1, 0.1, 0.9)[1].mean().backward()
torch.cuda.synchronize()
print("done", what)
