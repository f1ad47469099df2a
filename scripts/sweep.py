"""Roofline sweeps on shapes where the bounds are meaningful: row kernels on 2^18 rows (HBM) and the
InfoNCE / prototype GEMMs on large batches and wide embeddings (tensor pipe).  Prints one JSON object per line.

  python scripts/sweep.py [rows|gemm|all]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import _lib  # noqa: E402
from stil_tta_b200._lib import ptr  # noqa: E402

PEAKS = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text()) \
    if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
HBM, TF = PEAKS["hbm_gbs"], PEAKS.get("bf16_tflops_sustained", PEAKS.get("bf16_tflops"))
dev = torch.device("cuda")
lib = _lib.load()


def time_fn(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def rows_sweep():
    g = torch.Generator().manual_seed(0)
    for rows, k in ((1 << 18, 286), (1 << 20, 2), (1 << 16, 1000)):
        ys = [(torch.randn(rows, k, generator=g) * 3).to(dev) for _ in range(3)]
        tl = torch.randn(rows, k, generator=g).to(dev)
        pl = torch.empty(rows, k, device=dev)
        mp = torch.empty(rows, device=dev); mi = torch.empty(rows, dtype=torch.int64, device=dev)
        fl = [torch.empty(rows, dtype=torch.bool, device=dev) for _ in range(5)]

        def cg():
            _lib.check(lib.stil_cgpl_pgls(ptr(ys[0]), ptr(ys[1]), ptr(ys[2]), 0, k, ptr(tl), k, rows, k, 0.1, 0.9, 0.9, 1,
                                          None, 0, ptr(pl), k, None, 0, ptr(mp), ptr(mi), ptr(fl[0]), ptr(fl[1]), ptr(fl[2]),
                                          ptr(fl[3]), ptr(fl[4]), None, None, None, _lib.stream_ptr(dev)))
        t = time_fn(cg)
        b = 5 * rows * k * 4 + rows * 17
        print(json.dumps({"kernel": "cgpl_pgls_kernel", "rows": rows, "k": k, "ms": t * 1e3, "alg_MB": b / 1e6,
                          "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / HBM}), flush=True)
        # masked soft CE fwd+grad
        mr = torch.rand(rows, generator=g).ge(0.5).to(dev)
        for f in fl:
            f.fill_(True)
        losses = torch.empty(3, device=dev)
        gr = [torch.empty(rows, k, device=dev) for _ in range(3)]
        ws = torch.empty(lib.stil_masked_softce_workspace_bytes(rows), dtype=torch.uint8, device=dev)
        pl.copy_(torch.softmax(ys[0], 1))

        def ce():
            _lib.check(lib.stil_masked_softce(ptr(ys[0]), ptr(ys[1]), ptr(ys[2]), 0, k, ptr(pl), k, ptr(fl[0]), ptr(fl[1]),
                                              ptr(fl[2]), ptr(fl[3]), ptr(fl[4]), ptr(mr), rows, k, ptr(losses), ptr(gr[0]),
                                              ptr(gr[1]), ptr(gr[2]), k, 1.0, ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        t = time_fn(ce)
        b = 7 * rows * k * 4 + rows * 6
        print(json.dumps({"kernel": "masked_softce_kernel", "rows": rows, "k": k, "ms": t * 1e3, "alg_MB": b / 1e6,
                          "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / HBM}), flush=True)
        del ys, tl, pl, gr


def gemm_sweep():
    g = torch.Generator().manual_seed(1)
    for n, d, dtype in ((4096, 128, torch.bfloat16), (4096, 512, torch.bfloat16), (8192, 512, torch.bfloat16),
                        (4096, 2048, torch.bfloat16), (4096, 512, torch.float32)):
        a = torch.randn(n, d, generator=g).to(dtype).to(dev).requires_grad_(True)
        b = torch.randn(n, d, generator=g).to(dtype).to(dev).requires_grad_(True)
        crit = S.CLIPLoss(0.1, 0.5, return_logits=False)

        def fwd_bwd():
            a.grad = b.grad = None
            loss, _, _ = crit(a, b)
            loss.backward()
        t = time_fn(fwd_bwd, iters=5, warm=2)
        fl = 6.0 * n * n * d
        print(json.dumps({"op": "CLIPLoss fwd+bwd", "n": n, "d": d, "dtype": str(dtype), "ms": t * 1e3,
                          "alg_GFLOP": fl / 1e9, "TFLOPs": fl / t / 1e12, "frac_tensor": fl / t / 1e12 / TF}), flush=True)
    # prototype / bank-sized logits: rows x 65536 x 512 (the GEMM of BASELINE config C5)
    for rows, k, d in ((448, 65536, 512), (512, 286, 128)):
        feat = torch.nn.functional.normalize(torch.randn(rows, d, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
        protos = (torch.randn(k, d, generator=g) * 0.1).to(dev)
        label = torch.softmax(torch.randn(rows, k, generator=g) * 8, 1).to(dev)
        crit = S.PrototypeLoss(0.1, 0.0)
        cls, conf, _ = S.label_argmax(label, 0.0)

        def fb():
            feat.grad = None
            crit.forward_hard(cls, conf, protos, feat).backward()
        t = time_fn(fb, iters=5, warm=2)
        fl = 4.0 * rows * k * d
        print(json.dumps({"op": "PrototypeLoss fwd+bwd", "rows": rows, "k": k, "d": d, "ms": t * 1e3,
                          "alg_GFLOP": fl / 1e9, "TFLOPs": fl / t / 1e12, "frac_tensor": fl / t / 1e12 / TF}), flush=True)


def bank_sweep():
    """BASELINE config C5 on one GPU: SimMatch bank block, bank 65536 x 512 bf16, 448 unlabelled rows, 286 classes."""
    g = torch.Generator().manual_seed(2)
    for rows, kb, d, c in ((448, 65536, 512, 286), (448, 2560, 128, 286)):
        unit = torch.nn.functional.normalize
        bank = S.alloc_bank(d, kb, torch.bfloat16)
        bank.copy_(unit(torch.randn(kb, d, generator=g)).t())
        labels = torch.randint(0, c, (kb,), generator=g).to(dev)
        fk = unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16).to(dev)
        fq = unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
        p = torch.softmax(torch.randn(rows, c, generator=g) * 3, 1).to(dev)

        def fb():
            fq.grad = None
            prob_ku, loss_in = S.simmatch_bank(fk, fq, p, bank, labels, 0.1, 0.1, 0.9)
            loss_in.mean().backward()
        t = time_fn(fb, iters=5, warm=2)
        fl = 6.0 * rows * kb * d
        print(json.dumps({"op": "simmatch_bank fwd+bwd (a7)", "rows": rows, "k_bank": kb, "d": d, "ms": t * 1e3,
                          "samples_per_s": rows / t, "alg_GFLOP": fl / 1e9, "TFLOPs": fl / t / 1e12,
                          "frac_tensor": fl / t / 1e12 / TF}), flush=True)


def kernel_sweep():
    """Tensor-pipe rate of the similarity GEMM kernel ITSELF (executed flops of one launch / its device duration from the
    torch profiler's CUPTI kernel records) on a shape where the tensor pipe is the bound: CLIPLoss fwd+bwd, n x n x d."""
    from torch.profiler import ProfilerActivity, profile
    g = torch.Generator().manual_seed(3)
    for n, d in ((4096, 2048), (8192, 1024), (4096, 512)):
        a = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
        b = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
        crit = S.CLIPLoss(0.1, 0.5, return_logits=False)
        for _ in range(3):
            a.grad = b.grad = None
            crit(a, b)[0].backward()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                a.grad = b.grad = None
                crit(a, b)[0].backward()
            torch.cuda.synchronize()
        per = {}
        for e in prof.events():
            if "gemm_tc05_kernel" in e.name:
                mode = e.name.split("gemm_tc05_kernel<")[1].split(",")[0]
                per.setdefault(mode, []).append(e.time_range.end - e.time_range.start)
        # every launch runs BOTH InfoNCE sides: 2 x (2 n^2 d) executed flops (statistics, recompute for G, dX)
        fl = 2 * 2.0 * n * n * d
        for mode, name in (("0", "STATS (logits + softmax statistics)"), ("2", "GRAD (recompute + dLogits)"), ("1", "STORE (dX = G.Y)")):
            if mode in per:
                us = sorted(per[mode])[len(per[mode]) // 2]
                print(json.dumps({"kernel": f"gemm_tc05_kernel<{name}>", "n": n, "d": d, "us_per_launch": us,
                                  "executed_GFLOP": fl / 1e9, "TFLOPs": fl / us / 1e6, "frac_tensor": fl / us / 1e6 / TF}),
                      flush=True)
        del a, b


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "hbm_peak_GBps": HBM, "bf16_peak_TFLOPs": TF}))
    if what in ("rows", "all"):
        rows_sweep()
    if what in ("gemm", "all"):
        gemm_sweep()
    if what in ("bank", "all"):
        bank_sweep()
    if what in ("kernel", "all"):
        kernel_sweep()
