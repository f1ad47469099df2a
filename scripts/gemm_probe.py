"""GPU diagnostic: exercise the tcgen05 GEMM through stil_proto_logits on small shapes and print error maps.
Not a test — a first-bring-up aid (python scripts/gemm_probe.py > gpurun_out/probe.log)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402


def probe(rows, k, d, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(rows, d, generator=g).to(dtype)
    protos = torch.randn(k, d, generator=g)
    ref = feat.double() @ protos.double().t()
    out = S.prototype_logits(feat.cuda(), protos.cuda())
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs()
    print(f"rows={rows} k={k} d={d} {dtype}: max_err={err.max():.3e} ref_max={ref.abs().max():.3f}", flush=True)
    if err.max() > 1e-3:
        # coarse map: max error per 32x32 block
        R, C = (rows + 31) // 32, (k + 31) // 32
        for r in range(R):
            line = " ".join(f"{err[r*32:(r+1)*32, c*32:(c+1)*32].max():8.2e}" for c in range(C))
            print("   ", line)
        print("   out[0,:8]", out[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
    return float(err.max())


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    worst = 0.0
    for args in [(128, 128, 64, torch.bfloat16), (128, 128, 128, torch.bfloat16), (256, 256, 128, torch.bfloat16),
                 (100, 30, 64, torch.bfloat16), (448, 286, 128, torch.bfloat16), (448, 286, 128, torch.float32),
                 (64, 286, 128, torch.float32), (300, 700, 512, torch.bfloat16), (33, 2, 128, torch.float32)]:
        worst = max(worst, probe(*args))
    print("WORST", worst)
