"""Build libstil_head.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_C"
LIB = OUT_DIR / "libstil_head.so"
SOURCES = ["api.cu", "gemm_tc05.cu", "row_kernels.cu", "bank_kernels.cu", "bank_sweep.cu", "threshold_kernels.cu", "p2p.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libstil_head.so cannot be built (there is no prebuilt or CPU fallback)")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                    [PKG.parent / "include" / "stil_head.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source into stil_tta_b200/_C/libstil_head.so (skipped when up to date)."""
    OUT_DIR.mkdir(exist_ok=True)
    stamp = OUT_DIR / "build.sha256"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    nvcc = _nvcc()
    objs = []
    for src in SOURCES:
        obj = OUT_DIR / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.run(cmd, check=True)
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
