"""N-rank correctness check of DistributedSTiLHead against the single-process oracle (the 2-rank pytest worker, any N).
python scripts/dist_check.py N [transport] [batch]"""
import sys
from pathlib import Path

import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from test_gpu_dist import _free_port, _worker  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1])
    transport = sys.argv[2] if len(sys.argv) > 2 else "fused"
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    for use_graph in (False, True):
        mp.spawn(_worker, args=(n, _free_port(), batch, use_graph, transport), nprocs=n, join=True)
        print(f"world {n} transport {transport} batch {batch} graph={use_graph}: OK", flush=True)
