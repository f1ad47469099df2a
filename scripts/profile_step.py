"""Run a few un-captured C2 head steps (for ncu: `ncu ... python scripts/profile_step.py`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import synth  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = synth.CONFIGS[cfg_name]()
head = S.STiLHead(cfg, device="cuda", use_graph=False)
head.load(synth.make_batch(cfg, seed=2022))
for _ in range(steps):
    head.run()
torch.cuda.synchronize()
print("losses", head.out["losses"].tolist())
