"""Generate tests/golden/ema_*.npz by EXECUTING the reference's ``STiLModel.momentum_update_ema``
(``models/Disentangle/STiLModel.py:154-168``) as an unbound function on a stand-in ``self`` that carries two small
modules (Linear + BatchNorm1d, so that ``num_batches_tracked`` and running statistics are in the state dict) — both
branches (``eman`` True / False), three consecutive updates.  TEST INFRASTRUCTURE; run once in the authoring container
(``python oracle/gen_golden_ema.py``).  Only tensors are stored, no reference source.
"""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle.gen_golden import OUT, load_reference  # noqa: E402


def make_net(seed: int) -> torch.nn.Module:
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.BatchNorm1d(64), torch.nn.ReLU(), torch.nn.Linear(64, 257),
                              torch.nn.BatchNorm1d(257), torch.nn.Linear(257, 10))
    with torch.no_grad():            # non-trivial BatchNorm statistics and step counters
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_()
                m.running_var.uniform_(0.5, 2.0)
                m.num_batches_tracked.fill_(seed * 7 + 3)
    return net


def main():
    _, _, STiLModel = load_reference()
    for eman in (True, False):
        model, ema = make_net(1), make_net(2)
        rec = {"meta_eman": int(eman), "meta_momentum": 0.999}
        for k, v in model.state_dict().items():
            rec["main_" + k] = v.detach().numpy().copy()
        for k, v in ema.state_dict().items():
            rec["ema0_" + k] = v.detach().numpy().copy()
        me = SimpleNamespace(eman=eman, model=model, ema=ema, momentum=0.999)
        for step in range(1, 4):
            STiLModel.momentum_update_ema(me)                         # the real method body
            if step in (1, 3):
                for k, v in ema.state_dict().items():
                    rec[f"ema{step}_" + k] = v.detach().numpy().copy()
        np.savez_compressed(OUT / f"ema_{'eman' if eman else 'params'}.npz", **rec)
        print("ema", "eman" if eman else "params", "written:", len(rec), "arrays")


if __name__ == "__main__":
    main()
