"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbols, gloo.
`-m gpu` runs on a B200: CUDA path vs oracle through the C-ABI.
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); skipped/deselected on the CPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Load a fixture written by oracle/gen_golden.py -> (inputs dict of tensors, ref dict, meta)."""
    z = np.load(GOLDEN / f"{name}.npz")
    ins, ref, meta = {}, {}, {}
    bf16 = bool(int(z["meta_embed_bf16"])) if "meta_embed_bf16" in z else False
    for k in z.files:
        v = z[k]
        if k.startswith("in_"):
            t = torch.from_numpy(np.array(v))
            if bf16 and k[3:] in ("feat_i", "feat_t", "feat_m", "feat_m_e"):
                t = t.view(torch.bfloat16)
            ins[k[3:]] = t
        elif k.startswith("ref_"):
            ref[k[4:]] = torch.from_numpy(np.array(v))
        elif k.startswith("meta_"):
            meta[k[5:]] = v.item() if v.shape == () else v
    return ins, ref, meta


STEP_CASES = ["step_c1_dvm_b64_f32", "step_c1_dvm_b64_bf16", "step_dvm_b128_edge", "step_dvm_b64_pre_start",
              "step_dvm_b64_zero_protos", "step_cardiac_b128", "step_cardiac_b64_ragged", "step_dvm_b200_k10"]


def cfg_for(name, meta):
    from stil_tta_b200 import synth
    table = {
        "step_c1_dvm_b64_f32": synth.dvm_config(64, embed_dtype="f32"),
        "step_c1_dvm_b64_bf16": synth.dvm_config(64),
        "step_dvm_b128_edge": synth.dvm_config(128, embed_dtype="f32"),
        "step_dvm_b64_pre_start": synth.dvm_config(64, past_start_epoch=False, repeat_ratio=1.0),
        "step_dvm_b64_zero_protos": synth.dvm_config(64),
        "step_cardiac_b128": synth.cardiac_config(128),
        "step_cardiac_b64_ragged": synth.cardiac_config(72, unlabelled_ratio=5),
        "step_dvm_b200_k10": synth.dvm_config(200, num_classes=10, proj_dim=64, unlabelled_ratio=3,
                                              embed_dtype="f32", th1=0.6),
    }
    return table[name]
