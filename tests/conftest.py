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


# ---------------------------------------------------------------------------------------------- measured parity errors
# Every tolerance comparison of a -m gpu test records what it measured; the table is printed at the end of the run
# (also with -q) and written to gpurun_out/parity_errors.jsonl when that directory exists.
PARITY_LOG = []
REL = 1e-3          # north_star: losses and gradients within 1e-3 relative of the reference (fp32 accumulate)


def rel_err(a, b):
    """max |a - b| / max |b| (gradient tensors span many orders of magnitude)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_rel(a, b, tol=REL, what=""):
    """|a - b| <= tol * max|b| element-wise.  A bf16-typed `a` (autograd hands a bf16 input a bf16 .grad) is allowed
    its own output rounding on top: half a bf16 ulp of the reference entry, 2^-9 |b| — the value is formed in fp32
    and rounded ONCE, which is what the reference's autograd does with bf16 leaves."""
    import os, inspect
    af, bf = a.detach().float().cpu(), b.detach().float().cpu()
    e = rel_err(af, bf)
    scale = float(bf.abs().max().clamp_min(1e-30))
    slack = bf.abs() * 2.0 ** -8 if a.dtype == torch.bfloat16 else torch.zeros_like(bf)
    # error beyond the output format's own rounding, relative to the largest reference entry
    excess = float(((af - bf).abs() - slack).clamp_min(0).max() / scale)
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split("::")[-1].split(" ")[0]
    PARITY_LOG.append({"test": test, "tensor": what, "rel_err": e, "excess_over_bf16_rounding": excess, "tol": tol,
                       "out_dtype": str(a.dtype).replace("torch.", "")})
    assert excess <= tol, f"{what}: rel err {e:.3e} (beyond output rounding: {excess:.3e}) > {tol}"


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    if not PARITY_LOG:
        return
    import json
    tr = terminalreporter
    tr.write_line("")
    tr.write_line("measured parity errors (max |ours - ref| / max |ref|; tolerance 1e-3 unless stated):")
    worst = {}
    for r in PARITY_LOG:
        k = (r["test"].split("[")[0], r["tensor"], r["out_dtype"])
        if k not in worst or r["excess_over_bf16_rounding"] > worst[k]["excess_over_bf16_rounding"]:
            worst[k] = r
    for (t, w, dt), r in sorted(worst.items()):
        tr.write_line(f"  {t:44s} {w:14s} {dt:9s} rel_err={r['rel_err']:.2e} beyond_rounding={r['excess_over_bf16_rounding']:.2e} tol={r['tol']:g}")
    out = REPO / "gpurun_out"
    if out.is_dir():
        with open(out / "parity_errors.jsonl", "w") as f:
            for r in PARITY_LOG:
                f.write(json.dumps(r) + "\n")


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


# hparams.DA == True fixtures (STiLModel.py:276-277): in_DA_queue / in_DA_ptr are the ring buffer before the step
DA_CASES = ["step_dvm_b64_da", "step_cardiac_b128_da"]


def da_state_of(ins):
    """The DA ring buffer of a DA fixture as the head_step oracle / the kernels take it (fresh copies)."""
    return {"DA_queue": ins["DA_queue"].clone(), "DA_ptr": ins["DA_ptr"].clone()}


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
        "step_dvm_b64_da": synth.dvm_config(64, embed_dtype="f32"),
        "step_cardiac_b128_da": synth.cardiac_config(128),
    }
    return table[name]
