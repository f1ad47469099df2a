"""GPU: the experimental fused backward kernel (gemm_bwd_kernel, opt-in with STIL_FUSED_BWD=1) is held to the same parity
tests as the default GRAD + STORE pair.  The switch is read once per process, so the parity tests are re-run in a child."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]


def test_parity_suite_with_fused_backward():
    env = dict(os.environ, STIL_FUSED_BWD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu", "-k",
                        "clip_loss or prototype_loss or head_step"], cwd=REPO, env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
