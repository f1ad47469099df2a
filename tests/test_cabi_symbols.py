"""CPU: the C-ABI library builds, loads and exports every symbol include/stil_head.h declares; argument
validation that needs no GPU returns the documented status codes (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="session")
def lib():
    from stil_tta_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from stil_tta_b200 import _lib
    header = (REPO / "include" / "stil_head.h").read_text()
    declared = set(re.findall(r"STIL_API\s+[\w\s\*]+?\b(stil_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.stil_version() == 100


def test_workspace_queries_are_pure_host(lib):
    a = lib.stil_infonce_workspace_bytes(512, 512, 128, 1)
    b = lib.stil_infonce_workspace_bytes(512, 4096, 128, 1)
    assert 0 < a < b
    assert lib.stil_infonce_workspace_bytes(512, 512, 128, 0) > a       # fp32 needs split operands
    assert lib.stil_head_step_workspace_bytes(512, 64, 286, 128, 1) > a
    assert lib.stil_masked_softce_workspace_bytes(448) > 0


def test_argument_validation_without_gpu(lib):
    from stil_tta_b200 import _lib
    # lambda_0 outside [0,1] -> STIL_E_ARG -> ValueError, like utils/clip_loss.py:22-23
    buf = ctypes.create_string_buffer(4096)
    p = (ctypes.addressof(buf) + 15) // 16 * 16
    rc = lib.stil_infonce_fwd(p, p, p, p, 1, 8, 8, 8, 8, 0, 0.1, 1.5, p, p, p, None, 0, p, 4096, None)
    assert rc == -6
    with pytest.raises(ValueError, match="lambda_0 must be a float between 0 and 1"):
        _lib.check(rc)
    # misaligned row length
    rc = lib.stil_infonce_fwd(p, p, p, p, 1, 8, 8, 12, 12, 0, 0.1, 0.5, p, p, p, None, 0, p, 4096, None)
    assert rc == -3
    # unsupported dtype
    rc = lib.stil_infonce_fwd(p, p, p, p, 7, 8, 8, 8, 8, 0, 0.1, 0.5, p, p, p, None, 0, p, 4096, None)
    assert rc == -2
    # too many classes for the row kernel
    rc = lib.stil_cgpl_pgls(p, p, p, 0, 4096, p, 4096, 4, 4096, 0.1, 0.9, 0.9, 1, p, 4096, None, 0, p, p, p, p, p, p,
                            p, None, None, None, None)
    assert rc == -1
    # workspace too small
    rc = lib.stil_infonce_fwd(p, p, p, p, 1, 8, 8, 8, 8, 0, 0.1, 0.5, p, p, p, None, 0, p, 16, None)
    assert rc == -7


def test_ops_refuse_cpu_tensors():
    import torch
    import stil_tta_b200 as S
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.CLIPLoss(0.1)(torch.randn(8, 8), torch.randn(8, 8))
    with pytest.raises(ValueError):
        S.CLIPLoss(0.1, lambda_0=-0.1)
