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
    rc = lib.stil_cgpl_pgls(p, p, p, 0, 4096, p, 4096, 4, 4096, 0.1, 0.9, 0.9, 1, None, 0, p, 4096, None, 0, p, p, p, p, p, p,
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


def test_struct_mirrors_match_the_header(lib):
    """The ctypes mirrors of the two ABI structs have the C compiler's sizes (field order / padding drift would corrupt
    every argument after the first mismatch)."""
    import ctypes
    from stil_tta_b200 import _lib
    assert lib.stil_abi_struct_bytes(0) == ctypes.sizeof(_lib.HeadStepArgs)
    assert lib.stil_abi_struct_bytes(1) == ctypes.sizeof(_lib.P2PChannel)
    assert lib.stil_abi_struct_bytes(7) == -1


def test_new_entry_points_validate_arguments_without_gpu(lib):
    buf = ctypes.create_string_buffer(8192)
    p = (ctypes.addressof(buf) + 15) // 16 * 16
    # bank smoothing: misaligned queue leading dimension (fp32 needs a multiple of 4)
    rc = lib.stil_bank_smooth(p, 16, 4, 16, p, 0, 8, 8, p, 6, p, 8, 6, 0.1, 0.9, 0.1, p, 16, 0.9, None, None, None, p, 8192,
                              None)
    assert rc == -3
    # bank smoothing without a bank (epoch gate) needs no workspace and no features
    assert lib.stil_bank_smooth_workspace_bytes(448, 640, 128, 286, 0) > 0
    # single-head CE: exactly one kind of target
    rc = lib.stil_weighted_softce(p, 0, 8, p, 8, p, None, 4, 8, p, None, 8, 1.0, p, 8192, None)
    assert rc == -6
    rc = lib.stil_weighted_softce(p, 5, 8, p, 8, None, None, 4, 8, p, None, 8, 1.0, p, 8192, None)
    assert rc == -2
    # CoMatch graphs: output leading dimension must hold rows + k_q columns
    rc = lib.stil_comatch_graphs_fwd(p, 16, 4, 16, p, 8, p, p, 0, 8, 8, p, 8, 8, 0.1, p, p, 8, p, 8192, None)
    assert rc == -6
    # CLUB: bad dtype / missing outputs
    assert lib.stil_club_fwd(p, p, 9, 4, 8, 8, p, p, p, None) == -2
    assert lib.stil_club_fwd(p, p, 0, 4, 8, 8, p, None, None, None) == -6
    # gathered InfoNCE is bf16-only
    rc = lib.stil_infonce_stats_gathered(p, p, p, p, 0, 8, 8, 8, 8, 0, 0.1, None, None, 0, p, 8192, None)
    assert rc == -2
    # peer exchange: world / rank / channel range
    bases = (ctypes.c_void_p * 8)(*([p] * 8))
    assert lib.stil_p2p_wait(bases, 9, 0, 0, 512, 0, None) == -6
    assert lib.stil_p2p_wait(bases, 2, 0, 0, 512, 8, None) == -6


def test_bank_and_club_dropins_refuse_cpu_tensors():
    import torch
    import stil_tta_b200 as S
    p = torch.softmax(torch.randn(4, 10), 1)
    f, q, qp = torch.randn(4, 8), torch.randn(8, 16), torch.rand(10, 16)
    for call in (lambda: S.bank_smooth(p, f, q, qp, 0.1, 0.9, 0.1, 0.9),
                 lambda: S.mmatch_pseudo_label(p, f, q, qp, 0.1, 0.9),
                 lambda: S.comatch_graphs(p, qp, f, f, q, 0.1),
                 lambda: S.graph_contrast_loss(torch.rand(4, 20), torch.rand(4, 20), 0.8),
                 lambda: S.masked_ce(torch.randn(4, 10), p, torch.ones(4, dtype=torch.bool)),
                 lambda: S.queue_enqueue(q, qp, torch.zeros(1, dtype=torch.int64), f, p),
                 lambda: S.club_bound(torch.randn(4, 8), torch.randn(4, 8)),
                 lambda: S.CLUBMean(8, 8, 16)(torch.randn(4, 8), torch.randn(4, 8))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_header_is_plain_c():
    """include/stil_head.h is the drop-in boundary: it must compile as C99 (no C++ in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", str(REPO / "include" / "stil_head.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
