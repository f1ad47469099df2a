"""ctypes binding of libstil_head.so (the C ABI declared in include/stil_head.h).

There is no CPU or pure-PyTorch fallback: if the shared library is missing the
import of any op fails loudly with the build command.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Dict, Optional

import torch

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "_C" / "libstil_head.so"

STIL_F32, STIL_BF16 = 0, 1
STIL_OK = 0
_ERR_NAMES = {-1: "STIL_E_SHAPE", -2: "STIL_E_DTYPE", -3: "STIL_E_ALIGN", -4: "STIL_E_ARCH", -5: "STIL_E_CUDA",
              -6: "STIL_E_ARG", -7: "STIL_E_WORKSPACE"}

i64, i32, f32, vp = C.c_int64, C.c_int, C.c_float, C.c_void_p


class P2PChannel(C.Structure):
    """Mirror of `stil_p2p_channel` (include/stil_head.h)."""
    _fields_ = [("bases", vp * 8), ("world", i32), ("rank", i32), ("flags_offset", i64), ("ctrl_offset", i64),
                ("channel", i32)]


class EmaEntry(C.Structure):
    """Mirror of `stil_ema_entry` (include/stil_head.h)."""
    _fields_ = [("ema", vp), ("main", vp), ("numel", i64), ("dtype", i32), ("kind", i32)]


class HeadStepArgs(C.Structure):
    """Mirror of `stil_head_step_args` (include/stil_head.h) — field order must match."""
    _fields_ = [
        ("batch", i64), ("b_l", i64), ("k", i64), ("dim", i64),
        ("embed_dtype", i32), ("logit_dtype", i32), ("grad_dtype", i32),
        ("temperature", f32), ("lambda0", f32), ("th1", f32), ("rate_pseudo", f32), ("repeat_ratio", f32),
        ("past_start_epoch", i32),
        ("feat_i", vp), ("feat_t", vp), ("feat_m", vp), ("feat_m_e", vp),
        ("y_m_ue", vp), ("y_i_ue", vp), ("y_t_ue", vp),
        ("y_m", vp), ("y_i", vp), ("y_t", vp),
        ("y_l", vp), ("prototypes", vp), ("mask_random", vp),
        ("losses", vp),
        ("d_feat_i", vp), ("d_feat_t", vp), ("d_feat_m", vp),
        ("d_y_m", vp), ("d_y_i", vp), ("d_y_t", vp),
        ("pseudo_label", vp),
        ("max_prob", vp), ("max_idx", vp), ("mask1", vp), ("case1", vp), ("case2_i", vp), ("case2_t", vp),
        ("case3", vp),
        ("class_sum", vp), ("class_count", vp),
        ("prototypes_sum", vp), ("prototypes_count_sum", vp),
        ("rate_uce_scale", f32),
        ("workspace", vp), ("workspace_bytes", i64), ("stream", vp),
        ("timing_events", vp), ("n_timing_events", i32), ("skip_infonce", i32), ("prototypes_prepared", i32),
        ("partials_push", C.POINTER(P2PChannel)), ("partials_dst_offset", i64),
        ("prediction_in", vp),
    ]


# name -> (restype, argtypes); every symbol of include/stil_head.h
SIGNATURES = {
    "stil_version": (i32, []),
    "stil_abi_struct_bytes": (i64, [i32]),
    "stil_last_error": (C.c_char_p, []),
    "stil_check_device": (i32, []),
    "stil_debug_trace": (i32, [vp]),
    "stil_debug_pdl": (i32, [i32]),
    "stil_infonce_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "stil_infonce_fwd": (i32, [vp, vp, vp, vp, i32, i64, i64, i64, i64, i64, f32, f32, vp, vp, vp, vp, i64, vp, i64, vp]),
    "stil_infonce_bwd": (i32, [vp, vp, vp, vp, i32, i64, i64, i64, i64, i64, f32, f32, vp, vp, vp, vp, vp, i32, i64,
                               vp, i64, vp]),
    "stil_infonce_bwd_after_fwd": (i32, [vp, vp, i32, i64, i64, i64, i64, i64, f32, f32, vp, vp, vp, vp, vp, i32, i64,
                                         vp, i64, vp]),
    "stil_proto_logits_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "stil_proto_logits": (i32, [vp, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, vp]),
    "stil_cgpl_pgls": (i32, [vp, vp, vp, i32, i64, vp, i64, i64, i64, f32, f32, f32, i32, vp, i64, vp, i64, vp, i64, vp,
                             vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "stil_label_argmax": (i32, [vp, i64, i64, i64, f32, vp, vp, vp, vp]),
    "stil_proto_ce_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "stil_proto_ce_fwd": (i32, [vp, i32, i64, i64, i64, vp, i64, vp, vp, f32, vp, vp, vp, vp, i64, vp]),
    "stil_proto_ce_bwd": (i32, [vp, i32, i64, i64, i64, vp, i64, vp, vp, vp, f32, vp, vp, i32, i64, vp, i64, vp]),
    "stil_proto_accumulate": (i32, [vp, i32, i64, i64, i64, vp, vp, i64, f32, i64, vp, vp, vp, vp, vp]),
    "stil_proto_add": (i32, [vp, vp, i64, i64, vp, vp, vp]),
    "stil_proto_finalize": (i32, [vp, vp, vp, i64, i64, vp, vp]),
    "stil_proto_add_gathered": (i32, [vp, i64, i64, i64, i64, vp, vp, vp, vp, vp]),
    "stil_p2p_alloc": (i32, [i64, C.POINTER(vp)]),
    "stil_p2p_free": (i32, [vp]),
    "stil_p2p_export": (i32, [vp, C.c_char_p]),
    "stil_p2p_import": (i32, [C.c_char_p, C.POINTER(vp)]),
    "stil_p2p_close": (i32, [vp]),
    "stil_p2p_exchange": (i32, [C.POINTER(vp), i32, i32, i64, i64, i32, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64),
                                vp]),
    "stil_p2p_push": (i32, [C.POINTER(vp), i32, i32, i64, i64, i32, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), vp]),
    "stil_p2p_wait": (i32, [C.POINTER(vp), i32, i32, i64, i64, i32, vp]),
    "stil_p2p_push_embeddings": (i32, [C.POINTER(vp), i32, i32, i64, i64, i32, vp, vp, i32, i64, i64, i64, i64, i64, i64, vp]),
    "stil_p2p_push_lse": (i32, [C.POINTER(vp), i32, i32, i64, i64, i32, vp, i64, i64, i64, i32, i64, i64, i64, vp]),
    "stil_infonce_stats_gathered": (i32, [vp, vp, vp, vp, i32, i64, i64, i64, i64, i64, f32, vp, vp, i64, vp, i64, vp]),
    "stil_infonce_loss_gathered": (i32, [vp, vp, vp, vp, i32, i64, i64, i64, i64, i64, f32, f32, vp, vp, vp, C.POINTER(vp),
                                         i32, i32, i64, vp, vp, i64, vp]),
    "stil_proto_add_gathered_wait": (i32, [vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "stil_infonce_bwd_gathered": (i32, [vp, vp, vp, vp, i32, i64, i64, i64, i64, i64, f32, f32, vp, vp, vp, vp, vp,
                                        vp, i32, i64, vp, i64, vp]),
    "stil_softmax_rows": (i32, [vp, i32, i64, i64, i64, vp, i64, vp]),
    "stil_da_batch_mean": (i32, [vp, i64, i64, i64, vp, vp]),
    "stil_threshold_workspace_bytes": (i64, [i64, i64]),
    "stil_freematch_stats": (i32, [vp, i64, i64, i64, i32, vp, vp, vp, vp, i64, vp, i64, vp]),
    "stil_freematch_update_mask": (i32, [vp, i64, i64, f32, f32, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "stil_threshold_rows": (i32, [vp, i64, i64, i64, f32, vp, i64, vp, vp, vp, vp]),
    "stil_freematch_entropy_fwd": (i32, [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, i64, vp]),
    "stil_freematch_entropy_bwd": (i32, [i64, i64, vp, vp, i64, vp, i64, vp]),
    "stil_da_apply": (i32, [vp, i64, i64, i64, vp, vp, i64, vp, vp, vp, i64, vp]),
    "stil_simmatch_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "stil_simmatch_fwd": (i32, [vp, vp, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, f32, f32, f32, vp, vp, i32, vp,
                                i64, vp]),
    "stil_simmatch_bwd": (i32, [vp, i32, i64, i64, vp, i64, i64, vp, vp, i32, i64, vp, i64, vp]),
    "stil_simmatch_shard_stats": (i32, [vp, vp, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, f32, f32, vp, vp, i64, vp]),
    "stil_simmatch_shard_finish": (i32, [vp, vp, i64, i64, f32, f32, vp, vp, vp, vp]),
    "stil_simmatch_shard_grad": (i32, [i32, i64, i64, vp, i64, vp, i64, vp, i64, f32, f32, vp, vp, i64, vp, i64, vp]),
    "stil_bank_smooth_workspace_bytes": (i64, [i64, i64, i64, i64, i32]),
    "stil_bank_smooth": (i32, [vp, i64, i64, i64, vp, i32, i64, i64, vp, i64, vp, i64, i64, f32, f32, f32, vp, i64, f32,
                               vp, vp, vp, vp, i64, vp]),
    "stil_comatch_graphs_workspace_bytes": (i64, [i64, i64, i64, i64, i32]),
    "stil_comatch_graphs_fwd": (i32, [vp, i64, i64, i64, vp, i64, vp, vp, i32, i64, i64, vp, i64, i64, f32, vp, vp, i64,
                                      vp, i64, vp]),
    "stil_comatch_sim_bwd_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "stil_comatch_sim_bwd": (i32, [vp, vp, i64, i64, i64, vp, i32, i64, i64, vp, i64, f32, vp, i32, i64, vp, i64, vp]),
    "stil_row_loss_workspace_bytes": (i64, [i64]),
    "stil_graph_contrast_loss": (i32, [vp, vp, i64, i64, i64, f32, vp, vp, f32, vp, i64, vp]),
    "stil_weighted_softce": (i32, [vp, i32, i64, vp, i64, vp, vp, i64, i64, vp, vp, i64, f32, vp, i64, vp]),
    "stil_queue_enqueue": (i32, [vp, i32, i64, vp, i64, i64, vp, vp, i32, i64, i64, i64, vp, i64, i64, vp]),
    "stil_bank_update": (i32, [vp, i32, i64, vp, vp, i32, i64, vp, vp, i64, i64, vp]),
    "stil_da_apply_hist": (i32, [vp, i64, i64, i64, vp, vp, i64, vp, vp, vp, i64, vp]),
    "stil_linear_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "stil_linear_fwd": (i32, [vp, i32, i64, i64, i64, vp, vp, i64, i32, vp, i64, vp, vp, vp, i64, vp]),
    "stil_linear_bwd": (i32, [vp, i32, i64, i64, i64, vp, i64, vp, vp, vp, i64, vp, i32, i64, vp, vp, vp, i64, vp]),
    "stil_ema_update": (i32, [vp, i64, vp, vp, i64, i64, C.c_double, vp]),
    "stil_club_fwd": (i32, [vp, vp, i32, i64, i64, i64, vp, vp, vp, vp]),
    "stil_club_bwd": (i32, [vp, vp, i32, i64, i64, i64, vp, vp, vp, vp, vp, i64, vp]),
    "stil_masked_softce_workspace_bytes": (i64, [i64]),
    "stil_masked_softce": (i32, [vp, vp, vp, i32, i64, vp, i64, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, i64,
                                 f32, vp, i64, vp]),
    "stil_head_step_workspace_bytes": (i64, [i64, i64, i64, i64, i32]),
    "stil_head_step": (i32, [C.POINTER(HeadStepArgs)]),
    "stil_head_step_launches": (i32, [C.POINTER(HeadStepArgs)]),
    "stil_head_prepare_prototypes": (i32, [C.POINTER(HeadStepArgs)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree extension; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the STiL head has no CPU/PyTorch fallback. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` or `python stil_tta_b200/build.py` (needs nvcc).")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == STIL_OK:
        return
    msg = load().stil_last_error().decode(errors="replace")
    text = f"{_ERR_NAMES.get(rc, rc)}: {msg}"
    if rc in (-1, -2, -3, -6):
        raise ValueError(text)     # reference convention: ValueError for bad arguments (clip_loss.py:22-23)
    raise RuntimeError(text)


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return STIL_F32
    if t.dtype == torch.bfloat16:
        return STIL_BF16
    raise ValueError(f"unsupported dtype {t.dtype}: the STiL head takes float32 or bfloat16")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("stil_tta_b200 ops run on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if dev is not None and t.device != dev:
            raise RuntimeError("all tensors must live on the same CUDA device")
        dev = t.device
    return dev


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


_ws_cache: Dict[tuple, torch.Tensor] = {}


def workspace(dev: torch.device, key: str, nbytes: int) -> torch.Tensor:
    """Scratch owned by the Python side (the extension allocates nothing), cached per device, op AND stream: two
    streams (or threads with their own streams) driving the same op concurrently never share a buffer."""
    k = (dev.index, key, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws_cache.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _ws_cache[k] = buf
    return buf


_checked = set()


def ensure_device(dev: torch.device) -> None:
    if dev.index in _checked:
        return
    with torch.cuda.device(dev):
        check(load().stil_check_device())
    _checked.add(dev.index)
