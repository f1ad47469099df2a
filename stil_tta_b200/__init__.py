"""stil_tta_b200 — B200-native (sm_100a) implementation of the STiL per-batch semi-supervised head.

Drop-ins for the reference's hot path (kgutjahr/STiL-TTA): ``CLIPLoss`` (utils/clip_loss.py),
``PrototypeLoss`` (utils/prototype_loss.py), ``cal_prototypes`` / ``cal_prototypes_separate`` and the
inline CGPL/PGLS block of ``models/Disentangle/STiLModel.py`` (``cgpl_pgls``), over a C-ABI CUDA
extension (include/stil_head.h).  Importing this package does not need a GPU; calling an op does,
and fails loudly when ``_C/libstil_head.so`` is missing (no CPU fallback).
"""
from .synth import HeadConfig, cardiac_config, dvm_config, make_batch  # noqa: F401

_LAZY = {
    "CLIPLoss": "losses", "PrototypeLoss": "losses", "masked_soft_ce": "losses", "label_argmax": "losses",
    "cgpl_pgls": "pseudo_label", "distribution_alignment": "pseudo_label", "prototype_logits": "pseudo_label", "PseudoLabels": "pseudo_label",
    "cal_prototypes": "prototypes", "cal_prototypes_separate": "prototypes", "PrototypeBank": "prototypes",
    "simmatch_bank": "bank", "alloc_bank": "bank", "ShardedSimMatchBank": "bank",
    "bank_smooth": "bank_blocks", "mmatch_pseudo_label": "bank_blocks", "comatch_smooth": "bank_blocks",
    "comatch_graphs": "bank_blocks", "graph_contrast_loss": "bank_blocks", "masked_ce": "bank_blocks",
    "queue_enqueue": "bank_blocks", "update_bank": "bank_blocks", "HistAlignment": "bank_blocks",
    "SmoothedLabels": "bank_blocks",
    "Linear": "linear", "linear": "linear",
    "EmaTeacher": "ema", "momentum_update_ema": "ema",
    "FreeMatchThreshold": "thresholds", "entropy_loss": "thresholds", "threshold_rows": "thresholds",
    "cotraining_pseudo_labels": "thresholds",
    "CLUBMean": "club", "club_bound": "club", "club_learning_loss": "club", "club_both": "club", "STiLHead": "head", "DistributedSTiLHead": "head", "GlobalBatch": "distributed", "P2PBuffer": "distributed", "all_reduce_prototype_partials": "distributed",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f".{_LAZY[name]}", __name__), name)
    raise AttributeError(name)
