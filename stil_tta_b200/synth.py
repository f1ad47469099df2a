"""Seeded synthetic STiL head batches with planted structure.

Random Gaussian logits never pass ``th1`` at K=286, so every generator here
plants a true class per row: logits are ``sigma*N(0,1) + mu*onehot(c)`` with the
modality agreement pattern drawn so the four CGPL cases of
``models/Disentangle/STiLModel.py:264-267`` all occur, embeddings are unit
vectors around a per-class direction, and prototypes are (un-normalised) class
means exactly like ``STiLModel.py:408-415`` produces them.

Everything is generated on the CPU with an explicit ``torch.Generator`` so the
same batch can be replayed in this container (golden fixtures), in the oracle
and on the GPU box.  Shapes follow ``trainers/evaluate.py:83-85``:
``B_l = B // (1 + unlabelled_ratio)``, ``B_u = B - B_l``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

SEED = 2022  # first entry of `seeds` in configs/config_dvm_STiL.yaml:11-16


@dataclass
class HeadConfig:
    """Head hyper-parameters (configs/config_dvm_STiL.yaml:52-59,74,152-175)."""
    name: str = "dvm"
    batch: int = 512
    num_classes: int = 286
    proj_dim: int = 128
    unlabelled_ratio: int = 7
    temperature: float = 0.1
    lambda_0: float = 0.5
    th1: float = 0.90
    rate_pseudo: float = 0.9
    repeat_ratio: float = 13.0
    embed_dtype: str = "bf16"      # C2: "bf16 embeddings"; "f32" = reference dtype
    past_start_epoch: bool = True  # STiLModel.py:317-320 gate

    @property
    def b_l(self) -> int:
        return self.batch // (1 + self.unlabelled_ratio)

    @property
    def b_u(self) -> int:
        return self.batch - self.b_l


def dvm_config(batch: int = 512, **kw) -> HeadConfig:
    return HeadConfig(name="dvm", batch=batch, **kw)


def cardiac_config(batch: int = 1024, **kw) -> HeadConfig:
    # configs/config_cardiac_STiL.yaml:57-64,79,157-168
    base = dict(name="cardiac", num_classes=2, th1=0.85, rate_pseudo=0.95, repeat_ratio=1.0)
    base.update(kw)
    return HeadConfig(batch=batch, **base)


CONFIGS = {
    "C1": lambda: dvm_config(64),
    "C2": lambda: dvm_config(512),
    "C3": lambda: cardiac_config(1024),
}


def _unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def _round_embed(x: torch.Tensor, embed_dtype: str) -> torch.Tensor:
    """Embeddings are *stored* in `embed_dtype`; the oracle upcasts the same values."""
    if embed_dtype == "bf16":
        return x.to(torch.bfloat16)
    return x.to(torch.float32)


def make_batch(cfg: HeadConfig, seed: int = SEED, rank: int = 0,
               zero_prototypes: bool = False, edge_rows: bool = False) -> Dict[str, torch.Tensor]:
    """One synthetic head batch (all CPU tensors).

    Keys: teacher logits ``y_m_ue,y_i_ue,y_t_ue`` [B_u,K] f32; student logits
    ``y_m,y_i,y_t`` [B,K] f32; ``feat_i,feat_t,feat_m,feat_m_e`` [B,P] in
    ``cfg.embed_dtype``; ``y_l`` [B_l] int64; ``prototypes`` [K,P] f32;
    ``mask_random`` [B_u] bool (STiLModel.py:299 draws it with torch RNG —
    kept outside the kernels).
    """
    g = torch.Generator().manual_seed(seed + 7919 * rank)
    B, K, P, B_l, B_u = cfg.batch, cfg.num_classes, cfg.proj_dim, cfg.b_l, cfg.b_u

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    def randint(hi, *shape):
        return torch.randint(0, hi, shape, generator=g)

    # class directions and prototypes = means of noisy unit vectors (norm < 1)
    proto_true = _unit(randn(K, P))
    per_class = 16
    cloud = _unit(proto_true[:, None, :] + 0.5 * randn(K, per_class, P) / P ** 0.5 * 4.0)
    prototypes = cloud.mean(dim=1) if not zero_prototypes else torch.zeros(K, P)

    y_true = randint(K, B)
    y_l = y_true[:B_l].clone()

    # agreement pattern over unlabelled rows: 50/15/15/20 %
    u = torch.rand(B, generator=g)
    flip_i = (u >= 0.50) & (u < 0.65) | (u >= 0.80)          # imaging disagrees
    flip_t = (u >= 0.65) & (u < 0.80) | (u >= 0.80)          # tabular disagrees
    c_m = y_true
    c_i = torch.where(flip_i, (y_true + 1 + randint(max(K - 1, 1), B)) % K, y_true)
    c_t = torch.where(flip_t, (y_true + 1 + randint(max(K - 1, 1), B)) % K, y_true)
    if K > 2:
        # for case 3 make i and t also differ from each other
        same = (c_i == c_t) & (u >= 0.80)
        c_t = torch.where(same, (c_t + 1) % K, c_t)
        c_t = torch.where((c_t == c_m) & (u >= 0.80), (c_t + 1) % K, c_t)
        c_t = torch.where((c_t == c_i) & (u >= 0.80), (c_t + 1) % K, c_t)

    mu_choices = torch.tensor([4.0, 8.0, 12.0]) if K > 2 else torch.tensor([1.0, 3.0, 6.0])
    mu = mu_choices[randint(3, B)]

    def planted(c):
        y = randn(B, K)
        y[torch.arange(B), c] += mu
        return y

    t_m, t_i, t_t = planted(c_m), planted(c_i), planted(c_t)       # teacher logits
    s_m = t_m + 0.3 * randn(B, K)                                   # student ≈ teacher + noise
    s_i = t_i + 0.3 * randn(B, K)
    s_t = t_t + 0.3 * randn(B, K)

    def embed(noise):
        return _unit(proto_true[y_true] + noise * randn(B, P) / P ** 0.5 * 4.0)

    feat_m_e = embed(0.35)
    feat_m = _unit(feat_m_e + 0.05 * randn(B, P))
    feat_i = embed(0.5)
    feat_t = _unit(feat_i + 0.3 * randn(B, P) / P ** 0.5 * 4.0)

    if edge_rows and B_u >= 8:
        r = B_l  # first unlabelled rows carry engineered edge cases
        # exact tie between classes 0 and 1 in all three heads -> argmax must be 0
        for y in (t_m, t_i, t_t):
            y[r] = 0.0
            y[r, 0] = 9.0
            y[r, 1 % K] = 9.0
        # constant rows (uniform softmax, argmax 0, never confident for K>1)
        for y in (t_m, t_i, t_t):
            y[r + 1] = 1.25
        # huge-magnitude logits (softmax must subtract the row max)
        t_m[r + 2] = 0.0
        t_m[r + 2, K - 1] = 80.0
        t_i[r + 2] = t_m[r + 2]
        t_t[r + 2] = -t_m[r + 2]
        # tie on the *last* two classes
        if K >= 3:
            for y in (t_m, t_i, t_t):
                y[r + 3] = -3.0
                y[r + 3, K - 2] = 7.5
                y[r + 3, K - 1] = 7.5

    out = {
        "y_m_ue": t_m[B_l:].contiguous(), "y_i_ue": t_i[B_l:].contiguous(), "y_t_ue": t_t[B_l:].contiguous(),
        "y_m": s_m, "y_i": s_i, "y_t": s_t,
        "feat_i": _round_embed(feat_i, cfg.embed_dtype),
        "feat_t": _round_embed(feat_t, cfg.embed_dtype),
        "feat_m": _round_embed(feat_m, cfg.embed_dtype),
        "feat_m_e": _round_embed(feat_m_e, cfg.embed_dtype),
        "y_l": y_l.to(torch.int64),
        "y_true": y_true.to(torch.int64),
        "prototypes": prototypes.contiguous(),
        "mask_random": torch.rand(B_u, generator=g) >= 0.5,
    }
    return out


def make_bank(k_bank: int, dim: int, num_classes: int, seed: int = SEED,
              dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """SimMatch-style memory bank (simmatch_model.py:68-70): unit rows + int64 labels.

    Stored row-major [K_b, D] (the reference stores the transpose [D, K_b])."""
    g = torch.Generator().manual_seed(seed + 101)
    bank = _unit(torch.randn(k_bank, dim, generator=g, dtype=torch.float32)).to(dtype)
    labels = torch.randint(0, num_classes, (k_bank,), generator=g).to(torch.int64)
    return {"bank": bank, "labels": labels}
