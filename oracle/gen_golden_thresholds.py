"""Generate tests/golden/freematch.npz by EXECUTING the reference: ``FreeMatchModel.masking`` / ``.update``
(``models/MatchModel/FreeMatchFolder/freematch_model.py:128-165``) as unbound functions on a stand-in ``self`` that carries the
three state tensors (the constructor builds encoders, which are upstream of this block), three consecutive batches, and
``entropy_loss`` (``FreeMatchFolder/freematch_utils.py:17-45``) with its autograd gradient; plus tests/golden/cotraining.npz
from the oracle restatement of ``models/SemiMultimodal/CoTraining.py:141-149`` (those lines sit inline in a Lightning
``training_step`` behind two encoders: parity for them is pinned by restatement only).  TEST INFRASTRUCTURE; run once in the
authoring container (``python oracle/gen_golden_thresholds.py``).  Only tensors are stored, no reference source.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle.gen_golden import OUT, REF, _stub_modules  # noqa: E402
from oracle import stil_head_oracle as O  # noqa: E402


def load_freematch():
    sys.path.insert(0, str(REF))
    _stub_modules()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    mod("pl_bolts.utils"); mod("pl_bolts.utils.self_supervised", torchvision_ssl_encoder=object)
    mod("models.MatchModel.multimodal_backbone", MultimodalBackbone=object)
    mod("models.pieces", DotDict=dict)
    import importlib
    fm = importlib.import_module("models.MatchModel.FreeMatchFolder.freematch_model")      # real file, stubbed encoders
    fu = importlib.import_module("models.MatchModel.FreeMatchFolder.freematch_utils")      # real file
    return fm.FreeMatchModel, fu.entropy_loss


def main():
    FreeMatchModel, entropy_loss = load_freematch()
    rec = {}
    for name, rows, c, scale in (("dvm", 96, 286, 6.0), ("cardiac", 256, 2, 2.0), ("small", 37, 10, 4.0)):
        g = torch.Generator().manual_seed(len(name) * 1000 + c)
        # a tensor has no .is_cuda setter: the reference's device moves (:148-153) see CPU tensors and call .to(cpu)
        me = SimpleNamespace(m=0.999, clip_thresh=0.0, use_ddp=False, p_model=torch.ones(c) / c, label_hist=torch.ones(c) / c)
        me.time_p = me.p_model.mean()
        me.update = types.MethodType(FreeMatchModel.update, me)
        for step in range(3):
            logits = torch.randn(rows, c, generator=g) * scale
            mask = FreeMatchModel.masking(me, logits)
            rec[f"{name}_logits{step}"] = logits.numpy().copy()
            rec[f"{name}_mask{step}"] = mask.numpy().copy()
            rec[f"{name}_time_p{step}"] = me.time_p.reshape(1).numpy().copy()
            rec[f"{name}_p_model{step}"] = me.p_model.numpy().copy()
            rec[f"{name}_label_hist{step}"] = me.label_hist.numpy().copy()
        # fairness loss on the strong-view logits of the last batch, with the state the last masking left
        logits_s = (logits + 0.7 * torch.randn(rows, c, generator=g)).requires_grad_(True)
        if mask.sum() == 0:
            mask[0] = 1.0
        loss, hist_mean = entropy_loss(mask, logits_s, me.p_model, me.label_hist)
        (gl,) = torch.autograd.grad(loss, logits_s)
        rec[f"{name}_ent_logits_s"] = logits_s.detach().numpy().copy()
        rec[f"{name}_ent_mask"] = mask.numpy().copy()
        rec[f"{name}_ent_loss"] = loss.detach().reshape(1).numpy().copy()
        rec[f"{name}_ent_hist_mean"] = hist_mean.reshape(1).numpy().copy()
        rec[f"{name}_ent_grad"] = gl.numpy().copy()
        # a clipped variant (clip_thresh is 0.0 in the reference constructor, :48; the branch :139-140 exists)
        me2 = SimpleNamespace(m=0.9, clip_thresh=1.0, use_ddp=False, p_model=torch.ones(c) / c, label_hist=torch.ones(c) / c)
        me2.time_p = torch.tensor(0.99)
        me2.update = types.MethodType(FreeMatchModel.update, me2)
        mask2 = FreeMatchModel.masking(me2, logits)
        rec[f"{name}_clip_mask"] = mask2.numpy().copy()
        rec[f"{name}_clip_time_p"] = me2.time_p.reshape(1).numpy().copy()
    np.savez_compressed(OUT / "freematch.npz", **rec)
    print("wrote", OUT / "freematch.npz", {k: v.shape for k, v in list(rec.items())[:6]})

    rec = {}
    for name, rows, c in (("dvm", 64, 286), ("cardiac", 256, 2)):
        g = torch.Generator().manual_seed(rows + c)
        ys = [(torch.randn(rows, c, generator=g) * 3) for _ in range(4)]
        yi, yt = ys[0].clone().requires_grad_(True), ys[1].clone().requires_grad_(True)
        out = O.cotraining_unsup(yi, yt, ys[2], ys[3], 0.9 if c > 2 else 0.8)
        gi, gt = torch.autograd.grad(out["loss_i_u"] + out["loss_t_u"], (yi, yt))
        for k, v in zip(("y_i", "y_t", "y_i_e", "y_t_e"), ys):
            rec[f"{name}_{k}"] = v.numpy().copy()
        rec[f"{name}_threshold"] = np.float32(0.9 if c > 2 else 0.8)
        for k in ("mask_i", "mask_t", "max_prob_i", "max_prob_t", "loss_i_u", "loss_t_u"):
            rec[f"{name}_{k}"] = out[k].detach().float().reshape(-1).numpy().copy()
        rec[f"{name}_d_y_i"], rec[f"{name}_d_y_t"] = gi.numpy().copy(), gt.numpy().copy()
    np.savez_compressed(OUT / "cotraining.npz", **rec)
    print("wrote", OUT / "cotraining.npz")


if __name__ == "__main__":
    main()
