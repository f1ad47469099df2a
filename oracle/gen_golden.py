"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE in the authoring container.

TEST INFRASTRUCTURE.  Run once here (``python oracle/gen_golden.py``); the GPU box
has no ``/root/reference`` and only reads the committed ``.npz`` files.

What is executed, unmodified, from ``$STIL_REF`` (default ``/root/reference``):

* ``utils/clip_loss.py``  ``CLIPLoss``   (real module, real autograd)
* ``utils/prototype_loss.py`` ``PrototypeLoss``
* ``models/Disentangle/STiLModel.py``: the bodies of ``STiLModel.training_step``
  (lines 228-386: CGPL 262-279, PGLS 291-299, masked CE 301-303, ITC 322, PT 339,
  prototype partials 374-381), ``cal_prototypes`` / ``cal_prototypes_separate``
  (199-226) and ``training_epoch_end`` (389-421) are called as *unbound functions*
  on a stand-in ``self`` that supplies planted tensors instead of the encoders.
  The packages the module imports but this image lacks (pytorch_lightning,
  lightly, pl_bolts, torchmetrics and the backbone module with its
  omegaconf/timm dependencies) are replaced by empty stubs in ``sys.modules`` — none
  of them is on the hot path.  ``training_step``'s locals (pseudo_label, mask1,
  cases, max_idx, ...) are captured with ``sys.settrace`` at its return.

Nothing from the reference is copied into the repo: only input/output tensors.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from stil_tta_b200 import synth  # noqa: E402

REF = Path(os.environ.get("STIL_REF", "/root/reference"))
OUT = REPO / "tests" / "golden"


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _LM(torch.nn.Module):
        pass

    mod("torchmetrics")
    mod("pytorch_lightning", LightningModule=_LM)
    mod("lightly"); mod("lightly.models"); mod("lightly.models.modules", SimCLRProjectionHead=object)
    mod("pl_bolts"); mod("pl_bolts.optimizers")
    mod("pl_bolts.optimizers.lr_scheduler", LinearWarmupCosineAnnealingLR=object)
    # the backbone (encoders) is upstream of the head; stub the module, not the head code
    mod("models.Disentangle.utils.STiLModel_backbone", DisCoAttentionBackbone=object)


def load_reference():
    sys.path.insert(0, str(REF))
    _stub_modules()
    from utils.clip_loss import CLIPLoss                       # real
    from utils.prototype_loss import PrototypeLoss             # real
    import importlib
    stil = importlib.import_module("models.Disentangle.STiLModel")   # real file, stubbed deps
    return CLIPLoss, PrototypeLoss, stil.STiLModel


class _Capture:
    """settrace hook: grab f_locals of `code` when it returns."""
    def __init__(self, code):
        self.code, self.locals = code, None

    def __call__(self, frame, event, arg):
        if frame.f_code is self.code:
            return self._local
        return None

    def _local(self, frame, event, arg):
        if event == "return":
            self.locals = {k: v for k, v in frame.f_locals.items()}
        return self._local


def _ensure_process_group():
    """STiLModel.distribution_alignment calls torch.distributed.all_reduce unguarded (STiLModel.py:174): a
    one-rank gloo group makes it the identity."""
    import tempfile
    import torch.distributed as dist
    if not dist.is_initialized():
        f = tempfile.NamedTemporaryFile(delete=False)
        dist.init_process_group("gloo", store=dist.FileStore(f.name, 1), rank=0, world_size=1)


def run_reference_step(STiLModel, CLIPLoss, PrototypeLoss, batch, cfg, prototypes_sum, prototypes_count_sum,
                       da_state=None):
    """Execute the real training_step on planted tensors; return its locals + grads."""
    f32 = lambda t: t.to(torch.float32)
    B_l, P = cfg.b_l, cfg.proj_dim
    # student tensors carry grad, as if produced by the student backbone
    y = {k: f32(batch[k]).clone().requires_grad_(True) for k in ("y_m", "y_i", "y_t")}
    feats = {k: f32(batch[k]).clone().requires_grad_(True) for k in ("feat_i", "feat_t", "feat_m")}
    # teacher logits: labelled rows are unused by the head; splice the planted unlabelled rows
    zeros_l = torch.zeros(B_l, cfg.num_classes)
    ye = {k: torch.cat((zeros_l, f32(batch[k + "_ue"]))) for k in ("y_m", "y_i", "y_t")}
    feat_m_e = f32(batch["feat_m_e"])
    p1 = P // 3
    split = lambda t: (t[:, :p1], t[:, p1:2 * p1], t[:, 2 * p1:])

    class Student:
        def forward_all(self, x):
            si, c, st = split(feats["feat_m"])
            # (y_hat_m, y_hat_i, y_hat_t, x_si_enhance, x_si, x_ai, x_st_enhance, x_st, x_at, x_c)
            return (y["y_m"], y["y_i"], y["y_t"], si, None, feats["feat_i"], st, None, feats["feat_t"], c)

    class Teacher:
        def eval(self):
            return self

        def forward_all(self, x):
            si, c, st = split(feat_m_e)
            return (ye["y_m"], ye["y_i"], ye["y_t"], si, None, None, st, None, None, c)

    # STiLModel.project_3features (182-192) = Linear/MLP + F.normalize; the head's inputs are its
    # outputs, so the projectors are identities here and the planted features are already unit-norm.
    def project_3features(feat_m=None, feat_i=None, feat_t=None):
        return feat_m, feat_i, feat_t

    class Club:
        def __call__(self, a, b):
            return torch.zeros(())

        def learning_loss(self, a, b):
            return torch.zeros(())

    noop = lambda *a, **k: None
    me = SimpleNamespace(
        current_epoch=(10 ** 6 if cfg.past_start_epoch else 0), start_epoch=35,
        model=Student(), ema=Teacher(), use_ema=True, momentum_update_ema=noop,
        project_3features=project_3features,
        hparams=SimpleNamespace(DA=False, num_classes=cfg.num_classes),
        criterion_ce=torch.nn.CrossEntropyLoss(),
        criterion_itc=CLIPLoss(temperature=cfg.temperature, lambda_0=cfg.lambda_0),
        criterion_pt=PrototypeLoss(temperature=cfg.temperature, threshold=cfg.th1),
        prototypes=f32(batch["prototypes"]).clone(),
        prototypes_sum=prototypes_sum, prototypes_count_sum=prototypes_count_sum,
        T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1, repeat_ratio=cfg.repeat_ratio,
        alpha=1.0, beta=1.0, gamma=1.0, rate_pt=1.0, rate_uce=1.0,
        CLUB_imaging=Club(), CLUB_tabular=Club(), use_ddp=False, log=noop,
        acc_train=noop, auc_train=noop, acc_train_unlabelled=noop, auc_train_unlabelled=noop,
    )
    me.sharpen_predictions = lambda logits, temperature: STiLModel.sharpen_predictions(me, logits, temperature)
    me.cal_prototypes = lambda label, feat: STiLModel.cal_prototypes(me, label, feat)
    me.cal_prototypes_separate = lambda label, feat, bl: STiLModel.cal_prototypes_separate(me, label, feat, bl)
    if da_state is not None:
        # hparams.DA == True: the real distribution_alignment (STiLModel.py:171-180) on the real buffers (:98-100)
        _ensure_process_group()
        me.hparams.DA = True
        me.DA_queue, me.DA_ptr, me.DA_len = da_state["DA_queue"], da_state["DA_ptr"], da_state["DA_queue"].shape[0]
        me.distribution_alignment = lambda probs: STiLModel.distribution_alignment(me, probs)

    B = cfg.batch
    ident_l, ident_u = torch.ones(B_l), torch.zeros(B - B_l)
    dummy = [None, torch.zeros(1)]
    batch_arg = {"l": (dummy, dummy, batch["y_l"], None, ident_l),
                 "u": (dummy, dummy, batch["y_true"][B_l:], None, ident_u)}
    # torch.cat((im_views_l[1], im_views_u[1])) must work: give 1-element tensors
    # mask_random: the reference draws it with torch.rand_like (line 299); pin the RNG so the
    # fixture records the very mask that was used.
    torch.manual_seed(1234)
    cap = _Capture(STiLModel.training_step.__code__)
    sys.settrace(cap)
    try:
        loss = STiLModel.training_step(me, batch_arg, 0)
    finally:
        sys.settrace(None)
    loc = cap.locals
    grads = {}
    # separate gradients of the three head losses (the total loss mixes them with weights)
    g_i, g_t = torch.autograd.grad(loc["loss_itc"], (feats["feat_i"], feats["feat_t"]), retain_graph=True)
    (g_m,) = torch.autograd.grad(loc["loss_pt"], (feats["feat_m"],), retain_graph=True)
    g_y = torch.autograd.grad(loc["loss_m_u"] + loc["loss_i_u"] + loc["loss_t_u"],
                              (y["y_m"], y["y_i"], y["y_t"]), retain_graph=True, allow_unused=True)
    grads.update(d_feat_i=g_i, d_feat_t=g_t, d_feat_m=g_m,
                 d_y_m=g_y[0], d_y_i=g_y[1], d_y_t=g_y[2])
    return loc, grads, me


def to_np(t):
    if isinstance(t, torch.Tensor):
        t = t.detach()
        if t.dtype == torch.bfloat16:
            return t.view(torch.int16).numpy().copy()   # raw bf16 bits
        return t.numpy().copy()
    return np.asarray(t)


def save_case(name, cfg, seed, STiLModel, CLIPLoss, PrototypeLoss, da=False, **mk):
    batch = synth.make_batch(cfg, seed=seed, **mk)
    K, P = cfg.num_classes, cfg.proj_dim
    psum, pcnt = torch.zeros(K, P), torch.zeros(K, 1)
    da_state, da_in = None, {}
    if da:
        # a ring buffer that already holds a few batch means and is about to wrap (ptr = len - 1)
        g = torch.Generator().manual_seed(seed + 1)
        q = torch.zeros(256, K)
        q[250:] = torch.softmax(torch.randn(6, K, generator=g), dim=1)
        q[:3] = torch.softmax(torch.randn(3, K, generator=g) * 0.5, dim=1)
        da_state = {"DA_queue": q, "DA_ptr": torch.tensor([255], dtype=torch.int64)}
        da_in = {"in_DA_queue": to_np(q.clone()), "in_DA_ptr": to_np(da_state["DA_ptr"].clone())}
    loc, grads, me = run_reference_step(STiLModel, CLIPLoss, PrototypeLoss, batch, cfg, psum, pcnt, da_state)
    # class partials as the reference computed them this step (accumulators started at zero)
    rec = {
        "meta_seed": seed, "meta_batch": cfg.batch, "meta_K": K, "meta_P": P,
        "meta_embed_bf16": int(cfg.embed_dtype == "bf16"), "meta_past_start": int(cfg.past_start_epoch),
        "meta_zero_protos": int(mk.get("zero_prototypes", False)), "meta_edge": int(mk.get("edge_rows", False)),
        "meta_cfg_name": cfg.name,
    }
    for k in ("pseudo_label", "max_prob", "max_idx", "mask1", "mask_random", "case1", "case2_i",
              "case2_t", "case3", "top1_m", "top1_i", "top1_t", "teacher_probs", "pseudo_label_orig",
              "pseudo_label_all", "loss_itc", "loss_pt", "loss_m_u", "loss_i_u", "loss_t_u", "logits"):
        rec["ref_" + k] = to_np(loc[k])
    # `prediction` after the epoch gate (line 317-320); pre-gate value = rows of pseudo_label_all when past start
    rec["ref_prediction_gated"] = to_np(loc["prediction"])
    rec["ref_class_sum"] = to_np(me.prototypes_sum)
    rec["ref_class_count"] = to_np(me.prototypes_count_sum)
    for k, v in grads.items():
        rec["ref_" + k] = to_np(v if v is not None else torch.zeros(1))
    # epoch end on the accumulated state (real training_epoch_end needs every class seen; run it
    # only when that holds, otherwise record the count of empty classes)
    empty = int((me.prototypes_count_sum < 1).sum())
    rec["ref_empty_classes"] = empty
    if da:
        rec.update(da_in)
        rec["ref_DA_queue"], rec["ref_DA_ptr"] = to_np(da_state["DA_queue"]), to_np(da_state["DA_ptr"])
        rec["meta_da"] = 1
    for k, v in batch.items():
        rec["in_" + k] = to_np(v)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **rec)
    print(f"{name}: loss_itc={float(loc['loss_itc']):.6f} loss_pt={float(loc['loss_pt']):.6f} "
          f"mask1={float(loc['mask1'].float().mean()):.3f} case1={float(loc['case1'].float().mean()):.3f} "
          f"case3={float(loc['case3'].float().mean()):.3f} empty={empty}")


def save_module_cases(CLIPLoss, PrototypeLoss):
    """Direct module-level vectors, incl. the reference's own __main__ smoke inputs
    (utils/prototype_loss.py:42-48: integer labels, an all-zero row)."""
    g = torch.Generator().manual_seed(7)
    rec = {}
    for tag, (n, d, t, lam) in {"a": (37, 24, 0.1, 0.5), "b": (128, 128, 0.07, 0.3), "c": (5, 8, 1.0, 1.0),
                                "d": (96, 512, 0.1, 0.0)}.items():
        a = torch.randn(n, d, generator=g).requires_grad_(True)
        b = torch.randn(n, d, generator=g).requires_grad_(True)
        loss, logits, labels = CLIPLoss(t, lam)(a, b)
        ga, gb = torch.autograd.grad(loss, (a, b))
        rec.update({f"clip_{tag}_a": to_np(a), f"clip_{tag}_b": to_np(b), f"clip_{tag}_T": t, f"clip_{tag}_lam": lam,
                    f"clip_{tag}_loss": to_np(loss), f"clip_{tag}_logits": to_np(logits),
                    f"clip_{tag}_labels": to_np(labels), f"clip_{tag}_ga": to_np(ga), f"clip_{tag}_gb": to_np(gb)})
    # prototype loss: smoke-block shaped input (int labels, zero row) + random soft labels
    label = torch.tensor([[0, 1], [1, 0], [0, 0]])
    protos = torch.nn.functional.normalize(torch.randn(2, 128, generator=g))
    feat = torch.nn.functional.normalize(torch.randn(3, 128, generator=g)).requires_grad_(True)
    loss = PrototypeLoss(0.1, 0.9)(label, protos, feat)
    (gf,) = torch.autograd.grad(loss, (feat,))
    rec.update(pt_smoke_label=to_np(label), pt_smoke_protos=to_np(protos), pt_smoke_feat=to_np(feat),
               pt_smoke_loss=to_np(loss), pt_smoke_gfeat=to_np(gf))
    for tag, (n, k, d, t, th) in {"a": (64, 286, 128, 0.1, 0.9), "b": (33, 2, 128, 0.1, 0.85),
                                  "c": (50, 10, 64, 0.5, 0.3)}.items():
        label = torch.softmax(torch.randn(n, k, generator=g) * 4, dim=1)
        protos = torch.randn(k, d, generator=g) * 0.3
        feat = torch.nn.functional.normalize(torch.randn(n, d, generator=g)).requires_grad_(True)
        loss = PrototypeLoss(t, th)(label, protos, feat)
        (gf,) = torch.autograd.grad(loss, (feat,))
        rec.update({f"pt_{tag}_label": to_np(label), f"pt_{tag}_protos": to_np(protos), f"pt_{tag}_feat": to_np(feat),
                    f"pt_{tag}_T": t, f"pt_{tag}_th": th, f"pt_{tag}_loss": to_np(loss), f"pt_{tag}_gfeat": to_np(gf)})
    try:
        CLIPLoss(0.1, 1.5)
        rec["clip_bad_lambda_raises"] = 0
    except ValueError:
        rec["clip_bad_lambda_raises"] = 1
    np.savez_compressed(OUT / "modules.npz", **rec)
    print("modules.npz written")


def main():
    torch.set_num_threads(1)   # deterministic reductions
    CLIPLoss, PrototypeLoss, STiLModel = load_reference()
    save_module_cases(CLIPLoss, PrototypeLoss)
    S = synth
    args = (STiLModel, CLIPLoss, PrototypeLoss)
    save_case("step_c1_dvm_b64_f32", S.dvm_config(64, embed_dtype="f32"), 2022, *args)
    save_case("step_c1_dvm_b64_bf16", S.dvm_config(64), 2022, *args)
    save_case("step_dvm_b128_edge", S.dvm_config(128, embed_dtype="f32"), 2023, *args, edge_rows=True)
    save_case("step_dvm_b64_pre_start", S.dvm_config(64, past_start_epoch=False, repeat_ratio=1.0), 2024, *args)
    save_case("step_dvm_b64_zero_protos", S.dvm_config(64), 2025, *args, zero_prototypes=True)
    save_case("step_cardiac_b128", S.cardiac_config(128), 2026, *args)
    save_case("step_cardiac_b64_ragged", S.cardiac_config(72, unlabelled_ratio=5), 2027, *args)
    save_case("step_dvm_b200_k10", S.dvm_config(200, num_classes=10, proj_dim=64, unlabelled_ratio=3,
                                                 embed_dtype="f32", th1=0.6), 2028, *args)
    # hparams.DA == True (STiLModel.py:276-277): prediction = distribution_alignment(softmax(y_hat_m_ue))
    save_case("step_dvm_b64_da", S.dvm_config(64, embed_dtype="f32"), 2029, *args, da=True)
    save_case("step_cardiac_b128_da", S.cardiac_config(128), 2030, *args, da=True)


if __name__ == "__main__":
    main()
