"""Generate tests/golden/bank_*.npz by EXECUTING THE REFERENCE's memory-bank baselines.

TEST INFRASTRUCTURE.  Run once in the authoring container (``python oracle/gen_golden_banks.py``); the GPU box has
no ``/root/reference`` and only reads the committed ``.npz`` files.

Executed, unmodified, from ``$STIL_REF`` (default ``/root/reference``), as *unbound functions* on stand-in ``self``
objects whose encoders return planted ``(logits, features)``:

* a7  ``SimMatchModel.forward`` + ``_update_bank`` + ``distribution_alignment``
      (``models/MatchModel/simmatch_model.py:225-292, 141-147, 150-163``) and the consumer lines of
      ``SimMatch.training_step`` (``models/MatchModel/SimMatch.py:75-92``)
* a8  ``CoMatchModel.forward`` + ``_dequeue_and_enqueue`` (``models/MatchModel/comatch_model.py:208-321, 117-146``)
      and ``CoMatch.training_step`` (``models/MatchModel/CoMatch.py:76-123``)
* a9  ``MMatch.training_step`` + ``_dequeue_and_enqueue`` + ``distribution_alignment``
      (``models/SemiMultimodal/MMatch.py:191-262, 102-117, 135-148``)

Packages the modules import but this image lacks (pytorch_lightning, pl_bolts, torchmetrics, the encoder modules)
are replaced by empty stubs in ``sys.modules`` — none is on the path.  Locals are captured with ``sys.settrace``.
Nothing from the reference is copied into the repo: only input/output tensors.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle.gen_golden import REF, OUT, _Capture, _stub_modules, to_np  # noqa: E402


def _stub_bank_modules():
    _stub_modules()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("models.self_supervised", torchvision_ssl_encoder=object)
    mod("models.MatchModel.multimodal_backbone", MultimodalBackbone=object)
    mod("models.pieces", DotDict=dict)
    mod("models.SemiMultimodal.Multimodal_model", MultimodalBackbone=object)


def load_bank_reference():
    sys.path.insert(0, str(REF))
    _stub_bank_modules()
    sim_model = importlib.import_module("models.MatchModel.simmatch_model")
    sim_pl = importlib.import_module("models.MatchModel.SimMatch")
    co_model = importlib.import_module("models.MatchModel.comatch_model")
    co_pl = importlib.import_module("models.MatchModel.CoMatch")
    mm_pl = importlib.import_module("models.SemiMultimodal.MMatch")
    return sim_model.SimMatchModel, sim_pl.SimMatch, co_model.CoMatchModel, co_pl.CoMatch, mm_pl.MMatch


def _unit(x, dim=1):
    return F.normalize(x, dim=dim)


def planted(g, rows, k, mu_choices=None):
    if mu_choices is None:
        mu_choices = (4.0, 8.0, 12.0) if k <= 100 else (7.0, 11.0, 15.0)
    c = torch.randint(0, k, (rows,), generator=g)
    y = torch.randn(rows, k, generator=g)
    mu = torch.tensor(mu_choices)[torch.randint(0, len(mu_choices), (rows,), generator=g)]
    y[torch.arange(rows), c] += mu
    return y, c


class _Seq:
    """Encoder stand-in: returns the planted outputs in call order."""
    def __init__(self, *outs):
        self.outs, self.i = list(outs), 0

    def __call__(self, x):
        o = self.outs[self.i]
        self.i += 1
        return o

    def eval(self):
        return self


noop = lambda *a, **k: None


# ---------------------------------------------------------------------------------------------------- a7
def simmatch_case(name, SimMatchModel, SimMatch, seed, b_l, b_u, k_cls, dim, k_bank, da, c_smooth, th=0.9):
    g = torch.Generator().manual_seed(seed)
    centers = _unit(torch.randn(k_cls, dim, generator=g))
    bank_labels = torch.randint(0, k_cls, (k_bank,), generator=g)
    bank = _unit(centers[bank_labels] + 0.6 * torch.randn(k_bank, dim, generator=g)).t().contiguous()   # [dim, K]
    logits_k, cls = planted(g, b_l + b_u, k_cls)
    logits_q = logits_k + 0.3 * torch.randn(b_l + b_u, k_cls, generator=g)
    feat_k = _unit(centers[cls] + 0.5 * torch.randn(b_l + b_u, dim, generator=g))
    feat_q = _unit(feat_k + 0.2 * torch.randn(b_l + b_u, dim, generator=g)).requires_grad_(True)
    logits_q = logits_q.requires_grad_(True)
    y_l = cls[:b_l].clone()
    index = torch.randperm(k_bank, generator=g)[:b_l]
    da_queue = torch.zeros(256, k_cls)
    da_ptr = torch.zeros(1, dtype=torch.long)
    if da:   # a queue that already holds a few batch means
        for r in range(5):
            da_queue[r] = torch.softmax(torch.randn(k_cls, generator=g) * 0.3, 0)
        da_ptr[0] = 5
    me = SimpleNamespace(eval_datatype="imaging", bank=bank.clone(), labels=bank_labels.clone(),
                         main=_Seq((logits_q, feat_q)), ema=_Seq((logits_k, feat_k)), momentum_update_ema=noop,
                         DA=da, DA_len=256, DA_queue=da_queue.clone(), DA_ptr=da_ptr.clone(), use_ddp=False,
                         tt=0.1, st=0.1, c_smooth=c_smooth, num_classes=k_cls)
    me.distribution_alignment = lambda p: SimMatchModel.distribution_alignment(me, p)
    me._update_bank = lambda k, l, i: SimMatchModel._update_bank(me, k, l, i)
    im = lambda n: torch.zeros(n, 1)
    cap = _Capture(SimMatchModel.forward.__code__)
    sys.settrace(cap)
    try:
        logits_qx, prob_ku, logits_qu, loss_in = SimMatchModel.forward(me, im(b_l), im(b_u), im(b_u), labels=y_l,
                                                                       index=index, start_unlabel=True)
    finally:
        sys.settrace(None)
    loc = cap.locals
    # consumer (SimMatch.py:75-92) on the real training_step
    pl = SimpleNamespace(current_epoch=10 ** 6, start_epoch=0, threshold=th, lambda_u=1.0, lambda_in=1.0,
                         hparams=SimpleNamespace(num_classes=k_cls), log=noop, acc_train=noop, auc_train=noop,
                         acc_train_unlabelled=noop, auc_train_unlabelled=noop,
                         model=lambda *a, **k: (logits_qx, prob_ku, logits_qu, loss_in))
    cap2 = _Capture(SimMatch.training_step.__code__)
    sys.settrace(cap2)
    try:
        SimMatch.training_step(pl, {"l": (im(b_l), y_l, index), "u": ((im(b_u), im(b_u)), cls[b_l:])}, 0)
    finally:
        sys.settrace(None)
    l2 = cap2.locals
    (d_feat_q,) = torch.autograd.grad(l2["loss_in"], (feat_q,), retain_graph=True)
    (d_logits_q,) = torch.autograd.grad(l2["loss_u"], (logits_q,), retain_graph=True)
    np.savez_compressed(
        OUT / f"{name}.npz",
        meta=np.array([b_l, b_u, k_cls, dim, k_bank, int(da)]), tt=0.1, st=0.1, c_smooth=c_smooth, threshold=th,
        bank=to_np(bank), bank_labels=to_np(bank_labels), logits_ku=to_np(logits_k[b_l:]),
        feat_k=to_np(feat_k), feat_q=to_np(feat_q), logits_q=to_np(logits_q), y_l=to_np(y_l), index=to_np(index),
        da_queue_in=to_np(da_queue), da_ptr_in=to_np(da_ptr),
        prob_ku_orig=to_np(loc["prob_ku_orig"]), prob_ku=to_np(prob_ku), loss_in_rows=to_np(loss_in),
        mask=to_np(l2["mask"]), loss_u=to_np(l2["loss_u"]), loss_in=to_np(l2["loss_in"]),
        d_feat_q=to_np(d_feat_q), d_logits_q=to_np(d_logits_q),
        bank_out=to_np(me.bank), labels_out=to_np(me.labels), da_queue_out=to_np(me.DA_queue), da_ptr_out=to_np(me.DA_ptr))
    print(f"{name}: mask rate {float(l2['mask'].mean()):.2f} loss_in {float(l2['loss_in']):.4f}")


# ---------------------------------------------------------------------------------------------------- a8
def comatch_case(name, CoMatchModel, CoMatch, seed, b_l, b_u, k_cls, dim, k_q, ptr_w, ptr_s, epoch=5, hist=3,
                 thr=0.9, contrast_th=0.8):
    g = torch.Generator().manual_seed(seed)
    centers = _unit(torch.randn(k_cls, dim, generator=g))
    q_cls_w = torch.randint(0, k_cls, (k_q,), generator=g)
    q_cls_s = torch.randint(0, k_cls, (k_q,), generator=g)
    queue_w = _unit(centers[q_cls_w] + 0.6 * torch.randn(k_q, dim, generator=g)).t().contiguous()
    queue_s = _unit(centers[q_cls_s] + 0.6 * torch.randn(k_q, dim, generator=g)).t().contiguous()
    soft = lambda c: torch.softmax(8.0 * F.one_hot(c, k_cls).float() + torch.randn(len(c), k_cls, generator=g), 1)
    probs_xu = soft(q_cls_w).t().contiguous()     # [C, K_q]
    probs_u = soft(q_cls_s).t().contiguous()
    outputs_m, cls = planted(g, b_l + 2 * b_u, k_cls)
    cls[b_l + b_u:] = cls[b_l:b_l + b_u]          # s1 view of the same unlabelled samples
    features_m = _unit(centers[cls] + 0.5 * torch.randn(b_l + 2 * b_u, dim, generator=g))
    outputs = (outputs_m[:b_l + b_u] + 0.3 * torch.randn(b_l + b_u, k_cls, generator=g)).requires_grad_(True)
    features = _unit(features_m[:b_l + b_u] + 0.2 * torch.randn(b_l + b_u, dim, generator=g)).requires_grad_(True)
    labels_x = cls[:b_l].clone()
    hist_prob = [torch.softmax(torch.randn(k_cls, generator=g) * 0.3, 0) for _ in range(hist)]
    me = SimpleNamespace(eval_datatype="imaging", encoder=_Seq((outputs, features)), m_encoder=_Seq((outputs_m, features_m)),
                         _update_momentum_encoder=noop, momentum=0.99, use_ddp=False, hist_prob=list(hist_prob),
                         start_epoch=0, temperature=0.1, alpha=0.9, K=k_q, num_classes=k_cls,
                         queue_w=queue_w.clone(), probs_xu=probs_xu.clone(), queue_ptr_w=torch.tensor([ptr_w]),
                         queue_s=queue_s.clone(), probs_u=probs_u.clone(), queue_ptr_s=torch.tensor([ptr_s]))
    me._dequeue_and_enqueue = lambda z, t, ws: CoMatchModel._dequeue_and_enqueue(me, z, t, ws)
    im = lambda n: torch.zeros(n, 1)
    cap = _Capture(CoMatchModel.forward.__code__)

    def model(labeled, unlabeled, epoch=0):
        sys.settrace(cap)
        try:
            return CoMatchModel.forward(me, labeled, unlabeled, epoch=epoch)
        finally:
            sys.settrace(cap2)

    pl = SimpleNamespace(current_epoch=epoch, start_epoch=0, thr=thr, contrast_th=contrast_th, lam_c=1.0, lam_u=1.0,
                         criterion=torch.nn.CrossEntropyLoss(), hparams=SimpleNamespace(num_classes=k_cls), log=noop,
                         acc_train=noop, auc_train=noop, acc_train_unlabelled=noop, auc_train_unlabelled=noop, model=model)
    cap2 = _Capture(CoMatch.training_step.__code__)
    sys.settrace(cap2)
    try:
        CoMatch.training_step(pl, {"l": (im(b_l), labels_x, None), "u": ((im(b_u), im(b_u), im(b_u)), cls[b_l:b_l + b_u])}, 0)
    finally:
        sys.settrace(None)
    loc, l2 = cap.locals, cap2.locals
    (d_feat,) = torch.autograd.grad(l2["loss_contrast"], (features,), retain_graph=True)
    (d_out,) = torch.autograd.grad(l2["loss_u"], (outputs,), retain_graph=True)
    np.savez_compressed(
        OUT / f"{name}.npz",
        meta=np.array([b_l, b_u, k_cls, dim, k_q, ptr_w, ptr_s, epoch]), temperature=0.1, alpha=0.9, thr=thr,
        contrast_th=contrast_th,
        queue_w=to_np(queue_w), probs_xu=to_np(probs_xu), queue_s=to_np(queue_s), probs_u=to_np(probs_u),
        outputs_m=to_np(outputs_m), features_m=to_np(features_m), outputs=to_np(outputs), features=to_np(features),
        labels_x=to_np(labels_x), hist_prob=to_np(torch.stack(hist_prob)),
        probs_orig=to_np(loc["probs_orig"]), probs=to_np(loc["probs"]), Q=to_np(loc["Q"]), sim=to_np(loc["sim"]),
        mask=to_np(l2["mask"]), pos_mask=to_np(l2["pos_mask"]), loss_u=to_np(l2["loss_u"]),
        loss_contrast=to_np(l2["loss_contrast"]), d_features=to_np(d_feat), d_outputs=to_np(d_out),
        queue_w_out=to_np(me.queue_w), probs_xu_out=to_np(me.probs_xu), queue_ptr_w_out=to_np(me.queue_ptr_w),
        queue_s_out=to_np(me.queue_s), probs_u_out=to_np(me.probs_u), queue_ptr_s_out=to_np(me.queue_ptr_s))
    print(f"{name}: mask rate {float(l2['mask'].mean()):.2f} pos edges/row {float(l2['pos_mask'].float().sum(1).mean()):.1f} "
          f"loss_contrast {float(l2['loss_contrast']):.4f}")


# ---------------------------------------------------------------------------------------------------- a9
def mmatch_case(name, MMatch, seed, b_l, b_u, k_cls, dim, k_q, ptr, da, epoch=5, th1=0.9):
    g = torch.Generator().manual_seed(seed)
    centers = _unit(torch.randn(k_cls, dim, generator=g))
    q_cls = torch.randint(0, k_cls, (k_q,), generator=g)
    embed_queue = _unit(centers[q_cls] + 0.6 * torch.randn(k_q, dim, generator=g)).t().contiguous()
    probs_queue = torch.softmax(8.0 * F.one_hot(q_cls, k_cls).float() + torch.randn(k_q, k_cls, generator=g), 1).t().contiguous()
    B = b_l + b_u
    y_m, cls = planted(g, B, k_cls)
    y_i = (y_m + 0.5 * torch.randn(B, k_cls, generator=g)).requires_grad_(True)
    y_t = (y_m + 0.5 * torch.randn(B, k_cls, generator=g)).requires_grad_(True)
    y_m = y_m.requires_grad_(True)
    x_m = centers[cls] + 0.5 * torch.randn(B, dim, generator=g)       # un-normalised multimodal embedding (:206)
    y_l = cls[:b_l].clone()
    da_queue = torch.zeros(256, k_cls)
    da_ptr = torch.zeros(1, dtype=torch.long)
    for r in range(4):
        da_queue[r] = torch.softmax(torch.randn(k_cls, generator=g) * 0.3, 0)
    da_ptr[0] = 4
    me = SimpleNamespace(current_epoch=epoch, start_epoch=0, forward=lambda x: (y_m, y_i, y_t, x_m),
                         criterion_ce=torch.nn.CrossEntropyLoss(), T=0.1, th1=th1, alpha=1.0, mmatch_lambda=1.0,
                         embed_queue=embed_queue.clone(), probs_queue=probs_queue.clone(),
                         embed_queue_ptr=torch.tensor([ptr]), K=k_q, use_ddp=False, DA_len=256,
                         DA_queue=da_queue.clone(), DA_ptr=da_ptr.clone(),
                         hparams=SimpleNamespace(num_classes=k_cls, DA=da), log=noop, acc_train=noop, auc_train=noop,
                         acc_train_unlabelled=noop, auc_train_unlabelled=noop)
    if da:
        me.distribution_alignment = lambda p: MMatch.distribution_alignment(me, p)
    else:
        # MMatch.__init__ binds an identity when DA is off
        me.distribution_alignment = lambda p: p
    me._dequeue_and_enqueue = lambda z, t, ws: MMatch._dequeue_and_enqueue(me, z, t, ws)
    dummy = [None, torch.zeros(1)]
    batch = {"l": (dummy, dummy, y_l, None, torch.ones(b_l)), "u": (dummy, dummy, cls[b_l:], None, torch.zeros(b_u))}
    cap = _Capture(MMatch.training_step.__code__)
    sys.settrace(cap)
    try:
        MMatch.training_step(me, batch, 0)
    finally:
        sys.settrace(None)
    loc = cap.locals
    g_i, g_t = torch.autograd.grad(loc["loss_i_u"] + loc["loss_t_u"], (y_i, y_t), retain_graph=True)
    np.savez_compressed(
        OUT / f"{name}.npz",
        meta=np.array([b_l, b_u, k_cls, dim, k_q, ptr, int(da), epoch]), T=0.1, th1=th1,
        embed_queue=to_np(embed_queue), probs_queue=to_np(probs_queue), y_m=to_np(y_m), y_i=to_np(y_i), y_t=to_np(y_t),
        x_m=to_np(x_m), y_l=to_np(y_l), da_queue_in=to_np(da_queue), da_ptr_in=to_np(da_ptr),
        feat_m=to_np(loc["feat_m"]), pseudo_label_orig=to_np(loc["pseudo_label_orig"]),
        pseudo_label=to_np(loc["pseudo_label"] if k_cls != 2 else loc["pseudo_label_all"][b_l:]),
        max_prob=to_np(loc["max_prob"]), max_idx=to_np(loc["max_idx"]), mask1=to_np(loc["mask1"]),
        loss_i_u=to_np(loc["loss_i_u"]), loss_t_u=to_np(loc["loss_t_u"]), d_y_i=to_np(g_i), d_y_t=to_np(g_t),
        embed_queue_out=to_np(me.embed_queue), probs_queue_out=to_np(me.probs_queue),
        embed_queue_ptr_out=to_np(me.embed_queue_ptr), da_queue_out=to_np(me.DA_queue), da_ptr_out=to_np(me.DA_ptr))
    print(f"{name}: mask rate {float(loc['mask1'].float().mean()):.2f}")


def main():
    torch.set_num_threads(1)
    SimMatchModel, SimMatch, CoMatchModel, CoMatch, MMatch = load_bank_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    simmatch_case("bank_simmatch_k10", SimMatchModel, SimMatch, 3101, 8, 56, 10, 128, 512, da=True, c_smooth=0.9)
    simmatch_case("bank_simmatch_k286", SimMatchModel, SimMatch, 3102, 8, 40, 286, 64, 1024, da=False, c_smooth=0.9)
    simmatch_case("bank_simmatch_nosmooth", SimMatchModel, SimMatch, 3103, 4, 28, 2, 32, 96, da=False, c_smooth=1.0, th=0.85)
    comatch_case("bank_comatch_k10", CoMatchModel, CoMatch, 3201, 8, 56, 10, 128, 256, ptr_w=64, ptr_s=0)
    comatch_case("bank_comatch_wrap", CoMatchModel, CoMatch, 3202, 8, 40, 286, 64, 320, ptr_w=300, ptr_s=296, contrast_th=0.5)
    comatch_case("bank_comatch_pre_start", CoMatchModel, CoMatch, 3203, 4, 28, 2, 32, 96, ptr_w=0, ptr_s=32, epoch=0)
    mmatch_case("bank_mmatch_k10", MMatch, 3301, 8, 56, 10, 128, 640, ptr=128, da=True)
    mmatch_case("bank_mmatch_k286_wrap", MMatch, 3302, 8, 40, 286, 64, 320, ptr=288, da=False)
    mmatch_case("bank_mmatch_epoch0", MMatch, 3303, 4, 28, 2, 32, 96, ptr=0, da=True, epoch=0, th1=0.85)


if __name__ == "__main__" and "--club" not in sys.argv:
    main()


# ---------------------------------------------------------------------------------------------------- f-2
def club_case(name, seed, b, d, hidden):
    """models/Disentangle/utils/club.py CLUBMean (real module): forward (the MI upper bound, :107-121) and learning_loss
    (:125-130) on planted samples; the fixture records mu = p_mu(x), so the GPU path is compared from mu on."""
    club = importlib.import_module("models.Disentangle.utils.club")
    torch.manual_seed(seed)
    m = club.CLUBMean(d, d, hidden)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, d, generator=g)
    y = (0.6 * x + 0.8 * torch.randn(b, d, generator=g)).requires_grad_(True)
    mu = m.p_mu(x).detach().requires_grad_(True)
    m.get_mu_logvar = lambda _x: (mu, 0)           # forward()/learning_loss() from the recorded mu on
    bound = m.forward(x, y)
    est = m.learning_loss(x, y)
    gb = torch.autograd.grad(bound, (mu, y), retain_graph=True)
    ge = torch.autograd.grad(est, (mu, y))
    np.savez_compressed(OUT / f"{name}.npz", mu=to_np(mu), y=to_np(y), bound=to_np(bound), est=to_np(est),
                        d_mu_bound=to_np(gb[0]), d_y_bound=to_np(gb[1]), d_mu_est=to_np(ge[0]), d_y_est=to_np(ge[1]))
    print(f"{name}: bound {float(bound):.5f} est {float(est):.4f}")


def main_club():
    sys.path.insert(0, str(REF))
    _stub_bank_modules()
    OUT.mkdir(parents=True, exist_ok=True)
    club_case("club_b64_d128", 3401, 64, 128, 64)
    club_case("club_b56_d512", 3402, 56, 512, None)
    club_case("club_b37_d24", 3403, 37, 24, 16)


if __name__ == "__main__" and "--club" in sys.argv:
    main_club()
