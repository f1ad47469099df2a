"""CPU oracle for the STiL per-batch head — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``stil_tta_b200``) never routes through it and has no CPU fallback.

This is an independently written restatement, in plain PyTorch ops, of the
reference algorithm.  Every function cites the reference lines it follows
(paths relative to the upstream repo kgutjahr/STiL-TTA):

* ``utils/clip_loss.py:27-40``          -> :func:`clip_loss`
* ``STiLModel.py:262-279`` (CGPL)       -> :func:`cgpl`
* ``STiLModel.py:291-299`` (PGLS)       -> :func:`pgls`
* ``STiLModel.py:301-303``              -> :func:`masked_soft_ce`
* ``STiLModel.py:317-321``              -> :func:`pseudo_label_all`
* ``utils/prototype_loss.py:24-40``     -> :func:`prototype_loss`
* ``STiLModel.py:199-226``              -> :func:`cal_prototypes`, :func:`cal_prototypes_separate`
* ``STiLModel.py:374-381, 408-415``     -> :func:`accumulate`, :func:`finalize`
* ``STiLModel.py:171-180``              -> :func:`distribution_alignment`
* ``simmatch_model.py:268-286``         -> :func:`simmatch_bank`
* ``MMatch.py:215-230``                 -> :func:`mmatch_bank`
(``STiLModel.py`` = ``models/Disentangle/STiLModel.py``.)

Pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against the REFERENCE ITSELF, executed in the authoring
container by ``oracle/gen_golden.py`` — the real ``CLIPLoss`` / ``PrototypeLoss``
modules and the real ``STiLModel.training_step`` / ``cal_prototypes*`` bodies (run
on planted inputs with the Lightning scaffolding stubbed out) — whose outputs
are committed under ``tests/golden/`` and compared in
``tests/test_oracle_golden.py``.

All functions are device- and dtype-agnostic (fp32 like the reference, fp64 to
classify numerically ambiguous rows).  Op order deliberately mirrors the
reference wherever an index/mask decision depends on fp32 rounding.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- a1
def clip_loss(out0: Tensor, out1: Tensor, temperature: float, lambda_0: float = 0.5
              ) -> Tuple[Tensor, Tensor, Tensor]:
    """Symmetric InfoNCE. Follows utils/clip_loss.py:27-40 (ctor check :22-23)."""
    if lambda_0 > 1 or lambda_0 < 0:
        raise ValueError("lambda_0 must be a float between 0 and 1.")
    a = out0 / out0.norm(dim=1, keepdim=True).clamp_min(1e-12)       # :29 (F.normalize eps)
    b = out1 / out1.norm(dim=1, keepdim=True).clamp_min(1e-12)       # :30
    logits = (a @ b.t()) / temperature                                # :33
    n = out0.shape[0]
    labels = torch.arange(n, device=out0.device)                      # :34
    diag = logits.diagonal()
    ce_rows = (torch.logsumexp(logits, dim=1) - diag).mean()          # :36 CE(logits, arange)
    ce_cols = (torch.logsumexp(logits, dim=0) - diag).mean()          # :37 CE(logits.T, arange)
    loss = lambda_0 * ce_rows + (1 - lambda_0) * ce_cols              # :36-38
    return loss, logits, labels


def clip_loss_global(out0_parts, out1_parts, temperature: float, lambda_0: float = 0.5):
    """Oracle of the global-batch extension: the reference
    CLIPLoss applied to the concatenation of all ranks' rows in one process."""
    return clip_loss(torch.cat(list(out0_parts)), torch.cat(list(out1_parts)), temperature, lambda_0)


# --------------------------------------------------------------------------- a2
def cgpl(y_m: Tensor, y_i: Tensor, y_t: Tensor, prediction_override: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Consensus pseudo-labelling. Follows STiLModel.py:262-279 (+ :195-196, T=1).

    Inputs are the *teacher* logits of the unlabelled rows, [B_u, K].  `prediction_override` is the
    distribution-aligned softmax(y_m) of :276-277 (hparams.DA == True)."""
    p_m, p_i, p_t = (torch.softmax(y, dim=1) for y in (y_m, y_i, y_t))            # :262
    top_m, top_i, top_t = (p.argmax(dim=1) for p in (p_m, p_i, p_t))              # :263 (on probs)
    mi, mt = top_m == top_i, top_m == top_t
    case1 = mi & mt                                                               # :264
    case2_i = mi & ~mt                                                            # :265
    case2_t = mt & ~mi                                                            # :266
    case3 = ~(case1 | case2_i | case2_t)                                          # :267
    avg3 = torch.softmax((y_m + y_i + y_t) / 3.0, dim=1)                          # :270
    avg_mi = torch.softmax((y_m + y_i) / 2.0, dim=1)                              # :271
    avg_mt = torch.softmax((y_m + y_t) / 2.0, dim=1)                              # :272
    only_m = torch.softmax(y_m, dim=1)                                            # :273
    f = lambda m: m[:, None].to(y_m.dtype)
    pl_orig = f(case1) * avg3 + f(case2_i) * avg_mi + f(case2_t) * avg_mt + f(case3) * only_m   # :274
    prediction = torch.softmax(y_m, dim=1)                                        # :279 (DA False)
    if prediction_override is not None:
        prediction = prediction_override                                          # :276-277 (DA True)
    return dict(pseudo_label_orig=pl_orig, prediction=prediction,
                case1=case1, case2_i=case2_i, case2_t=case2_t, case3=case3,
                top1_m=top_m, top1_i=top_i, top1_t=top_t)


# --------------------------------------------------------------------------- a3
def pgls(feat_m_ue: Tensor, prototypes: Tensor, pseudo_label_orig: Tensor, prediction: Tensor,
         temperature: float, rate_pseudo: float, th1: float) -> Dict[str, Tensor]:
    """Prototype-guided label smoothing. Follows STiLModel.py:291-298."""
    teacher_logits = feat_m_ue @ prototypes.t()                                   # :293
    teacher_probs = torch.softmax(teacher_logits / temperature, dim=1)            # :294
    pseudo_label = rate_pseudo * pseudo_label_orig + (1 - rate_pseudo) * teacher_probs   # :295
    prediction = rate_pseudo * prediction + (1 - rate_pseudo) * teacher_probs     # :296
    max_prob, max_idx = prediction.max(dim=1)                                     # :297
    mask1 = max_prob >= th1                                                       # :298
    return dict(teacher_logits=teacher_logits, teacher_probs=teacher_probs, pseudo_label=pseudo_label,
                prediction=prediction, max_prob=max_prob, max_idx=max_idx, mask1=mask1)


def cgpl_pgls(y_m, y_i, y_t, feat_m_ue, prototypes, *, T, rate_pseudo, th1, prediction_override=None
              ) -> Dict[str, Tensor]:
    """a2 followed by a3 — the inline block STiLModel.py:262-298."""
    a = cgpl(y_m, y_i, y_t, prediction_override)
    b = pgls(feat_m_ue, prototypes, a["pseudo_label_orig"], a["prediction"], T, rate_pseudo, th1)
    out = dict(a)
    out.update(b)
    return out


# ---------------------------------------------------------------------- f-1
def masked_soft_ce(y_m_u: Tensor, y_i_u: Tensor, y_t_u: Tensor, pseudo_label: Tensor, mask1: Tensor,
                   case1: Tensor, case2_i: Tensor, case2_t: Tensor, case3: Tensor, mask_random: Tensor
                   ) -> Tuple[Tensor, Tensor, Tensor]:
    """Masked soft-target CE on the student logits. Follows STiLModel.py:301-303."""
    def soft_ce(y):                                         # F.cross_entropy(prob target, 'none')
        return -(pseudo_label * torch.log_softmax(y, dim=1)).sum(dim=1)
    f = lambda m: m.to(y_m_u.dtype)
    w_m = f(mask1) * f(case1)                                                       # :301
    w_i = f(mask1) * (f(case1) + f(case2_t) + f(case3) * f(mask_random))            # :302
    w_t = f(mask1) * (f(case1) + f(case2_i) + f(case3) * f(~mask_random))           # :303
    return (soft_ce(y_m_u) * w_m).mean(), (soft_ce(y_i_u) * w_i).mean(), (soft_ce(y_t_u) * w_t).mean()


def pseudo_label_all(y_l: Tensor, num_classes: int, prediction: Tensor, past_start_epoch: bool) -> Tensor:
    """One-hot labelled rows over (gated) `prediction`. Follows STiLModel.py:317-321."""
    if not past_start_epoch:
        prediction = torch.zeros_like(prediction)
    onehot = torch.nn.functional.one_hot(y_l, num_classes).to(prediction.dtype)
    return torch.cat((onehot, prediction), dim=0)


# --------------------------------------------------------------------------- a4
def prototype_loss(label: Tensor, prototypes: Tensor, feat: Tensor, temperature: float, threshold: float
                   ) -> Tensor:
    """PGLS prototype loss. Follows utils/prototype_loss.py:24-40.

    NB ``log(softmax + 1e-7)`` (not log_softmax) and ``.mean()`` over ALL rows."""
    p = torch.softmax((feat @ prototypes.t()) / temperature, dim=1)                # :26-27
    logp = torch.log(p + 1e-7)                                                     # :28
    max_prob, max_id = label.max(dim=1)                                            # :31
    conf = max_prob >= threshold                                                   # :32
    picked = logp.gather(1, max_id[:, None]).squeeze(1)                            # :34-37 (one-hot dot)
    return (-(picked) * conf.to(logp.dtype)).mean()                                # :37-39


# --------------------------------------------------------------------------- a5
def cal_prototypes(label: Tensor, feat: Tensor, th1: float) -> Tuple[Tensor, Tensor]:
    """Per-class sums / counts over confident rows. Follows STiLModel.py:199-214."""
    max_prob, max_id = label.max(dim=1)                                            # :204
    conf = max_prob >= th1                                                         # :205
    K = label.shape[1]
    w = conf.to(feat.dtype)
    class_sum = torch.zeros(K, feat.shape[1], dtype=feat.dtype, device=feat.device)
    class_sum.index_add_(0, max_id, feat * w[:, None])                             # :212 hard_labelᵀ @ feat
    class_count = torch.zeros(K, dtype=feat.dtype, device=feat.device).index_add_(0, max_id, w)   # :213
    return class_sum, class_count[:, None]


def cal_prototypes_separate(label: Tensor, feat: Tensor, b_l: int, th1: float, repeat_ratio: float
                            ) -> Tuple[Tensor, Tensor]:
    """Labelled rows down-weighted by repeat_ratio. Follows STiLModel.py:216-226."""
    ls, lc = cal_prototypes(label[:b_l], feat[:b_l], th1)
    us, uc = cal_prototypes(label[b_l:], feat[b_l:], th1)
    return ls / repeat_ratio + us, lc / repeat_ratio + uc                          # :224-225


def accumulate(prototypes_sum: Tensor, prototypes_count_sum: Tensor, class_sum: Tensor, class_count: Tensor
               ) -> None:
    """Running accumulators. Follows STiLModel.py:380-381 (all-reduce :377-379 happens before)."""
    prototypes_sum.add_(class_sum)
    prototypes_count_sum.add_(class_count)


def finalize(prototypes: Tensor, prototypes_sum: Tensor, prototypes_count_sum: Tensor) -> int:
    """Epoch-end replace + zero. Follows STiLModel.py:408-415. Returns #classes with count < 1
    (the reference asserts this is 0, :411-412)."""
    empty = int((prototypes_count_sum < 1).sum())
    prototypes.copy_(prototypes_sum / prototypes_count_sum)
    prototypes_sum.zero_()
    prototypes_count_sum.zero_()
    return empty


# --------------------------------------------------------------------------- a6
def distribution_alignment(probs: Tensor, da_queue: Tensor, da_ptr: Tensor, world_mean: Optional[Tensor] = None
                           ) -> Tensor:
    """Follows STiLModel.py:171-180; `world_mean` stands in for all_reduce(mean)/world_size."""
    mean = probs.mean(0) if world_mean is None else world_mean
    ptr = int(da_ptr)
    da_queue[ptr] = mean
    da_ptr[0] = (ptr + 1) % da_queue.shape[0]
    out = probs / da_queue.mean(0)
    return out / out.sum(dim=1, keepdim=True)


# --------------------------------------------------------------------------- a7
def simmatch_bank(feat_ku: Tensor, feat_qu: Tensor, prob_ku_orig: Tensor, bank_rows: Tensor, labels: Tensor,
                  tt: float, st: float, c_smooth: float) -> Dict[str, Tensor]:
    """SimMatch bank block. Follows models/MatchModel/simmatch_model.py:268-286.

    `bank_rows` is [K_b, D] row-major (the reference keeps the transpose [D, K_b], :68-69)."""
    K = prob_ku_orig.shape[1]
    teacher = torch.softmax((feat_ku @ bank_rows.t()) / tt, dim=1)                 # :270-271
    factor = prob_ku_orig.gather(1, labels[None, :].expand(feat_ku.shape[0], -1))  # :272
    teacher_f = teacher * factor                                                   # :273
    teacher_f = teacher_f / teacher_f.sum(dim=1, keepdim=True)                     # :274
    prob_ku = prob_ku_orig
    if c_smooth < 1:
        agg = torch.zeros(feat_ku.shape[0], K, dtype=teacher.dtype, device=teacher.device)
        agg.scatter_add_(1, labels[None, :].expand(feat_ku.shape[0], -1), teacher)  # :276-279 (un-reweighted)
        prob_ku = c_smooth * prob_ku_orig + (1 - c_smooth) * agg                   # :280
    student = torch.softmax((feat_qu @ bank_rows.t()) / st, dim=1)                 # :284-285
    loss_in = torch.sum(-teacher_f * torch.log(student), dim=1)                    # :286
    return dict(prob_ku=prob_ku, loss_in=loss_in, teacher=teacher_f)


# --------------------------------------------------------------------------- a9
def mmatch_bank(prob: Tensor, feat_u: Tensor, embed_rows: Tensor, probs_rows: Tensor, temperature: float,
                th1: float) -> Dict[str, Tensor]:
    """MMatch bank smoothing. Follows models/SemiMultimodal/MMatch.py:215-230.

    `embed_rows` [K_q, D], `probs_rows` [K_q, C] (reference keeps both transposed)."""
    a = torch.exp((feat_u @ embed_rows.t()) / temperature)
    a = a / a.sum(dim=1, keepdim=True)
    pseudo = 0.9 * prob + 0.1 * (a @ probs_rows)
    max_prob, max_idx = pseudo.max(dim=1)
    return dict(pseudo=pseudo, max_prob=max_prob, max_idx=max_idx, mask1=max_prob >= th1)


# ----------------------------------------------------------------- whole step
def head_step(batch: Dict[str, Tensor], cfg, *, dtype=torch.float32, with_grads: bool = True,
              state: Optional[Dict[str, Tensor]] = None, da_state: Optional[Dict[str, Tensor]] = None
              ) -> Dict[str, Tensor]:
    """The full hot path of STiLModel.training_step (lines 262-303, 317-322, 339, 374-381)
    on one synthetic batch (see stil_tta_b200/synth.py).  Embeddings stored in bf16 are
    upcast — the oracle computes on the *same values* in `dtype`.  `da_state` = {"DA_queue", "DA_ptr"}
    switches hparams.DA on (:276-277): the buffers are updated in place like the reference's."""
    up = lambda t: t.to(dtype)
    B_l = cfg.b_l
    feat_i = up(batch["feat_i"]).clone().requires_grad_(with_grads)
    feat_t = up(batch["feat_t"]).clone().requires_grad_(with_grads)
    feat_m = up(batch["feat_m"]).clone().requires_grad_(with_grads)
    y_s = [up(batch[k]).clone().requires_grad_(with_grads) for k in ("y_m", "y_i", "y_t")]
    feat_m_e = up(batch["feat_m_e"])
    protos = up(batch["prototypes"])
    with torch.no_grad():
        pred_da = None
        if da_state is not None:
            pred_da = distribution_alignment(torch.softmax(up(batch["y_m_ue"]), dim=1), da_state["DA_queue"],
                                             da_state["DA_ptr"])
        pl = cgpl_pgls(up(batch["y_m_ue"]), up(batch["y_i_ue"]), up(batch["y_t_ue"]), feat_m_e[B_l:], protos,
                       T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1, prediction_override=pred_da)
        label_all = pseudo_label_all(batch["y_l"], cfg.num_classes, pl["prediction"], cfg.past_start_epoch)
    l_m, l_i, l_t = masked_soft_ce(y_s[0][B_l:], y_s[1][B_l:], y_s[2][B_l:], pl["pseudo_label"], pl["mask1"],
                                   pl["case1"], pl["case2_i"], pl["case2_t"], pl["case3"], batch["mask_random"])
    loss_itc, logits, _ = clip_loss(feat_i, feat_t, cfg.temperature, cfg.lambda_0)
    loss_pt = prototype_loss(label_all, protos, feat_m, cfg.temperature, cfg.th1)
    out = dict(pl)
    out.update(loss_itc=loss_itc.detach(), loss_pt=loss_pt.detach(), logits=logits.detach(),
               loss_m_u=l_m.detach(), loss_i_u=l_i.detach(), loss_t_u=l_t.detach(), label_all=label_all)
    if with_grads:
        # three independent backward passes so each gradient can be checked separately
        g_i, g_t = torch.autograd.grad(loss_itc, (feat_i, feat_t))
        (g_m,) = torch.autograd.grad(loss_pt, (feat_m,))
        g_y = torch.autograd.grad(l_m + l_i + l_t, y_s)
        out.update(d_feat_i=g_i, d_feat_t=g_t, d_feat_m=g_m, d_y_m=g_y[0], d_y_i=g_y[1], d_y_t=g_y[2])
    with torch.no_grad():
        cs, cc = cal_prototypes_separate(label_all, feat_m_e, B_l, cfg.th1, cfg.repeat_ratio)
        out.update(class_sum=cs, class_count=cc)
        if state is not None:
            accumulate(state["prototypes_sum"], state["prototypes_count_sum"], cs, cc)
    return out


def ambiguous_rows(batch: Dict[str, Tensor], cfg, tol: float = 1e-5, da_state=None) -> Tensor:
    """Rows whose index/mask decisions are not determined at fp32 resolution:
    |max_prob - th1| < tol, or a top-2 gap < tol in any of the four argmaxes.  Evaluated in fp64.
    Exact ties (gap == 0) are NOT ambiguous: the first-index rule decides them."""
    if da_state is not None:
        da_state = {"DA_queue": da_state["DA_queue"].to(torch.float64), "DA_ptr": da_state["DA_ptr"].clone()}
    o = head_step(batch, cfg, dtype=torch.float64, with_grads=False, da_state=da_state)
    up = lambda t: t.to(torch.float64)

    def near_tie(p):
        top2 = p.topk(min(2, p.shape[1]), dim=1).values
        if top2.shape[1] < 2:
            return torch.zeros(p.shape[0], dtype=torch.bool)
        gap = top2[:, 0] - top2[:, 1]
        return (gap < tol) & (gap > 0)

    amb = (o["max_prob"] - cfg.th1).abs() < tol
    amb |= near_tie(o["prediction"])
    for k in ("y_m_ue", "y_i_ue", "y_t_ue"):
        amb |= near_tie(torch.softmax(up(batch[k]), dim=1))
    return amb


# ---------------------------------------------------------------------- f-4
def momentum_update_ema(main_state: Dict[str, Tensor], ema_state: Dict[str, Tensor], momentum: float, eman: bool,
                        param_names=None) -> None:
    """EMA teacher update, in place on `ema_state`. Follows STiLModel.py:154-168: with `eman` every state-dict
    entry (``num_batches_tracked`` copied, :163-164); otherwise only the parameters (`param_names`, :167-168)."""
    for k, v_main in main_state.items():
        v_ema = ema_state[k]
        if eman:
            if "num_batches_tracked" in k:
                v_ema.copy_(v_main)                                                    # :164
            else:
                v_ema.mul_(momentum).add_((1.0 - momentum) * v_main)                  # :166
        elif param_names is None or k in param_names:
            v_ema.mul_(momentum).add_((1.0 - momentum) * v_main)                      # :168


# ------------------------------------------------------------------------------------------------------------------
# FreeMatch self-adaptive threshold / fairness loss, CoTraining cross pseudo labels
# ------------------------------------------------------------------------------------------------------------------
def freematch_masking(state: Dict[str, Tensor], logits_x_ulb: Tensor, m: float = 0.999, clip_thresh: float = 0.0,
                      softmax_x_ulb: bool = True) -> Dict[str, Tensor]:
    """``FreeMatchModel.masking`` incl. its ``update`` — models/MatchModel/FreeMatchFolder/freematch_model.py:128-165.
    ``state`` = {time_p, p_model, label_hist} is updated IN PLACE like the reference's attributes."""
    probs = torch.softmax(logits_x_ulb.detach(), dim=-1) if softmax_x_ulb else logits_x_ulb.detach()      # :153-157
    max_probs, max_idx = torch.max(probs, dim=-1, keepdim=True)                                           # :132
    state["time_p"] = state["time_p"] * m + (1 - m) * max_probs.mean()                                    # :137
    if clip_thresh:
        state["time_p"] = torch.clip(state["time_p"], 0.0, 0.95)                                          # :139-140
    state["p_model"] = state["p_model"] * m + (1 - m) * probs.mean(dim=0)                                 # :142
    hist = torch.bincount(max_idx.reshape(-1), minlength=state["p_model"].shape[0]).to(state["p_model"].dtype)
    state["label_hist"] = state["label_hist"] * m + (1 - m) * (hist / hist.sum())                         # :143-144
    max_probs, max_idx = probs.max(dim=-1)                                                                # :161
    mod = state["p_model"] / torch.max(state["p_model"], dim=-1)[0]                                       # :162
    thr = state["time_p"] * mod[max_idx]
    return {"mask": max_probs.ge(thr).to(max_probs.dtype), "max_probs": max_probs, "max_idx": max_idx, "thr": thr, "probs": probs}


def freematch_entropy_loss(mask: Tensor, logits_s: Tensor, prob_model: Tensor, label_hist: Tensor) -> Tuple[Tensor, Tensor]:
    """``entropy_loss`` — FreeMatchFolder/freematch_utils.py:17-45."""
    def inf_to_zero(v):                                                                                   # :12-14
        v = v.clone()
        v[v == float("inf")] = 0.0
        return v
    logits_s = logits_s[mask.bool()]                                                                      # :18-21
    prob_s = logits_s.softmax(dim=-1)
    _, pred = torch.max(prob_s, dim=-1)
    hist_s = torch.bincount(pred, minlength=logits_s.shape[1]).to(logits_s.dtype)
    hist_s = hist_s / hist_s.sum()                                                                        # :26-27
    mod_prob_model = prob_model.reshape(1, -1) * inf_to_zero(1 / label_hist.reshape(1, -1)).detach()     # :30-34
    mod_prob_model = mod_prob_model / mod_prob_model.sum(dim=-1, keepdim=True)
    mod_mean = prob_s.mean(dim=0, keepdim=True) * inf_to_zero(1 / hist_s).detach()                        # :38-41
    mod_mean = mod_mean / mod_mean.sum(dim=-1, keepdim=True)
    loss = (mod_prob_model * torch.log(mod_mean + 1e-12)).sum(dim=1)                                      # :43-44
    return loss.mean(), hist_s.mean()


def cotraining_unsup(y_hat_i_u: Tensor, y_hat_t_u: Tensor, y_hat_i_e_u: Tensor, y_hat_t_e_u: Tensor, threshold: float
                     ) -> Dict[str, Tensor]:
    """The unsupervised part of ``CoTraining.training_step`` — models/SemiMultimodal/CoTraining.py:141-149 (arguments are the
    unlabelled rows ``[B_l:]`` of the student and teacher logits)."""
    pl_i = torch.softmax(y_hat_i_e_u.detach(), dim=1)                                                     # :141
    pl_t = torch.softmax(y_hat_t_e_u.detach(), dim=1)                                                     # :142
    max_i, _ = torch.max(pl_i, dim=1)
    max_t, _ = torch.max(pl_t, dim=1)
    mask_i, mask_t = max_i.ge(threshold), max_t.ge(threshold)                                             # :145-146
    loss_i_u = (F.cross_entropy(y_hat_i_u, pl_t, reduction="none") * mask_t).mean()                       # :148
    loss_t_u = (F.cross_entropy(y_hat_t_u, pl_i, reduction="none") * mask_i).mean()                       # :149
    return {"pseudo_label_i": pl_i, "pseudo_label_t": pl_t, "mask_i": mask_i, "mask_t": mask_t, "max_prob_i": max_i,
            "max_prob_t": max_t, "loss_i_u": loss_i_u, "loss_t_u": loss_t_u}
