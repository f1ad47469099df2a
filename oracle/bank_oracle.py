"""CPU oracle for the memory-bank baselines' head blocks — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

Independently written restatement, in plain PyTorch ops, of (paths relative to kgutjahr/STiL-TTA):

* ``models/MatchModel/simmatch_model.py:141-147``   -> :func:`update_bank`
* ``models/MatchModel/SimMatch.py:86-92``           -> :func:`masked_soft_ce_single`
* ``models/MatchModel/comatch_model.py:271-285``    -> :func:`hist_alignment`
* ``models/MatchModel/comatch_model.py:288-293``, ``models/SemiMultimodal/MMatch.py:222-227`` -> :func:`bank_smooth`
* ``models/MatchModel/comatch_model.py:298-312``    -> :func:`comatch_graphs`
* ``models/MatchModel/comatch_model.py:117-146``, ``MMatch.py:102-117`` -> :func:`queue_enqueue`
* ``models/MatchModel/CoMatch.py:92-110``           -> :func:`comatch_losses`
* ``models/SemiMultimodal/MMatch.py:215-235``       -> :func:`mmatch_block`

Pinning: ``oracle/gen_golden_banks.py`` executes the real ``SimMatchModel.forward``, ``CoMatchModel.forward``,
``CoMatch.training_step`` and ``MMatch.training_step`` on planted inputs in the authoring container; the outputs are
committed as ``tests/golden/bank_*.npz`` and compared in ``tests/test_oracle_banks.py``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor


def update_bank(bank: Tensor, labels: Tensor, k: Tensor, y: Tensor, index: Tensor) -> None:
    """simmatch_model.py:141-147 (single process): bank is [dim, K]."""
    bank[:, index] = k.t()
    labels[index] = y


def masked_soft_ce_single(logits: Tensor, target: Tensor, mask: Tensor) -> Tensor:
    """SimMatch.py:91 / CoMatch.py:96-97: mean_i( -sum_c log_softmax(logits)_ic * target_ic * mask_i )."""
    return (torch.sum(-torch.log_softmax(logits, dim=1) * target, dim=1) * mask.to(logits.dtype)).mean()


def hist_alignment(probs: Tensor, hist_prob: Tensor) -> Tensor:
    """comatch_model.py:271-285: `hist_prob` [n<=128, C] already contains this batch's mean as its last row."""
    probs = probs / hist_prob.mean(0)
    return probs / probs.sum(dim=1, keepdim=True)


def bank_smooth(probs: Tensor, feat: Tensor, queue_feat: Tensor, queue_probs: Tensor, temperature: float,
                keep: float) -> Tensor:
    """`keep*probs + (1-keep) * rownorm(exp(feat @ queue_feat / T)) @ queue_probs.T`
    (comatch_model.py:288-293 with keep = alpha; MMatch.py:222-227 with the literals 0.9 / 0.1).
    queue_feat [dim, K_q], queue_probs [C, K_q] (reference layouts)."""
    a = torch.exp(torch.mm(feat, queue_feat) / temperature)
    a = a / a.sum(1, keepdim=True)
    return keep * probs + (1 - keep) * torch.mm(a, queue_probs.t())


def comatch_graphs(probs: Tensor, probs_u: Tensor, feat_s0: Tensor, feat_s1: Tensor, queue_s: Tensor,
                   temperature: float) -> Tuple[Tensor, Tensor]:
    """Pseudo-label graph Q and embedding graph sim, comatch_model.py:298-312."""
    q_self = torch.mm(probs, probs.t())
    q_self.fill_diagonal_(1)
    q = torch.cat([q_self, torch.mm(probs, probs_u)], dim=1)
    sim = torch.cat([torch.exp(torch.mm(feat_s0, feat_s1.t()) / temperature),
                     torch.exp(torch.mm(feat_s0, queue_s) / temperature)], dim=1)
    return q, sim


def queue_enqueue(queue_feat: Tensor, queue_probs: Tensor, ptr: Tensor, z: Tensor, t: Tensor) -> None:
    """FIFO write with truncation at the wrap point (comatch_model.py:117-146, MMatch.py:102-117)."""
    k_q = queue_feat.shape[1]
    n = z.shape[0]
    p = int(ptr)
    if p + n > k_q:
        n = k_q - p
    queue_feat[:, p:p + n] = z[:n].t()
    queue_probs[:, p:p + n] = t[:n].t()
    ptr[0] = (p + n) % k_q


def graph_contrast_loss(q: Tensor, sim: Tensor, contrast_th: float) -> Tuple[Tensor, Tensor]:
    """CoMatch.py:100-110 -> (loss_contrast, pos_mask)."""
    pos_mask = q >= contrast_th
    q_mask = q * pos_mask
    q_mask = q_mask / q_mask.sum(1, keepdim=True)
    positives = sim * pos_mask
    pos_probs = positives / sim.sum(1, keepdim=True)
    log_probs = torch.log(pos_probs + 1e-7) * pos_mask
    return (-(log_probs * q_mask).sum(1)).mean(), pos_mask


def comatch_block(outputs_u_w: Tensor, feat_u_w: Tensor, feat_s0: Tensor, feat_s1: Tensor, outputs_u_s0: Tensor,
                  hist_prob: Tensor, queue_w: Tensor, probs_xu: Tensor, queue_s: Tensor, probs_u: Tensor,
                  temperature: float, alpha: float, smooth: bool, thr: float, contrast_th: float) -> Dict[str, Tensor]:
    """comatch_model.py:268-312 + CoMatch.py:92-110 (no queue writes; `hist_prob` holds the earlier batch means)."""
    probs = torch.softmax(outputs_u_w, dim=1)
    hist = torch.cat([hist_prob, probs.mean(0, keepdim=True)])[-128:]
    probs = hist_alignment(probs, hist)
    probs_orig = probs.clone()
    if smooth:
        probs = bank_smooth(probs, feat_u_w, queue_w, probs_xu, temperature, alpha)
    q, sim = comatch_graphs(probs, probs_u, feat_s0, feat_s1, queue_s, temperature)
    scores, _ = probs.max(dim=1)
    mask = scores >= thr
    loss_u = masked_soft_ce_single(outputs_u_s0, probs, mask)
    loss_c, pos_mask = graph_contrast_loss(q, sim, contrast_th)
    return dict(probs_orig=probs_orig, probs=probs, Q=q, sim=sim, mask=mask, pos_mask=pos_mask, loss_u=loss_u,
                loss_contrast=loss_c, hist=hist)


def mmatch_block(pseudo_label_orig: Tensor, feat_m_u: Tensor, embed_queue: Tensor, probs_queue: Tensor,
                 y_i_u: Tensor, y_t_u: Tensor, temperature: float, th1: float, smooth: bool) -> Dict[str, Tensor]:
    """MMatch.py:215-235: bank smoothing, max/argmax, mask, hard-label CE on the two unimodal heads."""
    pseudo = pseudo_label_orig
    if smooth:
        # the reference writes the literals 0.9 / 0.1 (:227); 1 - 0.9 differs from 0.1 in the last bit
        a = torch.exp(torch.mm(feat_m_u, embed_queue) / temperature)
        a = a / a.sum(dim=1, keepdim=True)
        pseudo = 0.9 * pseudo_label_orig + 0.1 * torch.mm(a, probs_queue.t())
    max_prob, max_idx = torch.max(pseudo, dim=1)
    mask1 = max_prob >= th1
    ce = lambda y: torch.nn.functional.cross_entropy(y, max_idx, reduction="none")
    m = mask1.to(y_i_u.dtype)
    return dict(pseudo_label=pseudo, max_prob=max_prob, max_idx=max_idx, mask1=mask1,
                loss_i_u=(ce(y_i_u) * m).mean(), loss_t_u=(ce(y_t_u) * m).mean())


# --------------------------------------------------------------------------- f-2
def club_mean(mu: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """CLUBMean.forward and .learning_loss from mu = p_mu(x) on (models/Disentangle/utils/club.py:107-121, 125-130;
    called STiLModel.py:327-330): the MI upper bound with its B x B x D broadcast, and the q(y|x) learning loss."""
    positive = -(mu - y) ** 2 / 2.0
    negative = -((y.unsqueeze(0) - mu.unsqueeze(1)) ** 2).mean(dim=1) / 2.0
    bound = (positive.sum(dim=-1) - negative.sum(dim=-1)).mean()
    est = -(-(mu - y) ** 2).sum(dim=1).mean(dim=0)
    return bound, est


# ----------------------------------------------------------------------------------------------------------------------
# Column-sharded SimMatch bank: torch restatement of the three shard kernels (include/stil_head.h,
# stil_simmatch_shard_stats / _finish / _grad).  The REFERENCE of the sharded sweep is stil_head_oracle.simmatch_bank on
# the whole bank; these functions only let tests run the sharding SCHEDULE (additive fixed-shift statistics, all-reduce,
# reduce-scatter) on CPU ranks over gloo.
# ----------------------------------------------------------------------------------------------------------------------
def simmatch_shard_stats(fk: Tensor, fq: Tensor, p: Tensor, bank_shard: Tensor, labels: Tensor, tt: float, st: float) -> Tensor:
    """bank_shard [dim, k_shard].  stats[row] = [sum e_t | sum e_s | sum e_t p[y_j] z_s/st | A_c], e = exp((z - 1)/T)."""
    zt, zs = fk @ bank_shard, fq @ bank_shard
    et, es = torch.exp((zt - 1) / tt), torch.exp((zs - 1) / st)
    f = p[:, labels]
    agg = torch.zeros_like(p).index_add_(1, labels, et)
    return torch.cat((et.sum(1, keepdim=True), es.sum(1, keepdim=True), (et * f * zs / st).sum(1, keepdim=True), agg), dim=1)


def simmatch_shard_finish(stats: Tensor, p: Tensor, st: float, c_smooth: float):
    sum_t, sum_s, num, agg = stats[:, 0], stats[:, 1], stats[:, 2], stats[:, 3:]
    den = (p * agg).sum(1)
    prob_ku = c_smooth * p + (1 - c_smooth) * agg / sum_t[:, None] if c_smooth < 1 else p.clone()
    loss_in = 1.0 / st + torch.log(sum_s) - num / den
    return prob_ku, loss_in, torch.stack((1 / sum_s, 1 / den), dim=1)


def simmatch_shard_grad(fk: Tensor, fq: Tensor, p: Tensor, bank_shard: Tensor, labels: Tensor, tt: float, st: float,
                        norms: Tensor) -> Tensor:
    zt, zs = fk @ bank_shard, fq @ bank_shard
    g = (torch.exp((zs - 1) / st) * norms[:, :1] - torch.exp((zt - 1) / tt) * p[:, labels] * norms[:, 1:]) / st
    return g @ bank_shard.t()
