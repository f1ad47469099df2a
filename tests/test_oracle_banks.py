"""Pin the bank-block oracles (rows a7-a9) against the reference's own outputs: tests/golden/bank_*.npz were made by
oracle/gen_golden_banks.py from the real SimMatchModel.forward / CoMatchModel.forward / CoMatch.training_step /
MMatch.training_step.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import bank_oracle as BO
from oracle import stil_head_oracle as O

SIM_CASES = ["bank_simmatch_k10", "bank_simmatch_k286", "bank_simmatch_nosmooth"]
CO_CASES = ["bank_comatch_k10", "bank_comatch_wrap", "bank_comatch_pre_start"]
MM_CASES = ["bank_mmatch_k10", "bank_mmatch_k286_wrap", "bank_mmatch_epoch0"]


def load(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: torch.from_numpy(np.array(z[k])) for k in z.files}


def close(a, b, rtol=2e-5, atol=1e-6):
    torch.testing.assert_close(a.to(torch.float32), b.to(torch.float32).reshape(a.shape), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", SIM_CASES)
def test_simmatch_oracle_matches_reference(name):
    z = load(name)
    torch.set_num_threads(1)
    b_l, b_u, k_cls, dim, k_bank, da = (int(v) for v in z["meta"])
    p = torch.softmax(z["logits_ku"], dim=-1)
    da_queue, da_ptr = z["da_queue_in"].clone(), z["da_ptr_in"].clone()
    if da:
        p = O.distribution_alignment(p, da_queue, da_ptr)
    close(p, z["prob_ku_orig"])
    fq = z["feat_q"][b_l:].clone().requires_grad_(True)
    o = O.simmatch_bank(z["feat_k"][b_l:], fq, p, z["bank"].t().contiguous(), z["bank_labels"], float(z["tt"]),
                        float(z["st"]), float(z["c_smooth"]))
    close(o["prob_ku"], z["prob_ku"])
    close(o["loss_in"], z["loss_in_rows"], rtol=1e-4, atol=1e-5)
    mask = o["prob_ku"].max(dim=-1)[0] >= float(z["threshold"])
    assert torch.equal(mask.float(), z["mask"])
    (g,) = torch.autograd.grad(o["loss_in"].mean(), (fq,))
    close(g, z["d_feat_q"][b_l:], rtol=1e-4, atol=1e-6)
    lq = z["logits_q"][b_l:].clone().requires_grad_(True)
    loss_u = BO.masked_soft_ce_single(lq, o["prob_ku"], mask)
    close(loss_u, z["loss_u"])
    close(torch.autograd.grad(loss_u, (lq,))[0], z["d_logits_q"][b_l:])
    bank, labels = z["bank"].clone(), z["bank_labels"].clone()
    BO.update_bank(bank, labels, z["feat_k"][:b_l], z["y_l"], z["index"])
    assert torch.equal(bank, z["bank_out"]) and torch.equal(labels, z["labels_out"])
    assert torch.equal(da_ptr, z["da_ptr_out"])
    close(da_queue, z["da_queue_out"])


@pytest.mark.parametrize("name", CO_CASES)
def test_comatch_oracle_matches_reference(name):
    z = load(name)
    torch.set_num_threads(1)
    b_l, b_u, k_cls, dim, k_q, ptr_w, ptr_s, epoch = (int(v) for v in z["meta"])
    feats = z["features"].clone().requires_grad_(True)
    outs = z["outputs"].clone().requires_grad_(True)
    fm = z["features_m"]
    o = BO.comatch_block(z["outputs_m"][b_l:b_l + b_u], fm[b_l:b_l + b_u], feats[b_l:], fm[b_l + b_u:], outs[b_l:],
                         z["hist_prob"], z["queue_w"], z["probs_xu"], z["queue_s"], z["probs_u"],
                         float(z["temperature"]), float(z["alpha"]), epoch > 0, float(z["thr"]), float(z["contrast_th"]))
    close(o["probs_orig"], z["probs_orig"])
    close(o["probs"], z["probs"])
    close(o["Q"], z["Q"])
    close(o["sim"], z["sim"], rtol=1e-4)
    assert torch.equal(o["mask"].float(), z["mask"])
    assert torch.equal(o["pos_mask"], z["pos_mask"])
    close(o["loss_u"], z["loss_u"])
    close(o["loss_contrast"], z["loss_contrast"], rtol=1e-4)
    close(torch.autograd.grad(o["loss_contrast"], (feats,), retain_graph=True)[0], z["d_features"], rtol=1e-4)
    close(torch.autograd.grad(o["loss_u"], (outs,))[0], z["d_outputs"])
    # queue writes (comatch_model.py:315-321): strong queue gets (features_u_s1, probs), weak gets (feature_xu_w, [onehot; probs_orig])
    qs, pu, ps = z["queue_s"].clone(), z["probs_u"].clone(), torch.tensor([ptr_s])
    BO.queue_enqueue(qs, pu, ps, fm[b_l + b_u:], o["probs"].detach())
    qw, pxu, pw = z["queue_w"].clone(), z["probs_xu"].clone(), torch.tensor([ptr_w])
    onehot = torch.nn.functional.one_hot(z["labels_x"], k_cls).float()
    BO.queue_enqueue(qw, pxu, pw, fm[:b_l + b_u], torch.cat([onehot, o["probs_orig"]]))
    assert torch.equal(qs, z["queue_s_out"]) and torch.equal(qw, z["queue_w_out"])
    close(pu, z["probs_u_out"]); close(pxu, z["probs_xu_out"])
    assert int(ps) == int(z["queue_ptr_s_out"]) and int(pw) == int(z["queue_ptr_w_out"])


@pytest.mark.parametrize("name", MM_CASES)
def test_mmatch_oracle_matches_reference(name):
    z = load(name)
    torch.set_num_threads(1)
    b_l, b_u, k_cls, dim, k_q, ptr, da, epoch = (int(v) for v in z["meta"])
    p = torch.softmax(z["y_m"][b_l:], dim=1)
    da_queue, da_ptr = z["da_queue_in"].clone(), z["da_ptr_in"].clone()
    if da:
        p = O.distribution_alignment(p, da_queue, da_ptr)
    close(p, z["pseudo_label_orig"])
    feat_m = torch.nn.functional.normalize(z["x_m"], dim=1)
    y_i = z["y_i"].clone().requires_grad_(True)
    y_t = z["y_t"].clone().requires_grad_(True)
    o = BO.mmatch_block(p, feat_m[b_l:], z["embed_queue"], z["probs_queue"], y_i[b_l:], y_t[b_l:], float(z["T"]),
                        float(z["th1"]), epoch > 0)
    ref_pl = z["pseudo_label"]
    close(o["pseudo_label"] if k_cls != 2 else o["pseudo_label"], ref_pl)
    assert torch.equal(o["max_idx"], z["max_idx"]) and torch.equal(o["mask1"], z["mask1"])
    close(o["max_prob"], z["max_prob"])
    close(o["loss_i_u"], z["loss_i_u"]); close(o["loss_t_u"], z["loss_t_u"])
    g_i, g_t = torch.autograd.grad(o["loss_i_u"] + o["loss_t_u"], (y_i, y_t))
    close(g_i, z["d_y_i"]); close(g_t, z["d_y_t"])
    # queue write (MMatch.py:259): all rows of feat_m with [onehot(y_l); pseudo_label]
    q, pq, pp = z["embed_queue"].clone(), z["probs_queue"].clone(), torch.tensor([ptr])
    onehot = torch.nn.functional.one_hot(z["y_l"], k_cls).float()
    BO.queue_enqueue(q, pq, pp, feat_m, torch.cat([onehot, o["pseudo_label"]]))
    close(q, z["embed_queue_out"]); close(pq, z["probs_queue_out"])
    assert int(pp) == int(z["embed_queue_ptr_out"])


CLUB_CASES = ["club_b64_d128", "club_b56_d512", "club_b37_d24"]


@pytest.mark.parametrize("name", CLUB_CASES)
def test_club_oracle_matches_reference(name):
    z = load(name)
    torch.set_num_threads(1)
    mu, y = z["mu"].clone().requires_grad_(True), z["y"].clone().requires_grad_(True)
    bound, est = BO.club_mean(mu, y)
    close(bound, z["bound"], rtol=1e-4, atol=1e-5)
    close(est, z["est"])
    gb = torch.autograd.grad(bound, (mu, y), retain_graph=True)
    ge = torch.autograd.grad(est, (mu, y))
    close(gb[0], z["d_mu_bound"], rtol=1e-4); close(gb[1], z["d_y_bound"], rtol=1e-4)
    close(ge[0], z["d_mu_est"]); close(ge[1], z["d_y_est"])
    # the closed form the CUDA path uses (SURVEY f-2): bound = sum_i mu_i.y_i / B - (sum_i mu_i).(sum_j y_j) / B^2
    b = mu.shape[0]
    closed = (mu * y).sum() / b - (mu.sum(0) * y.sum(0)).sum() / b ** 2
    close(closed, z["bound"], rtol=1e-4, atol=2e-5)
