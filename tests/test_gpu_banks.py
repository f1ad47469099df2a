"""GPU parity of the memory-bank blocks (SURVEY §8 rows a7-a9 + queue maintenance) through the C ABI:
against the reference's own outputs (tests/golden/bank_*.npz, made from the executed reference forwards) and against
the pinned oracle at the baselines' full shapes.  Decisions (argmax, masks, graph edges) bit-exact on rows / entries
that are not within fp32 noise of their threshold; values within 1e-3 relative."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, REL, assert_rel, rel_err
from test_oracle_banks import CO_CASES, MM_CASES, SIM_CASES, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import stil_tta_b200 as S
    return S


@pytest.fixture(scope="module")
def BO():
    from oracle import bank_oracle as BO
    return BO


def dev(t):
    return t.cuda()


def padded_queue(S, q, dtype=torch.float32):
    out = S.alloc_bank(q.shape[0], q.shape[1], dtype)
    out.copy_(q)
    return out


# ------------------------------------------------------------------------------------------ a7 vs the reference
@pytest.mark.parametrize("name", SIM_CASES)
def test_simmatch_matches_reference_fixture(S, name):
    z = load(name)
    b_l, b_u, k_cls, dim, k_bank, da = (int(v) for v in z["meta"])
    p = torch.softmax(dev(z["logits_ku"]), dim=-1)
    if da:
        q, ptr = dev(z["da_queue_in"].clone()), dev(z["da_ptr_in"].clone())
        p = S.distribution_alignment(p, q, ptr)
        assert torch.equal(ptr.cpu(), z["da_ptr_out"])
        assert_rel(q, z["da_queue_out"], 1e-5, "DA_queue")
    assert_rel(p, z["prob_ku_orig"], 1e-5, "prob_ku_orig")
    bank = padded_queue(S, z["bank"])
    labels = dev(z["bank_labels"].clone())
    fq = dev(z["feat_q"][b_l:]).requires_grad_(True)
    prob_ku, loss_in = S.simmatch_bank(dev(z["feat_k"][b_l:]), fq, p, bank, labels, float(z["tt"]), float(z["st"]),
                                       float(z["c_smooth"]))
    assert float((prob_ku.cpu() - z["prob_ku"]).abs().max()) <= 2e-5
    assert_rel(loss_in, z["loss_in_rows"], REL, "loss_in rows")
    th = float(z["threshold"])
    mp = z["prob_ku"].max(1).values
    keep = (mp - th).abs() > 1e-5
    mask = prob_ku.max(1).values >= th
    assert torch.equal(mask.cpu()[keep].float(), z["mask"][keep])
    (g,) = torch.autograd.grad(loss_in.mean(), fq)
    assert_rel(g, z["d_feat_q"][b_l:], REL, "d_feat_q")
    lq = dev(z["logits_q"][b_l:]).requires_grad_(True)
    loss_u = S.masked_ce(lq, prob_ku, mask)                       # SimMatch.py:91
    assert_rel(loss_u, z["loss_u"], REL, "loss_u")
    assert_rel(torch.autograd.grad(loss_u, lq)[0], z["d_logits_q"][b_l:], REL, "d_logits_q")
    S.update_bank(bank, labels, dev(z["feat_k"][:b_l]), dev(z["y_l"]), dev(z["index"]))    # simmatch_model.py:141-147
    assert torch.equal(bank.cpu(), z["bank_out"]) and torch.equal(labels.cpu(), z["labels_out"])


# ------------------------------------------------------------------------------------------ a9 vs the reference
@pytest.mark.parametrize("name", MM_CASES)
def test_mmatch_matches_reference_fixture(S, name):
    z = load(name)
    b_l, b_u, k_cls, dim, k_q, ptr0, da, epoch = (int(v) for v in z["meta"])
    p = dev(z["pseudo_label_orig"])
    feat_m = dev(z["feat_m"])
    eq, pq = padded_queue(S, z["embed_queue"]), padded_queue(S, z["probs_queue"])
    out = S.mmatch_pseudo_label(p, feat_m[b_l:], eq, pq, float(z["T"]), float(z["th1"]), current_epoch=epoch)
    assert float((out.probs.cpu() - z["pseudo_label"].reshape(b_u, k_cls)).abs().max()) <= 5e-6
    assert_rel(out.max_prob, z["max_prob"], 1e-5, "max_prob")
    assert torch.equal(out.max_idx.cpu(), z["max_idx"])           # fixtures hold no ambiguous rows (checked below)
    assert torch.equal(out.mask.cpu(), z["mask1"])
    y_i, y_t = dev(z["y_i"][b_l:]).requires_grad_(True), dev(z["y_t"][b_l:]).requires_grad_(True)
    li, lt = S.masked_ce(y_i, out.max_idx, out.mask), S.masked_ce(y_t, out.max_idx, out.mask)    # MMatch.py:231-235
    assert_rel(li, z["loss_i_u"], REL, "loss_i_u"); assert_rel(lt, z["loss_t_u"], REL, "loss_t_u")
    g_i, g_t = torch.autograd.grad(li + lt, (y_i, y_t))
    assert_rel(g_i, z["d_y_i"][b_l:], REL, "d_y_i"); assert_rel(g_t, z["d_y_t"][b_l:], REL, "d_y_t")
    ptr = dev(torch.tensor([ptr0]))
    onehot = F.one_hot(dev(z["y_l"]), k_cls).float()
    S.queue_enqueue(eq, pq, ptr, feat_m, torch.cat([onehot, out.probs]))                          # MMatch.py:259
    assert int(ptr) == int(z["embed_queue_ptr_out"])
    assert torch.equal(eq.cpu(), z["embed_queue_out"])
    assert float((pq.cpu() - z["probs_queue_out"]).abs().max()) <= 5e-6


@pytest.mark.parametrize("name", MM_CASES)
def test_mmatch_fixture_has_no_ambiguous_rows(name):
    z = load(name)
    p = z["pseudo_label"].double().reshape(int(z["meta"][1]), -1)
    top2 = p.topk(min(2, p.shape[1]), dim=1).values
    assert ((top2[:, 0] - float(z["th1"])).abs() > 1e-5).all()
    if p.shape[1] > 1:
        assert ((top2[:, 0] - top2[:, 1]) > 1e-5).all()


# ------------------------------------------------------------------------------------------ a8 vs the reference
@pytest.mark.parametrize("name", CO_CASES)
def test_comatch_matches_reference_fixture(S, name):
    z = load(name)
    b_l, b_u, k_cls, dim, k_q, ptr_w, ptr_s, epoch = (int(v) for v in z["meta"])
    T, alpha, thr, cth = float(z["temperature"]), float(z["alpha"]), float(z["thr"]), float(z["contrast_th"])
    fm = dev(z["features_m"])
    # distribution alignment over the list of earlier batch means (comatch_model.py:271-285)
    da = S.HistAlignment(k_cls, "cuda")
    n_hist = z["hist_prob"].shape[0]
    da.hist[:n_hist] = dev(z["hist_prob"]); da.count[0] = n_hist
    probs_orig = da(torch.softmax(dev(z["outputs_m"][b_l:b_l + b_u]), dim=1))
    assert_rel(probs_orig, z["probs_orig"], 1e-5, "probs_orig")
    qw, pxu = padded_queue(S, z["queue_w"]), padded_queue(S, z["probs_xu"])
    qs, pu = padded_queue(S, z["queue_s"]), padded_queue(S, z["probs_u"])
    sm = S.comatch_smooth(probs_orig, fm[b_l:b_l + b_u], qw, pxu, T, alpha, thr, smooth=epoch > 0)
    assert float((sm.probs.cpu() - z["probs"]).abs().max()) <= 5e-6
    mp = z["probs"].max(1).values
    keep = (mp - thr).abs() > 1e-5
    assert torch.equal(sm.mask.cpu()[keep].float(), z["mask"][keep])
    feats = dev(z["features"]).requires_grad_(True)
    Q, sim = S.comatch_graphs(sm.probs, pu, feats[b_l:], fm[b_l + b_u:], qs, T)
    assert float((Q.cpu() - z["Q"]).abs().max()) <= 5e-6
    assert_rel(sim, z["sim"], REL, "sim")
    far = (z["Q"] - cth).abs() > 1e-5
    assert torch.equal((Q.cpu() >= cth)[far], z["pos_mask"][far])
    loss_c = S.graph_contrast_loss(Q, sim, cth)
    assert_rel(loss_c, z["loss_contrast"], REL, "loss_contrast")
    (g,) = torch.autograd.grad(loss_c, feats)
    assert_rel(g, z["d_features"], REL, "d_features")
    outs = dev(z["outputs"]).requires_grad_(True)
    loss_u = S.masked_ce(outs[b_l:], sm.probs, sm.mask)           # CoMatch.py:96-97
    assert_rel(loss_u, z["loss_u"], REL, "loss_u")
    assert_rel(torch.autograd.grad(loss_u, outs)[0], z["d_outputs"], REL, "d_outputs")
    # queue writes (comatch_model.py:315-321)
    p_s, p_w = dev(torch.tensor([ptr_s])), dev(torch.tensor([ptr_w]))
    S.queue_enqueue(qs, pu, p_s, fm[b_l + b_u:], sm.probs)
    onehot = F.one_hot(dev(z["labels_x"]), k_cls).float()
    S.queue_enqueue(qw, pxu, p_w, fm[:b_l + b_u], torch.cat([onehot, probs_orig]))
    assert int(p_s) == int(z["queue_ptr_s_out"]) and int(p_w) == int(z["queue_ptr_w_out"])
    assert torch.equal(qs.cpu(), z["queue_s_out"]) and torch.equal(qw.cpu(), z["queue_w_out"])
    assert float((pu.cpu() - z["probs_u_out"]).abs().max()) <= 5e-6
    assert float((pxu.cpu() - z["probs_xu_out"]).abs().max()) <= 5e-6


# ------------------------------------------------------------------------------------------ full shapes vs the oracle
def _planted_probs(g, rows, k):
    c = torch.randint(0, k, (rows,), generator=g)
    y = torch.randn(rows, k, generator=g)
    y[torch.arange(rows), c] += torch.tensor([4.0, 8.0, 12.0])[torch.randint(0, 3, (rows,), generator=g)]
    return torch.softmax(y, 1), c


@pytest.mark.parametrize("rows,k,d,kq,dtype", [
    (448, 286, 128, 640, torch.float32), (448, 286, 128, 2560, torch.bfloat16), (896, 2, 128, 640, torch.float32),
    (37, 10, 64, 100, torch.float32), (448, 286, 512, 2560, torch.bfloat16),
])
def test_bank_smooth_matches_oracle(S, BO, rows, k, d, kq, dtype):
    g = torch.Generator().manual_seed(rows + kq)
    centers = F.normalize(torch.randn(k, d, generator=g))
    p, c = _planted_probs(g, rows, k)
    qc = torch.randint(0, k, (kq,), generator=g)
    queue = F.normalize(centers[qc] + 0.6 * torch.randn(kq, d, generator=g)).t().contiguous().to(dtype)
    qprobs = torch.softmax(6.0 * F.one_hot(qc, k).float() + torch.randn(kq, k, generator=g), 1).t().contiguous()
    feat = F.normalize(centers[c] + 0.5 * torch.randn(rows, d, generator=g)).to(dtype)
    ref = BO.bank_smooth(p, feat.float(), queue.float(), qprobs, 0.1, 0.9)
    ref64 = BO.bank_smooth(p.double(), feat.double(), queue.double(), qprobs.double(), 0.1, 0.9)
    out = S.bank_smooth(dev(p), dev(feat), padded_queue(S, queue, dtype), padded_queue(S, qprobs), 0.1, 0.9, 1 - 0.9, 0.9)
    assert float((out.probs.cpu() - ref).abs().max()) <= 1e-5
    top2 = ref64.topk(2, dim=1).values
    clear = ((top2[:, 0] - 0.9).abs() > 1e-5) & ((top2[:, 0] - top2[:, 1]) > 1e-5)
    assert int((~clear).sum()) <= 2
    assert torch.equal(out.max_idx.cpu()[clear], ref.max(1).indices[clear])
    assert torch.equal(out.mask.cpu()[clear], (ref.max(1).values >= 0.9)[clear])
    # the outputs are self-consistent on every row
    assert torch.equal(out.max_prob, out.probs.max(1).values) and torch.equal(out.max_idx, out.probs.argmax(1))
    assert torch.equal(out.mask, out.max_prob >= 0.9)


@pytest.mark.parametrize("rows,k,d,kq,dtype", [
    (448, 286, 128, 2560, torch.float32), (448, 286, 128, 2560, torch.bfloat16), (56, 10, 64, 96, torch.float32),
    (130, 2, 128, 640, torch.bfloat16),
])
def test_comatch_graphs_and_loss_match_oracle(S, BO, rows, k, d, kq, dtype):
    g = torch.Generator().manual_seed(rows * 3 + kq)
    centers = F.normalize(torch.randn(k, d, generator=g))
    probs, c = _planted_probs(g, rows, k)
    qc = torch.randint(0, k, (kq,), generator=g)
    queue_s = F.normalize(centers[qc] + 0.6 * torch.randn(kq, d, generator=g)).t().contiguous().to(dtype)
    probs_u = torch.softmax(6.0 * F.one_hot(qc, k).float() + torch.randn(kq, k, generator=g), 1).t().contiguous()
    f1 = F.normalize(centers[c] + 0.5 * torch.randn(rows, d, generator=g)).to(dtype)
    f0 = F.normalize(f1.float() + 0.2 * torch.randn(rows, d, generator=g)).to(dtype)
    f0r = f0.float().requires_grad_(True)
    Qr, simr = BO.comatch_graphs(probs, probs_u, f0r, f1.float(), queue_s.float(), 0.1)
    lr, pos_r = BO.graph_contrast_loss(Qr, simr, 0.8)
    (gr,) = torch.autograd.grad(lr, f0r)
    f0c = dev(f0).requires_grad_(True)
    Q, sim = S.comatch_graphs(dev(probs), padded_queue(S, probs_u), f0c, dev(f1), padded_queue(S, queue_s, dtype), 0.1)
    assert float((Q.cpu() - Qr).abs().max()) <= 5e-6
    assert_rel(sim, simr.detach(), REL, "sim")
    far = (Qr - 0.8).abs() > 1e-5
    assert torch.equal((Q.cpu() >= 0.8)[far], pos_r[far])
    loss = S.graph_contrast_loss(Q, sim, 0.8)
    assert_rel(loss, lr.detach(), REL, "loss_contrast")
    (gc,) = torch.autograd.grad(loss, f0c)
    assert_rel(gc, gr, REL, "d_feat_s0")


def test_masked_ce_variants(S, BO):
    g = torch.Generator().manual_seed(9)
    for rows, k in ((448, 286), (896, 2), (5, 7)):
        y = torch.randn(rows, k, generator=g) * 3
        t = torch.softmax(torch.randn(rows, k, generator=g) * 2, 1)
        m = torch.rand(rows, generator=g) > 0.4
        yr = y.clone().requires_grad_(True)
        lr = BO.masked_soft_ce_single(yr, t, m)
        (gr,) = torch.autograd.grad(lr, yr)
        yc = dev(y).requires_grad_(True)
        l = S.masked_ce(yc, dev(t), dev(m))
        assert_rel(l, lr.detach(), REL, "soft ce")
        assert_rel(torch.autograd.grad(l, yc)[0], gr, REL, "d soft ce")
        idx = torch.randint(0, k, (rows,), generator=g)
        yr = y.clone().requires_grad_(True)
        lr = (F.cross_entropy(yr, idx, reduction="none") * m).mean()
        (gr,) = torch.autograd.grad(lr, yr)
        yc = dev(y).requires_grad_(True)
        l = S.masked_ce(yc, dev(idx), dev(m))
        assert_rel(l, lr.detach(), REL, "hard ce")
        assert_rel(torch.autograd.grad(l, yc)[0], gr, REL, "d hard ce")


def test_queue_enqueue_wraps_like_the_reference(S, BO):
    g = torch.Generator().manual_seed(11)
    d, c, kq = 64, 10, 200
    q_ref, p_ref, ptr_ref = torch.randn(d, kq, generator=g), torch.rand(c, kq, generator=g), torch.tensor([0])
    q, p, ptr = dev(q_ref.clone()), dev(p_ref.clone()), dev(ptr_ref.clone())
    for n in (64, 64, 64, 64, 30, 200, 7):          # 4th write is truncated at the wrap point, then restarts at 0
        z, t = torch.randn(n, d, generator=g), torch.rand(n, c, generator=g)
        BO.queue_enqueue(q_ref, p_ref, ptr_ref, z, t)
        S.queue_enqueue(q, p, ptr, dev(z), dev(t))
        assert int(ptr) == int(ptr_ref)
        assert torch.equal(q.cpu(), q_ref) and torch.equal(p.cpu(), p_ref)


# ------------------------------------------------------------------------------------------ f-2 CLUBMean
from test_oracle_banks import CLUB_CASES  # noqa: E402


@pytest.mark.parametrize("name", CLUB_CASES)
def test_club_matches_reference_fixture(S, name):
    z = load(name)
    mu, y = dev(z["mu"]).requires_grad_(True), dev(z["y"]).requires_grad_(True)
    bound, est = S.club_both(mu, y)
    assert abs(float(bound) - float(z["bound"])) <= REL * abs(float(z["bound"])) + 2e-5
    assert_rel(est, z["est"], REL, "learning_loss")
    gb = torch.autograd.grad(bound, (mu, y), retain_graph=True)
    ge = torch.autograd.grad(est, (mu, y))
    assert_rel(gb[0], z["d_mu_bound"], REL, "d_mu bound"); assert_rel(gb[1], z["d_y_bound"], REL, "d_y bound")
    assert_rel(ge[0], z["d_mu_est"], REL, "d_mu est"); assert_rel(ge[1], z["d_y_est"], REL, "d_y est")


def test_club_module_matches_oracle_at_full_size(S, BO):
    """B = D = 512 (the reference's broadcast temporary would be 512 MB): the drop-in module against the closed form in
    fp64 and, through autograd, the parameter gradients of p_mu."""
    torch.manual_seed(5)
    m = S.CLUBMean(512, 512, 512).cuda()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(512, 512, generator=g).cuda()
    y = (0.5 * x.cpu() + torch.randn(512, 512, generator=g)).cuda().requires_grad_(True)
    loss = m(x, y) + m.learning_loss(x, y)
    loss.backward()
    mu64 = m.p_mu(x).detach().double().cpu().requires_grad_(True)
    y64 = y.detach().double().cpu().requires_grad_(True)
    b = 512
    ref = (mu64 * y64).sum() / b - (mu64.sum(0) * y64.sum(0)).sum() / b ** 2 + ((mu64 - y64) ** 2).sum() / b
    ref.backward()
    assert abs(float(loss) - float(ref)) <= REL * abs(float(ref))
    assert_rel(y.grad, y64.grad, REL, "d_y")
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


# ----------------------------------------------------------------------------------------------- forward -> update -> backward
def test_simmatch_backward_after_bank_update(S):
    """The reference overwrites bank columns right after the bank block and BEFORE loss.backward()
    (simmatch_model.py:291; it reads a bank.clone(), :237).  The drop-in must give the gradient of the bank the forward
    saw, whatever update_bank did in between (raw kernels do not move torch's version counter)."""
    from oracle import stil_head_oracle as O
    g = torch.Generator().manual_seed(21)
    rows, kb, d, c = 96, 1024, 128, 10
    bank_rows = F.normalize(torch.randn(kb, d, generator=g)).to(torch.bfloat16)
    labels = torch.randint(0, c, (kb,), generator=g)
    fk = F.normalize(bank_rows.float()[torch.randint(0, kb, (rows,), generator=g)] + 0.3 * torch.randn(rows, d, generator=g)).to(torch.bfloat16)
    fq = F.normalize(fk.float() + 0.2 * torch.randn(rows, d, generator=g)).to(torch.bfloat16)
    p = torch.softmax(torch.randn(rows, c, generator=g) * 3, 1)
    fqr = fq.float().requires_grad_(True)
    ref = O.simmatch_bank(fk.float(), fqr, p, bank_rows.float(), labels, 0.1, 0.1, 0.9)
    (g_ref,) = torch.autograd.grad(ref["loss_in"].mean(), fqr)
    bank = S.alloc_bank(d, kb, torch.bfloat16)
    bank.copy_(bank_rows.t())
    lab = dev(labels)
    fqc = dev(fq).requires_grad_(True)
    prob_ku, loss_in = S.simmatch_bank(dev(fk), fqc, dev(p), bank, lab, 0.1, 0.1, 0.9)
    # _update_bank with the current batch's teacher features (the reference order), hitting HALF of the bank
    idx = torch.randperm(kb, generator=g)[:kb // 2]
    newk = F.normalize(torch.randn(kb // 2, d, generator=g))
    S.update_bank(bank, lab, dev(newk), dev(torch.randint(0, c, (kb // 2,), generator=g)), dev(idx))
    assert not torch.equal(bank.cpu().float(), bank_rows.t().float())
    (g_c,) = torch.autograd.grad(loss_in.mean(), fqc)
    assert_rel(g_c, g_ref, REL, "d_feat_qu after update_bank")


def test_comatch_backward_after_queue_enqueue(S, BO):
    """Same ordering hazard for CoMatch: _dequeue_and_enqueue overwrites queue_s between the forward and
    loss.backward(); the reference passes queue_s.clone().detach() (comatch_model.py:310)."""
    g = torch.Generator().manual_seed(22)
    rows, k, d, kq = 64, 10, 64, 128
    probs, c = _planted_probs(g, rows, k)
    queue_s = F.normalize(torch.randn(kq, d, generator=g)).t().contiguous()
    probs_u = torch.softmax(torch.randn(kq, k, generator=g), 1).t().contiguous()
    f1 = F.normalize(torch.randn(rows, d, generator=g))
    f0 = F.normalize(f1 + 0.2 * torch.randn(rows, d, generator=g))
    f0r = f0.clone().requires_grad_(True)
    Qr, simr = BO.comatch_graphs(probs, probs_u, f0r, f1, queue_s, 0.1)
    lr, _ = BO.graph_contrast_loss(Qr, simr, 0.8)
    (gr,) = torch.autograd.grad(lr, f0r)
    f0c = dev(f0).requires_grad_(True)
    qs, pu = padded_queue(S, queue_s), padded_queue(S, probs_u)
    Q, sim = S.comatch_graphs(dev(probs), pu, f0c, dev(f1), qs, 0.1)
    loss = S.graph_contrast_loss(Q, sim, 0.8)
    ptr_ = torch.zeros(1, dtype=torch.int64, device="cuda")
    S.queue_enqueue(qs, pu, ptr_, dev(F.normalize(torch.randn(rows, d, generator=g))), dev(probs))   # overwrites columns 0..rows-1
    (gc,) = torch.autograd.grad(loss, f0c)
    assert_rel(gc, gr, REL, "d_feat_s0 after queue_enqueue")


# ----------------------------------------------------------------------------------------------- column-sharded bank (e-C5)
@pytest.mark.parametrize("rows,kb,d,c,dtype,shards", [
    (96, 1024, 128, 10, torch.bfloat16, 1), (96, 1024, 128, 10, torch.bfloat16, 4), (64, 640, 64, 10, torch.float32, 2),
    (448, 4096, 128, 286, torch.bfloat16, 8), (448, 65536, 512, 286, torch.bfloat16, 8),      # the last one is BASELINE config C5
    # ragged shapes of the persistent sweep kernels (csrc/bank_sweep.cu): partial row blocks, shard widths that are not a
    # multiple of 64 / 128 columns (two-box TMA path, clipped bulk stores), every supported dim, several chunks per CTA
    (130, 1000, 256, 7, torch.bfloat16, 1), (64, 200, 384, 5, torch.bfloat16, 1), (300, 2112, 512, 10, torch.bfloat16, 2),
    (1, 136, 128, 3, torch.bfloat16, 1), (2100, 1024, 128, 10, torch.bfloat16, 1),
    (96, 1024, 192, 10, torch.bfloat16, 1),                                                   # dim % 128 != 0: tiled-GEMM path
])
def test_sharded_simmatch_bank_matches_whole_bank_oracle(S, rows, kb, d, c, dtype, shards):
    """The column-sharded sweep (fixed-shift additive statistics; `shards` emulated one after the other on this GPU — the same
    kernels, column offsets and sums as `shards` ranks) against the oracle on the WHOLE bank, incl. the C5 shape."""
    from oracle import stil_head_oracle as O
    g = torch.Generator().manual_seed(rows + kb + shards)
    unit = F.normalize
    bank_rows = unit(torch.randn(kb, d, generator=g)).to(dtype)
    labels = torch.randint(0, c, (kb,), generator=g)
    fk = unit(bank_rows.float()[torch.randint(0, kb, (rows,), generator=g)] + 0.3 * torch.randn(rows, d, generator=g)).to(dtype)
    fq = unit(fk.float() + 0.2 * torch.randn(rows, d, generator=g)).to(dtype)
    p = torch.softmax(torch.randn(rows, c, generator=g) * 3, 1)
    fqr = fq.float().requires_grad_(True)
    ref = O.simmatch_bank(fk.float(), fqr, p, bank_rows.float(), labels, 0.1, 0.1, 0.9)
    (g_ref,) = torch.autograd.grad(ref["loss_in"].mean(), fqr)
    sb = S.ShardedSimMatchBank(d, kb, c, dtype=dtype, device="cuda", emulate_shards=shards)
    sb.load(bank_rows, labels)
    fqc = dev(fq).requires_grad_(True)
    prob_ku, loss_in = sb(dev(fk), fqc, dev(p), 0.1, 0.1, 0.9)
    # the reference order: bank columns are overwritten before loss.backward() (simmatch_model.py:291)
    idx = torch.randperm(kb, generator=g)[:64]
    sb.update(dev(unit(torch.randn(64, d, generator=g))), dev(torch.randint(0, c, (64,), generator=g)), dev(idx))
    (g_c,) = torch.autograd.grad(loss_in.mean(), fqc)
    assert float((prob_ku.cpu() - ref["prob_ku"]).abs().max()) <= 2e-5
    assert_rel(loss_in, ref["loss_in"].detach(), REL, "loss_in (sharded)")
    assert_rel(g_c, g_ref, REL, "d_feat_qu (sharded)")
    mp_ref = ref["prob_ku"].max(1).values
    keep = (mp_ref - 0.95).abs() > 1e-4
    assert torch.equal((prob_ku.cpu().max(1).values >= 0.95)[keep], (mp_ref >= 0.95)[keep])
    # the update reached the owning shards: the next sweep sees the new columns
    whole = torch.cat([b.float().cpu() for b in sb.bank], dim=1)
    assert float((whole[:, idx].t().norm(dim=1) - 1).abs().max()) < 1e-2
    assert not torch.equal(whole.t().to(dtype), bank_rows)


@pytest.mark.parametrize("rows,kb,d,c", [(96, 1024, 128, 10), (448, 8192, 512, 286)])
def test_sharded_simmatch_bank_cuda_graph_replays(S, rows, kb, d, c):
    """use_graph=True: the captured sweep gives the eager sweep's results on NEW inputs, and sees bank updates made between
    replays (the graph reads the bank and label buffers in place)."""
    g = torch.Generator().manual_seed(7 * rows + kb)
    unit = F.normalize
    bank_rows = unit(torch.randn(kb, d, generator=g)).to(torch.bfloat16)
    labels = torch.randint(0, c, (kb,), generator=g)
    eager = S.ShardedSimMatchBank(d, kb, c, device="cuda")
    graphed = S.ShardedSimMatchBank(d, kb, c, device="cuda", use_graph=True)
    for sb in (eager, graphed):
        sb.load(bank_rows, labels)
    for it in range(3):
        fk = dev(unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16))
        fq = unit(torch.randn(rows, d, generator=g)).to(torch.bfloat16)
        p = dev(torch.softmax(torch.randn(rows, c, generator=g) * 3, 1))
        outs = []
        for sb in (eager, graphed):
            fqc = dev(fq.float()).requires_grad_(True)      # fp32 leaf: the gradient comes back unrounded
            prob_ku, loss_in = sb(fk, fqc, p, 0.1, 0.1, 0.9)
            (gq,) = torch.autograd.grad(loss_in.mean(), fqc)
            outs.append((prob_ku, loss_in, gq))
        assert len(graphed._graphs) == 1
        assert float((outs[0][0] - outs[1][0]).abs().max()) <= 1e-6
        assert_rel(outs[1][1], outs[0][1].detach().cpu(), 1e-5, "loss_in (graph replay vs eager)")
        # the dX partials are reduced in arrival order (bulk fp32 reductions): equal up to summation order
        assert_rel(outs[1][2], outs[0][2].cpu(), 1e-4, "d_feat_qu (graph replay vs eager)")
        idx = dev(torch.randperm(kb, generator=g)[:32])
        k_new, y_new = dev(unit(torch.randn(32, d, generator=g))), dev(torch.randint(0, c, (32,), generator=g))
        for sb in (eager, graphed):
            sb.update(k_new, y_new, idx)
    with pytest.raises(ValueError):                 # a class id outside [0, C) is refused where it enters the bank
        eager.update(k_new, torch.full_like(y_new, c), idx)
    # without a gradient: a second signature, a second graph
    with torch.no_grad():
        pe, le = eager(fk, dev(fq), p, 0.1, 0.1, 0.9)
        pg, lg = graphed(fk, dev(fq), p, 0.1, 0.1, 0.9)
    assert len(graphed._graphs) == 2
    assert_rel(lg, le.cpu(), 1e-5, "loss_in (no-grad graph vs eager)")
