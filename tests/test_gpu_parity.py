"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.

Bars (BASELINE.json north_star): pseudo-label indices and masks BIT-EXACT; losses and gradients within
1e-3 relative (fp32 accumulate).  Gradients are compared with max-abs error relative to the largest
reference entry (grad tensors span many orders of magnitude)."""
import numpy as np
import pytest
import torch

from conftest import DA_CASES, GOLDEN, REL, STEP_CASES, assert_rel, cfg_for, da_state_of, load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import stil_tta_b200 as S
    return S


@pytest.fixture(scope="module")
def O():
    from oracle import stil_head_oracle as O
    return O


def dev(t):
    return t.cuda()


# ----------------------------------------------------------------------------------------------- a1
@pytest.mark.parametrize("n,d,dtype,T,lam", [
    (64, 128, torch.float32, 0.1, 0.5), (512, 128, torch.bfloat16, 0.1, 0.5), (200, 64, torch.float32, 0.07, 0.3),
    (37, 24, torch.float32, 0.1, 0.5), (129, 128, torch.bfloat16, 0.1, 1.0), (1024, 128, torch.bfloat16, 0.1, 0.0),
    (300, 512, torch.bfloat16, 0.2, 0.5), (96, 2048, torch.float32, 0.1, 0.5), (1, 8, torch.float32, 0.1, 0.5),
])
def test_clip_loss_matches_oracle(S, O, n, d, dtype, T, lam):
    g = torch.Generator().manual_seed(n * 7 + d)
    a = torch.randn(n, d, generator=g).to(dtype)
    b = (a.float() * 0.7 + 0.7 * torch.randn(n, d, generator=g)).to(dtype)
    ar, br = a.float().requires_grad_(True), b.float().requires_grad_(True)
    loss_r, logits_r, labels_r = O.clip_loss(ar, br, T, lam)
    ga_r, gb_r = torch.autograd.grad(loss_r, (ar, br))
    ac, bc = dev(a).requires_grad_(True), dev(b).requires_grad_(True)
    loss, logits, labels = S.CLIPLoss(T, lam)(ac, bc)
    ga, gb = torch.autograd.grad(loss, (ac, bc))
    assert abs(float(loss) - float(loss_r)) <= REL * abs(float(loss_r)) + 5e-5   # abs floor: |logit| <= 1/T = 10
    assert torch.equal(labels.cpu(), labels_r)
    assert float((logits.cpu() - logits_r).abs().max()) <= 2e-3     # |logit| <= 1/T
    # bf16 inputs: the gradient is formed in fp32 and rounded once to the input dtype (assert_rel allows that rounding)
    assert_rel(ga, ga_r, REL, "d_out0")
    assert_rel(gb, gb_r, REL, "d_out1")


def test_clip_loss_golden_modules(S):
    z = np.load(GOLDEN / "modules.npz")
    t = lambda k: torch.from_numpy(np.array(z[k]))
    for tag in "abcd":
        a, b = dev(t(f"clip_{tag}_a")).requires_grad_(True), dev(t(f"clip_{tag}_b")).requires_grad_(True)
        loss, logits, labels = S.CLIPLoss(float(z[f"clip_{tag}_T"]), float(z[f"clip_{tag}_lam"]))(a, b)
        ga, gb = torch.autograd.grad(loss, (a, b))
        ref = float(z[f"clip_{tag}_loss"])
        assert abs(float(loss) - ref) <= REL * abs(ref) + 1e-6
        assert float((logits.cpu() - t(f"clip_{tag}_logits")).abs().max()) <= 2e-3 / float(z[f"clip_{tag}_T"]) * 0.1 + 1e-3
        assert_rel(ga, t(f"clip_{tag}_ga"), REL, "ga")
        assert_rel(gb, t(f"clip_{tag}_gb"), REL, "gb")


def test_clip_loss_upstream_grad_and_no_logits(S, O):
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(96, 128, generator=g), torch.randn(96, 128, generator=g)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    (3.0 * O.clip_loss(ar, br, 0.1, 0.5)[0]).backward()
    ac, bc = dev(a).requires_grad_(True), dev(b).requires_grad_(True)
    loss, logits, _ = S.CLIPLoss(0.1, 0.5, return_logits=False)(ac, bc)
    assert logits is None
    (3.0 * loss).backward()
    assert_rel(ac.grad, ar.grad, REL, "scaled grad")
    assert_rel(bc.grad, br.grad, REL, "scaled grad")


# ----------------------------------------------------------------------------------------------- a4
@pytest.mark.parametrize("n,k,d,dtype,T,th", [
    (64, 286, 128, torch.float32, 0.1, 0.9), (512, 286, 128, torch.bfloat16, 0.1, 0.9),
    (1024, 2, 128, torch.bfloat16, 0.1, 0.85), (50, 10, 64, torch.float32, 0.5, 0.3), (3, 2, 128, torch.float32, 0.1, 0.9),
    (130, 700, 256, torch.bfloat16, 0.1, 0.5),
])
def test_prototype_loss_matches_oracle(S, O, n, k, d, dtype, T, th):
    g = torch.Generator().manual_seed(n + k)
    label = torch.softmax(torch.randn(n, k, generator=g) * 6, dim=1)
    protos = torch.randn(k, d, generator=g) * 0.08
    feat = torch.nn.functional.normalize(torch.randn(n, d, generator=g)).to(dtype)
    fr = feat.float().requires_grad_(True)
    loss_r = O.prototype_loss(label, protos, fr, T, th)
    (g_r,) = torch.autograd.grad(loss_r, fr)
    fc = dev(feat).requires_grad_(True)
    loss = S.PrototypeLoss(T, th)(dev(label), dev(protos), fc)
    (gc,) = torch.autograd.grad(loss, fc)
    assert abs(float(loss) - float(loss_r)) <= REL * abs(float(loss_r)) + 5e-5   # abs floor: |logit| <= 1/T = 10
    assert_rel(gc, g_r, REL, "d_feat")


def test_prototype_loss_golden_and_smoke_input(S):
    z = np.load(GOLDEN / "modules.npz")
    t = lambda k: torch.from_numpy(np.array(z[k]))
    # the reference's own __main__ input: integer labels with an all-zero row (prototype_loss.py:42-48)
    feat = dev(t("pt_smoke_feat")).requires_grad_(True)
    loss = S.PrototypeLoss(0.1, 0.9)(dev(t("pt_smoke_label")), dev(t("pt_smoke_protos")), feat)
    (gf,) = torch.autograd.grad(loss, feat)
    ref = float(z["pt_smoke_loss"])
    assert abs(float(loss) - ref) <= REL * abs(ref) + 1e-6
    assert_rel(gf, t("pt_smoke_gfeat"), REL, "smoke grad")
    for tag in "abc":
        feat = dev(t(f"pt_{tag}_feat")).requires_grad_(True)
        loss = S.PrototypeLoss(float(z[f"pt_{tag}_T"]), float(z[f"pt_{tag}_th"]))(
            dev(t(f"pt_{tag}_label")), dev(t(f"pt_{tag}_protos")), feat)
        (gf,) = torch.autograd.grad(loss, feat)
        ref = float(z[f"pt_{tag}_loss"])
        assert abs(float(loss) - ref) <= REL * abs(ref) + 1e-6
        assert_rel(gf, t(f"pt_{tag}_gfeat"), REL, f"pt_{tag} grad")


# ----------------------------------------------------------------------------------------------- a2 + a3
DECISIONS = ("max_idx", "mask1", "case1", "case2_i", "case2_t", "case3")


@pytest.mark.parametrize("name", STEP_CASES)
def test_cgpl_pgls_golden_bit_exact(S, name):
    ins, ref, meta = load_golden(name)
    cfg = cfg_for(name, meta)
    B_l = cfg.b_l
    out = S.cgpl_pgls(dev(ins["y_m_ue"]), dev(ins["y_i_ue"]), dev(ins["y_t_ue"]), dev(ins["feat_m_e"][B_l:]),
                      dev(ins["prototypes"]), T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1)
    for k in DECISIONS:
        assert torch.equal(getattr(out, k).cpu(), ref[k]), k                     # bit-exact
    for j, k in enumerate(("top1_m", "top1_i", "top1_t")):
        assert torch.equal(out.top1[j].cpu(), ref[k]), k
    pl_ref = ref["pseudo_label"]
    pl = out.pseudo_label.cpu()
    if cfg.num_classes == 2:
        pl = pl[:, 1]                                                             # STiLModel.py:349-354
    assert float((pl - pl_ref).abs().max()) <= 2e-6
    assert float((out.max_prob.cpu() - ref["max_prob"]).abs().max()) <= 2e-6


@pytest.mark.parametrize("cfg_name,batch,seed", [("dvm", 512, 2022), ("dvm", 512, 11), ("cardiac", 1024, 2022),
                                                 ("dvm", 4096, 5), ("cardiac", 256, 9)])
def test_cgpl_pgls_matches_oracle_full_size(S, O, cfg_name, batch, seed):
    from stil_tta_b200 import synth
    cfg = synth.dvm_config(batch) if cfg_name == "dvm" else synth.cardiac_config(batch)
    b = synth.make_batch(cfg, seed=seed, edge_rows=True)
    B_l = cfg.b_l
    o = O.cgpl_pgls(b["y_m_ue"], b["y_i_ue"], b["y_t_ue"], b["feat_m_e"][B_l:].float(), b["prototypes"],
                    T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1)
    amb = O.ambiguous_rows(b, cfg)
    out = S.cgpl_pgls(dev(b["y_m_ue"]), dev(b["y_i_ue"]), dev(b["y_t_ue"]), dev(b["feat_m_e"][B_l:]),
                      dev(b["prototypes"]), T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1)
    keep = ~amb
    assert int(amb.sum()) <= 2, "planted batches should have (almost) no ambiguous rows"
    for k in DECISIONS:
        assert torch.equal(getattr(out, k).cpu()[keep], o[k][keep]), k
    assert float((out.pseudo_label.cpu() - o["pseudo_label"]).abs().max()) <= 2e-6
    assert float((out.prediction.cpu() - o["prediction"]).abs().max()) <= 2e-6
    # structural properties (size independent): cases partition rows, rows of pseudo_label sum to 1
    s = out.case1.int() + out.case2_i.int() + out.case2_t.int() + out.case3.int()
    assert torch.all(s == 1)
    assert float((out.pseudo_label.sum(1) - 1).abs().max()) <= 1e-5


def test_cgpl_pgls_bf16_logits_and_odd_k(S, O):
    g = torch.Generator().manual_seed(3)
    for k, rows in ((7, 33), (286, 100), (1000, 17), (33, 64)):
        ys = [(torch.randn(rows, k, generator=g) * 3).to(torch.bfloat16) for _ in range(3)]
        feat = torch.nn.functional.normalize(torch.randn(rows, 64, generator=g))
        protos = torch.randn(k, 64, generator=g) * 0.2
        o = O.cgpl_pgls(*[y.float() for y in ys], feat, protos, T=0.1, rate_pseudo=0.9, th1=0.5)
        out = S.cgpl_pgls(*[dev(y) for y in ys], dev(feat), dev(protos), T=0.1, rate_pseudo=0.9, th1=0.5)
        assert float((out.pseudo_label.cpu() - o["pseudo_label"]).abs().max()) <= 5e-6
        agree = (out.max_idx.cpu() == o["max_idx"]).float().mean()
        assert agree >= 0.99      # bf16 logits tie often; ties resolved identically unless fp32-ambiguous


# ----------------------------------------------------------------------------------------------- a7
@pytest.mark.parametrize("rows,kb,d,c,dtype,cs", [
    (56, 706, 128, 286, torch.bfloat16, 0.9), (448, 4096, 128, 286, torch.bfloat16, 0.9),
    (64, 640, 64, 10, torch.float32, 0.9), (33, 2560, 128, 2, torch.bfloat16, 1.0), (128, 16384, 512, 286, torch.bfloat16, 0.9),
])
def test_simmatch_bank_matches_oracle(S, O, rows, kb, d, c, dtype, cs):
    g = torch.Generator().manual_seed(rows + kb)
    unit = torch.nn.functional.normalize
    bank_rows = unit(torch.randn(kb, d, generator=g)).to(dtype)           # [K_b, D]; the module keeps the transpose
    labels = torch.randint(0, c, (kb,), generator=g)
    fk = unit(bank_rows.float()[torch.randint(0, kb, (rows,), generator=g)] + 0.3 * torch.randn(rows, d, generator=g)).to(dtype)
    fq = unit(fk.float() + 0.2 * torch.randn(rows, d, generator=g)).to(dtype)
    p = torch.softmax(torch.randn(rows, c, generator=g) * 3, 1)
    fqr = fq.float().requires_grad_(True)
    ref = O.simmatch_bank(fk.float(), fqr, p, bank_rows.float(), labels, 0.1, 0.1, cs)
    (g_ref,) = torch.autograd.grad(ref["loss_in"].mean(), fqr)
    bank = S.alloc_bank(d, kb, dtype)                                      # reference layout [D, K_b]
    bank.copy_(bank_rows.t())
    fqc = dev(fq).requires_grad_(True)
    prob_ku, loss_in = S.simmatch_bank(dev(fk), fqc, dev(p), bank, dev(labels), 0.1, 0.1, cs)
    (g_c,) = torch.autograd.grad(loss_in.mean(), fqc)
    assert float((prob_ku.cpu() - ref["prob_ku"]).abs().max()) <= 2e-5
    assert_rel(loss_in, ref["loss_in"].detach(), REL, "loss_in")
    assert_rel(g_c, g_ref, REL, "d_feat_qu")
    # consumer decision (SimMatch.py:87-88) on non-ambiguous rows
    mp_ref = ref["prob_ku"].max(1).values
    keep = (mp_ref - 0.95).abs() > 1e-4
    assert torch.equal((prob_ku.cpu().max(1).values >= 0.95)[keep], (mp_ref >= 0.95)[keep])


# ----------------------------------------------------------------------------------------------- a6
def test_distribution_alignment_matches_oracle(S, O):
    g = torch.Generator().manual_seed(4)
    for k, rows in ((286, 448), (2, 896), (10, 33)):
        q_ref, p_ref = torch.zeros(256, k), torch.zeros(1, dtype=torch.int64)
        q, p = dev(q_ref.clone()), dev(p_ref.clone())
        p_ref[0] = p[0] = 254                       # exercise the ring-buffer wrap-around
        for step in range(4):
            probs = torch.softmax(torch.randn(rows, k, generator=g) * 2, dim=1)
            ref = O.distribution_alignment(probs, q_ref, p_ref)
            out = S.distribution_alignment(dev(probs), q, p)
            assert int(p) == int(p_ref) == (254 + step + 1) % 256
            assert float((q.cpu() - q_ref).abs().max()) <= 1e-7
            assert float(((out.cpu() - ref).abs() / ref.abs().clamp_min(1e-6)).max()) <= 1e-4
            assert float((out.sum(1) - 1).abs().max()) <= 1e-5


# ----------------------------------------------------------------------------------------------- a5
@pytest.mark.parametrize("name", STEP_CASES)
def test_cal_prototypes_separate_golden(S, name):
    ins, ref, meta = load_golden(name)
    cfg = cfg_for(name, meta)
    label_all = dev(ref["pseudo_label_all"])
    cs, cc = S.cal_prototypes_separate(label_all, dev(ins["feat_m_e"]), cfg.b_l, cfg.th1, cfg.repeat_ratio)
    assert float((cs.cpu() - ref["class_sum"]).abs().max()) <= 1e-5
    assert float((cc.cpu() - ref["class_count"]).abs().max()) <= 1e-6


def test_prototype_bank_accumulate_finalize(S, O):
    from stil_tta_b200 import synth
    cfg = synth.cardiac_config(256)
    bank = S.PrototypeBank(cfg.num_classes, cfg.proj_dim, cfg.th1, cfg.repeat_ratio).cuda()
    ps, pc, pr = torch.zeros(2, 128), torch.zeros(2, 1), torch.zeros(2, 128)
    for step in range(3):
        b = synth.make_batch(cfg, seed=100 + step)
        o = O.head_step(b, cfg, with_grads=False, state={"prototypes_sum": ps, "prototypes_count_sum": pc})
        bank.update(dev(o["label_all"]), dev(b["feat_m_e"]), cfg.b_l)
    assert float((bank.prototypes_sum.cpu() - ps).abs().max()) <= 1e-4
    assert float((bank.prototypes_count_sum.cpu() - pc).abs().max()) <= 1e-5
    empty_ref = O.finalize(pr, ps, pc)
    empty = bank.finalize()
    assert int(empty) == empty_ref
    assert float((bank.prototypes.cpu() - pr).abs().max()) <= 1e-6
    assert float(bank.prototypes_sum.abs().max()) == 0 and float(bank.prototypes_count_sum.abs().max()) == 0
    # empty-class flag instead of the reference's host assert (STiLModel.py:411-412)
    bank2 = S.PrototypeBank(5, 8, 0.5).cuda()
    bank2.prototypes_count_sum[:3] = 2.0
    assert int(bank2.finalize()) == 2


def test_cal_prototypes_edge_cases(S, O):
    # no confident row at all / all rows confident in one class / B_l == 0 and B_l == B
    k, d = 6, 16
    feat = torch.randn(40, d)
    label = torch.full((40, k), 1.0 / k)
    cs, cc = S.cal_prototypes(dev(label), dev(feat), 0.9)
    assert float(cs.abs().max()) == 0 and float(cc.abs().max()) == 0
    label = torch.zeros(40, k)
    label[:, 4] = 1.0
    for b_l in (0, 13, 40):
        cs, cc = S.cal_prototypes_separate(dev(label), dev(feat), b_l, 0.9, 3.0)
        rs, rc = O.cal_prototypes_separate(label, feat, b_l, 0.9, 3.0)
        assert float((cs.cpu() - rs).abs().max()) <= 1e-5 and float((cc.cpu() - rc).abs().max()) <= 1e-6


# ----------------------------------------------------------------------------------------------- f-1
@pytest.mark.parametrize("name", STEP_CASES)
def test_masked_soft_ce_golden(S, name):
    ins, ref, meta = load_golden(name)
    cfg = cfg_for(name, meta)
    if cfg.num_classes == 2:
        pytest.skip("golden pseudo_label is column 1 only for binary tasks (STiLModel.py:349-354); covered below")
    B_l = cfg.b_l
    ys = [dev(ins[k][B_l:]).requires_grad_(True) for k in ("y_m", "y_i", "y_t")]
    flags = [dev(ref[k]) for k in ("mask1", "case1", "case2_i", "case2_t", "case3", "mask_random")]
    losses = S.masked_soft_ce(*ys, dev(ref["pseudo_label"]), *flags)
    for l, k in zip(losses, ("loss_m_u", "loss_i_u", "loss_t_u")):
        assert abs(float(l) - float(ref[k])) <= REL * abs(float(ref[k])) + 1e-7, k
    grads = torch.autograd.grad(losses[0] + losses[1] + losses[2], ys)
    for gct, k in zip(grads, ("d_y_m", "d_y_i", "d_y_t")):
        r = ref[k][B_l:] if ref[k].numel() > 1 else torch.zeros_like(gct.cpu())
        if float(r.abs().max()) == 0:
            assert float(gct.abs().max()) == 0
        else:
            assert_rel(gct, r, REL, k)


def test_masked_soft_ce_vs_oracle_binary_and_scaled(S, O):
    from stil_tta_b200 import synth
    cfg = synth.cardiac_config(512)
    b = synth.make_batch(cfg, seed=77)
    B_l = cfg.b_l
    o = O.head_step(b, cfg)
    ys = [dev(b[k][B_l:]).requires_grad_(True) for k in ("y_m", "y_i", "y_t")]
    flags = [dev(o[k]) for k in ("mask1", "case1", "case2_i", "case2_t", "case3")] + [dev(b["mask_random"])]
    losses = S.masked_soft_ce(*ys, dev(o["pseudo_label"]), *flags)
    for l, k in zip(losses, ("loss_m_u", "loss_i_u", "loss_t_u")):
        assert abs(float(l) - float(o[k])) <= REL * abs(float(o[k])) + 1e-7
    (0.2 * (losses[0] + losses[1] + losses[2])).backward()
    for y, k in zip(ys, ("d_y_m", "d_y_i", "d_y_t")):
        assert_rel(y.grad, 0.2 * o[k][B_l:], REL, k)


# ----------------------------------------------------------------------------------------------- whole step
def run_head(S, cfg, batch, use_graph):
    head = S.STiLHead(cfg, device="cuda", use_graph=use_graph)
    head.load(batch)
    head.run()
    torch.cuda.synchronize()
    return head


def check_step(head, o, cfg, amb, grad_tol):
    out = head.out
    keep = ~amb
    for k in DECISIONS:
        assert torch.equal(out[k].cpu()[keep], o[k][keep]), k
    assert float((out["pseudo_label"].cpu() - o["pseudo_label"]).abs().max()) <= 2e-6
    L = out["losses"].cpu()
    for j, k in enumerate(("loss_itc", "loss_pt", "loss_m_u", "loss_i_u", "loss_t_u")):
        assert abs(float(L[j]) - float(o[k])) <= REL * abs(float(o[k])) + 1e-6, (k, float(L[j]), float(o[k]))
    for k in ("d_feat_i", "d_feat_t", "d_feat_m"):
        assert_rel(out[k], o[k], grad_tol, k)
    for k in ("d_y_m", "d_y_i", "d_y_t"):
        if float(o[k].abs().max()) > 0:
            assert_rel(out[k], o[k], REL, k)
    assert float((out["class_sum"].cpu() - o["class_sum"]).abs().max()) <= 1e-4
    assert float((out["class_count"].cpu() - o["class_count"]).abs().max()) <= 1e-5


@pytest.mark.parametrize("name", STEP_CASES)
def test_head_step_golden(S, O, name):
    ins, ref, meta = load_golden(name)
    cfg = cfg_for(name, meta)
    ins["mask_random"] = ref["mask_random"]
    o = O.head_step(ins, cfg)
    head = run_head(S, cfg, ins, use_graph=False)
    check_step(head, o, cfg, torch.zeros(cfg.b_u, dtype=torch.bool), REL)
    # and straight against the reference's recorded outputs
    for k in DECISIONS:
        assert torch.equal(head.out[k].cpu(), ref[k]), k
    assert abs(float(head.out["losses"][0]) - float(ref["loss_itc"])) <= REL * abs(float(ref["loss_itc"]))
    assert abs(float(head.out["losses"][1]) - float(ref["loss_pt"])) <= REL * abs(float(ref["loss_pt"])) + 1e-6


@pytest.mark.parametrize("cfg_name,batch,graph", [("dvm", 512, True), ("cardiac", 1024, True), ("dvm", 64, False),
                                                  ("dvm", 2048, True)])
def test_head_step_full_size_vs_oracle(S, O, cfg_name, batch, graph):
    from stil_tta_b200 import synth
    cfg = synth.dvm_config(batch) if cfg_name == "dvm" else synth.cardiac_config(batch)
    b = synth.make_batch(cfg, seed=2022)
    o = O.head_step(b, cfg)
    amb = O.ambiguous_rows(b, cfg)
    head = run_head(S, cfg, b, use_graph=graph)
    check_step(head, o, cfg, amb, REL)       # fp32 gradients (the head's default) also for bf16 embeddings: C2, C3
    # replay is idempotent on outputs and accumulates the prototype partials (STiLModel.py:380-381)
    first = {k: v.clone() for k, v in head.out.items()}
    head.run()
    torch.cuda.synchronize()
    for k in ("losses", "pseudo_label", "mask1", "d_feat_i", "class_sum"):
        assert torch.equal(first[k], head.out[k]), k
    assert float((head.prototypes_sum.cpu() - 2 * o["class_sum"]).abs().max()) <= 2e-4
    assert float((head.prototypes_count_sum.cpu() - 2 * o["class_count"]).abs().max()) <= 2e-5


def test_head_prototype_operand_cache_follows_changes(S, O):
    """The cached tensor-core form of the prototypes is refreshed after in-place torch writes and after the
    epoch-end finalise (STiLModel.py:408-415) — graph replays included."""
    from stil_tta_b200 import synth
    cfg = synth.cardiac_config(256)
    b = synth.make_batch(cfg, seed=3)
    head = S.STiLHead(cfg, device="cuda", use_graph=True)
    head.load(b)
    head.run()
    torch.cuda.synchronize()
    l0 = head.out["losses"].clone()
    # 1. in-place torch write
    b2 = dict(b)
    b2["prototypes"] = b["prototypes"] * 0.5
    head.prototypes.copy_(b2["prototypes"])
    head.run()
    torch.cuda.synchronize()
    o2 = O.head_step(b2, cfg, with_grads=False)
    assert abs(float(head.out["losses"][1]) - float(o2["loss_pt"])) <= REL * abs(float(o2["loss_pt"])) + 1e-6
    assert not torch.equal(l0, head.out["losses"])
    # 2. epoch-end finalise writes the prototypes from a kernel
    head.prototypes_sum.copy_(torch.randn(cfg.num_classes, cfg.proj_dim) * 0.1)
    head.prototypes_count_sum.fill_(2.0)
    expect = (head.prototypes_sum / head.prototypes_count_sum).cpu()
    head.finalize_prototypes()
    head.run()
    torch.cuda.synchronize()
    b3 = dict(b)
    b3["prototypes"] = expect
    o3 = O.head_step(b3, cfg, with_grads=False)
    assert abs(float(head.out["losses"][1]) - float(o3["loss_pt"])) <= REL * abs(float(o3["loss_pt"])) + 1e-6
    assert torch.equal(head.out["mask1"].cpu(), o3["mask1"]) or int(O.ambiguous_rows(b3, cfg).sum()) > 0


def test_head_step_host_end_to_end(S, O):
    from stil_tta_b200 import synth
    cfg = synth.dvm_config(512)
    b = synth.make_batch(cfg, seed=1)
    o = O.head_step(b, cfg, with_grads=False)
    head = S.STiLHead(cfg, device="cuda")
    head.prototypes.copy_(b["prototypes"])
    pinned = head.pin(b)
    losses = head.step_host(pinned)
    torch.cuda.synchronize()
    assert abs(float(losses[0]) - float(o["loss_itc"])) <= REL * float(o["loss_itc"])
    assert abs(float(losses[1]) - float(o["loss_pt"])) <= REL * float(o["loss_pt"]) + 1e-6


# ----------------------------------------------------------------------------------------------- DA == True (a6 -> a2/a3)
@pytest.mark.parametrize("name", DA_CASES)
def test_cgpl_pgls_prediction_override_da_golden(S, name):
    """hparams.DA == True (STiLModel.py:276-277): prediction = distribution_alignment(softmax(y_hat_m_ue)) feeds the
    smoothing mix and the threshold (:296-298); fixtures recorded from the executed reference, decisions bit-exact."""
    ins, ref, meta = load_golden(name)
    cfg = cfg_for(name, meta)
    B_l = cfg.b_l
    da = da_state_of(ins)
    q, p = dev(da["DA_queue"]), dev(da["DA_ptr"])
    y_m = dev(ins["y_m_ue"])
    aligned = S.distribution_alignment(torch.softmax(y_m, dim=1), q, p)          # the trainer's own line :277
    assert torch.equal(p.cpu(), ref["DA_ptr"])
    assert float((q.cpu() - ref["DA_queue"]).abs().max()) <= 1e-7
    out = S.cgpl_pgls(y_m, dev(ins["y_i_ue"]), dev(ins["y_t_ue"]), dev(ins["feat_m_e"][B_l:]), dev(ins["prototypes"]),
                      T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1, prediction_override=aligned)
    for k in DECISIONS:
        assert torch.equal(getattr(out, k).cpu(), ref[k]), k
    pl = out.pseudo_label.cpu()
    if cfg.num_classes == 2:
        pl = pl[:, 1]
    assert float((pl - ref["pseudo_label"]).abs().max()) <= 2e-6
    assert float((out.max_prob.cpu() - ref["max_prob"]).abs().max()) <= 2e-6
    # without the override the thresholded decisions are those of the un-aligned prediction — they must differ somewhere,
    # or the fixture would not be testing anything
    plain = S.cgpl_pgls(y_m, dev(ins["y_i_ue"]), dev(ins["y_t_ue"]), dev(ins["feat_m_e"][B_l:]), dev(ins["prototypes"]),
                        T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1)
    assert not torch.equal(plain.max_prob.cpu(), out.max_prob.cpu())


@pytest.mark.parametrize("name", DA_CASES)
@pytest.mark.parametrize("graph", [False, True])
def test_head_step_da_golden(S, O, name, graph):
    """The whole step with DA on: softmax + batch mean + ring-buffer update + alignment run inside the step (and inside
    its CUDA graph); state buffers carry the reference names DA_queue / DA_ptr."""
    ins, ref, meta = load_golden(name)
    cfg = cfg_for(name, meta)
    ins["mask_random"] = ref["mask_random"]
    da = da_state_of(ins)
    head = S.STiLHead(cfg, device="cuda", use_graph=graph, da=True)
    head.DA_queue.copy_(da["DA_queue"])
    head.DA_ptr.copy_(da["DA_ptr"])
    head.load(ins)
    head.run()
    torch.cuda.synchronize()
    o = O.head_step(ins, cfg, da_state=da_state_of(ins))
    check_step(head, o, cfg, torch.zeros(cfg.b_u, dtype=torch.bool), REL)
    for k in DECISIONS:
        assert torch.equal(head.out[k].cpu(), ref[k]), k
    assert torch.equal(head.DA_ptr.cpu(), ref["DA_ptr"])          # exactly one ring-buffer step, capture left no trace
    assert float((head.DA_queue.cpu() - ref["DA_queue"]).abs().max()) <= 1e-7
    assert abs(float(head.out["losses"][1]) - float(ref["loss_pt"])) <= REL * abs(float(ref["loss_pt"])) + 1e-6


# ----------------------------------------------------------------------------------------------- same-device oracle
@pytest.mark.parametrize("cfg_name,batch,seed", [("dvm", 512, 2022), ("cardiac", 1024, 2022), ("dvm", 4096, 5)])
def test_cgpl_pgls_vs_oracle_on_the_same_cuda_device(S, O, cfg_name, batch, seed):
    """the reference's decisions are taken on probabilities produced by torch's CUDA softmax, so the
    oracle is also run ON THE SAME DEVICE (eager torch CUDA, fp32, TF32 off).  Every decision must agree on every row
    that is not within fp32 noise of a boundary (fp64 classification); the counts are recorded in the parity log."""
    from conftest import PARITY_LOG
    from stil_tta_b200 import synth
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = synth.dvm_config(batch) if cfg_name == "dvm" else synth.cardiac_config(batch)
    b = synth.make_batch(cfg, seed=seed, edge_rows=True)
    B_l = cfg.b_l
    args = [dev(b[k]) for k in ("y_m_ue", "y_i_ue", "y_t_ue")] + [dev(b["feat_m_e"][B_l:]), dev(b["prototypes"])]
    kw = dict(T=cfg.temperature, rate_pseudo=cfg.rate_pseudo, th1=cfg.th1)
    o = O.cgpl_pgls(args[0], args[1], args[2], args[3].float(), args[4], **kw)         # eager torch on cuda:0
    out = S.cgpl_pgls(*args, **kw)
    amb = O.ambiguous_rows(b, cfg)
    differ = torch.zeros(cfg.b_u, dtype=torch.bool)
    for k in DECISIONS:
        differ |= (getattr(out, k).cpu() != o[k].cpu())
    n_diff, n_amb = int(differ.sum()), int(amb.sum())
    PARITY_LOG.append({"test": f"same_device_oracle[{cfg_name}-{batch}]", "tensor": f"rows_differ={n_diff},ambiguous={n_amb}",
                       "rel_err": float(n_diff), "excess_over_bf16_rounding": float(int((differ & ~amb).sum())), "tol": 0,
                       "out_dtype": "decision"})
    assert int((differ & ~amb).sum()) == 0, f"{n_diff} rows differ from the CUDA oracle, {n_amb} ambiguous"
    assert float((out.pseudo_label - o["pseudo_label"]).abs().max()) <= 2e-6
    assert float((out.max_prob - o["max_prob"]).abs().max()) <= 2e-6


# ----------------------------------------------------------------------------------------------- hyper-parameter changes
def test_head_set_hparams_recaptures(S, O):
    """The epoch gate of STiLModel.py:317-320 flips during training; kernel parameters are baked into the captured graph,
    so set_hparams drops it and the next run re-captures."""
    from stil_tta_b200 import synth
    cfg = synth.dvm_config(64, past_start_epoch=False)
    b = synth.make_batch(cfg, seed=8)
    head = S.STiLHead(cfg, device="cuda", use_graph=True)
    head.load(b)
    head.run()
    torch.cuda.synchronize()
    o0 = O.head_step(b, cfg, with_grads=False)
    assert abs(float(head.out["losses"][1]) - float(o0["loss_pt"])) <= REL * abs(float(o0["loss_pt"])) + 1e-6
    head.set_hparams(past_start_epoch=True, th1=0.8)
    head.run()
    torch.cuda.synchronize()
    cfg1 = synth.dvm_config(64, past_start_epoch=True, th1=0.8)
    o1 = O.head_step(b, cfg1, with_grads=False)
    assert abs(float(head.out["losses"][1]) - float(o1["loss_pt"])) <= REL * abs(float(o1["loss_pt"])) + 1e-6
    assert float(o1["loss_pt"]) != float(o0["loss_pt"])
    keep = ~O.ambiguous_rows(b, cfg1)
    assert torch.equal(head.out["mask1"].cpu()[keep], o1["mask1"][keep])
    with pytest.raises(ValueError):
        head.set_hparams(no_such_parameter=1)


@pytest.mark.parametrize("rows", [33, 64, 1])
def test_masked_soft_ce_binary_two_rows_per_thread(S, O, rows):
    """K = 2 takes the two-rows-per-thread kernel (16-byte loads); odd row counts end in a single-row tail."""
    g = torch.Generator().manual_seed(rows)
    ys = [torch.randn(rows, 2, generator=g) * 3 for _ in range(3)]
    pl = torch.softmax(torch.randn(rows, 2, generator=g) * 2, 1)
    flags = [torch.rand(rows, generator=g) > 0.4 for _ in range(6)]
    yr = [y.clone().requires_grad_(True) for y in ys]
    lr = O.masked_soft_ce(*yr, pl, *flags)
    gr = torch.autograd.grad(lr[0] + 2 * lr[1] + 3 * lr[2], yr)
    yc = [dev(y).requires_grad_(True) for y in ys]
    lc = S.masked_soft_ce(*yc, dev(pl), *[dev(f) for f in flags])
    gc = torch.autograd.grad(lc[0] + 2 * lc[1] + 3 * lc[2], yc)
    for a, b in zip(lc, lr):
        assert abs(float(a) - float(b)) <= REL * abs(float(b)) + 1e-7
    for a, b, nm in zip(gc, gr, ("d_y_m", "d_y_i", "d_y_t")):
        if float(b.abs().max()) > 0:
            assert_rel(a, b, REL, nm + " (K=2)")
