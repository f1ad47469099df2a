"""FreeMatch self-adaptive threshold + fairness loss and CoTraining cross pseudo labels against
fixtures recorded by EXECUTING the reference (``oracle/gen_golden_thresholds.py``: ``FreeMatchModel.update`` / ``.masking``,
``freematch_utils.entropy_loss``) — CPU: the oracle restatement; GPU: the CUDA drop-ins (``stil_tta_b200/thresholds.py``)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, REL, assert_rel

CASES = ["dvm", "cardiac", "small"]


def _t(z, k):
    return torch.from_numpy(z[k])


# --------------------------------------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("name", CASES)
def test_oracle_freematch_matches_reference(name):
    from oracle import stil_head_oracle as O
    z = np.load(GOLDEN / "freematch.npz")
    c = z[f"{name}_p_model0"].shape[0]
    state = {"p_model": torch.ones(c) / c, "label_hist": torch.ones(c) / c}
    state["time_p"] = state["p_model"].mean()
    for step in range(3):
        out = O.freematch_masking(state, _t(z, f"{name}_logits{step}"))
        assert torch.equal(out["mask"], _t(z, f"{name}_mask{step}"))
        assert torch.equal(state["time_p"].reshape(1), _t(z, f"{name}_time_p{step}"))
        assert torch.equal(state["p_model"], _t(z, f"{name}_p_model{step}"))
        assert torch.equal(state["label_hist"], _t(z, f"{name}_label_hist{step}"))
    ls = _t(z, f"{name}_ent_logits_s").requires_grad_(True)
    loss, hm = O.freematch_entropy_loss(_t(z, f"{name}_ent_mask"), ls, state["p_model"], state["label_hist"])
    (g,) = torch.autograd.grad(loss, ls)
    assert torch.allclose(loss.reshape(1), _t(z, f"{name}_ent_loss"), rtol=1e-6, atol=0)
    assert torch.allclose(hm.reshape(1), _t(z, f"{name}_ent_hist_mean"), rtol=1e-6, atol=0)
    assert torch.allclose(g, _t(z, f"{name}_ent_grad"), rtol=1e-5, atol=1e-9)
    st2 = {"p_model": torch.ones(c) / c, "label_hist": torch.ones(c) / c, "time_p": torch.tensor(0.99)}
    out = O.freematch_masking(st2, _t(z, f"{name}_logits2"), m=0.9, clip_thresh=1.0)
    assert torch.equal(out["mask"], _t(z, f"{name}_clip_mask")) and torch.equal(st2["time_p"].reshape(1), _t(z, f"{name}_clip_time_p"))


def test_oracle_cotraining_fixture_is_self_consistent():
    from oracle import stil_head_oracle as O
    z = np.load(GOLDEN / "cotraining.npz")
    for name in ("dvm", "cardiac"):
        out = O.cotraining_unsup(_t(z, f"{name}_y_i"), _t(z, f"{name}_y_t"), _t(z, f"{name}_y_i_e"), _t(z, f"{name}_y_t_e"),
                                 float(z[f"{name}_threshold"]))
        assert torch.equal(out["mask_i"].float(), _t(z, f"{name}_mask_i")) and torch.equal(out["mask_t"].float(), _t(z, f"{name}_mask_t"))
        assert torch.allclose(out["loss_i_u"].reshape(1), _t(z, f"{name}_loss_i_u"), rtol=1e-6)


def test_freematch_batch_statistics_are_additive_over_ranks():
    """The CUDA drop-in all-reduces ``[column sums | arg-max histogram | sum of row maxima | rows]`` instead of gathering the
    probabilities of all ranks (``freematch_model.py:129-130``): the update computed from the summed statistics of two half
    batches equals the reference update on the concatenated batch."""
    from oracle import stil_head_oracle as O
    g = torch.Generator().manual_seed(5)
    c, m = 7, 0.999
    logits = torch.randn(64, c, generator=g) * 3
    state = {"p_model": torch.rand(c, generator=g), "label_hist": torch.rand(c, generator=g), "time_p": torch.tensor(0.4)}
    ref = {k: v.clone() for k, v in state.items()}
    O.freematch_masking(ref, logits, m=m)
    stats = torch.zeros(2 * c + 2, dtype=torch.float64)
    for part in (logits[:40], logits[40:]):                       # two "ranks"
        p = torch.softmax(part, -1)
        mp, mi = p.max(-1)
        stats[:c] += p.sum(0).double()
        stats[c:2 * c] += torch.bincount(mi, minlength=c).double()
        stats[2 * c] += mp.sum().double()
        stats[2 * c + 1] += part.shape[0]
    n = stats[2 * c + 1]
    time_p = state["time_p"] * m + (1 - m) * (stats[2 * c] / n).float()
    p_model = state["p_model"] * m + (1 - m) * (stats[:c] / n).float()
    label_hist = state["label_hist"] * m + (1 - m) * (stats[c:2 * c] / n).float()
    assert torch.allclose(time_p, ref["time_p"], rtol=1e-6) and torch.allclose(p_model, ref["p_model"], rtol=1e-6)
    assert torch.allclose(label_hist, ref["label_hist"], rtol=1e-6)


def test_threshold_dropins_refuse_cpu_tensors():
    import stil_tta_b200 as S
    with pytest.raises((RuntimeError, ValueError)):
        S.threshold_rows(torch.randn(4, 3), 0.9)
    with pytest.raises((RuntimeError, ValueError)):
        S.entropy_loss(torch.ones(4), torch.randn(4, 3), torch.ones(3) / 3, torch.ones(3) / 3)
    with pytest.raises(RuntimeError):
        S.FreeMatchThreshold(3, device="cpu")


# ---------------------------------------------------------------------------------------------------- GPU: CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_freematch_threshold_matches_reference(name):
    import stil_tta_b200 as S
    from oracle import stil_head_oracle as O
    z = np.load(GOLDEN / "freematch.npz")
    c = z[f"{name}_p_model0"].shape[0]
    fm = S.FreeMatchThreshold(c, momentum=0.999, clip_thresh=0.0, device="cuda")
    ref_state = {"p_model": torch.ones(c) / c, "label_hist": torch.ones(c) / c}
    ref_state["time_p"] = ref_state["p_model"].mean()
    for step in range(3):
        logits = _t(z, f"{name}_logits{step}")
        mask = fm.masking(logits.cuda())
        ref = O.freematch_masking(ref_state, logits)
        assert_rel(fm.time_p, _t(z, f"{name}_time_p{step}"), 1e-5, "time_p")
        assert_rel(fm.p_model, _t(z, f"{name}_p_model{step}"), 1e-5, "p_model")
        assert torch.allclose(fm.label_hist.cpu(), _t(z, f"{name}_label_hist{step}"), rtol=1e-6, atol=1e-9)
        assert torch.equal(fm.max_idx.cpu(), ref["max_idx"])                      # integer part: bit-exact
        assert_rel(fm.max_probs, ref["max_probs"], 1e-5, "max_probs")
        keep = (ref["max_probs"] - ref["thr"]).abs() > 1e-5                        # decisions: exact outside the ambiguity band
        assert torch.equal(mask.cpu()[keep], _t(z, f"{name}_mask{step}")[keep])
        assert int(keep.sum()) >= int(0.95 * keep.numel())
    # fairness loss with the state the reference had at this point
    ls = _t(z, f"{name}_ent_logits_s").cuda().requires_grad_(True)
    loss, hm = S.entropy_loss(_t(z, f"{name}_ent_mask").cuda(), ls, ref_state["p_model"].cuda(), ref_state["label_hist"].cuda())
    (g,) = torch.autograd.grad(loss * 1.7, ls)
    assert_rel(loss.reshape(1), _t(z, f"{name}_ent_loss"), 1e-4, "entropy_loss")
    assert_rel(hm.reshape(1), _t(z, f"{name}_ent_hist_mean"), 1e-6, "hist_s.mean()")
    assert_rel(g, 1.7 * _t(z, f"{name}_ent_grad"), REL, "d entropy_loss / d logits_s")
    unsel = _t(z, f"{name}_ent_mask") == 0
    assert float(g.cpu()[unsel].abs().max()) == 0.0 if bool(unsel.any()) else True
    # the clip branch (:139-140)
    fm2 = S.FreeMatchThreshold(c, momentum=0.9, clip_thresh=1.0, device="cuda")
    fm2.time_p.fill_(0.99)
    fm2.masking(_t(z, f"{name}_logits2").cuda())
    assert_rel(fm2.time_p, _t(z, f"{name}_clip_time_p"), 1e-6, "clipped time_p")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dvm", "cardiac"])
def test_cotraining_cross_pseudo_labels(name):
    import stil_tta_b200 as S
    z = np.load(GOLDEN / "cotraining.npz")
    thr = float(z[f"{name}_threshold"])
    pl_i, pl_t, mask_i, mask_t = S.cotraining_pseudo_labels(_t(z, f"{name}_y_i_e").cuda(), _t(z, f"{name}_y_t_e").cuda(), thr)
    for m, mp, tag in ((mask_i, pl_i.max(1).values, "i"), (mask_t, pl_t.max(1).values, "t")):
        ref_mp = _t(z, f"{name}_max_prob_{tag}")
        assert_rel(mp, ref_mp, 1e-5, f"max_prob_{tag}")
        keep = (ref_mp - thr).abs() > 1e-5
        assert torch.equal(m.cpu()[keep], _t(z, f"{name}_mask_{tag}")[keep])
    yi = _t(z, f"{name}_y_i").cuda().requires_grad_(True)
    yt = _t(z, f"{name}_y_t").cuda().requires_grad_(True)
    loss_i_u = S.masked_ce(yi, pl_t, mask_t)                                        # CoTraining.py:148
    loss_t_u = S.masked_ce(yt, pl_i, mask_i)                                        # :149
    gi, gt = torch.autograd.grad(loss_i_u + loss_t_u, (yi, yt))
    assert_rel(loss_i_u.reshape(1), _t(z, f"{name}_loss_i_u"), REL, "loss_i_u")
    assert_rel(loss_t_u.reshape(1), _t(z, f"{name}_loss_t_u"), REL, "loss_t_u")
    assert_rel(gi, _t(z, f"{name}_d_y_i"), REL, "d loss / d y_hat_i")
    assert_rel(gt, _t(z, f"{name}_d_y_t"), REL, "d loss / d y_hat_t")
