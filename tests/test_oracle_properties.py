"""Structural properties of the head checked on the CPU oracle over seeded synthetic batches: they hold
for the reference by construction and give size-independent checks beside the golden vectors."""
import pytest
import torch

from oracle import stil_head_oracle as O
from stil_tta_b200 import synth


@pytest.mark.parametrize("cfg,seed", [(synth.dvm_config(64), 11), (synth.dvm_config(128, embed_dtype="f32"), 12),
                                      (synth.cardiac_config(128), 13), (synth.dvm_config(96, num_classes=10, th1=0.6), 14)])
def test_head_step_invariants(cfg, seed):
    torch.set_num_threads(1)
    b = synth.make_batch(cfg, seed=seed)
    o = O.head_step(b, cfg)
    b_l, b_u, k = cfg.b_l, cfg.b_u, cfg.num_classes
    # the four agreement cases partition the unlabelled rows (STiLModel.py:268)
    s = o["case1"].int() + o["case2_i"].int() + o["case2_t"].int() + o["case3"].int()
    assert torch.all(s == 1)
    # pseudo labels and predictions are distributions; max_prob / max_idx / mask1 are consistent with `prediction`
    for key in ("pseudo_label", "prediction", "teacher_probs"):
        torch.testing.assert_close(o[key].sum(1), torch.ones(b_u), rtol=0, atol=2e-5)
    assert torch.equal(o["max_idx"], o["prediction"].argmax(1))
    assert torch.equal(o["mask1"], o["max_prob"] >= cfg.th1)
    # prototype partials: counts = labelled rows / repeat_ratio + confident unlabelled rows; sums have the counts' support
    label_all = o["label_all"]
    conf_u = (label_all[b_l:].max(1).values >= cfg.th1).sum()
    expect = b_l / cfg.repeat_ratio + float(conf_u)
    assert abs(float(o["class_count"].sum()) - expect) <= 1e-4 * max(1.0, expect)
    empty = o["class_count"].reshape(-1) == 0
    assert torch.all(o["class_sum"][empty] == 0)
    # gradients of the masked CE vanish on rows the mask removes
    dead = ~o["mask1"]
    assert torch.all(o["d_y_m"][b_l:][dead] == 0) and torch.all(o["d_y_m"][:b_l] == 0)


def test_prototype_loss_is_zero_when_nothing_is_confident():
    cfg = synth.dvm_config(64)
    b = synth.make_batch(cfg, seed=21)
    label = torch.full((64, cfg.num_classes), 1.0 / cfg.num_classes)            # max prob 1/K < threshold
    feat = b["feat_m"].float().requires_grad_(True)
    loss = O.prototype_loss(label, b["prototypes"], feat, cfg.temperature, cfg.th1)
    assert float(loss) == 0.0
    (g,) = torch.autograd.grad(loss, feat, allow_unused=True)
    assert g is None or float(g.abs().max()) == 0.0


def test_infonce_is_symmetric_and_bounded():
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(48, 32, generator=g), torch.randn(48, 32, generator=g)
    l_ab, logits, labels = O.clip_loss(a, b, 0.1, 0.5)
    l_ba, logits_t, _ = O.clip_loss(b, a, 0.1, 0.5)
    torch.testing.assert_close(l_ab, l_ba)                       # lambda = 0.5: swapping the modalities changes nothing
    torch.testing.assert_close(logits, logits_t.t())
    assert float(logits.abs().max()) <= 10.0 + 1e-4             # unit vectors / T
    assert torch.equal(labels, torch.arange(48))
    # the loss of a perfectly aligned pair is below log(B) and tends to 0 with the temperature
    l_same, _, _ = O.clip_loss(a, a, 0.01, 0.5)
    assert float(l_same) < 1e-3
