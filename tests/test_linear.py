"""f-3 — the Linear layers either side of the head (projectors + F.normalize, classifiers) against fixtures recorded from
the executed reference (oracle/gen_golden_linear.py: the real STiLModel.project_3features, the classifiers applied like
STiLModel_backbone.forward_all)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, REL, assert_rel

CASES = [("proj_dvm_i", True), ("proj_small_i", True), ("proj_small_t", True), ("cls_dvm", False), ("cls_cardiac", False)]


def load(tag):
    z = np.load(GOLDEN / "linear_heads.npz")
    return {k[len(tag) + 1:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith(tag + "_")}


@pytest.mark.parametrize("tag,normalize", CASES)
def test_fixture_is_self_consistent(tag, normalize):
    """CPU: the recorded outputs are what plain torch gives for the recorded inputs (guards the fixture itself)."""
    z = load(tag)
    y = torch.nn.functional.linear(z["x"], z["w"], z["b"])
    if normalize:
        y = torch.nn.functional.normalize(y)
    torch.testing.assert_close(y, z["y"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,normalize", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_matches_reference(tag, normalize, dtype):
    import stil_tta_b200 as S
    z = load(tag)
    x = z["x"].to(dtype)
    if dtype == torch.bfloat16:
        if x.shape[1] % 8:
            pytest.skip("bf16 rows must be 16-byte granular")
        # reference on the SAME (bf16-rounded) inputs, fp32 arithmetic
        xr = x.float().requires_grad_(True)
        w, b = z["w"].clone().requires_grad_(True), z["b"].clone().requires_grad_(True)
        yr = torch.nn.functional.linear(xr, w, b)
        if normalize:
            yr = torch.nn.functional.normalize(yr)
        (yr * z["cot"]).sum().backward()
        ref = {"y": yr.detach(), "dx": xr.grad, "dw": w.grad, "db": b.grad}
    else:
        ref = z
    dout, din = z["w"].shape
    lin = S.Linear(din, dout, normalize=normalize, device="cuda")
    with torch.no_grad():
        lin.weight.copy_(z["w"]); lin.bias.copy_(z["b"])
    xc = x.cuda().requires_grad_(True)
    y = lin(xc)
    (y * z["cot"].cuda()).sum().backward()
    assert_rel(y, ref["y"], REL, f"{tag} y")
    assert_rel(xc.grad, ref["dx"], REL, f"{tag} d_x")
    assert_rel(lin.weight.grad, ref["dw"], REL, f"{tag} d_weight")
    assert_rel(lin.bias.grad, ref["db"], REL, f"{tag} d_bias")
    if normalize:
        assert float((y.norm(dim=1) - 1).abs().max()) <= 1e-5


@pytest.mark.gpu
def test_linear_feeds_the_head_like_project_3features():
    """projector -> F.normalize -> CLIPLoss, end to end with gradients to the projector parameters (STiLModel.py:188-191, :322)."""
    import stil_tta_b200 as S
    from oracle import stil_head_oracle as O
    g = torch.Generator().manual_seed(2)
    B, din, P = 256, 512, 128
    xi, xt = torch.randn(B, din, generator=g), torch.randn(B, din, generator=g)
    pi, pt = torch.nn.Linear(din, P), torch.nn.Linear(din, P)
    loss_r, _, _ = O.clip_loss(torch.nn.functional.normalize(pi(xi)), torch.nn.functional.normalize(pt(xt)), 0.1, 0.5)
    loss_r.backward()
    si, st = S.Linear(din, P, normalize=True, device="cuda"), S.Linear(din, P, normalize=True, device="cuda")
    si.load_state_dict(pi.state_dict()); st.load_state_dict(pt.state_dict())       # nn.Linear state dicts round-trip
    loss, _, _ = S.CLIPLoss(0.1, 0.5, return_logits=False)(si(xi.cuda()), st(xt.cuda()))
    loss.backward()
    assert abs(float(loss) - float(loss_r)) <= REL * abs(float(loss_r))
    assert_rel(si.weight.grad, pi.weight.grad, REL, "projector_imaging.weight.grad")
    assert_rel(st.bias.grad, pt.bias.grad, REL, "projector_tabular.bias.grad")
    with pytest.raises(ValueError):
        S.Linear(16, 256, normalize=True)
