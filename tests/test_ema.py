"""f-4 — EMA teacher update (STiLModel.momentum_update_ema, STiLModel.py:154-168): the oracle restatement and the
one-launch CUDA kernel against fixtures recorded from the executed reference method (oracle/gen_golden_ema.py)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import stil_head_oracle as O

CASES = ["ema_eman", "ema_params"]
PARAMS = lambda names: {k for k in names if not any(s in k for s in ("running_mean", "running_var", "num_batches_tracked"))}


def load(name):
    z = np.load(GOLDEN / f"{name}.npz")
    t = lambda k: torch.from_numpy(np.array(z[k]))
    names = [k[5:] for k in z.files if k.startswith("main_")]
    return z, t, names


@pytest.mark.parametrize("name", CASES)
def test_oracle_ema_matches_reference(name):
    z, t, names = load(name)
    eman, m = bool(int(z["meta_eman"])), float(z["meta_momentum"])
    main = {k: t("main_" + k) for k in names}
    ema = {k: t("ema0_" + k).clone() for k in names}
    for step in range(1, 4):
        O.momentum_update_ema(main, ema, m, eman, PARAMS(names))
        if step in (1, 3):
            for k in names:
                assert torch.equal(ema[k], t(f"ema{step}_" + k)), (step, k)       # bit-exact


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_ema_kernel_bit_exact_vs_reference(name):
    import stil_tta_b200 as S
    z, t, names = load(name)
    eman, m = bool(int(z["meta_eman"])), float(z["meta_momentum"])
    main = {k: t("main_" + k).cuda() for k in names}
    ema = {k: t("ema0_" + k).cuda() for k in names}
    keys = names if eman else sorted(PARAMS(names), key=names.index)
    upd = S.EmaTeacher([(ema[k], main[k], eman and "num_batches_tracked" in k) for k in keys], m)
    assert upd.n_entries == len(keys)
    for step in range(1, 4):
        upd.step()
        if step in (1, 3):
            torch.cuda.synchronize()
            for k in names:
                assert torch.equal(ema[k].cpu(), t(f"ema{step}_" + k)), (step, k)  # bit-exact, untouched entries included


@pytest.mark.gpu
def test_ema_from_modules_graph_and_bf16():
    """from_modules on real nn.Modules (both branches), replayed from a CUDA graph; bf16 parameters follow the eager
    rounding (each product and the sum rounded to bf16)."""
    import stil_tta_b200 as S

    def net(seed, dtype=torch.float32):
        torch.manual_seed(seed)
        n = torch.nn.Sequential(torch.nn.Linear(33, 129), torch.nn.BatchNorm1d(129), torch.nn.Linear(129, 7)).cuda().to(dtype)
        n[1].num_batches_tracked.fill_(seed)
        return n
    for eman in (True, False):
        model, ema, ref = net(1), net(2), net(2)
        upd = S.EmaTeacher.from_modules(model, ema, 0.99, eman=eman)
        g = torch.cuda.CUDAGraph()
        upd.step()                                                # warm-up (also applied to the reference below)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            upd.step()
        g.replay()
        torch.cuda.synchronize()
        sm, sr = {k: v.cpu() for k, v in model.state_dict().items()}, {k: v.cpu() for k, v in ref.state_dict().items()}
        for _ in range(2):
            O.momentum_update_ema(sm, sr, 0.99, eman, {k for k, _ in model.named_parameters()})
        for k, v in ema.state_dict().items():
            assert torch.equal(v.cpu(), sr[k]), (eman, k)
    # bf16
    model, ema = net(3, torch.bfloat16), net(4, torch.bfloat16)
    ref = {k: v.clone() for k, v in ema.state_dict().items()}
    S.EmaTeacher.from_modules(model, ema, 0.9, eman=False).step()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        exp = ref[k].mul_(0.9).add_((1.0 - 0.9) * p.data)
        assert torch.equal(dict(ema.named_parameters())[k].data, exp), k
    with pytest.raises(ValueError):
        S.EmaTeacher([(torch.zeros(3, device="cuda"), torch.zeros(4, device="cuda"), False)], 0.9)
