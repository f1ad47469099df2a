"""Generate tests/golden/linear_*.npz by EXECUTING the reference's ``STiLModel.project_3features``
(``models/Disentangle/STiLModel.py:182-192``: projector + ``F.normalize``) as an unbound function on a stand-in ``self`` whose
projectors are the very modules the reference constructs for DVM (``nn.Linear(multimodal_embedding_dim, projection_dim)``,
``:57-59``), and the classifier ``nn.Linear``s of ``STiLModel_backbone.py:66-68`` applied like ``forward_all`` does (``:153-155``:
``classifier(torch.cat([...], dim=1))``) — outputs and the gradients of a random cotangent w.r.t. x, weight and bias.
TEST INFRASTRUCTURE; run once in the authoring container.  Only tensors are stored.
"""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle.gen_golden import OUT, load_reference  # noqa: E402


def main():
    torch.set_num_threads(1)
    _, _, STiLModel = load_reference()
    g = torch.Generator().manual_seed(11)
    rec = {}
    # ---- projection heads + F.normalize (DVM: Linear 512 -> 128; a small odd-sized case too)
    for tag, (rows, din, dout) in {"proj_dvm": (64, 512, 128), "proj_small": (37, 40, 24)}.items():
        torch.manual_seed(3)
        me = SimpleNamespace(projector_imaging=torch.nn.Linear(din, dout), projector_tabular=torch.nn.Linear(din, dout),
                             projector_multimodal=None)
        xi = torch.randn(rows, din, generator=g).requires_grad_(True)
        xt = torch.randn(rows, din, generator=g).requires_grad_(True)
        _, fi, ft = STiLModel.project_3features(me, None, xi, xt)            # the real method body
        ci, ct = torch.randn(rows, dout, generator=g), torch.randn(rows, dout, generator=g)
        ((fi * ci).sum() + (ft * ct).sum()).backward()
        heads = (("i", me.projector_imaging, xi, fi, ci), ("t", me.projector_tabular, xt, ft, ct))
        for nm, mod, x, f, c in (heads[:1] if tag == "proj_dvm" else heads):      # one DVM-sized head keeps the fixture small
            p = f"{tag}_{nm}_"
            rec.update({p + "x": xi.detach().numpy() if nm == "i" else xt.detach().numpy(), p + "w": mod.weight.detach().numpy(),
                        p + "b": mod.bias.detach().numpy(), p + "y": f.detach().numpy(), p + "cot": c.numpy(),
                        p + "dx": x.grad.numpy(), p + "dw": mod.weight.grad.numpy(), p + "db": mod.bias.grad.numpy()})
    # ---- classifiers: Linear(hidden * 3 -> K) on the concatenation, K = 286 (DVM) and 2 (cardiac)
    for tag, (rows, hid, k) in {"cls_dvm": (64, 3 * 64, 286), "cls_cardiac": (50, 2 * 32, 2)}.items():
        torch.manual_seed(5)
        clf = torch.nn.Linear(hid, k)
        parts = [torch.randn(rows, hid // (3 if "dvm" in tag else 2), generator=g) for _ in range(3 if "dvm" in tag else 2)]
        x = torch.cat(parts, dim=1).requires_grad_(True)                     # STiLModel_backbone.py:153-155
        y = clf(x)
        c = torch.randn(rows, k, generator=g)
        (y * c).sum().backward()
        p = tag + "_"
        rec.update({p + "x": x.detach().numpy(), p + "w": clf.weight.detach().numpy(), p + "b": clf.bias.detach().numpy(),
                    p + "y": y.detach().numpy(), p + "cot": c.numpy(), p + "dx": x.grad.numpy(), p + "dw": clf.weight.grad.numpy(),
                    p + "db": clf.bias.grad.numpy()})
    np.savez_compressed(OUT / "linear_heads.npz", **rec)
    print("linear_heads.npz written:", len(rec), "arrays")


if __name__ == "__main__":
    main()
