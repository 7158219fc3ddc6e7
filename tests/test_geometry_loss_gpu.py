"""nrc_geometry_losses / nrc_mask_loss against the oracle's restatement of internal/loss_utils.py:127-199 and
internal/train_utils.py:785-836 (values and gradients w.r.t. weights, analytic normals, predicted normals, acc)."""
import numpy as np
import pytest
import torch

from oracle import loss_utils as oloss
from neural_radiance_caching_b200 import loss_utils as nloss
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


@pytest.mark.parametrize("R,n", [(37, 32), (256, 48), (5, 7)])
def test_geometry_losses_match_oracle(cuda_device, R, n):
    g = gen(1200 + R)
    w = f32(g.uniform(0, 0.2, size=(R, n)))
    nrm = f32(_unit(g.normal(size=(R, n, 3))))
    npred = f32(_unit(g.normal(size=(R, n, 3))))
    nrm[0, 0] = float("nan")      # nan_to_num path
    vd = f32(_unit(g.normal(size=(R, 3))))
    mults = (0.01, 0.001, 0.01)
    def run(lib, dev):
        W, N, NP = (t.clone().to(dev).requires_grad_(True) for t in (w, nrm, npred))
        ls = lib.geometry_losses({"viewdirs": vd.to(dev)}, {"weights": W, "normals": N, "normals_pred": NP}, *mults, 0.1)
        tot = sum(ls)
        tot.backward()
        return tot.detach().cpu(), W.grad.cpu(), N.grad.cpu(), NP.grad.cpu()
    lo, gwo, gno, gpo = run(oloss, "cpu")
    ln, gwn, gnn, gpn = run(nloss, cuda_device)
    assert abs(float(ln) - float(lo)) <= 1e-5 * abs(float(lo))
    assert rel_err(gwn, gwo) <= 1e-5
    assert rel_err(gpn, gpo) <= 1e-5
    gno = torch.nan_to_num(gno)   # the oracle's gradient at the NaN normal is NaN * 0; the kernel writes the finite value
    gnn[0, 0] = 0.0
    gno[0, 0] = 0.0
    assert rel_err(gnn, gno) <= 1e-5


@pytest.mark.parametrize("backward", [False, True])
def test_mask_loss_matches_oracle(cuda_device, backward):
    g = gen(1300)
    R = 333
    acc = f32(g.uniform(0, 1.2, size=(R,)))
    A = acc.clone().requires_grad_(True)
    if backward:   # the backward-mask call: zero masks (train_utils.py:2936-2938)
        lo = oloss.compute_mask_loss(A, torch.zeros(R, 1), 0.001, empty_loss_weight=0.1, backward=True)
    else:
        lo = oloss.compute_mask_loss(A, None, 0.001, 1.0, 1.0)
    lo.backward()
    B = acc.clone().to(cuda_device).requires_grad_(True)
    if backward:
        ln = nloss.compute_mask_loss(B, None, 0.001, empty_loss_weight=0.1, backward=True)
    else:
        ln = nloss.compute_mask_loss(B, None, 0.001, 1.0, 1.0)
    ln.backward()
    assert abs(float(ln) - float(lo)) <= 1e-5 * abs(float(lo))
    assert rel_err(B.grad, A.grad) <= 1e-5
