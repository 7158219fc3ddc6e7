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


@pytest.mark.parametrize("R,n", [(64, 32), (7, 100), (3, 1)])
def test_distortion_loss_matches_oracle(cuda_device, R, n):
    """nrc_distortion_loss against loss_utils.distortion_loss over stepfun.lossfun_distortion on metric distances
    through power_ladder(-0.25, 1e4).  The curve compresses [2, 6] into a 0.03-wide interval near 4.6, so the pairwise
    differences lose ~3 digits to cancellation in fp32: tolerance 2e-4 of the output scale."""
    g = gen(1400 + n)
    t = f32(np.sort(g.uniform(2.0, 6.0, size=(R, n + 1)), axis=-1))
    w = f32(g.uniform(0, 1, size=(R, n)) ** 3 * 0.2)
    Wo = w.clone().requires_grad_(True)
    lo = oloss.distortion_loss([{"tdist": t, "weights": Wo}], 0.01, -0.25, 10000.0)
    lo.backward()
    Wn = w.clone().to(cuda_device).requires_grad_(True)
    ln = nloss.distortion_loss([{"tdist": t.to(cuda_device), "weights": Wn}], 0.01, -0.25, 10000.0)
    ln.backward()
    assert abs(float(ln.detach()) - float(lo.detach())) <= 2e-4 * abs(float(lo.detach()))
    assert rel_err(Wn.grad, Wo.grad) <= 2e-4


def test_grid_regularizer_matches_oracle(cuda_device):
    """nrc_grid_regularizer against param_regularizer_loss (internal/train_utils.py:1169-1216), every level table."""
    import ctypes as C
    from neural_radiance_caching_b200 import _lib, grid_utils as ng
    g = gen(1500)
    enc = ng.HashEncoding(hash_map_size=4096, max_grid_size=128, num_features=2, scale_supersample=1.0)
    gen_t = torch.Generator(device=cuda_device)
    gen_t.manual_seed(5)
    _, arena = enc.init(cuda_device, generator=gen_t, init_range=0.3)
    tables_o = [t.detach().cpu().clone().requires_grad_(True) for t in enc.tables(enc.views(arena))]
    lo = oloss.param_regularizer_loss(tables_o, 1.0)
    lo.backward()
    grad = torch.zeros_like(arena)
    loss = torch.zeros((), device=cuda_device)
    d = enc._descriptor(enc.tables(enc.views(arena)), enc.tables(enc.views(grad)))
    _lib.call("nrc_grid_regularizer", _lib.stream_ptr(), C.byref(d), 1.0, _lib.ptr(loss))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(lo.detach())) <= 1e-5 * float(lo.detach())
    for a, b in zip(enc.tables(enc.views(grad)), tables_o):
        assert rel_err(a, b.grad) <= 1e-5
    _lib.call("nrc_grid_regularizer", _lib.stream_ptr(), C.byref(d), 1.0, _lib.ptr(loss))   # the kernel ACCUMULATES
    torch.cuda.synchronize()
    for a, b in zip(enc.tables(enc.views(grad)), tables_o):
        assert rel_err(a, 2 * b.grad) <= 1e-5
    # init mode: the gradient tables are OVERWRITTEN (stale contents, here 2x the gradient, are replaced), loss accumulated
    loss.zero_()
    _lib.call("nrc_grid_regularizer_init", _lib.stream_ptr(), C.byref(d), 1.0, _lib.ptr(loss))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(lo.detach())) <= 1e-5 * float(lo.detach())
    for a, b in zip(enc.tables(enc.views(grad)), tables_o):
        assert rel_err(a, b.grad) <= 1e-5
