"""spline interlevel loss (blur_and_resample_weights over the linspline helpers) and the
Charbonnier-sRGB data term -- CUDA bodies vs the oracle restatement and its autograd."""
import numpy as np
import pytest
import torch

from oracle import loss_utils as oloss
from neural_radiance_caching_b200 import loss_utils as nloss, workload
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu


def _stepfun(g, R, n, total):
    t = np.sort(g.uniform(0, 1, size=(R, n + 1)).astype(np.float32), -1)
    t[:, 0], t[:, -1] = 0.0, 1.0
    t[0, 3] = t[0, 4]                       # a zero-width interval (weight_to_pdf's tiny branch)
    w = g.uniform(size=(R, n)).astype(np.float32) ** 3
    w = w / w.sum(-1, keepdims=True) * total
    return f32(t), f32(w)


@pytest.mark.parametrize("blur", [0.03, 0.003])
@pytest.mark.parametrize("m,nq", [(32, 64), (8, 5), (64, 128)])
def test_blur_and_resample_weights(cuda_device, blur, m, nq):
    g = gen(700 + m + nq)
    R = 300
    c, w = _stepfun(g, R, m, 0.8)
    cp, wp = _stepfun(g, R, nq, 0.7)
    want = oloss.blur_and_resample_weights(cp, c, w, blur)
    got = nloss.blur_and_resample_weights(cp.to(cuda_device), c.to(cuda_device), w.to(cuda_device), blur)
    assert rel_err(got, want) <= 1e-5
    # property: blurring moves mass but creates none
    assert float(got.sum(-1).max()) <= 0.8 * (1 + 1e-4)


def test_spline_interlevel_loss_and_gradient(cuda_device):
    g = gen(710)
    R = 257
    c, w = _stepfun(g, R, 32, 0.9)
    hist_o, hist_n = [], []
    for nq in (64, 64):
        cp, wp = _stepfun(g, R, nq, 0.6)
        wo, wn = wp.clone().requires_grad_(True), wp.to(cuda_device).requires_grad_(True)
        hist_o.append(dict(sdist=cp, weights=wo))
        hist_n.append(dict(sdist=cp.to(cuda_device), weights=wn))
    hist_o.append(dict(sdist=c, weights=w))
    hist_n.append(dict(sdist=c.to(cuda_device), weights=w.to(cuda_device)))
    lo = oloss.spline_interlevel_loss(hist_o)
    ln = nloss.spline_interlevel_loss(hist_n)
    sum(lo).backward()
    sum(ln).backward()
    for a, b in zip(ln, lo):
        assert abs(float(a) - float(b)) <= 1e-5 * abs(float(b))
    for hn, ho in zip(hist_n[:-1], hist_o[:-1]):
        assert rel_err(hn["weights"].grad, ho["weights"].grad) <= 1e-5


def test_cache_loss_matches_oracle(cuda_device):
    """workload.cache_loss on the device (CUDA interlevel loss) == the same objective with the oracle's."""
    g = gen(720)
    R = 64
    c, w = _stepfun(g, R, 32, 0.9)
    cp0, wp0 = _stepfun(g, R, 64, 0.6)
    cp1, wp1 = _stepfun(g, R, 64, 0.7)
    rgb, tgt = f32(g.uniform(0, 1.2, size=(R, 3))), f32(g.uniform(size=(R, 3)))
    rgb[0, 0] = 0.001                      # linear branch of linear_to_srgb
    res_o = dict(render=dict(rgb=rgb), sampler=[dict(sdist=cp0, weights=wp0), dict(sdist=cp1, weights=wp1), dict(sdist=c, weights=w)])
    d = lambda t: t.to(cuda_device)
    res_n = dict(render=dict(rgb=d(rgb)), sampler=[dict(sdist=d(h["sdist"]), weights=d(h["weights"])) for h in res_o["sampler"]])
    want = workload.cache_loss(res_o, tgt, interlevel_fn=oloss.spline_interlevel_loss)
    got = workload.cache_loss(res_n, d(tgt))
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
