"""Parity of the proposal sampling loop (BASELINE config 1/2, cache stage):
3 levels x (sample_intervals -> ray cast -> contract+encode+MLP -> alpha weights).

fp32 rounding differences in the CDF inversion move samples by ~1e-5; with white-noise
("trained-like" stress) tables the field has O(1) variation per finest cell (1/2048), so those
position differences are amplified by ~N into 1e-3..1e-2 differences downstream.  Per-level
parity is therefore checked on bit-identical sample positions (the oracle's fenceposts are fed
to the CUDA level through `sdist_override`), which isolates the kernels from that conditioning;
a separate end-to-end test bounds the compounded difference.
"""
import numpy as np
import pytest
import torch

from oracle import sampling as osamp
from neural_radiance_caching_b200 import sampling as nsamp
from tests.util import f32, gen, make_rays, rel_err, to_dev

pytestmark = pytest.mark.gpu


def _setup(g, R, table_range, use_raydist, device):
    o = osamp.ProposalVolumeSampler()
    n = nsamp.ProposalVolumeSampler()
    po = o.init(g, table_init_range=table_range, bias_range=0.05)
    pn = n.from_oracle(po, device)
    rays = make_rays(g, R, near=0.05, far=2.0, radius=0.7) if use_raydist else make_rays(g, R)
    u = [f32(g.uniform(size=(R, 1))) for _ in range(3)]
    return o, n, po, pn, rays, u


@pytest.mark.parametrize("table_range,use_raydist", [(0.1, False), (None, False), (0.1, True)])
def test_sampler_levels_forward(cuda_device, table_range, use_raydist):
    g = gen(100)
    R = 256
    o, n, po, pn, rays, u = _setup(g, R, table_range, use_raydist, cuda_device)
    ho = o(po, rays, u, use_raydist_fn=use_raydist)
    override = [h["sdist"].to(cuda_device) for h in ho]
    hn = n(pn, to_dev(rays, cuda_device), to_dev(u, cuda_device), use_raydist_fn=use_raydist,
           sdist_override=override)
    for lvl, (a, b) in enumerate(zip(hn, ho)):
        for k in ("tdist", "means", "density", "weights", "alphas", "trans", "feature"):
            # the power-ladder ray warp goes through powf (CUDA) vs pow (MKL): tdist agrees to
            # 1e-5 but is no longer bit-identical, and the white-noise field amplifies that.
            tol = 1e-5 if (not use_raydist or k in ("tdist", "means")) else 1e-4
            assert rel_err(a[k], b[k]) <= tol, (lvl, k, rel_err(a[k], b[k]))
        # interval resampling from the previous level's (CUDA) weights: the weights agree to
        # 1e-5, the inverse CDF amplifies that by its slope (bounded by the 1e-5 padding)
        assert float((a["sdist_sampled"].cpu() - b["sdist"]).abs().max()) <= 5e-5, lvl
    for k in ("normals_pred", "normals"):
        d = (hn[2][k].cpu() - ho[2][k].detach()).abs().max(dim=-1).values
        # unit vectors: where the (raw) gradient is tiny, normalisation amplifies rounding
        assert float(d.median()) <= 1e-5 and float((d > 1e-3).float().mean()) <= 2e-3, (k, float(d.max()))
    assert hn[0]["normals"] is None and hn[1]["normals"] is None


def test_sampler_end_to_end_forward(cuda_device):
    """Free-running pipeline (no override): differences compound but stay small."""
    g = gen(102)
    R = 256
    o, n, po, pn, rays, u = _setup(g, R, None, False, cuda_device)  # reference init: smooth field
    ho = o(po, rays, u)
    hn = n(pn, to_dev(rays, cuda_device), to_dev(u, cuda_device), return_covs=True)
    for lvl, (a, b) in enumerate(zip(hn, ho)):
        for k in ("sdist", "tdist", "means", "density", "weights"):
            assert rel_err(a[k], b[k]) <= 1e-4, (lvl, k, rel_err(a[k], b[k]))
        # the reference's `covs` entry (cast_rays diag=False, sampling.py:361-368), emitted on request
        assert a["covs"].shape == b["covs"].shape and rel_err(a["covs"], b["covs"]) <= 1e-4, (lvl, rel_err(a["covs"], b["covs"]))
    assert "covs" not in n(pn, to_dev(rays, cuda_device), to_dev(u, cuda_device))[0]


def test_sampler_train_gradients(cuda_device):
    """Backward of the cache-stage sampler: d(sum_l <weights_l, G_l>)/d(params) for all three
    proposal levels (tables + MLP weights) vs oracle autograd, on identical sample positions."""
    g = gen(101)
    R = 128
    o, n, po, pn, rays, u = _setup(g, R, 0.1, False, cuda_device)
    G = [f32(g.normal(size=(R, ns))) for (_, _, ns) in o.sampling_strategy]

    def leaves(p):
        out = []
        for i in range(3):
            m = p[f"MLP_{i}"]
            for name in sorted(k for k in m["density_grid"].keys() if k != "_arena"):
                out.append((f"MLP_{i}/density_grid/{name}", m["density_grid"], name))
            for k in ("density_layers_0", "density_layers_1", "output_density_layer"):
                for kk in ("kernel", "bias"):
                    out.append((f"MLP_{i}/{k}/{kk}", m[k], kk))
        return out

    lo = leaves(po)
    for _, d, k in lo:
        d[k] = d[k].clone().requires_grad_(True)
    ho = o(po, rays, u)
    sum((h["weights"] * Gl).sum() for h, Gl in zip(ho, G)).backward()
    override = [h["sdist"].to(cuda_device) for h in ho]

    # CUDA: make the arena the leaf so table grads land in one contiguous buffer
    for i, m in enumerate(n.mlps):
        p = pn[f"MLP_{i}"]
        arena = p["density_grid"]["_arena"].clone().requires_grad_(True)
        p["density_grid"] = dict(m.grid.views(arena.detach()), _arena=arena)
    ln = leaves(pn)
    for name, d, k in ln:
        if "density_grid" not in name:
            d[k] = d[k].clone().requires_grad_(True)
    hn = n(pn, to_dev(rays, cuda_device), to_dev(u, cuda_device), train=True, sdist_override=override)
    sum((h["weights"] * Gl.to(cuda_device)).sum() for h, Gl in zip(hn, G)).backward()
    for i, m in enumerate(n.mlps):
        gviews = m.grid.views(pn[f"MLP_{i}"]["density_grid"]["_arena"].grad)
        for name in gviews:
            ref = po[f"MLP_{i}"]["density_grid"][name].grad
            assert rel_err(gviews[name], ref) <= 2e-5, (i, name, rel_err(gviews[name], ref))
    for (name, dn, kn), (_, do, ko) in zip(ln, lo):
        if "density_grid" in name:
            continue
        assert rel_err(dn[kn].grad, do[ko].grad) <= 2e-5, (name, rel_err(dn[kn].grad, do[ko].grad))
