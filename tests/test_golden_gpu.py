"""GPU: the CUDA path (through the C ABI) against the committed golden fixtures of the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import grid_utils as og
from neural_radiance_caching_b200 import coord as ncoord, grid_utils as ng, models as nmodels, nerf as nnerf
from neural_radiance_caching_b200 import render as nrender, stepfun as nstep
from neural_radiance_caching_b200.inverse_render import render_utils as nru
from tests.golden import make_golden
from tests.util import f32, rel_err

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz"))
KW = dict(hash_map_size=524288, num_features=4, scale_supersample=1.0, max_grid_size=2048, bbox_scaling=1.0)


def test_encode_indices_and_features(cuda_device):
    n = ng.HashEncoding(**KW)
    po = og.HashEncoding(**KW).init(np.random.Generator(np.random.PCG64(make_golden.SEED + 1)), init_range=0.1)
    pn = {k: v.to(cuda_device) for k, v in po.items()}
    x = f32(GOLD["enc_x"]).to(cuda_device)
    for l in range(8):
        assert np.array_equal(n.corner_indices(pn, x, l).cpu().numpy(), GOLD[f"enc_idx_l{l}"]), l   # bit-exact
    assert rel_err(n(pn, x), f32(GOLD["enc_feat"])) <= 1e-5
    assert rel_err(ncoord.contract_radius_2(x * 3.0), f32(GOLD["contract_c2"])) <= 1e-6


def test_sample_intervals_bins_and_fenceposts(cuda_device):
    d = lambda k: f32(GOLD[k]).to(cuda_device)
    sd, bins = nstep.sample_intervals_from_weights(d("si_u01"), d("si_t"), d("si_w"), 12, anneal=0.4, padding=1e-5,
                                                   domain=(0.0, 1.0), return_bins=True)
    assert np.array_equal(bins.cpu().numpy(), GOLD["si_bins"])       # bit-exact bin indices
    assert rel_err(sd, f32(GOLD["si_sdist"])) <= 1e-5


def test_alpha_weights_rendering_resample(cuda_device):
    d = lambda k: f32(GOLD[k]).to(cuda_device)
    w, a, t = nrender.compute_alpha_weights(d("aw_density"), d("aw_tdist"), d("aw_dirs"))
    for got, k in ((w, "aw_weights"), (a, "aw_alpha"), (t, "aw_trans")):
        assert rel_err(got, f32(GOLD[k])) <= 1e-5, k
    ren = nrender.volumetric_rendering(d("vr_rgbs"), w, w, d("aw_tdist"), 1.0, True)
    for k in ("rgb", "acc", "distance_mean", "distance_median", "distance_percentile_5", "distance_percentile_95"):
        assert rel_err(ren[k], f32(GOLD["vr_" + k])) <= 1e-5, k
    inds, w_new = nmodels._ResampleWeightsFn.apply(w.contiguous(), d("rs_gumbel"), 0.0, 1.0)
    assert np.array_equal(inds.cpu().numpy(), GOLD["rs_inds"])       # bit-exact given the Gumbel noise
    assert rel_err(w_new, f32(GOLD["rs_w"])) <= 1e-5


def test_ide_and_ggx(cuda_device):
    d = lambda k: f32(GOLD[k]).to(cuda_device)
    assert rel_err(nnerf.generate_ide_fn(4)(d("ide_dirs"), d("ide_kinv")), f32(GOLD["ide4"])) <= 1e-5
    got5 = nnerf.generate_ide_fn(5)(d("ide_dirs"), d("ide_kinv"))
    assert rel_err(got5, torch.from_numpy(GOLD["ide5_f64"])) <= 1e-5
    samples = {k[6:]: d(k) for k in GOLD.files if k.startswith("ggx_s_")}
    material = {k[6:]: d(k) for k in GOLD.files if k.startswith("ggx_m_")}
    res = nru.integrate_reflect_rays("microfacet", False, material, samples, max_radiance=10000.0)
    for k in ("radiance_out", "irradiance", "indirect_occ"):
        assert rel_err(res[k], f32(GOLD["ggx_" + k])) <= 1e-5, k
