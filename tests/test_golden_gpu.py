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


# ------------------------------------------------------------------ SURVEY 8f-1 pieces (oracle_v2.npz)
def test_geometry_mask_losses_and_second_order_against_golden(cuda_device):
    from oracle import geometry as ogeo
    from neural_radiance_caching_b200 import geometry as ngeo, loss_utils as nloss
    from tests.golden import make_golden_v2 as mg2
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v2.npz"))
    d = lambda k: f32(gold[k]).to(cuda_device)
    # geometry losses
    W, N, NP = (d(k).requires_grad_(True) for k in ("gl_w", "gl_n", "gl_np"))
    total = sum(nloss.geometry_losses({"viewdirs": d("gl_vd")}, {"weights": W, "normals": N, "normals_pred": NP},
                                      *mg2.MULTS, 0.1))
    total.backward()
    assert abs(float(total.detach()) - float(gold["gl_terms"].sum())) <= 1e-5 * float(gold["gl_terms"].sum())
    assert rel_err(W.grad, f32(gold["gl_gw"])) <= 1e-5
    assert rel_err(N.grad, f32(gold["gl_gn"])) <= 1e-5
    assert rel_err(NP.grad, f32(gold["gl_gnp"])) <= 1e-5
    # mask losses
    A = d("ml_acc").requires_grad_(True)
    l = nloss.compute_mask_loss(A, None, 0.001, 1.0, 1.0)
    l.backward()
    assert abs(float(l.detach()) - float(gold["ml_loss"])) <= 1e-5 * float(gold["ml_loss"])
    assert rel_err(A.grad, f32(gold["ml_g"])) <= 1e-5
    A.grad = None
    l = nloss.compute_mask_loss(A, None, 0.001, empty_loss_weight=0.1, backward=True)
    l.backward()
    assert abs(float(l.detach()) - float(gold["mlb_loss"])) <= 1e-5 * float(gold["mlb_loss"])
    assert rel_err(A.grad, f32(gold["mlb_g"])) <= 1e-5
    # distortion loss
    Wd = d("dl_w").requires_grad_(True)
    l = nloss.distortion_loss([{"tdist": d("dl_t"), "weights": Wd}], 0.01, -0.25, 10000.0)
    l.backward()
    assert abs(float(l.detach()) - float(gold["dl_loss"])) <= 2e-5 * float(gold["dl_loss"])
    assert rel_err(Wd.grad, f32(gold["dl_g"])) <= 2e-5
    # second-order path
    o = ogeo.DensityMLP(grid_params=mg2.GRID, enable_pred_normals=True)
    po = o.init(np.random.Generator(np.random.PCG64(mg2.SEED + 1)), table_init_range=0.5, bias_range=0.1)
    n = ngeo.DensityMLP(grid_params=mg2.GRID, enable_pred_normals=True)
    pn = n.from_oracle(po, cuda_device)
    arena = pn["density_grid"]["_arena"].requires_grad_(True)
    for k in ("density_layers_0", "density_layers_1", "output_density_layer", "pred_normals_layer"):
        for kk in pn[k]:
            pn[k][kk].requires_grad_(True)
    rg = n.raw_grad_density(pn, d("so_means"))
    assert rel_err(rg, f32(gold["so_raw_grad"])) <= 1e-5
    (rg * d("so_G")).sum().backward()
    for k in ("density_layers_0", "density_layers_1", "output_density_layer"):
        assert rel_err(pn[k]["kernel"].grad, f32(gold[f"so_d_{k}"])) <= 2e-5, k
    names = sorted(po["density_grid"].keys())
    views = n.grid.views(arena.grad)
    got = torch.cat([views[k].reshape(-1) for k in names])
    assert rel_err(got, f32(gold["so_d_tables"])) <= 2e-5
