"""The CUDA kernels against the REFERENCE'S OWN SOURCE: tests/golden/reference_np.npz (outputs of the reference's
functions under the NumPy stand-in for jax, see tests/test_reference_vectors.py) compared with the C-ABI entry points
directly - no oracle in between.  Corner indices / interpolation / contraction bit-exact, the rest fp32 rel 1e-5
(BASELINE north star), per-level conditioning notes as in tests/test_ray_gpu.py."""
import os

import numpy as np
import pytest
import torch

from neural_radiance_caching_b200 import coord as ncoord, grid_utils as ng, render as nrender, stepfun as nstep
from tests.util import rel_err

pytestmark = pytest.mark.gpu
V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_np.npz"))


def D(key, dev):
    return torch.from_numpy(V[key]).to(dev).contiguous()


def test_contract_bit_exact(cuda_device):
    x = D("coord_x", cuda_device)
    assert np.array_equal(ncoord.contract(x).cpu().numpy(), V["coord_contract"])
    assert np.array_equal(ncoord.contract_radius_2(x).cpu().numpy(), V["coord_contract_radius_2"])
    assert np.array_equal(ncoord.contract_radius_5(x).cpu().numpy(), V["coord_contract_radius_5"])


def test_trilerp_dense_bit_exact(cuda_device):
    got = ng.trilerp(D("grid_dense", cuda_device), D("grid_loc", cuda_device) * 16.0, "grid")
    assert np.array_equal(got.cpu().numpy(), V["grid_dense_16"])


@pytest.mark.parametrize("res", [64, 256])
def test_trilerp_hash_bit_exact(cuda_device, res):
    """One hash level (N^3 > T) of HashEncoding whose bbox maps voxel units back to [0,1] (exact for power-of-two N):
    the spatial hash, the uint32 wrap-around and the interpolation order of jax_hash_resample_3d."""
    table = D("grid_table", cuda_device)
    enc = ng.HashEncoding(hash_map_size=table.shape[0], num_features=table.shape[1], scale_supersample=1.0,
                          min_grid_size=res, max_grid_size=res, precondition_scaling=1.0,
                          bbox_scaling=((0.0, 0.0, 0.0), (float(res),) * 3))
    got = ng._EncodeFn.apply(enc, D("grid_loc", cuda_device) * float(res), False, table)
    assert np.array_equal(got.cpu().numpy(), V[f"grid_hash_{res}"])


def test_sample_intervals(cuda_device):
    # logits = 0.4 * safe_log(w + 1e-5) in the generator = the sampler's annealed logits (sampling.py:340)
    got = nstep.sample_intervals_from_weights(D("step_u01", cuda_device), D("step_t", cuda_device), D("step_w", cuda_device),
                                              32, anneal=0.4, padding=1e-5, domain=(0.0, 1.0))
    assert float((got.cpu() - torch.from_numpy(V["step_sample_intervals"])).abs().max()) <= 1e-5


def test_alpha_weights_and_rendering(cuda_device):
    tm, dens, dirs = D("dist_t", cuda_device), D("render_density", cuda_device), D("render_dirs", cuda_device)
    for op in (0, 1):
        w = nrender.compute_alpha_weights(dens, tm, dirs, opaque_background=bool(op))[0]
        assert rel_err(w, torch.from_numpy(V[f"render_weights_{op}"])) <= 1e-5
    wn = D("dist_w", cuda_device)
    vr = nrender.volumetric_rendering(D("render_rgbs", cuda_device), wn, wn, tm, D("render_bg", cuda_device), True)
    for k in ("rgb", "acc", "distance_mean", "distance_median", "distance_percentile_5", "distance_percentile_95"):
        assert rel_err(vr[k], torch.from_numpy(V["render_vr_" + k])) <= 2e-5, k


def test_cast_rays_means(cuda_device):
    from neural_radiance_caching_b200.sampling import ProposalVolumeSampler

    # the kernel takes normalised fenceposts and the ray warp; with near = 0, far = 1 and the linear warp s == t
    tm = D("dist_t", cuda_device)
    R = tm.shape[0]
    rays = dict(origins=D("render_origins", cuda_device), directions=D("render_dirs", cuda_device),
                near=torch.zeros((R, 1), device=cuda_device), far=torch.ones((R, 1), device=cuda_device))
    t_n, means = ProposalVolumeSampler()._cast(tm, rays, False)
    assert np.array_equal(t_n.cpu().numpy(), V["dist_t"])
    assert rel_err(means, torch.from_numpy(V["render_means"])) <= 1e-6


def test_cast_rays_covs(cuda_device):
    """nrc_ray_cast_covs against render.cast_rays of the reference (render.py:26-131): cone / cylinder, full / diagonal,
    and the sampler's `covs` entry on request."""
    from neural_radiance_caching_b200 import render as nrender
    tm = D("dist_t", cuda_device)
    o, dirs = D("render_origins", cuda_device), D("render_dirs", cuda_device)
    m, c = nrender.cast_rays(tm, o, dirs, D("render_radii", cuda_device), "cone", diag=False)
    assert rel_err(m, torch.from_numpy(V["render_means"])) <= 1e-6 and rel_err(c, torch.from_numpy(V["render_covs"])) <= 2e-6
    assert rel_err(c, c.transpose(-1, -2)) <= 1e-6    # d (d / |d|^2)^T is symmetric up to rounding, in the reference as here
    rv = D("render_radii_v", cuda_device)
    assert rel_err(nrender.cast_rays(tm, o, dirs, rv, "cone", diag=True)[1], torch.from_numpy(V["render_covs_diag"])) <= 2e-6
    cm, cc = nrender.cast_rays(tm, o, dirs, rv, "cylinder", diag=False)
    assert rel_err(cm, torch.from_numpy(V["render_cyl_means"])) <= 1e-6 and rel_err(cc, torch.from_numpy(V["render_cyl_covs"])) <= 2e-6
    assert rel_err(nrender.cast_rays(tm, o, dirs, rv, "cylinder", diag=True)[1], torch.from_numpy(V["render_cyl_covs_diag"])) <= 2e-6
    assert nrender.cast_rays(tm, o, dirs, rv, "cone", want_covs=False)[1] is None
    with pytest.raises(ValueError):
        nrender.cast_rays(tm, o, dirs, rv, "sphere")


def _history(dev):
    return [dict(sdist=D("step_t", dev), weights=D("loss_w0", dev)), dict(sdist=D("blur_tq", dev), weights=D("loss_w1", dev)),
            dict(sdist=D("blur_t", dev), weights=D("dist_w", dev), tdist=D("dist_t", dev))]


def test_proposal_losses(cuda_device):
    """nrc_interlevel_loss / nrc_distortion_loss against loss_utils.spline_interlevel_loss (internal/loss_utils.py:74-108)
    and distortion_loss through power_ladder(-0.25, 1e4) (:108-123) as run from the reference's source."""
    from neural_radiance_caching_b200 import loss_utils as nloss

    hist = _history(cuda_device)
    got = torch.stack(nloss.spline_interlevel_loss(hist, mults=(0.01, 0.01), blurs=(0.03, 0.003))).cpu()
    assert rel_err(got, torch.from_numpy(V["loss_spline_interlevel"])) <= 2e-5
    # the curve compresses [2, 6] into a 0.03-wide interval: fp32 cancellation in either implementation
    d = nloss.distortion_loss(hist, mult=0.01, p=-0.25, premult=1e4, target="tdist").cpu()
    assert abs(float(d) - float(V["loss_distortion"])) <= 3e-4 * float(V["loss_distortion"])


def test_ggx_integration(cuda_device):
    """nrc_ggx_integrate_fwd against integrate_reflect_rays('microfacet') of the reference's render_utils.py:1102-1193
    (GGX_D :480-482, get_lobe :566-695), incl. below-horizon samples, negative weights and near-zero pdfs."""
    from neural_radiance_caching_b200.inverse_render import render_utils as nru

    material = {k: D("ggx_mat_" + k, cuda_device) for k in ("albedo", "roughness", "F_0", "metalness")}
    samples = {k: D("ggx_smp_" + k, cuda_device) for k in ("local_lightdirs", "local_viewdirs", "pdf", "weight",
                                                           "radiance_in", "indirect_occ")}
    res = nru.integrate_reflect_rays("microfacet", False, material, samples)
    for k in ("radiance_out", "indirect_occ", "irradiance"):
        assert rel_err(res[k], torch.from_numpy(V["ggx_int_" + k])) <= 2e-5, k


@pytest.mark.parametrize("tag", ["a", "b"])
def test_hash_encoding_call_bit_exact(cuda_device, tag):
    """nrc_encode_fwd through HashEncoding.__call__ against the reference's own class (internal/grid_utils.py:738-905):
    dense and hash levels, cubic and non-cubic bbox, multisample mean, precondition scaling, points outside the box."""
    from tests.util import ENC_CONFIGS, level_table

    enc = ng.HashEncoding(**ENC_CONFIGS[tag])
    layout = enc.level_layout
    assert [int(n) for n in enc.grid_sizes] == [int(n) for n in V[f"enc_{tag}_grid_sizes"]]
    assert [name for (name, _, _, _) in layout] == list(V[f"enc_{tag}_names"])
    params = {name: torch.from_numpy(level_table(shape, i + 1)).to(cuda_device)
              for i, (name, _, _, shape) in enumerate(layout)}
    got = enc(params, D("enc_x", cuda_device), per_level_fn=lambda f: f.mean(dim=-2))
    assert np.array_equal(got.cpu().numpy(), V[f"enc_{tag}_features"])
