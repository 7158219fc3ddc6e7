"""The oracle against the REFERENCE'S OWN SOURCE.  tests/golden/reference_np.npz holds outputs of the functions of
/root/reference/internal/{math,coord,stepfun,render,grid_utils,ref_utils,image,loss_utils}.py and
internal/inverse_render/render_utils.py, executed in the build container under a
NumPy stand-in for jax (tests/golden/make_reference_vectors.py + jax_numpy_shim.py; JAX itself is not installable
there), on seeded float32 inputs stored in the same file.  This pins the restatement in oracle/ to the reference's code:
corner indices / trilinear interpolation / contraction / cast_rays / l2_normalize BIT-EXACT, everything else to float32
rounding (different primitive implementations: NumPy vs PyTorch reductions and libm).  What it cannot pin is XLA's own
rounding inside a primitive - DESIGN.md section 3 lists those assumptions."""
import os

import numpy as np
import pytest
import torch

from oracle import coord as ocoord, grid_utils as ogrid, loss_utils as oloss, nerf as onerf, ref_math, render as orender
from oracle import stepfun as ostep
from tests.util import ENC_CONFIGS, level_table as _level_table

V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_np.npz"))
T = lambda k: torch.from_numpy(V[k])


def _np(x):
    return x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def exact(got, key):
    assert np.array_equal(_np(got), V[key], equal_nan=True), key


def close(got, key, tol, per_element=False):
    got, want = _np(got).astype(np.float64), V[key].astype(np.float64)
    assert got.shape == want.shape, key
    assert np.array_equal(np.isfinite(got), np.isfinite(want)), key
    m = np.isfinite(want)
    err = np.abs(got[m] - want[m])
    scale = np.maximum(np.abs(want[m]), 1e-30) if per_element else max(np.abs(want[m]).max(), 1e-30)
    assert float((err / scale).max()) <= tol, (key, float((err / scale).max()))


def test_math():
    x = T("math_x")
    close(ref_math.safe_exp(x), "math_safe_exp", 1e-6, per_element=True)     # internal/math.py:186-192
    close(ref_math.safe_log(x.abs()), "math_safe_log", 1e-6)                 # :177-183
    exact(ref_math.safe_sign(x), "math_safe_sign")                           # :122-124
    for p, pre in ((-1.5, 2.0), (-0.25, 1e4)):                               # :295-341 (both configured ladders)
        close(ref_math.power_ladder(T(f"math_pl_x_{p}"), p, premult=pre), f"math_pl_y_{p}", 1e-6)
        close(ref_math.inv_power_ladder(T(f"math_pl_y_{p}"), p, premult=pre), f"math_ipl_{p}", 1e-6)


def test_coord_contract_bit_exact():
    pts = T("coord_x")
    exact(ocoord.contract(pts), "coord_contract")                            # internal/coord.py:63-69
    exact(ocoord.contract_radius(pts, 2.0), "coord_contract_radius_2")       # :37-38
    exact(ocoord.contract_radius(pts, 5.0), "coord_contract_radius_5")       # :33-34


def test_stepfun():
    t, lg, u01 = T("step_t"), T("step_logits"), T("step_u01")
    close(ostep.integrate_weights(torch.softmax(lg, -1)), "step_integrate_weights", 1e-5)           # stepfun.py:125-144
    # the CDF inversion amplifies last-ulp differences of softmax / cumsum (NumPy vs PyTorch) into ~1e-6 shifts
    close(ostep.sample_intervals(u01, t, lg, 32, single_jitter=True, domain=(0.0, 1.0)), "step_sample_intervals", 1e-5)
    close(ostep.sample(u01, t, lg, 32, single_jitter=True), "step_sample_centres", 1e-5)             # :158-204
    tm, wn = T("dist_t"), T("dist_w")
    close(oloss.lossfun_distortion(tm, wn), "step_lossfun_distortion", 1e-6)                         # :253-269
    close(ostep.weighted_percentile(tm, wn, [5, 50, 95]), "step_weighted_percentile", 1e-5)          # :306-314
    for hw in (0.03, 0.003):   # :463-483 over linspline.py: piecewise-quadratic evaluation with fp32 cancellation
        close(oloss.blur_and_resample_weights(T("blur_tq"), T("blur_t"), wn, hw), f"step_blur_and_resample_{hw}", 2e-4)


def test_render():
    tm, wn, dens, dirs = T("dist_t"), T("dist_w"), T("render_density"), T("render_dirs")
    for op in (0, 1):                                                                                # render.py:134-169
        w, a, tr = orender.compute_alpha_weights(dens, tm, dirs, opaque_background=bool(op))[:3]
        close(w, f"render_weights_{op}", 1e-6)
        close(a, f"render_alpha_{op}", 1e-6)
        close(tr, f"render_trans_{op}", 1e-6)
    means, _ = orender.cast_rays(tm, T("render_origins"), dirs, torch.full((64, 1), 5e-4), "cone", diag=False)
    exact(means, "render_means")                                                                     # :26-131
    # covariances (full and diagonal; cone and cylinder rays with per-ray radii)
    _, covs = orender.cast_rays(tm, T("render_origins"), dirs, T("render_radii"), "cone", diag=False)
    close(covs, "render_covs", 1e-6)
    rv = T("render_radii_v")
    close(orender.cast_rays(tm, T("render_origins"), dirs, rv, "cone", diag=True)[1], "render_covs_diag", 1e-6)
    cm, cc = orender.cast_rays(tm, T("render_origins"), dirs, rv, "cylinder", diag=False)
    close(cm, "render_cyl_means", 1e-6)
    close(cc, "render_cyl_covs", 1e-6)
    close(orender.cast_rays(tm, T("render_origins"), dirs, rv, "cylinder", diag=True)[1], "render_cyl_covs_diag", 1e-6)
    vr = orender.volumetric_rendering(T("render_rgbs"), wn, wn, tm, T("render_bg"), True)            # :172-247
    for k in ("rgb", "acc", "distance_mean", "distance_median", "distance_percentile_5", "distance_percentile_95"):
        close(vr[k], "render_vr_" + k, 1e-6)


def test_trilerp_bit_exact():
    """Hash indices (spatial hash, uint32 wrap-around) and dense padded-volume indices + interpolation order."""
    loc = T("grid_loc")
    for res in (64, 256):
        exact(ogrid.trilerp(T("grid_table"), loc * res, "hash"), f"grid_hash_{res}")   # grid_utils.py:41-121, 679-726
    exact(ogrid.trilerp(T("grid_dense"), loc * 16, "grid"), "grid_dense_16")           # :352-445


def test_ref_utils():
    exact(ref_math.l2_normalize(T("ref_l2n_x")), "ref_l2n")                            # ref_utils.py:45-70
    # integrated directional encoding (ref_utils.py:131-192); the degree-5 basis has l = 16 harmonics whose fp32
    # evaluation carries ~1e-4 noise in the reference itself
    close(onerf.generate_ide_fn(4)(T("ide_dirs"), T("ide_kappa_inv")), "ref_ide_4", 1e-5)
    close(onerf.generate_ide_fn(5)(T("ide_dirs"), T("ide_kappa_inv")), "ref_ide_5", 2e-4)


def test_image_and_losses():
    from oracle import light_sampler as olight

    close(olight.linear_to_srgb(T("image_linear")), "image_linear_to_srgb", 1e-6)                    # image.py:192-200
    hist = [dict(sdist=T("step_t"), weights=T("loss_w0")), dict(sdist=T("blur_tq"), weights=T("loss_w1")),
            dict(sdist=T("blur_t"), weights=T("dist_w"), tdist=T("dist_t"))]
    got = torch.stack(oloss.spline_interlevel_loss(hist, mults=(0.01, 0.01), blurs=(0.03, 0.003)))   # loss_utils.py:74-108
    close(got, "loss_spline_interlevel", 1e-4)
    # loss_utils.py:108-123 with curve_fn = power_ladder(-0.25, 1e4): the curve compresses [2, 6] into a 0.03-wide
    # interval, so the fp32 loss carries cancellation noise in either implementation
    close(oloss.distortion_loss(hist, mult=0.01, p=-0.25, premult=1e4, target="tdist"), "loss_distortion", 2e-4)


def test_ggx_and_frames():
    from oracle import material as omat, render_utils as oru

    close(oru.GGX_D(T("ggx_costheta"), T("ggx_a")), "ggx_D", 1e-5, per_element=True)   # render_utils.py:480-482
    material = {k: T("ggx_mat_" + k) for k in ("albedo", "roughness", "F_0", "metalness")}
    samples = {k: T("ggx_smp_" + k) for k in ("local_lightdirs", "local_viewdirs", "brdf_correction", "pdf", "weight",
                                               "radiance_in", "indirect_occ")}
    res = oru.integrate_reflect_rays("microfacet", material, samples)                  # :566-695, :1102-1193
    for k in ("radiance_out", "indirect_occ", "irradiance"):
        close(res[k], "ggx_int_" + k, 1e-5)
    close(omat.get_rotation_matrix(T("rot_normal")), "rot_matrix", 1e-6)               # :145-168
    close(omat.eval_vmf(T("vmf_x"), T("vmf_means"), T("vmf_kappa")), "vmf_eval", 1e-5, per_element=True)   # :1335-1346


def test_importance_samplers_and_light_loss():
    from oracle import light_sampler as olight, material as omat

    u1, u2, wo, alpha, wi = T("smp_u1"), T("smp_u2"), T("smp_wo"), T("smp_alpha"), T("smp_wi")
    for name, smp in (("cosine", omat.CosineSampler()), ("microfacet", omat.MicrofacetSampler())):   # render_utils.py:417-546
        dirs, pdf = smp.sample_directions(u1, u2, wo, alpha, None)
        close(dirs, f"smp_{name}_dirs", 1e-5)
        close(pdf, f"smp_{name}_pdf", 1e-5)
        close(smp.pdf(wo, wi, alpha, None), f"smp_{name}_pdf_of_wi", 1e-5)
    aux = dict(vmf_means=T("light_means"), vmf_kappas=T("light_kappas"), vmf_logits=T("light_logits"))
    close(omat.LightSampler().pdf(wo, wi, alpha, aux), "smp_light_pdf_of_wi", 1e-5)                # :1470-1490
    v = (aux["vmf_means"], aux["vmf_kappas"], aux["vmf_logits"])
    for srgb in (True, False):                                                                       # :1493-1550
        got = olight.vmf_loss_fn(v, T("light_normals"), wi, T("light_pdf")[..., 0], T("light_weight")[..., 0], T("light_fv"),
                                 T("light_lossmult"), srgb=srgb)
        close(got, f"light_vmf_loss_{int(srgb)}", 1e-5)


def test_transient():
    """render.volumetric_transient_rendering (:250-449), shift_direct (:452-490: flat-index splat - a bin past the end of
    an inner ray lands in the next ray, past the end of the array it is dropped), shift_map_coordinates (:493-507),
    render_utils.zero_invalid_bins (:1699-1767)."""
    from oracle import transient as otr

    w, direct, ind = T("tr_weights"), T("tr_direct"), T("tr_indirect")
    ray, light = T("tr_ray_dists")[..., 0], T("tr_light_dists")[..., 0]
    R, n, B, C = ind.shape
    close(otr.shift_direct((ray + light) / 0.01, direct, w, B, C), "tr_shift_direct", 1e-6)
    close(otr.shift_map_coordinates(ind.reshape(-1, B, C), ray.reshape(-1), 0.01, B), "tr_shift_map", 1e-6)
    for tag, shift in (("tr_s0_", 0.0), ("tr_s1_", 0.0137)):
        res = otr.volumetric_transient_rendering(direct, ind, w, ray, light, B, exposure_time=0.01, shift=shift, dark_level=0.001)
        for k in ("transient_direct", "transient_indirect", "rgb"):
            close(res[k], tag + k, 2e-6)
    means = T("tr_means")
    light_d = torch.linalg.norm(T("tr_lights")[:, None, :] - means, dim=-1, keepdim=True)
    cam_d = (torch.linalg.norm(T("tr_origins")[:, None, :] - means, dim=-1, keepdim=True)
             + torch.linalg.norm(T("tr_origins") - T("tr_cam_origins"), dim=-1, keepdim=True)[:, None, :])
    for lz in (False, True):
        zd, zs = otr.zero_invalid_bins(ind, T("tr_spec"), light_d, cam_d, B, 0.01, 2.0, lz, 0.12)
        exact(zd, f"tr_zero_diffuse_{int(lz)}")
        exact(zs, f"tr_zero_specular_{int(lz)}")


def test_maybe_resample():
    """Model.maybe_resample (internal/models.py:193-292), resample_argmax off: categorical draw = argmax(logits + Gumbel),
    indices BIT-EXACT, importance-corrected weights, gathered points / features."""
    from oracle import models as omodels

    w, gum = T("rs_weights"), T("rs_gumbel")
    for k, bias in ((1, 0.0), (4, 1e-3)):
        inds, nw = omodels.maybe_resample(w, gum[..., :k], k, weights_bias=bias)
        exact(inds.to(torch.int32), f"rs_inds_{k}")
        close(nw, f"rs_new_weights_{k}", 1e-6)
        take = lambda x: torch.gather(x, 1, inds[..., None].expand(inds.shape + (x.shape[-1],)))
        exact(take(T("rs_points")), f"rs_new_points_{k}")
        exact(take(T("rs_feature")), f"rs_new_feature_{k}")


def test_geometry_and_mask_losses():
    """orientation_loss / predicted_normal_loss (internal/loss_utils.py:127-199) in the three configured uses
    (train_utils.py:1027-1093) and compute_mask_loss (train_utils.py:785-836) incl. the backward-mask call (:2929-2945)."""
    geo = dict(weights=T("gl_weights"), normals=T("gl_normals"), normals_pred=T("gl_normals_pred"))
    rays = dict(viewdirs=T("gl_viewdirs"))
    lo, lp, lr = oloss.geometry_losses(rays, geo, orientation_mult=0.01, predicted_normal_mult=0.001,
                                       predicted_normal_reverse_mult=0.01, stopgrad_weight=0.1)
    close(lo, "gl_orientation", 1e-6)
    close(lp, "gl_predicted_normal", 1e-6)
    close(lr, "gl_predicted_normal_reverse", 1e-6)
    acc, masks = T("gl_acc"), T("gl_masks")
    close(oloss.compute_mask_loss(acc, masks, 0.001, 1.0, 10.0), "gl_mask_loss", 1e-6)
    close(oloss.compute_mask_loss(acc, None, 0.001, 1.0, 10.0), "gl_mask_loss_none", 1e-6)
    close(oloss.compute_mask_loss(acc, torch.zeros_like(masks), 0.001, 1.0, 0.5, backward=True), "gl_mask_loss_backward", 1e-6)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_hash_encoding_call(tag):
    """HashEncoding.__call__ (internal/grid_utils.py:738-905) executed from the reference's class: level schedule, dense /
    hash choice per level, parameter names, bbox map (cubic and non-cubic), per-level multisample mean, precondition
    scaling - features BIT-EXACT."""
    enc = ogrid.HashEncoding(**ENC_CONFIGS[tag])
    assert list(enc.grid_sizes) == list(V[f"enc_{tag}_grid_sizes"])
    assert enc.param_names() == list(V[f"enc_{tag}_names"])
    params = {name: torch.from_numpy(_level_table(shape, i + 1))
              for i, (name, (_, _, shape)) in enumerate(zip(enc.param_names(), enc.layout))}
    exact(enc(params, T("enc_x"), per_level_mean=True), f"enc_{tag}_features")


@pytest.mark.parametrize("ns", [2, 3])
def test_importance_sample_rays(ns):
    """importance_sample_rays (render_utils.py:722-924) run from the reference with pre-drawn randoms: Microfacet (16) +
    Cosine (8) [+ vMF-mixture Light (8)] samplers, shading frame, MIS power heuristic and the energy correction."""
    from oracle import material as omat

    counts = (16, 8, 8)[:ns]
    samplers = [(omat.MicrofacetSampler(), 16), (omat.CosineSampler(), 8), (omat.LightSampler(), 8)][:ns]
    uniforms = [(T(f"is_uh_{j}"), T(f"is_uw_{j}")) for j in range(ns)]
    P = V["is_viewdirs"].shape[0]
    aux = None
    if ns == 3:
        aux = dict(vmf_means=T("light_means")[:P], vmf_kappas=T("light_kappas")[:P], vmf_logits=T("light_logits")[:P],
                   latent=T("is_light_latent").long(), normal2=T("is_light_normal2"), u=T("is_light_u"))
    res = omat.importance_sample_rays(T("is_viewdirs"), T("is_normals"), T("is_roughness"), samplers, uniforms, aux=aux)
    assert res["pdf"].shape[1] == sum(counts)
    for k in ("local_lightdirs", "local_viewdirs", "global_lightdirs"):
        close(res[k], f"is{ns}_{k}", 2e-6)
    # GGX pdfs at low roughness are ill conditioned in fp32: relative to the largest pdf / weight
    close(res["pdf"], f"is{ns}_pdf", 1e-5)
    close(res["weight"], f"is{ns}_weight", 1e-5)


def test_density_mlp():
    """DensityMLP.predict_density / run_network / convert_raw_density (internal/geometry.py:155-341) executed from the
    reference's class (flax Dense = x @ kernel + bias) in the configured shape: contract_radius_2 -> HashEncoding with the
    multisample mean -> 2 x 64 ReLU -> 1, safe_exp(raw - 1).  Encoding bit-exact (test_hash_encoding_call); the matmuls
    differ by summation order (OpenBLAS vs MKL sgemm)."""
    from oracle import geometry as ogeo
    from tests.util import dense_params

    mlp = ogeo.DensityMLP({k: v for k, v in ENC_CONFIGS["a"].items() if k != "scale_supersample"}, net_depth=2, net_width=64,
                          density_bias=-1.0, warp_c=2.0, bbox_scaling=2.0)
    p = {"density_grid": {name: torch.from_numpy(_level_table(shape, i + 1))
                          for i, (name, (_, _, shape)) in enumerate(zip(mlp.grid.param_names(), mlp.grid.layout))}}
    d_in = mlp.in_dim
    for i, (name, d_out) in enumerate([("density_layers_0", 64), ("density_layers_1", 64), ("output_density_layer", 1)]):
        k, b = dense_params(d_in, d_out, 100 + i)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
        d_in = d_out
    means = T("dmlp_means")
    raw, feat = mlp.predict_density(p, means)
    close(feat, "dmlp_feature", 2e-6)
    close(raw, "dmlp_raw_density", 2e-6)
    close(mlp.convert_raw_density(raw, means), "dmlp_density", 2e-6)


class _Field:
    """Closed-form rational density field of the generator (IEEE +, *, / only) standing in for a DensityMLP."""
    normals_for_filter_only = True
    disable_density_normals = True

    def __init__(self, scale, k, c):
        self.scale, self.k, self.c = scale, k, c

    def __call__(self, p, means, **kw):
        dx, dy, dz = means[..., 0] - self.c[0], means[..., 1] - self.c[1], means[..., 2] - self.c[2]
        return dict(density=self.scale / (1.0 + self.k * (dx * dx + dy * dy + dz * dz)))


@pytest.mark.parametrize("tag,use_raydist", [("id", False), ("pl", True)])
def test_proposal_sampler_loop(tag, use_raydist):
    """ProposalVolumeSampler.__call__ (internal/sampling.py:142-649) executed from the reference's class with the configured
    strategy (64, 64, 32), annealing, padding and ray warps (identity / power_ladder(-1.5, 2)); the density MLPs replaced by
    the same closed-form fields on both sides, so the level loop itself is what is compared."""
    from oracle import sampling as osamp

    s = osamp.ProposalVolumeSampler()
    s.mlps = [_Field(4.0, 3.0, (0.1, -0.2, 0.3)), _Field(9.0, 6.0, (0.0, -0.1, 0.2)), _Field(40.0, 14.0, (0.05, -0.15, 0.25))]
    rays = {k: T("pvs_" + k) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")}
    hist = s({f"MLP_{i}": None for i in range(3)}, rays, [T("pvs_u01")] * 3, use_raydist_fn=use_raydist)
    # the CDF inversion amplifies last-ulp differences (NumPy vs PyTorch softmax / cumsum, libm pow in the ladder) level by
    # level: tolerances are relative to the ray length / largest weight
    for lvl, h in enumerate(hist):
        for k in ("sdist", "tdist", "means", "weights"):
            tol = 1e-4 if k == "weights" else (2e-6 if lvl == 0 else 2e-5)     # measured: 1e-6 / 8e-6 / 3e-5 (weights)
            close(h[k], f"pvs_{tag}_{lvl}_{k}", tol)


def test_cache_shader_pieces():
    """BaseShader.predict_appearance_feature with net_depth 0 (internal/shading.py:133-220: [density feature | appearance grid
    of the contracted mean]) BIT-EXACT; NeRFMLP.get_integrated_brdf (internal/nerf.py:423-434,461-482) and _get_refdirs
    (:1344-1358) executed from the reference's classes."""
    from oracle import nerf as onerf2
    from tests.util import dense_params

    sh = onerf2.NeRFMLP(warp_c=2.0)
    sh.grid = ogrid.HashEncoding(hash_map_size=2 ** 15, num_features=4, scale_supersample=1.0, max_grid_size=256, bbox_scaling=2.0)
    p = {"appearance_grid": {name: torch.from_numpy(_level_table(shape, i + 1))
                             for i, (name, (_, _, shape)) in enumerate(zip(sh.grid.param_names(), sh.grid.layout))}}
    exact(sh.predict_appearance_feature(p, T("shd_density_feature"), T("shd_means")), "shd_appearance_feature")
    d_in = 129
    for i, (name, d_out) in enumerate([("integrated_brdf_layers_0", 64), ("integrated_brdf_layers_1", 64),
                                       ("output_integrated_brdf_layer", 1)]):
        k, b = dense_params(d_in, d_out, 200 + i)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
        d_in = d_out
    close(sh.get_integrated_brdf(p, T("shd_normals"), T("shd_viewdirs"), T("shd_bottleneck")), "shd_integrated_brdf", 2e-6)
    close(onerf2.reflect(-T("shd_viewdirs")[..., None, :], T("shd_normals")), "shd_refdirs", 1e-6)


def test_light_mlp_get_vmfs():
    """LightMLP.get_vmfs (internal/light_sampler.py:135-160) executed from the reference's class with its own bias and
    activation tables (:73-84): means scale + pre-drawn offsets, softplus kappas clamped at 50, logits clamped at -50."""
    from oracle import light_sampler as olight

    means_random = T("vmfs_normal") * 20.0 / 2.0
    got = olight.get_vmfs(T("vmfs_raw"), means_random)
    close(got["vmf_means"], "vmfs_vmf_means", 1e-6)
    close(got["vmf_kappas"], "vmfs_vmf_kappas", 1e-6)
    exact(got["vmf_logits"], "vmfs_vmf_logits")


def test_microfacet_material_head():
    """MaterialMLP._get_microfacet_material (internal/material.py:1276-1322) over the class's own property table
    (:957-1023) with the fields of configs/ngp_yobo.gin:256-303: slices, biases, sigmoid, the roughness floor, constant
    Fresnel, diffuseness / mirrorness off."""
    from oracle import material as omat

    got = omat.microfacet_material(T("mat_brdf_params"), min_roughness=0.01, default_F_0=0.04)
    for k in ("albedo", "specular_albedo", "roughness", "F_0", "metalness", "diffuseness", "mirrorness"):
        close(got[k], "mat_" + k, 1e-6)


def test_transient_head():
    """TransientNeRFMLP._compute_indirect_lighting + get_indirect (internal/nerf.py:1660-1777) executed from the reference's
    class: pos_enc(lights) -> irradiance stack -> n_bins*3 bins, softplus(. - 2) * scale; tint * F * ref_rgb * scale;
    zero_invalid_bins; clip; per-bin and summed outputs."""
    from oracle import coord as ocoord, geometry as ogeo, nerf as onerf2, transient as otr
    from tests.util import dense_params

    feat, means, nrm, ref = T("th_feature"), T("th_means"), T("th_normals"), T("th_ref_rgb")
    R, n, B = ref.shape[0], ref.shape[1], ref.shape[2] // 3
    p, names = {}, (("irradiance_layers_0", 64), ("irradiance_layers_1", 64), ("transient_indirect_layer", B * 3))
    for group, d_in, salt in ((names, 96 + 15, 300), ((("integrated_brdf_layers_0", 64), ("integrated_brdf_layers_1", 64),
                                                        ("output_integrated_brdf_layer", 1)), 129, 200), ((("tint_layer", 3),), 96, 320)):
        for i, (name, d_out) in enumerate(group):
            k, b = dense_params(d_in, d_out, salt + i)
            p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
            d_in = d_out
    lights = T("th_lights")[:, None, :] * torch.ones_like(nrm)
    x = torch.cat([feat, ocoord.pos_enc(lights, 0, 2, True)], dim=-1)                       # get_indirect :1757-1777
    x = torch.relu(ogeo.dense(p["irradiance_layers_0"], x))
    x = torch.relu(ogeo.dense(p["irradiance_layers_1"], x))
    diffuse_raw = ogeo.dense(p["transient_indirect_layer"], x).reshape(R, n, B, 3)
    F = onerf2.NeRFMLP().get_integrated_brdf(p, nrm, T("th_viewdirs"), T("th_bottleneck"))
    tint = torch.sigmoid(ogeo.dense(p["tint_layer"], feat))
    light_d = torch.linalg.norm(T("th_lights")[:, None, :] - means, dim=-1)
    cam_d = (torch.linalg.norm(T("th_origins")[:, None, :] - means, dim=-1)
             + torch.linalg.norm(T("th_origins") - T("th_cam_origins"), dim=-1)[:, None])
    diffuse, spec = otr.transient_head(diffuse_raw, ref.reshape(R, n, B, 3), tint * F, light_d, cam_d, B, exposure_time=0.01,
                                       diffuse_bias=-2.0, indirect_scale=0.7, bin_zero_threshold_light=2.0, rgb_max=1.5)
    close(diffuse, "th_transient_indirect_diffuse", 2e-6)
    close(spec, "th_transient_indirect_specular", 2e-6)
    close(diffuse + spec, "th_transient_indirect", 2e-6)
    close(diffuse.sum(-2), "th_indirect_diffuse", 2e-6)
    close(spec.sum(-2), "th_indirect_specular", 2e-6)


def test_light_sampling_loss_call_site():
    """train_utils.light_sampling_loss (internal/train_utils.py:1985-2071) executed from the reference: function values =
    |radiance_in|, lossmult / S, only the specular suffix present (multiplier 2, / 2 inside the loop)."""
    from oracle import light_sampler as olight

    vmfs = dict(vmf_means=T("light_means"), vmf_kappas=T("light_kappas"), vmf_logits=T("light_logits"),
                vmf_normals=T("light_normals")[:, None, :])
    for srgb in (True, False):
        got = olight.light_sampling_loss(vmfs, T("smp_wi"), T("light_pdf")[..., 0], T("light_weight")[..., 0],
                                         T("light_radiance_in"), srgb=srgb)
        close(got, f"light_sampling_loss_{int(srgb)}", 1e-5)


def test_density_mlp_module_call():
    """DensityMLP.__call__ -> predict_density_normals -> get_predict_density_kwargs -> coord.compute_control_points
    (internal/geometry.py:343-584, coord.py:568-611) executed from the reference's class on the branch without
    jax.value_and_grad: the 'mean' basis yields exactly one control point (= the mean), predicted-normals head,
    normals_to_use = normals_pred, ray distances."""
    from oracle import geometry as ogeo
    from tests.util import dense_params

    mlp = ogeo.DensityMLP({k: v for k, v in ENC_CONFIGS["a"].items() if k != "scale_supersample"}, net_depth=2, net_width=64,
                          density_bias=-1.0, warp_c=2.0, bbox_scaling=2.0, enable_pred_normals=True, disable_density_normals=True)
    p = {"density_grid": {name: torch.from_numpy(_level_table(shape, i + 1))
                          for i, (name, (_, _, shape)) in enumerate(zip(mlp.grid.param_names(), mlp.grid.layout))}}
    for name, d_in, d_out, salt in (("density_layers_0", mlp.in_dim, 64, 100), ("density_layers_1", 64, 64, 101),
                                    ("output_density_layer", 64, 1, 102), ("pred_normals_layer", 64, 3, 110)):
        k, b = dense_params(d_in, d_out, salt)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
    means = T("dmlp_means")[:696].reshape(58, 12, 3)
    res = mlp(p, means, viewdirs=T("dmlp_call_viewdirs"), origins=T("dmlp_call_origins"))
    assert res["normals"] is None and res["raw_grad_density"] is None
    for k in ("feature", "density", "grad_pred", "normals_pred", "normals_to_use", "ray_dists"):
        close(res[k], "dmlp_call_" + k, 2e-6 if k != "normals_pred" and k != "normals_to_use" else 1e-5)


def _slf_params(deg_view, use_bottleneck, salt):
    """Parameters of tests/golden/make_reference_vectors.py:make_slf (closed-form dense_params)."""
    from tests.util import dense_params
    n_in = (128 if use_bottleneck else 0) + {4: 38, 5: 72}[deg_view]
    p, d_in = {}, n_in
    for i, name in enumerate(("layer_0", "layer_1", "layer_2", "layer_bottleneck")):
        k, b = dense_params(d_in, 128, salt + i)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
        d_in = 128 + (n_in if (i % 2 == 0 and i > 0) else 0)
    k, b = dense_params(d_in, 3, salt + 10)
    p["output_ambient_rgb_layer"] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
    return p


def _shader_params():
    from tests.util import dense_params
    p = {}
    for name, d_in, d_out, salt in (("bottleneck_layer", 96, 128, 440), ("roughness_layer", 96, 1, 441),
                                    ("ambient_irradiance_layer", 96, 3, 442), ("irradiance_layer", 96, 3, 443),
                                    ("tint_layer", 96, 3, 444), ("integrated_brdf_layers_0", 129, 64, 445),
                                    ("integrated_brdf_layers_1", 64, 64, 455), ("output_integrated_brdf_layer", 64, 1, 465)):
        k, b = dense_params(d_in, d_out, salt)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
    p["SurfaceLightField"], p["EnvMap"] = _slf_params(5, True, 400), _slf_params(4, False, 420)
    return p


def test_surface_light_field_call():
    """SurfaceLightFieldMLP.__call__ (internal/surface_light_field.py:782-1069) EXECUTED FROM THE REFERENCE'S CLASS (its own
    setup(): encodings + every Dense) in the two configured shapes - `SurfaceLightField` (IDE_5 + shader bottleneck, skip
    after layer 2) and the shader-level `EnvMap` (IDE_4): incoming_ambient_rgb and incoming_acc == 1.  The degree-5 IDE
    is evaluated in float32 by both (the oracle follows the reference's op order), so the outputs agree to summation order."""
    from oracle import nerf as onerf2

    for tag, deg, use_b, salt, tol in (("slf5", 5, True, 400, 5e-6), ("slf4", 4, False, 420, 5e-6)):
        net = onerf2.SurfaceLightFieldMLP(deg, use_b)
        got = net(_slf_params(deg, use_b, salt), T("slf_refdirs"), T("slf_roughness"), T("slf_bottleneck") if use_b else None)
        close(got["incoming_ambient_rgb"], tag + "_incoming_ambient_rgb", tol)   # measured 3e-7 / 2e-7
        exact(got["incoming_acc"], tag + "_incoming_acc")


def test_predict_appearance_passive():
    """NeRFMLP.get_bottleneck_feature (internal/nerf.py:385-408), the roughness head (:633-634) and
    _predict_appearance_passive (:940-1090) EXECUTED FROM THE REFERENCE'S CLASS with both SurfaceLightFieldMLP instances
    attached, on the feature BaseShader.predict_appearance_feature (internal/shading.py:133-220) produced for an 8-level
    appearance grid: every output of the cache shader."""
    from oracle import nerf as onerf2

    sh = onerf2.NeRFMLP(warp_c=2.0)
    sh.grid = ogrid.HashEncoding(hash_map_size=2 ** 15, num_features=4, scale_supersample=1.0, max_grid_size=2048, bbox_scaling=2.0)
    p = _shader_params()
    p["appearance_grid"] = {name: torch.from_numpy(_level_table(shape, i + 1))
                            for i, (name, (_, _, shape)) in enumerate(zip(sh.grid.param_names(), sh.grid.layout))}
    feature = sh.predict_appearance_feature(p, T("shp_density_feature"), T("shp_means"))
    exact(feature, "shp_feature")
    bott, rough = sh.heads(p, feature)
    close(bott, "shp_bottleneck", 2e-6)
    close(rough, "shp_roughness", 2e-6)
    got = sh.predict_appearance_passive(p, feature, bott, rough, T("shp_normals"), T("shp_viewdirs"))
    for k in ("rgb", "diffuse_rgb", "specular_rgb", "ambient_rgb", "indirect_rgb", "albedo_rgb"):
        close(got[k], "shp_" + k, 2e-4 if k not in ("albedo_rgb", "diffuse_rgb") else 2e-6)
    whole = sh(p, T("shp_viewdirs"), T("shp_means"), T("shp_density_feature"), T("shp_normals"))
    close(whole["rgb"], "shp_rgb", 2e-4)


def fd_quantiles(got, key):
    """|got - central differences of the reference| relative to the largest gradient: (median, 65 % quantile)."""
    ref = torch.from_numpy(V[key]).double()
    d = (got.detach().double().cpu() - ref).abs().flatten() / float(ref.abs().max())
    return float(torch.quantile(d, 0.5)), float(torch.quantile(d, 0.65))


def test_analytic_normals_against_reference_differences():
    """Analytic normals (internal/geometry.py:442-460): the reference takes jax.value_and_grad of predict_density, which
    cannot run without JAX; its value d raw / d mean is pinned by central differences of the REFERENCE'S OWN predict_density
    (tests/golden/make_reference_vectors.py, h = 2^-11).  A difference that straddles a grid-cell face or a ReLU kink
    averages two slopes, hence quantiles: a wrong Jacobian (contraction, bbox map, level scale) would move the median by
    O(1)."""
    from oracle import geometry as ogeo
    from tests.util import dense_params

    mlp = ogeo.DensityMLP({k: v for k, v in ENC_CONFIGS["a"].items() if k != "scale_supersample"}, net_depth=2, net_width=64,
                          density_bias=-1.0, warp_c=2.0, bbox_scaling=2.0)
    p = {"density_grid": {name: torch.from_numpy(_level_table(shape, i + 1))
                          for i, (name, (_, _, shape)) in enumerate(zip(mlp.grid.param_names(), mlp.grid.layout))}}
    d_in = mlp.in_dim
    for i, (name, d_out) in enumerate([("density_layers_0", 64), ("density_layers_1", 64), ("output_density_layer", 1)]):
        k, b = dense_params(d_in, d_out, 100 + i)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
        d_in = d_out
    means = T("dnrm_means").clone().requires_grad_(True)
    raw, _ = mlp.predict_density(p, means)
    (grad,) = torch.autograd.grad(raw.sum(), means)
    med, q80 = fd_quantiles(grad, "dnrm_fd_raw_grad")
    assert med <= 3e-4 and q80 <= 3e-3, (med, q80)    # measured 7e-5 / ...; ~30 % of the differences straddle a face of
    # the finest level (cell 1/64 of the contracted box against 2h = 1e-3, three axes, five levels)


def test_pixels_to_rays():
    """camera_utils.pixels_to_rays (internal/camera_utils.py:896-1073) and get_pixtocam (:749-763) executed from the
    reference: perspective camera, pixel centres, OpenCV -> OpenGL flip, rotation, unit view directions, cone radii."""
    from oracle import camera_utils as ocam

    Wc, Hc = (int(v) for v in V["cam_size"])
    ys, xs = np.meshgrid(np.arange(Hc, dtype=np.float32), np.arange(Wc, dtype=np.float32), indexing="ij")
    got = ocam.pixels_to_rays(xs, ys, V["cam_pixtocam"], V["cam_camtoworld"])
    assert np.allclose(ocam.get_pixtocam(17.3, Wc, Hc).astype(np.float32), V["cam_pixtocam"], rtol=0, atol=0)
    for name, g_ in zip(("origins", "directions", "viewdirs", "radii", "imageplane"), got):
        close(torch.from_numpy(np.ascontiguousarray(g_)), "cam_" + name, 1e-6)


def test_temporal_filter_and_transient_integration():
    """The temporal filter of volumetric_transient_rendering (internal/render.py:397-415: Gaussian of tfilter_sigma bins,
    jax.scipy.signal.convolve mode='same', on the direct histogram and - filter_indirect - on the indirect one) and
    transient_integrate_reflect_rays with direct=False (internal/inverse_render/render_utils.py:1195-1302), both executed
    from the reference."""
    from oracle import render_utils as oru, transient as otr

    w, direct, ind = T("tr_weights"), T("tr_direct"), T("tr_indirect")
    ray, light = T("tr_ray_dists")[..., 0], T("tr_light_dists")[..., 0]
    B = ind.shape[2]
    base = otr.volumetric_transient_rendering(direct, ind, w, ray, light, B, exposure_time=0.01, shift=0.0137, dark_level=0.0)
    filt = otr.gaussian_tfilter(1.5)
    for fi in (0, 1):
        td = otr.temporal_filter(base["transient_direct"], filt)
        ti = otr.temporal_filter(base["transient_indirect"], filt) if fi else base["transient_indirect"]
        close(td, f"tr_f{fi}_transient_direct", 2e-6)
        close(ti, f"tr_f{fi}_transient_indirect", 2e-6)
        close(td + ti + 0.001, f"tr_f{fi}_rgb", 2e-6)
    material = {k: T("ggx_mat_" + k) for k in ("albedo", "roughness", "F_0", "metalness")}
    ld = T("ggx_smp_local_lightdirs")
    normal = torch.cat([torch.zeros_like(ld[..., :2]), torch.ones_like(ld[..., :1])], dim=-1)
    lobe = oru.get_lobe(ld, T("ggx_smp_local_viewdirs"), normal, material, T("ggx_smp_brdf_correction"), "microfacet")
    got = otr.transient_integrate_reflect_rays(lobe, T("ggx_smp_weight"), T("ggx_smp_pdf"), T("ggx_smp_local_lightdirs"),
                                               T("ggxt_radiance_in"), T("ggx_smp_indirect_occ"))
    for k in ("radiance_out", "irradiance", "indirect_occ"):
        close(got[k], "ggxt_" + k, 1e-5)


# ---- surface-light-field memory variant: oracle.surface_light_field against the reference's own class
VS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_slf.npz"))


@pytest.mark.parametrize("tag,n,kw", [("slfm", 8, {}), ("slfm1", 1, dict(near=0.07, far=0.13))])
def test_slf_memory(tag, n, kw):
    from oracle import surface_light_field as oslf
    from tests.util import SLF_GRID, slf_mem_params
    net = oslf.SurfaceLightFieldMemMLP(num_distance_samples=n, grid=dict(SLF_GRID), reflectance_grid=dict(SLF_GRID, bbox_scaling=2.0))
    assert net.grid.param_names() == list(VS[tag + "_grid_names"])
    assert net.reflectance_grid.param_names() == list(VS[tag + "_ref_grid_names"])
    p = slf_mem_params(net)
    o, d = torch.from_numpy(VS[tag + "_origins"]), torch.from_numpy(VS[tag + "_viewdirs"])
    res = net(p, o, d, **kw)
    R = o.shape[0]
    for k, tol in (("bottleneck", 2e-5), ("dist_net_outputs", 5e-5)):
        ref = torch.from_numpy(VS[tag + "_" + k]).reshape(R, -1)
        assert float((res[k] - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max())), k
    # predict_points on the REFERENCE's own network outputs: a fold of s (floor parity) cannot flip on a 1e-5 input difference
    raw = torch.from_numpy(VS[tag + "_dist_net_outputs"]).reshape(R, -1)
    pp = oslf.predict_points(raw, o, d, n, net.distance_near, net.distance_far, kw.get("near", 0.0), kw.get("far", float("inf")))
    for k, gk, tol in (("points_raw", "points", 2e-6), ("ref_weights", "incoming_weights", 2e-6), ("s_dist", "incoming_s_dist", 2e-6),
                       ("distances", "incoming_dist", 2e-6), ("env_rgba", "incoming_env_rgba", 2e-6), ("ref_mask", "ref_mask", 0.0),
                       ("s_distances", "s_distances", 2e-6), ("raw_weights", "raw_weights", 0.0)):
        ref = torch.from_numpy(VS[tag + "_" + gk]).reshape(pp[k].shape)
        assert float((pp[k] - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max())), k
    for k, tol in (("incoming_rgb", 2e-5), ("incoming_ambient_rgb", 2e-5), ("incoming_alpha", 2e-5), ("incoming_weights", 2e-5),
                   ("incoming_s_dist", 2e-5), ("incoming_dist", 2e-5), ("incoming_env_rgba", 2e-5), ("incoming_acc", 2e-5)):
        ref = torch.from_numpy(VS[tag + "_" + k]).reshape(res[k].shape)
        assert float((res[k] - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max())), k


def test_slf_variate_merge():
    """surface_light_field.integrate_slf_variate (host logic of the product; no kernel involved) against the reference's
    own material._integrate_slf_variate executed on a stand-in `self` (material.py:2433-2513): same keys, differences for
    the radiance keys, `_cache` / `_slf` copies (None where the light-field pass has no such entry)."""
    from neural_radiance_caching_b200 import surface_light_field as nslf
    cache = {k[len("variate_cache_"):]: torch.from_numpy(VS[k]) for k in VS.files if k.startswith("variate_cache_")}
    slf = {k[len("variate_slf_"):]: torch.from_numpy(VS[k]) for k in VS.files if k.startswith("variate_slf_")}
    got = nslf.integrate_slf_variate(cache, slf)
    assert sorted(got.keys()) == list(VS["variate_keys"])
    for k, v in got.items():
        name = "variate_out_" + k
        if v is None:
            assert name not in VS.files, k
        else:
            assert np.array_equal(v.numpy(), VS[name]), k
