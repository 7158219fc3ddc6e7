"""More CUDA kernels held DIRECTLY to the reference's own source (tests/golden/reference_np.npz: outputs of the
reference's functions / classes under the NumPy stand-in for jax) through the C ABI, no oracle in between:
importance_sample_rays with 2 and 3 samplers, the microfacet material head, LightMLP.get_vmfs, the time-resolved
rendering path and zero_invalid_bins, the transient indirect head, ProposalVolumeSampler.__call__ (closed-form density
fields), Model.maybe_resample and DensityMLP.predict_density / the module call.  tests/test_reference_vectors.py holds
the oracle to the same arrays on CPU."""
import os

import numpy as np
import pytest
import torch

from tests.util import ENC_CONFIGS, dense_params, level_table, rel_err

pytestmark = pytest.mark.gpu
V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_np.npz"))


def D(key, dev):
    return torch.from_numpy(np.ascontiguousarray(V[key])).to(dev).contiguous()


# ----------------------------------------------------------------------------- row 19
@pytest.mark.parametrize("ns", [2, 3])
def test_importance_sample_rays(cuda_device, ns):
    """nrc_secondary_sample against importance_sample_rays (internal/inverse_render/render_utils.py:722-924) run from the
    reference with pre-drawn randoms: Microfacet (16) + Cosine (8) [+ vMF-mixture Light (8)], shading frame, MIS power
    heuristic, energy correction."""
    from neural_radiance_caching_b200.inverse_render import render_utils as nru

    dev = cuda_device
    counts = (16, 8, 8)[:ns]
    S = sum(counts)
    P = V["is_viewdirs"].shape[0]
    u = torch.zeros((P, S, 2), device=dev)
    off = 0
    for j, c in enumerate(counts):
        u[:, off:off + c, 0] = D(f"is_uh_{j}", dev)
        u[:, off:off + c, 1] = D(f"is_uw_{j}", dev)
        off += c
    rng = dict(u=u)
    lsr = None
    samplers = [(nru.MicrofacetSampler, 16), (nru.CosineSampler, 8), (nru.LightSampler, 8)][:ns]
    if ns == 3:
        u[:, 24:, 0] = D("is_light_u", dev)     # the vMF sampler's own uniform (the kernel reads it from u[..., 0])
        rng.update(latent=D("is_light_latent", dev).to(torch.int32), normal2=D("is_light_normal2", dev))
        lsr = dict(vmf_means=D("light_means", dev)[:P], vmf_kappas=D("light_kappas", dev)[:P], vmf_logits=D("light_logits", dev)[:P])
    # importance_sample_rays takes the direction TOWARDS the viewer; get_secondary_rays negates its `viewdirs`
    _, got = nru.get_secondary_rays(rng, None, torch.zeros((P, 3), device=dev), -D("is_viewdirs", dev), D("is_normals", dev),
                                    dict(roughness=D("is_roughness", dev)), samplers=samplers, num_secondary_samples=S,
                                    light_sampler_results=lsr, far=2.0)
    for k in ("local_lightdirs", "local_viewdirs", "global_lightdirs"):
        assert rel_err(got[k], torch.from_numpy(V[f"is{ns}_{k}"])) <= 2e-5, (k, rel_err(got[k], torch.from_numpy(V[f"is{ns}_{k}"])))
    # GGX pdfs at roughness 0.01 are ill conditioned in fp32: the MIS denominator re-derives the
    # half-vector from (wo, wi), and a 1e-7 perturbation of it moves a peaked pdf by 1e-3..1e-2 in EITHER fp32 evaluation
    # (the reference's NumPy one and the kernel's).  The bulk is held tightly, the peaked tail to the conditioning bound;
    # tests/test_material_gpu.py holds the kernel to the float64 value of the same expressions.
    for k in ("pdf", "weight"):
        ref = torch.from_numpy(V[f"is{ns}_{k}"]).double()
        d = (got[k].cpu().double() - ref).abs() / (ref.abs() + 1e-3 * float(ref.abs().max()))
        q99, worst = float(torch.quantile(d.flatten(), 0.99)), float(d.max())
        assert q99 <= 1e-3 and worst <= 3e-2, (k, q99, worst)


# ----------------------------------------------------------------------------- row 18
def test_microfacet_material_head(cuda_device):
    """nrc_material_head against MaterialMLP._get_microfacet_material (internal/material.py:1276-1322) over the class's
    own property table with the fields of configs/ngp_yobo.gin:256-303."""
    from neural_radiance_caching_b200.inverse_render import render_utils as nru

    got = nru.microfacet_material(D("mat_brdf_params", cuda_device), min_roughness=0.01, default_F_0=0.04)
    for k in ("albedo", "specular_albedo", "roughness", "F_0", "metalness"):
        assert rel_err(got[k], torch.from_numpy(V["mat_" + k])) <= 1e-6, k


# ----------------------------------------------------------------------------- 8f-4
def test_light_mlp_get_vmfs(cuda_device):
    """nrc_vmf_head_fwd against LightMLP.get_vmfs (internal/light_sampler.py:135-160) with the class's own bias and
    activation tables; positions = 0 removes the recentring of predict_lighting (:203-204)."""
    from neural_radiance_caching_b200.light_sampler import _VmfHeadFn

    raw = D("vmfs_raw", cuda_device)
    P, K = raw.shape[0], raw.shape[1]
    means_random = D("vmfs_normal", cuda_device) * 20.0 / 2.0
    vm, vk, vl = _VmfHeadFn.apply(raw.reshape(P, K * 5), means_random, torch.zeros((P, 3), device=cuda_device), K, 20.0)
    assert rel_err(vm, torch.from_numpy(V["vmfs_vmf_means"])) <= 1e-6
    assert rel_err(vk.reshape(P, K, 1), torch.from_numpy(V["vmfs_vmf_kappas"])) <= 1e-6
    assert np.array_equal(vl.reshape(P, K, 1).cpu().numpy(), V["vmfs_vmf_logits"])


# ----------------------------------------------------------------------------- row 22
def _no_mask(dev, like):
    return torch.full_like(like, -1e9)


def test_transient_rendering(cuda_device):
    """nrc_transient_render_fwd against render.volumetric_transient_rendering (internal/render.py:250-449) incl.
    shift_direct's flat-index spill (:452-490) and shift_map_coordinates (:493-507): the head post-processing of the fused
    kernel is switched off (scale 1, no clip, validity thresholds out of reach)."""
    from neural_radiance_caching_b200 import render as nrender

    dev = cuda_device
    w, direct, ind = D("tr_weights", dev), D("tr_direct", dev), D("tr_indirect", dev)
    assert float(ind.min()) >= 0.0     # the fused kernel clips at 0; the vectors are non-negative radiance
    ray, light = D("tr_ray_dists", dev)[..., 0].contiguous(), D("tr_light_dists", dev)[..., 0].contiguous()
    B = ind.shape[2]
    for tag, shift in (("tr_s0_", 0.0), ("tr_s1_", 0.0137)):
        got = nrender.volumetric_transient_rendering(direct, None, ind, torch.ones_like(direct), w, ray, light, _no_mask(dev, ray),
                                                     n_bins=B, exposure_time=0.01, shift=shift, indirect_scale=1.0,
                                                     bin_zero_threshold_light=1e9, rgb_max=3e38, dark_level=0.001)
        for k in ("transient_direct", "transient_indirect", "rgb"):
            assert rel_err(got[k], torch.from_numpy(V[tag + k])) <= 2e-5, (tag, k, rel_err(got[k], torch.from_numpy(V[tag + k])))


def _per_sample(dev, fn, R, n):
    """The fused kernel reduces over the samples of a ray; one-hot weights and zero ray distance (no time shift) expose
    the per-sample, per-bin values it reduces."""
    outs = []
    for s in range(n):
        w = torch.zeros((R, n), device=dev)
        w[:, s] = 1.0
        outs.append(fn(w))
    return torch.stack(outs, dim=1)     # [R, n, B, C]


@pytest.mark.parametrize("light_zero", [False, True])
def test_zero_invalid_bins(cuda_device, light_zero):
    """The validity masks inside nrc_transient_render_fwd against render_utils.zero_invalid_bins
    (internal/inverse_render/render_utils.py:1699-1767): BIT-EXACT masked tensors."""
    from neural_radiance_caching_b200 import render as nrender

    dev = cuda_device
    means = D("tr_means", dev)
    light_d = torch.linalg.norm(D("tr_lights", dev)[:, None, :] - means, dim=-1)
    cam_d = (torch.linalg.norm(D("tr_origins", dev)[:, None, :] - means, dim=-1)
             + torch.linalg.norm(D("tr_origins", dev) - D("tr_cam_origins", dev), dim=-1)[:, None])
    # the distances as the reference computed them (NumPy): the masks compare them against bin edges
    light_np = np.linalg.norm(V["tr_lights"][:, None, :] - V["tr_means"], axis=-1).astype(np.float32)
    cam_np = (np.linalg.norm(V["tr_origins"][:, None, :] - V["tr_means"], axis=-1)
              + np.linalg.norm(V["tr_origins"] - V["tr_cam_origins"], axis=-1)[:, None]).astype(np.float32)
    light_d, cam_d = torch.from_numpy(light_np).to(dev), torch.from_numpy(cam_np).to(dev)
    R, n = light_d.shape
    zero = torch.zeros((R, n), device=dev)
    for key, want in (("tr_indirect", "tr_zero_diffuse"), ("tr_spec", "tr_zero_specular")):
        x = D(key, dev)
        B = x.shape[2]
        fn = lambda w: nrender.volumetric_transient_rendering(
            torch.zeros((R, n, 3), device=dev), None, x, torch.ones((R, n, 3), device=dev), w, zero, light_d, cam_d, n_bins=B,
            exposure_time=0.01, shift=0.0, indirect_scale=1.0, bin_zero_threshold_light=2.0, light_zero=light_zero,
            light_near=0.12, rgb_max=3e38)["transient_indirect"]
        got = _per_sample(dev, fn, R, n)
        assert np.array_equal(got.cpu().numpy(), V[f"{want}_{int(light_zero)}"]), want


def test_transient_indirect_head(cuda_device):
    """TransientIndirectHead + the fused post-processing against TransientNeRFMLP._compute_indirect_lighting / get_indirect
    (internal/nerf.py:1660-1777) executed from the reference's class: pos_enc(lights) -> irradiance stack -> n_bins * 3
    bins, softplus(. - 2) * scale, tint * F * ref_rgb * scale, zero_invalid_bins, clip."""
    from neural_radiance_caching_b200 import nerf as nnerf, render as nrender

    dev = cuda_device
    feat, means, ref = D("th_feature", dev), D("th_means", dev), D("th_ref_rgb", dev)
    R, n, B = ref.shape[0], ref.shape[1], ref.shape[2] // 3
    head = nnerf.TransientIndirectHead(n_bins=B)
    p, d_in = {}, 96 + 15
    for i, (name, d_out) in enumerate((("irradiance_layers_0", 64), ("irradiance_layers_1", 64), ("transient_indirect_layer", B * 3))):
        k, b = dense_params(d_in, d_out, 300 + i)
        p[name] = {"kernel": torch.from_numpy(k).to(dev), "bias": torch.from_numpy(b).to(dev)}
        d_in = d_out
    lights = (D("th_lights", dev)[:, None, :] * torch.ones_like(means)).reshape(R * n, 3)
    diffuse_raw = head(p, feat.reshape(R * n, 96), lights).reshape(R, n, B, 3)
    # tint * F of the specular term, from the reference's own outputs: indirect_specular = clip(tint F ref_rgb scale) summed
    # over bins cannot be inverted, so the factor is recomputed exactly as the CPU test does (Dense stacks in float32)
    from oracle import geometry as ogeo, nerf as onerf2
    pc = {}
    for group, din, salt in (((("integrated_brdf_layers_0", 64), ("integrated_brdf_layers_1", 64),
                               ("output_integrated_brdf_layer", 1)), 129, 200), ((("tint_layer", 3),), 96, 320)):
        for i, (name, d_out) in enumerate(group):
            k, b = dense_params(din, d_out, salt + i)
            pc[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
            din = d_out
    T = lambda key: torch.from_numpy(V[key])
    F = onerf2.NeRFMLP().get_integrated_brdf(pc, T("th_normals"), T("th_viewdirs"), T("th_bottleneck"))
    scale = (torch.sigmoid(ogeo.dense(pc["tint_layer"], T("th_feature"))) * F).to(dev).contiguous()
    light_np = np.linalg.norm(V["th_lights"][:, None, :] - V["th_means"], axis=-1).astype(np.float32)
    cam_np = (np.linalg.norm(V["th_origins"][:, None, :] - V["th_means"], axis=-1)
              + np.linalg.norm(V["th_origins"] - V["th_cam_origins"], axis=-1)[:, None]).astype(np.float32)
    light_d, cam_d = torch.from_numpy(light_np).to(dev), torch.from_numpy(cam_np).to(dev)
    zero = torch.zeros((R, n), device=dev)
    common = dict(n_bins=B, exposure_time=0.01, shift=0.0, diffuse_bias=-2.0, indirect_scale=0.7, bin_zero_threshold_light=2.0,
                  rgb_max=1.5)
    blank = torch.zeros((R, n, 3), device=dev)
    diff = _per_sample(dev, lambda w: nrender.volumetric_transient_rendering(
        blank, diffuse_raw, None, None, w, zero, light_d, cam_d, **common)["transient_indirect"], R, n)
    spec = _per_sample(dev, lambda w: nrender.volumetric_transient_rendering(
        blank, None, ref.reshape(R, n, B, 3), scale, w, zero, light_d, cam_d, **common)["transient_indirect"], R, n)
    assert rel_err(diff, T("th_transient_indirect_diffuse")) <= 1e-5, rel_err(diff, T("th_transient_indirect_diffuse"))
    assert rel_err(spec, T("th_transient_indirect_specular")) <= 1e-5, rel_err(spec, T("th_transient_indirect_specular"))
    assert rel_err(diff + spec, T("th_transient_indirect")) <= 1e-5
    assert rel_err(diff.sum(-2), T("th_indirect_diffuse")) <= 1e-5
    assert rel_err(spec.sum(-2), T("th_indirect_specular")) <= 1e-5


# ----------------------------------------------------------------------------- rows 10-14
class _Field:
    """Closed-form rational density field of the generator (IEEE +, *, / only: the same bits on NumPy and on the
    device) standing in for a DensityMLP."""
    normals_for_filter_only = True
    disable_density_normals = True
    enable_pred_normals = False

    def __init__(self, scale, k, c):
        self.scale, self.k, self.c = scale, k, c

    def query(self, p, means, want_feat=True, want_normals=False):
        dx, dy, dz = means[..., 0] - self.c[0], means[..., 1] - self.c[1], means[..., 2] - self.c[2]
        dens = self.scale / (1.0 + self.k * (dx * dx + dy * dy + dz * dz))
        return dict(density=dens.contiguous(), feature=None, raw_density=None, raw_grad_density=None, grad_pred=None)


@pytest.mark.parametrize("tag,use_raydist", [("id", False), ("pl", True)])
def test_proposal_sampler_loop(cuda_device, tag, use_raydist):
    """The sampler kernels (nrc_ray_sample_intervals, nrc_ray_cast, nrc_ray_alpha_weights_fwd) inside
    sampling.ProposalVolumeSampler.__call__ against the reference's own class (internal/sampling.py:142-649) with the
    configured strategy (64, 64, 32), annealing, padding and both ray warps; densities from the same closed-form fields on
    both sides."""
    from neural_radiance_caching_b200.sampling import ProposalVolumeSampler

    dev = cuda_device
    s = ProposalVolumeSampler()
    s.mlps = [_Field(4.0, 3.0, (0.1, -0.2, 0.3)), _Field(9.0, 6.0, (0.0, -0.1, 0.2)), _Field(40.0, 14.0, (0.05, -0.15, 0.25))]
    rays = {k: D("pvs_" + k, dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")}
    hist = s({f"MLP_{i}": None for i in range(3)}, rays, [D("pvs_u01", dev)] * 3, use_raydist_fn=use_raydist)
    # the CDF inversion amplifies last-ulp differences (softmax / cumsum order, libm pow in the ladder) level by level:
    # tolerances relative to the ray length / largest weight, as in the CPU test of the oracle
    for lvl, h in enumerate(hist):
        for k in ("sdist", "tdist", "means", "weights"):
            tol = 1e-4 if k == "weights" else (2e-6 if lvl == 0 else 2e-5)
            e = rel_err(h[k], torch.from_numpy(V[f"pvs_{tag}_{lvl}_{k}"]))
            assert e <= tol, (tag, lvl, k, e)


# ----------------------------------------------------------------------------- row 15
def test_maybe_resample(cuda_device):
    """nrc_ray_resample / nrc_ray_resample_gather against Model.maybe_resample (internal/models.py:193-292): indices and
    gathers BIT-EXACT, importance-corrected weights 1e-6."""
    from neural_radiance_caching_b200 import models as nmodels

    dev = cuda_device
    w, gum = D("rs_weights", dev), D("rs_gumbel", dev)
    for k, bias in ((1, 0.0), (4, 1e-3)):
        inds, nw = nmodels._ResampleWeightsFn.apply(w, gum[..., :k].contiguous(), bias, 1.0)
        assert np.array_equal(inds.cpu().numpy().astype(np.int32), V[f"rs_inds_{k}"].astype(np.int32))
        assert rel_err(nw, torch.from_numpy(V[f"rs_new_weights_{k}"])) <= 1e-6
        for f in ("points", "feature"):
            got = nmodels._GatherFn.apply(D("rs_" + f, dev), inds)
            assert np.array_equal(got.cpu().numpy(), V[f"rs_new_{f}_{k}"]), f


# ----------------------------------------------------------------------------- rows 8-9
def _dmlp(dev, pred_normals):
    from neural_radiance_caching_b200 import geometry as ngeo

    mlp = ngeo.DensityMLP({k: v for k, v in ENC_CONFIGS["a"].items() if k != "scale_supersample"}, net_depth=2, net_width=64,
                          density_bias=-1.0, warp_c=2.0, bbox_scaling=2.0, enable_pred_normals=pred_normals,
                          disable_density_normals=True)
    p = {"density_grid": {name: torch.from_numpy(level_table(shape, i + 1))
                          for i, (name, _, _, shape) in enumerate(mlp.grid.level_layout)}}
    layers = [("density_layers_0", mlp.in_dim, 64, 100), ("density_layers_1", 64, 64, 101), ("output_density_layer", 64, 1, 102)]
    if pred_normals:
        layers.append(("pred_normals_layer", 64, 3, 110))
    for name, d_in, d_out, salt in layers:
        k, b = dense_params(d_in, d_out, salt)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
    return mlp, mlp.from_oracle(p, dev)


def test_density_mlp(cuda_device):
    """nrc_density_query_fwd (exact mode) against DensityMLP.predict_density / run_network / convert_raw_density
    (internal/geometry.py:155-341) executed from the reference's class: contraction -> HashEncoding -> 2 x 64 ReLU -> 1,
    safe_exp(raw - 1).  The encoding is bit-exact (test_hash_encoding_call_bit_exact); the Dense layers differ by
    summation order."""
    mlp, p = _dmlp(cuda_device, False)
    q = mlp.query(p, D("dmlp_means", cuda_device), want_feat=True)
    assert rel_err(q["feature"], torch.from_numpy(V["dmlp_feature"])) <= 1e-5, rel_err(q["feature"], torch.from_numpy(V["dmlp_feature"]))
    assert rel_err(q["raw_density"], torch.from_numpy(V["dmlp_raw_density"])) <= 1e-5
    assert rel_err(q["density"], torch.from_numpy(V["dmlp_density"])) <= 1e-5


def test_density_mlp_module_call(cuda_device):
    """The fused query + normals kernels against DensityMLP.__call__ (internal/geometry.py:343-584) executed from the
    reference's class on the branch without jax.value_and_grad: features, density, predicted-normals head and its
    l2_normalize."""
    from neural_radiance_caching_b200.sampling import _NormalsFn

    mlp, p = _dmlp(cuda_device, True)
    means = D("dmlp_means", cuda_device)[:696].reshape(58, 12, 3)
    q = mlp.query(p, means, want_feat=True)
    for k in ("feature", "density", "grad_pred"):
        assert rel_err(q[k], torch.from_numpy(V["dmlp_call_" + k])) <= 1e-5, (k, rel_err(q[k], torch.from_numpy(V["dmlp_call_" + k])))
    npred = _NormalsFn.apply(q["grad_pred"])
    assert rel_err(npred, torch.from_numpy(V["dmlp_call_normals_pred"])) <= 2e-5


# ----------------------------------------------------------------------------- row 16
def _slf_params(deg_view, use_bottleneck, salt, dev):
    n_in = (128 if use_bottleneck else 0) + {4: 38, 5: 72}[deg_view]
    p, d_in = {}, n_in
    for i, name in enumerate(("layer_0", "layer_1", "layer_2", "layer_bottleneck")):
        k, b = dense_params(d_in, 128, salt + i)
        p[name] = {"kernel": torch.from_numpy(k).to(dev), "bias": torch.from_numpy(b).to(dev)}
        d_in = 128 + (n_in if (i % 2 == 0 and i > 0) else 0)
    k, b = dense_params(d_in, 3, salt + 10)
    p["output_ambient_rgb_layer"] = {"kernel": torch.from_numpy(k).to(dev), "bias": torch.from_numpy(b).to(dev)}
    return p


@pytest.mark.parametrize("bf16", [False, True])
def test_surface_light_field_call(cuda_device, bf16):
    """The SurfaceLightField / EnvMap stacks (fp32 GEMM path; tcgen05 chain kernel) against SurfaceLightFieldMLP.__call__
    (internal/surface_light_field.py:782-1069) executed from the reference's own class in both configured shapes."""
    from neural_radiance_caching_b200 import nerf as nnerf

    dev = cuda_device
    for tag, deg, use_b, salt in (("slf5", 5, True, 400), ("slf4", 4, False, 420)):
        net = nnerf.SurfaceLightFieldMLP(deg, use_b, bf16=bf16)
        with torch.no_grad():
            got = net(_slf_params(deg, use_b, salt, dev), D("slf_refdirs", dev), D("slf_roughness", dev),
                      D("slf_bottleneck", dev) if use_b else None)
        # the fp32 kernels evaluate the l = 16 Legendre sums in float64, the reference in float32 (~1e-3 noise)
        tol = 2e-2 if bf16 else 2e-4
        e = rel_err(got["incoming_ambient_rgb"], torch.from_numpy(V[tag + "_incoming_ambient_rgb"]))
        assert e <= tol, (tag, bf16, e)
        assert np.array_equal(got["incoming_acc"].cpu().numpy(), V[tag + "_incoming_acc"])


@pytest.mark.parametrize("bf16", [False, True])
def test_predict_appearance_passive(cuda_device, bf16):
    """The WHOLE cache shader (appearance grid -> bottleneck / heads -> integrated BRDF, EnvMap, SurfaceLightField ->
    composition) against the reference's own classes: BaseShader.predict_appearance_feature, NeRFMLP.get_bottleneck_feature,
    the roughness head and NeRFMLP._predict_appearance_passive (internal/nerf.py:385-408,633-634,940-1090) with both
    SurfaceLightFieldMLP instances attached.  fp32 kernels at the conditioning of the degree-5 IDE, the bf16 tensor-core
    path (trunk -> mid -> one launch of three stacks -> out) at the north star's 2e-2."""
    from neural_radiance_caching_b200 import grid_utils as ng, nerf as nnerf

    dev = cuda_device
    sh = nnerf.NeRFMLP(warp_c=2.0, bf16=bf16)
    sh.grid = ng.HashEncoding(hash_map_size=2 ** 15, num_features=4, scale_supersample=1.0, max_grid_size=2048, bbox_scaling=2.0)
    p = {}
    for name, d_in, d_out, salt in (("bottleneck_layer", 96, 128, 440), ("roughness_layer", 96, 1, 441),
                                    ("ambient_irradiance_layer", 96, 3, 442), ("irradiance_layer", 96, 3, 443),
                                    ("tint_layer", 96, 3, 444), ("integrated_brdf_layers_0", 129, 64, 445),
                                    ("integrated_brdf_layers_1", 64, 64, 455), ("output_integrated_brdf_layer", 64, 1, 465)):
        k, b = dense_params(d_in, d_out, salt)
        p[name] = {"kernel": torch.from_numpy(k).to(dev), "bias": torch.from_numpy(b).to(dev)}
    p["SurfaceLightField"], p["EnvMap"] = _slf_params(5, True, 400, dev), _slf_params(4, False, 420, dev)
    arena = torch.cat([torch.from_numpy(level_table(shape, i + 1)).reshape(-1)
                       for i, (name, _, _, shape) in enumerate(sh.grid.level_layout)]).to(dev)
    p["appearance_grid"] = dict(sh.grid.views(arena), _arena=arena)
    with torch.no_grad():
        got = sh(p, D("shp_viewdirs", dev), D("shp_means", dev), D("shp_density_feature", dev), D("shp_normals", dev),
                 return_feature=True)
    assert np.array_equal(got["feature"].cpu().numpy(), V["shp_feature"])       # appearance feature: bit-exact
    tol = 2e-2 if bf16 else 2e-4
    for k in ("rgb", "diffuse_rgb", "specular_rgb", "ambient_rgb", "indirect_rgb", "albedo_rgb"):
        e = rel_err(got[k], torch.from_numpy(V["shp_" + k]))
        assert e <= tol, (k, bf16, e)
    if got["bottleneck"] is not None:
        assert rel_err(got["bottleneck"], torch.from_numpy(V["shp_bottleneck"])) <= (2e-2 if bf16 else 1e-5)
    assert rel_err(got["roughness"], torch.from_numpy(V["shp_roughness"])) <= (2e-2 if bf16 else 1e-5)


# ----------------------------------------------------------------------------- row 9
def test_analytic_normals_against_reference_differences(cuda_device):
    """The in-kernel back-propagation of nrc_density_query_fwd (d raw / d mean, internal/geometry.py:442-460) against
    central differences of the REFERENCE'S OWN predict_density (h = 2^-11, tests/golden/make_reference_vectors.py).
    Differences that straddle a grid-cell face or a ReLU kink average two slopes: quantiles, as in the CPU test of the
    oracle's autograd."""
    mlp, p = _dmlp(cuda_device, False)
    q = mlp.query(p, D("dnrm_means", cuda_device), want_feat=False, want_normals=True)
    ref = torch.from_numpy(V["dnrm_fd_raw_grad"]).double()
    d = (q["raw_grad_density"].double().cpu() - ref).abs().flatten() / float(ref.abs().max())
    med, q65 = float(torch.quantile(d, 0.5)), float(torch.quantile(d, 0.65))
    assert med <= 3e-4 and q65 <= 3e-3, (med, q65)
