"""Row 22 (BASELINE config 4): fused transient rendering kernel vs the oracle restatement of
volumetric_transient_rendering + zero_invalid_bins + the transient heads' post-processing."""
import numpy as np
import pytest
import torch

from oracle import transient as otr
from neural_radiance_caching_b200 import render as nrender
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu


def _inputs(g, R, n, B, C=3, exposure=0.01):
    w = g.uniform(size=(R, n)) ** 3
    w = w / w.sum(-1, keepdims=True) * g.uniform(0.2, 1.0, size=(R, 1))
    span = B * exposure
    ray_d = np.sort(g.uniform(0.05 * span, 0.6 * span, size=(R, n)), axis=-1)
    light_d = ray_d + g.uniform(-0.01, 0.01, size=(R, n)) * span
    cam_d = ray_d + 0.02 * span
    ray_d[0, 0] = 3.0 * exposure          # integer bin shift: floor == ceil in shift_direct
    light_d[0, 0] = 2.0 * exposure
    light_d[1, :] = 0.999 * span          # direct splat beyond the last bin: spills into the next ray's histogram
    ray_d[2, :] = 0.45 * span
    return dict(direct=f32(g.uniform(size=(R, n, C))), diffuse_raw=f32(g.normal(size=(R, n, B, C))),
                specular=f32(g.gamma(1.0, 0.5, size=(R, n, B, C))), spec_scale=f32(g.uniform(size=(R, n, C))),
                weights=f32(w), ray_dists=f32(ray_d), light_dists=f32(light_d), cam_dists=f32(cam_d))


@pytest.mark.parametrize("R,n,B", [(20, 8, 50), (7, 32, 700)])
@pytest.mark.parametrize("light_zero", [False, True])
def test_transient_render(cuda_device, R, n, B, light_zero):
    g = gen(600 + R + n)
    x = _inputs(g, R, n, B)
    kw = dict(exposure_time=0.01, shift=0.003, diffuse_bias=-1.0, indirect_scale=0.7, bin_zero_threshold_light=1.5,
              light_zero=light_zero, light_near=0.08 * B * 0.01, rgb_max=2.0, dark_level=0.01)
    want = otr.transient_render(x["direct"], x["diffuse_raw"], x["specular"], x["spec_scale"], x["weights"], x["ray_dists"],
                                x["light_dists"], x["cam_dists"], B, **kw)
    d = {k: v.to(cuda_device) for k, v in x.items()}
    got = nrender.volumetric_transient_rendering(d["direct"], d["diffuse_raw"], d["specular"], d["spec_scale"], d["weights"],
                                                 d["ray_dists"], d["light_dists"], d["cam_dists"], n_bins=B, **kw)
    for k in ("transient_direct", "transient_indirect", "rgb"):
        assert got[k].shape == want[k].shape
        assert rel_err(got[k], want[k]) <= 2e-5, (k, rel_err(got[k], want[k]))
    # conservation: the direct histogram holds all of sum_s w * direct that landed inside the array
    assert float(got["transient_direct"].sum()) <= float((x["weights"][..., None] * x["direct"]).sum()) * (1 + 1e-5)


def test_transient_shift_is_linear_interpolation(cuda_device):
    """Property: with one sample of weight 1 and an impulse in bin k, a shift of s + f bins (0 <= f < 1) puts
    (1 - f) of the impulse in bin k + s and f in bin k + s + 1 (order-1 map_coordinates)."""
    R, n, B, C = 1, 1, 40, 3
    dev = cuda_device
    spec = torch.zeros((R, n, B, C), device=dev)
    spec[0, 0, 10] = 1.0
    one = torch.ones((R, n), device=dev)
    ray = torch.full((R, n), (5 + 0.25) * 0.01, device=dev)
    got = nrender.volumetric_transient_rendering(
        torch.zeros((R, n, C), device=dev), None, spec, torch.ones((R, n, C), device=dev), one, ray, torch.zeros_like(ray),
        torch.zeros_like(ray), n_bins=B, exposure_time=0.01, bin_zero_threshold_light=0.0)["transient_indirect"][0, :, 0].cpu()
    want = torch.zeros(B)
    want[15], want[16] = 0.75, 0.25
    assert torch.allclose(got, want, atol=1e-5)


def test_transient_indirect_head(cuda_device):
    """get_indirect (nerf.py:1757-1777): 111 -> 64 -> 64 -> n_bins*3, fp32 parity vs the same Dense stack in torch."""
    from oracle import coord as ocoord, geometry as ogeo
    from neural_radiance_caching_b200 import nerf as nnerf
    g = gen(640)
    P, B = 300, 700
    head = nnerf.TransientIndirectHead(n_bins=B)
    gen_t = torch.Generator(device=cuda_device)
    gen_t.manual_seed(5)
    p = head.init(cuda_device, gen_t)
    feat, lights = f32(g.normal(size=(P, 96))), f32(g.normal(size=(P, 3)) * 2)
    pc = {k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in p.items()}
    x = torch.cat([feat, ocoord.pos_enc(lights, 0, 2, True)], dim=-1)
    x = torch.relu(ogeo.dense(pc["irradiance_layers_0"], x))
    x = torch.relu(ogeo.dense(pc["irradiance_layers_1"], x))
    want = ogeo.dense(pc["transient_indirect_layer"], x).reshape(P, B, 3)
    got = head(p, feat.to(cuda_device), lights.to(cuda_device))
    assert got.shape == (P, B, 3)
    assert rel_err(got, want) <= 1e-5


# ----------------------------------------------------------------------------- fused heads (config 4)
V = np.load(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden",
                                        "reference_np.npz"))


@pytest.mark.parametrize("R,n,B", [(9, 8, 50), (5, 32, 700)])
@pytest.mark.parametrize("heads", ["both", "diffuse", "specular"])
def test_fused_heads_match_unfused(cuda_device, R, n, B, heads):
    """nrc_transient_head_render_fwd (last layer of both transient heads on tensor cores + activation + masks + shift +
    weighted reduction, per-sample histograms never written) against the unfused path that the other tests hold to the
    oracle and to the reference's vectors: fp32 last layer -> [R,n,B,3] -> nrc_transient_render_fwd.  bf16 operands of
    the fused GEMM: north-star tolerance of the bf16-MLP variant (2e-2), measured ~3e-3."""
    dev = cuda_device
    g = gen(700 + R + n)
    x = {k: v.to(dev) for k, v in _inputs(g, R, n, B).items()}
    hd = f32(np.maximum(g.normal(size=(R, n, 64)), 0.0)).to(dev)
    hs = f32(np.maximum(g.normal(size=(R, n, 128)), 0.0)).to(dev)
    ld = {"kernel": f32(g.normal(size=(64, B * 3)) * 0.2).to(dev), "bias": f32(g.normal(size=(B * 3,)) * 0.3).to(dev)}
    ls = {"kernel": f32(g.normal(size=(128, B * 3 + 1)) * 0.15).to(dev), "bias": f32(g.normal(size=(B * 3 + 1,)) * 0.3).to(dev)}
    kw = dict(exposure_time=0.01, shift=0.003, indirect_scale=0.7, bin_zero_threshold_light=1.5, light_zero=True,
              light_near=0.08 * B * 0.01, rgb_max=2.0, dark_level=0.01)
    use_d, use_s = heads in ("both", "diffuse"), heads in ("both", "specular")
    # unfused reference path in fp32
    raw_d = (hd.reshape(-1, 64) @ ld["kernel"] + ld["bias"]).reshape(R, n, B, 3) if use_d else None
    ref = None
    if use_s:
        raw_s = (hs.reshape(-1, 128) @ ls["kernel"] + ls["bias"])[:, :B * 3].reshape(R, n, B, 3)
        ref = torch.clamp(torch.nn.functional.softplus(raw_s - 2.0), 0.0, 5.0)        # SLF incoming_rgb: softplus(raw + rgb_bias), clip
    want = nrender.volumetric_transient_rendering(x["direct"], raw_d, ref, x["spec_scale"] if use_s else None, x["weights"],
                                                  x["ray_dists"], x["light_dists"], x["cam_dists"], n_bins=B, diffuse_bias=-2.0, **kw)
    got = nrender.volumetric_transient_rendering_fused(
        x["direct"], hd if use_d else None, ld if use_d else None, hs if use_s else None, ls if use_s else None,
        x["spec_scale"] if use_s else None, x["weights"], x["ray_dists"], x["light_dists"], x["cam_dists"], n_bins=B,
        diffuse_bias=-2.0, spec_bias=-2.0, spec_max=5.0, **kw)
    assert torch.equal(got["transient_direct"], want["transient_direct"]) or rel_err(got["transient_direct"], want["transient_direct"]) <= 1e-6
    for k in ("transient_indirect", "rgb"):
        e = rel_err(got[k], want[k])
        assert e <= 2e-2, (heads, k, e)


def test_temporal_filter_and_transient_integration_against_reference(cuda_device):
    """nrc_transient_filter against the temporal filter of the reference's volumetric_transient_rendering
    (internal/render.py:397-415, Gaussian of 1.5 bins, direct only / direct + indirect), and
    nrc_ggx_integrate_transient_fwd against transient_integrate_reflect_rays with direct=False
    (internal/inverse_render/render_utils.py:1195-1302), both executed from the reference."""
    from neural_radiance_caching_b200.inverse_render import render_utils as nru

    dev = cuda_device
    D = lambda k: torch.from_numpy(V[k]).to(dev).contiguous()
    w, direct, ind = D("tr_weights"), D("tr_direct"), D("tr_indirect")
    ray, light = D("tr_ray_dists")[..., 0].contiguous(), D("tr_light_dists")[..., 0].contiguous()
    B = ind.shape[2]
    base = nrender.volumetric_transient_rendering(direct, None, ind, torch.ones_like(direct), w, ray, light,
                                                  torch.full_like(ray, -1e9), n_bins=B, exposure_time=0.01, shift=0.0137,
                                                  indirect_scale=1.0, bin_zero_threshold_light=1e9, rgb_max=3e38)
    filt = nrender.gaussian_tfilter(1.5, dev)
    for fi in (0, 1):
        td = nrender.temporal_filter(base["transient_direct"], filt)
        ti = nrender.temporal_filter(base["transient_indirect"], filt) if fi else base["transient_indirect"]
        assert rel_err(td, torch.from_numpy(V[f"tr_f{fi}_transient_direct"])) <= 2e-5
        assert rel_err(ti, torch.from_numpy(V[f"tr_f{fi}_transient_indirect"])) <= 2e-5
        assert rel_err(td + ti + 0.001, torch.from_numpy(V[f"tr_f{fi}_rgb"])) <= 2e-5
    material = {k: D("ggx_mat_" + k) for k in ("albedo", "roughness", "F_0", "metalness")}
    samples = {k: D("ggx_smp_" + k) for k in ("local_lightdirs", "local_viewdirs", "pdf", "weight", "indirect_occ")}
    samples["radiance_in"] = D("ggxt_radiance_in")
    res = nru.transient_integrate_reflect_rays("microfacet", False, material, samples, direct=False)
    for k in ("radiance_out", "irradiance", "indirect_occ"):
        e = rel_err(res[k], torch.from_numpy(V["ggxt_" + k]))
        assert e <= 2e-5, (k, e)


@pytest.mark.parametrize("R,n,B,heads", [(12, 8, 50, "both"), (5, 32, 700, "both"), (9, 16, 120, "diffuse"), (9, 16, 120, "specular")])
@pytest.mark.parametrize("light_zero", [False, True])
def test_transient_render_backward(cuda_device, R, n, B, heads, light_zero):
    """nrc_transient_render_bwd (training the time-resolved cache): gradients of a random linear functional of the three
    outputs, temporal filter included, with respect to direct_rgbs / diffuse_raw / specular / spec_scale / weights against
    autograd through the oracle's restatement of volumetric_transient_rendering (render.py:250-449)."""
    g = gen(700 + R + n)
    x = _inputs(g, R, n, B)
    if heads == "diffuse":
        x["specular"] = None
    if heads == "specular":
        x["diffuse_raw"] = None
    kw = dict(exposure_time=0.01, shift=0.003, diffuse_bias=-1.0, indirect_scale=0.7, bin_zero_threshold_light=1.5,
              light_zero=light_zero, light_near=0.08 * B * 0.01, rgb_max=2.0, dark_level=0.01)
    ups = {k: f32(g.normal(size=(R, B, 3))) for k in ("transient_direct", "transient_indirect", "rgb")}
    filt = otr.gaussian_tfilter(1.5)
    names = ("direct", "diffuse_raw", "specular", "spec_scale", "weights")

    def run(lib, dev):
        t = {k: (v.clone().to(dev).requires_grad_(True) if (v is not None and k in names) else (v.to(dev) if v is not None else None))
             for k, v in x.items()}
        zeros = torch.zeros((R, n, B, 3), device=dev)
        if lib is otr:
            res = otr.transient_render(t["direct"], t["diffuse_raw"] if t["diffuse_raw"] is not None else zeros - 1e4,
                                       t["specular"] if t["specular"] is not None else zeros, t["spec_scale"], t["weights"],
                                       t["ray_dists"], t["light_dists"], t["cam_dists"], B, **kw)
            fd = otr.temporal_filter(res["transient_direct"], filt)
        else:
            res = nrender.volumetric_transient_rendering(t["direct"], t["diffuse_raw"], t["specular"], t["spec_scale"], t["weights"],
                                                         t["ray_dists"], t["light_dists"], t["cam_dists"], n_bins=B, **kw)
            fd = nrender.temporal_filter(res["transient_direct"], filt.to(dev))
        loss = sum((res[k] * ups[k].to(dev)).sum() for k in ups) + (fd * ups["rgb"].to(dev)).sum() * 0.5
        loss.backward()
        return {k: t[k].grad for k in names if t[k] is not None}, res

    want, _ = run(otr, "cpu")
    got, res = run(nrender, cuda_device)
    for k in want:
        if k == "spec_scale" and x["specular"] is None:
            assert got[k] is None or float(got[k].abs().max()) == 0.0
            continue
        assert got[k] is not None and got[k].shape == want[k].shape, k
        assert rel_err(got[k], want[k]) <= 5e-5, (k, rel_err(got[k], want[k]))
    # bins that zero_invalid_bins removed get exactly zero gradient
    if x["diffuse_raw"] is not None:
        assert bool((got["diffuse_raw"][want["diffuse_raw"].to(cuda_device) == 0] == 0).all()) or rel_err(got["diffuse_raw"], want["diffuse_raw"]) <= 5e-5


def test_config4_training_path(cuda_device):
    """workload.TransientRenderStep (BASELINE config 4): the differentiable unfused path equals the fused render path, and
    loss_and_grads back-propagates through integrator, heads, shader and sampler - checked by central differences of the
    objective along random directions of a few parameter tensors (fp32 variant)."""
    from neural_radiance_caching_b200 import workload
    R, B = 24, 60
    stage = workload.TransientRenderStep(cuda_device, n_bins=B, bf16=False, table_init_range=0.05)
    stage.cfg["exposure_time"] = 0.01 * 700 / B          # the same metric span on 60 bins
    g = np.random.Generator(np.random.PCG64(8100))
    rn = stage.make_rays(g, R)
    rays = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(cuda_device) for k, v in rn.items()}
    u01 = [f32(g.uniform(size=(R, 1))).to(cuda_device) for _ in range(3)]
    fused = stage.render(rays, u01)
    plain = stage.render_unfused(rays, u01)
    for k in ("transient_direct", "transient_indirect", "rgb"):
        assert float(fused[k].abs().max()) > 0
        assert rel_err(plain[k], fused[k]) <= 2e-2, (k, rel_err(plain[k], fused[k]))     # the fused heads run on bf16 operands
    target = (plain["rgb"] * 0.5 + 0.01 * f32(g.uniform(size=(R, B, 3))).to(cuda_device)).detach()
    loss, grads = stage.loss_and_grads(rays, u01, target)
    assert float(loss) > 0
    leaves = stage.trainable()
    picks = ["Transient/transient_indirect_layer/bias", "Transient/albedo_layer/kernel",
             "Transient/TransientSurfaceLightField/layer_0/kernel", "Shader/bottleneck_layer/kernel", "Shader/appearance_grid",
             "Sampler/MLP_2/output_density_layer/kernel"]
    for name in picks:
        assert name in grads and float(grads[name].abs().max()) > 0, name
        t = leaves[name]
        d = torch.from_numpy(g.normal(size=tuple(t.shape)).astype(np.float32)).to(cuda_device)
        d = d / d.norm() * max(float(t.norm()), 1.0) * 0.02
        want = float((grads[name].double() * d.double()).sum())

        def at(sign):
            with torch.no_grad():
                t.add_(d, alpha=sign)
            try:
                return float(torch.mean((stage.render_unfused(rays, u01)["rgb"] - target).double() ** 2))
            finally:
                with torch.no_grad():
                    t.sub_(d, alpha=sign)

        fd = (at(1.0) - at(-1.0)) / 2.0
        assert abs(fd - want) <= 5e-2 * abs(want) + 1e-9, (name, fd, want)


def test_config4_render_chain_programs(cuda_device):
    """bf16 variant of workload.TransientRenderStep: the render path runs the shader-side stacks as tcgen05 chain programs
    (trunk heads, integrated BRDF, transient SurfaceLightField up to its last activation, irradiance stack); the result must
    agree with the per-layer bf16 GEMM path (render_unfused) within the bf16 tolerance."""
    from neural_radiance_caching_b200 import workload
    R, B = 64, 60
    g = np.random.Generator(np.random.PCG64(8200))
    outs = {}
    for bf16 in (True, False):
        stage = workload.TransientRenderStep(cuda_device, n_bins=B, bf16=bf16, table_init_range=0.05)
        stage.cfg["exposure_time"] = 0.01 * 700 / B
        if not outs:
            rn = stage.make_rays(g, R)
            rays = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(cuda_device) for k, v in rn.items()}
            u01 = [f32(g.uniform(size=(R, 1))).to(cuda_device) for _ in range(3)]
        outs[bf16] = (stage.render(rays, u01), stage.render_unfused(rays, u01))
    chains, layers = outs[True]
    for k in ("transient_direct", "transient_indirect", "rgb"):
        assert float(chains[k].abs().max()) > 0
        assert rel_err(chains[k], layers[k]) <= 2e-2, (k, rel_err(chains[k], layers[k]))
        # against the fp32 variant only a sanity bound: the bf16 density MLPs move the sampler's fenceposts, which shifts
        # energy between neighbouring time bins (measured 6e-2 of the histogram's maximum)
        assert rel_err(chains[k], outs[False][1][k]) <= 0.15, (k, rel_err(chains[k], outs[False][1][k]))
