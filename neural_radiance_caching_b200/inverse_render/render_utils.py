"""Host-side mirror of internal/inverse_render/render_utils.py for the hot path (CUDA bodies)."""
import torch

from .. import _lib

_LOBE_KINDS = {"microfacet": 0, "microfacet_diffuse": 1, "microfacet_specular": 2, "lambertian": 3}
DENOMINATOR_EPS = 1e-5  # internal/inverse_render/render_utils.py:41


class _IntegrateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, rgb_max, wi, wo, radiance, weight, pdf, occ, albedo, rough, metal, f0):
        R, S = wi.shape[0], wi.shape[1]
        dev = wi.device
        out = torch.empty((R, 3), device=dev, dtype=torch.float32)
        irr = torch.empty((R, 3), device=dev, dtype=torch.float32)
        occ_out = torch.empty((R,), device=dev, dtype=torch.float32) if occ is not None else None
        _lib.call("nrc_ggx_integrate_fwd", _lib.stream_ptr(), _lib.ptr(wi), _lib.ptr(wo), _lib.ptr(radiance),
                  _lib.ptr(weight), _lib.ptr(pdf), _lib.ptr(occ), _lib.ptr(albedo), _lib.ptr(rough),
                  _lib.ptr(metal), _lib.ptr(f0), R, S, kind, float(rgb_max), _lib.ptr(out), _lib.ptr(irr),
                  _lib.ptr(occ_out))
        ctx.save_for_backward(wi, wo, radiance, weight, pdf, albedo, rough, metal, f0)
        ctx.meta = (kind, rgb_max, R, S)
        return out, irr, occ_out

    @staticmethod
    def backward(ctx, g_out, g_irr, _g_occ):
        wi, wo, radiance, weight, pdf, albedo, rough, metal, f0 = ctx.saved_tensors
        kind, rgb_max, R, S = ctx.meta
        g_rad = torch.empty_like(radiance)
        c = lambda g: g.contiguous() if g is not None else None
        _lib.call("nrc_ggx_integrate_bwd", _lib.stream_ptr(), _lib.ptr(wi), _lib.ptr(wo), _lib.ptr(radiance),
                  _lib.ptr(weight), _lib.ptr(pdf), _lib.ptr(albedo), _lib.ptr(rough), _lib.ptr(metal), _lib.ptr(f0),
                  _lib.ptr(c(g_out)), _lib.ptr(c(g_irr)), R, S, kind, float(rgb_max), _lib.ptr(g_rad))
        return (None, None, None, None, g_rad) + (None,) * 7


def integrate_reflect_rays(
    material_type,
    use_brdf_correction,
    material,
    samples,
    use_diffuseness=False,
    use_mirrorness=False,
    use_specular_albedo=False,
    max_radiance=float("inf"),
):
    """internal/inverse_render/render_utils.py:1102-1193.

    material: dict albedo [R,3], roughness [R,1], metalness [R,1], F_0 [R,1];
    samples: dict local_lightdirs / local_viewdirs / radiance_in [R,S,3], pdf / weight /
    indirect_occ [R,S,1].  The brdf-correction / diffuseness / mirrorness / specular-albedo
    branches are disabled in every BASELINE config (configs/ngp_yobo.gin:256-303) and raise.
    Gradients flow to `radiance_in` (the cache); material gradients are a later row.
    """
    if use_brdf_correction or use_diffuseness or use_mirrorness or use_specular_albedo:
        raise NotImplementedError("brdf correction / diffuseness / mirrorness / specular albedo are out of scope")
    if material_type not in _LOBE_KINDS:
        raise ValueError(f"unsupported material_type {material_type}")
    wi = samples["local_lightdirs"].contiguous()
    R, S = wi.shape[0], wi.shape[1]
    wo = samples["local_viewdirs"].expand(R, S, 3).contiguous()
    f = lambda k: samples[k].reshape(R, S).contiguous()
    m = lambda k: material[k].reshape(R).contiguous() if k in material else None
    occ = f("indirect_occ") if "indirect_occ" in samples else None
    out, irr, occ_out = _IntegrateFn.apply(
        _LOBE_KINDS[material_type], max_radiance if max_radiance != float("inf") else 3.4e38, wi, wo,
        samples["radiance_in"].contiguous(), f("weight"), f("pdf"), occ,
        material["albedo"].reshape(R, 3).contiguous(), m("roughness"), m("metalness"), m("F_0"))
    res = dict(radiance_out=out, irradiance=irr)
    if occ_out is not None:
        res["indirect_occ"] = occ_out[:, None]
    if "brdf_correction" in samples:
        res["integrated_multiplier"] = samples["brdf_correction"][:, 0]
        res["integrated_multiplier_irradiance"] = samples["brdf_correction"][:, 0, :1]
    return res


# ----------------------------------------------------------------------------- secondary rays
class CosineSampler:
    """internal/inverse_render/render_utils.py:417-444 (body: nrc_secondary_sample)."""
    global_dirs = False


class MicrofacetSampler:
    """internal/inverse_render/render_utils.py:485-546 (body: nrc_secondary_sample)."""
    global_dirs = False


class LightSampler:
    """vMF-mixture sampler, internal/inverse_render/render_utils.py:1419-1490 (body: nrc_secondary_sample)."""
    global_dirs = True


_SAMPLER_ORDER = (MicrofacetSampler, CosineSampler, LightSampler)


def get_secondary_rays(
    rng,
    rays,
    means,
    viewdirs,
    normals,
    material,
    normal_eps=1e-2,
    refdir_eps=1e-2,
    random_generator_2d=None,
    stratified_sampling=False,
    use_mis=True,
    samplers=None,
    num_secondary_samples=None,
    light_sampler_results=None,
    offset_origins=False,
    light_rotation=None,
    far=None,
):
    """Secondary rays for one sampler set (internal/inverse_render/render_utils.py:927-1056 over
    importance_sample_rays :722-924), one shaded point per ray.

    `rng` carries the random draws instead of a JAX key (the kernels never generate randomness):
    dict(u=[R,S,2] uniforms; latent=[R] int32 and normal2=[R,n_light,2] when a LightSampler is present).
    `samplers` = [(sampler, count)] with sampler classes in the order Microfacet, Cosine, Light and counts
    summing to num_secondary_samples.  means / viewdirs / normals [R,3]; material['roughness'] [R,1];
    light_sampler_results = dict(vmf_means [R,K,3], vmf_kappas [R,K,1], vmf_logits [R,K,1]).
    Returns (ref_rays, ref_samples) like the reference (rays as a dict of [R,S,...] tensors)."""
    if not use_mis or offset_origins or light_rotation is not None or stratified_sampling:
        raise NotImplementedError("the CUDA path covers use_mis=True without origin offsets / light rotation")
    counts = {MicrofacetSampler: 0, CosineSampler: 0, LightSampler: 0}
    order = []
    for s_, c_ in samplers:
        cls = s_ if isinstance(s_, type) else type(s_)
        if cls not in counts:
            raise ValueError(f"unsupported sampler {cls}")
        counts[cls] += int(c_)
        order.append(cls)
    if order != [c for c in _SAMPLER_ORDER if c in order]:
        raise ValueError("samplers must be ordered Microfacet, Cosine, Light")
    S = sum(counts.values())
    if num_secondary_samples is not None and S != num_secondary_samples:
        raise NotImplementedError("resampling of secondary samples (sum of counts != num_secondary_samples)")
    R = means.shape[0]
    dev = means.device
    c = lambda t: t.contiguous() if t is not None else None
    new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    origins, dirs, lld, lvd, pdf, weight = new(R, S, 3), new(R, S, 3), new(R, S, 3), new(R, 3), new(R, S), new(R, S)
    nl = counts[LightSampler]
    lsr = light_sampler_results if nl > 0 else None
    vm = c(lsr["vmf_means"]) if lsr else None
    vk = c(lsr["vmf_kappas"].reshape(R, -1)) if lsr else None
    vl = c(lsr["vmf_logits"].reshape(R, -1)) if lsr else None
    lat = c(rng["latent"].to(torch.int32)) if nl > 0 else None
    n2 = c(rng["normal2"]) if nl > 0 else None
    u = c(rng["u"])
    rough = c(material["roughness"].reshape(R))
    m2, v2, nn = c(means), c(viewdirs), c(normals)
    _lib.call("nrc_secondary_sample", _lib.stream_ptr(), _lib.ptr(m2), _lib.ptr(v2), _lib.ptr(nn), _lib.ptr(rough), R,
              counts[MicrofacetSampler], counts[CosineSampler], nl, _lib.ptr(u), _lib.ptr(vm), _lib.ptr(vk), _lib.ptr(vl),
              int(vk.shape[1]) if vk is not None else 0, _lib.ptr(lat), _lib.ptr(n2), float(normal_eps), _lib.ptr(origins),
              _lib.ptr(dirs), _lib.ptr(lld), _lib.ptr(lvd), _lib.ptr(pdf), _lib.ptr(weight))
    far_v = float(far) if far is not None else None
    ref_rays = dict(origins=origins, directions=dirs, viewdirs=dirs, radii=torch.ones((R, S, 1), device=dev),
                    near=torch.full((R, S, 1), float(refdir_eps), device=dev),
                    far=torch.full((R, S, 1), far_v, device=dev) if far_v is not None
                    else rays["far"].reshape(R, 1, 1).expand(R, S, 1).contiguous())
    ref_samples = dict(local_lightdirs=lld, local_viewdirs=lvd[:, None, :].expand(R, S, 3),
                       global_lightdirs=dirs, global_viewdirs=(-v2)[:, None, :].expand(R, S, 3),
                       pdf=pdf[..., None], weight=weight[..., None])
    return ref_rays, ref_samples


def microfacet_material(brdf_params, min_roughness=0.01, default_F_0=0.04):
    """MaterialMLP._get_microfacet_material (internal/material.py:1290-1322) for pred_brdf_layer's raw
    output [..., 10]: dict albedo [...,3], roughness / metalness / F_0 / specular_albedo [...,1]."""
    lead = brdf_params.shape[:-1]
    raw = brdf_params.reshape(-1, brdf_params.shape[-1]).contiguous()
    P = raw.shape[0]
    dev = raw.device
    new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    albedo, rough, metal, f0, spec = new(P, 3), new(P), new(P), new(P), new(P)
    _lib.call("nrc_material_head", _lib.stream_ptr(), _lib.ptr(raw), raw.shape[1], P, float(min_roughness),
              float(default_F_0), _lib.ptr(albedo), _lib.ptr(rough), _lib.ptr(metal), _lib.ptr(f0), _lib.ptr(spec))
    r = lambda t, *s: t.reshape(lead + s)
    return dict(albedo=r(albedo, 3), roughness=r(rough, 1), metalness=r(metal, 1), F_0=r(f0, 1),
                specular_albedo=r(spec, 1))


def transient_integrate_reflect_rays(material_type, use_brdf_correction, material, samples, use_diffuseness=False,
                                     use_mirrorness=False, use_specular_albedo=False, direct=True, max_radiance=float("inf")):
    """internal/inverse_render/render_utils.py:1195-1302.  direct=True is integrate_reflect_rays on [R,S,3] radiance;
    direct=False takes a HISTOGRAM of incoming radiance per secondary ray, samples['radiance_in'] [R,S,n_bins,3], and returns
    radiance_out / irradiance [R,n_bins,3] (nrc_ggx_integrate_transient_fwd).  Forward path."""
    if direct:
        return integrate_reflect_rays(material_type, use_brdf_correction, material, samples, use_diffuseness, use_mirrorness,
                                      use_specular_albedo, max_radiance)
    if use_brdf_correction or use_diffuseness or use_mirrorness or use_specular_albedo:
        raise NotImplementedError("brdf correction / diffuseness / mirrorness / specular albedo are out of scope")
    if material_type not in _LOBE_KINDS:
        raise ValueError(f"unsupported material_type {material_type}")
    wi = samples["local_lightdirs"].contiguous()
    R, S = wi.shape[0], wi.shape[1]
    wo = samples["local_viewdirs"].expand(R, S, 3).contiguous()
    rad = samples["radiance_in"].contiguous()
    B = rad.shape[2]
    f = lambda k: samples[k].reshape(R, S).contiguous()
    m = lambda k: material[k].reshape(R).contiguous() if k in material else None
    occ = f("indirect_occ") if "indirect_occ" in samples else None
    dev = wi.device
    out = torch.empty((R, B, 3), device=dev, dtype=torch.float32)
    irr = torch.empty((R, B, 3), device=dev, dtype=torch.float32)
    occ_out = torch.empty((R,), device=dev, dtype=torch.float32) if occ is not None else None
    _lib.call("nrc_ggx_integrate_transient_fwd", _lib.stream_ptr(), _lib.ptr(wi), _lib.ptr(wo), _lib.ptr(rad), _lib.ptr(f("weight")),
              _lib.ptr(f("pdf")), _lib.ptr(occ), _lib.ptr(material["albedo"].reshape(R, 3).contiguous()), _lib.ptr(m("roughness")),
              _lib.ptr(m("metalness")), _lib.ptr(m("F_0")), R, S, B, _LOBE_KINDS[material_type],
              float(max_radiance if max_radiance != float("inf") else 3.4e38), _lib.ptr(out), _lib.ptr(irr), _lib.ptr(occ_out))
    res = dict(radiance_out=out, irradiance=irr, indirect_occ=occ_out[:, None] if occ_out is not None else None)
    if "brdf_correction" in samples:
        res["integrated_multiplier"] = samples["brdf_correction"][:, 0]
        res["integrated_multiplier_irradiance"] = samples["brdf_correction"][:, 0, :1]
    return res
