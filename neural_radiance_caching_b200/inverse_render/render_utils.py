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
