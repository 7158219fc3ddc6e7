"""Oracle restatement of internal/inverse_render/render_utils.py pieces on the path
(TEST INFRASTRUCTURE ONLY): Disney-GGX lobe and the Monte-Carlo integration of cache
radiance against it.  Material configuration per configs/ngp_yobo.gin:256-303:
use_diffuseness = use_mirrorness = use_specular_albedo = False, brdf_correction = ones.
"""
import numpy as np
import torch

from . import ref_math

EPS = float(np.finfo(np.float32).eps)
DENOMINATOR_EPS = 1e-5  # internal/inverse_render/render_utils.py:41


def _normalize(v):
    """internal/inverse_render/math.py:87-88."""
    return v / ref_math.sqrt(1e-10 + torch.sum(v**2, dim=-1, keepdim=True))


def _dot(x, y):
    """internal/inverse_render/math.py:72-74."""
    return (x * y).sum(dim=-1, keepdim=True)


def GGX_D(costheta, a):
    """internal/inverse_render/render_utils.py:480-482."""
    return a**2 / torch.clamp(np.pi * ((costheta**2 * (a**2 - 1.0) + 1.0)) ** 2, min=EPS)


def get_lobe(wi, wo, normal, materials, brdf_correction, shading):
    """internal/inverse_render/render_utils.py:566-695 for shading in
    {'lambertian', 'microfacet', 'microfacet_diffuse', 'microfacet_specular'}."""
    lobe = torch.clamp(wi[..., 2:], min=0.0) * materials["albedo"][..., None, :] / np.pi
    if "microfacet" in shading:
        roughness = materials["roughness"][..., None, :]
        F_0 = materials["F_0"][..., None, :]
        albedo = materials["albedo"][..., None, :]
        metalness = materials["metalness"][..., None, :]
        specular_albedo = albedo
        mirrorness = torch.ones_like(metalness)
        diffuseness = 1.0 - metalness
        F_0 = specular_albedo * metalness + F_0 * (1.0 - metalness)
        halfdirs = _normalize(wi + wo)
        n_dot_v = torch.clamp(_dot(normal, wo), min=0.0)
        n_dot_l = torch.clamp(_dot(normal, wi), min=0.0)
        n_dot_h = torch.clamp(_dot(normal, halfdirs), min=0.0)
        l_dot_h = torch.clamp(_dot(wi, halfdirs), min=0.0)
        a = roughness
        F = F_0 + (1.0 - F_0) * torch.clamp(1.0 - l_dot_h, 0.0, 1.0) ** 5
        D = GGX_D(n_dot_h, a)
        k = a / 2
        G = (n_dot_v / torch.clamp(n_dot_v * (1.0 - k) + k, min=EPS)) * (
            n_dot_l / torch.clamp(n_dot_l * (1.0 - k) + k, min=EPS)
        )
        ggx_lobe = D * F * G / torch.clamp(4.0 * n_dot_v, min=EPS)
        lambertian_lobe = n_dot_l * albedo / np.pi
        if shading == "microfacet":
            lobe = ggx_lobe * brdf_correction[..., 0:1] * mirrorness + lambertian_lobe * brdf_correction[
                ..., 1:2
            ] * diffuseness
        elif shading == "microfacet_diffuse":
            lobe = (lambertian_lobe * brdf_correction[..., 1:2]) * diffuseness
        elif shading == "microfacet_specular":
            lobe = (ggx_lobe * brdf_correction[..., 0:1]) * mirrorness
    return lobe


def integrate_reflect_rays(material_type, material, samples, max_radiance=float("inf")):
    """internal/inverse_render/render_utils.py:1102-1193 with use_brdf_correction=False."""
    ld = samples["local_lightdirs"]
    local_normals = torch.cat([torch.zeros_like(ld[..., 0:1]), torch.zeros_like(ld[..., 0:1]),
                               torch.ones_like(ld[..., 0:1])], dim=-1)
    lobe = get_lobe(ld, samples["local_viewdirs"], local_normals, material, samples["brdf_correction"],
                    material_type)
    denominator = torch.clamp(samples["pdf"], min=DENOMINATOR_EPS)
    weight = torch.clamp(samples["weight"], min=0.0)
    weight = torch.where(ld[..., 2:] > 0.0, weight, torch.zeros_like(weight))
    radiance_out = (torch.clamp(samples["radiance_in"] * lobe, 0.0, max_radiance) * weight / denominator).mean(1)
    indirect_occ = samples["indirect_occ"].mean(1)
    diffuse_lobe = torch.clamp(ld[..., 2:], min=0.0) / np.pi
    irradiance = (torch.clamp(samples["radiance_in"] * diffuse_lobe, 0.0, max_radiance) * weight / denominator
                  ).mean(1)
    return dict(radiance_out=radiance_out, indirect_occ=indirect_occ, irradiance=irradiance)
