"""K5 parity: GGX lobe + Monte-Carlo integration vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import render_utils as oru
from neural_radiance_caching_b200.inverse_render import render_utils as nru
from tests.util import f32, gen, rel_err, to_dev

pytestmark = pytest.mark.gpu


def _inputs(g, R, S):
    def unit(v):
        return v / np.linalg.norm(v, axis=-1, keepdims=True)
    wi = unit(g.normal(size=(R, S, 3)))           # both hemispheres: below-horizon samples get weight 0
    wo = unit(np.abs(g.normal(size=(R, 1, 3))) + [0, 0, 0.05]) * np.ones((1, S, 1))
    samples = dict(
        local_lightdirs=f32(wi), local_viewdirs=f32(wo), radiance_in=f32(g.gamma(1.0, 1.0, size=(R, S, 3))),
        pdf=f32(g.gamma(1.0, 0.5, size=(R, S, 1))), weight=f32(g.uniform(-0.1, 2.0, size=(R, S, 1))),
        indirect_occ=f32(g.uniform(size=(R, S, 1))), brdf_correction=torch.ones(R, S, 2))
    samples["pdf"][0] = 0.0  # exercises the DENOMINATOR_EPS clamp
    material = dict(albedo=f32(g.uniform(size=(R, 3))), roughness=f32(g.uniform(0.01, 1.0, size=(R, 1))),
                    metalness=f32(g.uniform(size=(R, 1))), F_0=torch.full((R, 1), 0.04))
    return material, samples


@pytest.mark.parametrize("kind", ["microfacet", "microfacet_diffuse", "microfacet_specular", "lambertian"])
@pytest.mark.parametrize("S", [16, 40])
def test_integrate_reflect_rays(cuda_device, kind, S):
    g = gen(200 + S)
    R = 513
    material, samples = _inputs(g, R, S)
    Lo = samples["radiance_in"].clone().requires_grad_(True)
    want = oru.integrate_reflect_rays(kind, material, dict(samples, radiance_in=Lo), max_radiance=10000.0)
    sd = to_dev(samples, cuda_device)
    Ln = sd["radiance_in"].clone().requires_grad_(True)
    got = nru.integrate_reflect_rays(kind, False, to_dev(material, cuda_device), dict(sd, radiance_in=Ln),
                                     max_radiance=10000.0)
    for k in ("radiance_out", "irradiance", "indirect_occ"):
        assert rel_err(got[k], want[k]) <= 1e-5, k
    go, gi = f32(g.normal(size=(R, 3))), f32(g.normal(size=(R, 3)))
    ((want["radiance_out"] * go).sum() + (want["irradiance"] * gi).sum()).backward()
    ((got["radiance_out"] * go.to(cuda_device)).sum() + (got["irradiance"] * gi.to(cuda_device)).sum()).backward()
    assert rel_err(Ln.grad, Lo.grad) <= 1e-5


def test_white_furnace(cuda_device):
    """Property test: a Lambertian surface under uniform radiance L with cosine-weighted
    sampling (pdf = cos/pi, weight 1) returns albedo * L exactly, for any sample set."""
    g = gen(210)
    R, S = 64, 16
    u1, u2 = g.uniform(size=(R, S)), g.uniform(size=(R, S))
    r, phi = np.sqrt(u1), 2 * np.pi * u2
    wi = np.stack([r * np.cos(phi), r * np.sin(phi), np.sqrt(1 - u1)], -1)
    samples = dict(local_lightdirs=f32(wi), local_viewdirs=f32(np.tile([0, 0, 1.0], (R, S, 1))),
                   radiance_in=torch.full((R, S, 3), 2.0), pdf=f32(wi[..., 2:] / np.pi),
                   weight=torch.ones(R, S, 1))
    material = dict(albedo=f32(g.uniform(size=(R, 3))))
    got = nru.integrate_reflect_rays("lambertian", False, to_dev(material, cuda_device), to_dev(samples, cuda_device))
    assert torch.allclose(got["radiance_out"].cpu(), 2.0 * material["albedo"], rtol=1e-5, atol=1e-6)
