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
