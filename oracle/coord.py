"""Oracle restatement of internal/coord.py pieces on the path (TEST INFRASTRUCTURE ONLY)."""
import numpy as np
import torch

from . import ref_math


def contract(x):
    """internal/coord.py:63-69."""
    x_mag_sq = torch.clamp(torch.sum(x**2, dim=-1, keepdim=True), min=1.0)
    scale = (2 * ref_math.sqrt(x_mag_sq) - 1) / x_mag_sq
    return scale * x


def contract_radius(x, c):
    """internal/coord.py:33-38 (contract_radius_5 / contract_radius_2)."""
    return contract(x / c)


def make_warp(c):
    if c is None:
        return None
    return lambda x: contract_radius(x, c)


def construct_ray_warps(fn, t_near, t_far, fn_inv=None):
    """internal/coord.py:223-260."""
    if fn is None:
        fn_fwd = lambda x: x
        fn_inv = lambda x: x
    else:
        fn_fwd = fn
        assert fn_inv is not None
    s_near, s_far = fn_fwd(t_near), fn_fwd(t_far)
    t_to_s = lambda t: (fn_fwd(t) - s_near) / (s_far - s_near)
    s_to_t = lambda s: fn_inv(s * s_far + (1 - s) * s_near)
    return t_to_s, s_to_t


def power_ladder_warps(t_near, t_far, p=-1.5, premult=2.0):
    """configs/ngp_yobo.gin:238-242: raydist_fn = (power_ladder, inv_power_ladder, {p, premult})."""
    fn = lambda x: ref_math.power_ladder(x, p, premult=premult)
    fn_inv = lambda y: ref_math.inv_power_ladder(y, p, premult=premult)
    return construct_ray_warps(fn, t_near, t_far, fn_inv)


def pos_enc(x, min_deg, max_deg, append_identity=True):
    """internal/coord.py:298-312."""
    scales = torch.tensor([2.0**i for i in range(min_deg, max_deg)], dtype=torch.float32)
    shape = x.shape[:-1] + (-1,)
    scaled_x = (x[..., None, :] * scales[:, None]).reshape(shape)
    four_feat = torch.sin(torch.cat([scaled_x, scaled_x + np.float32(0.5 * np.pi)], dim=-1))
    if append_identity:
        return torch.cat([x, four_feat], dim=-1)
    return four_feat
