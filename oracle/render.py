"""Oracle restatement of internal/render.py (TEST INFRASTRUCTURE ONLY)."""
import numpy as np
import torch

from . import ref_math, stepfun

EPS = float(np.finfo(np.float32).eps)


def gaussianize_frustum(t0, t1):
    """internal/render.py:49-59."""
    s = t0 + t1
    d = t1 - t0
    eps = EPS**2
    ratio = d**2 / torch.clamp(3 * s**2 + d**2, min=eps)
    t_mean = s * (1 / 2 + ratio)
    t_var = (1 / 12) * d**2 - (1 / 15) * ratio**2 * (12 * s**2 - d**2)
    r_var = (1 / 16) * s**2 + d**2 * (5 / 48 - (1 / 15) * ratio)
    return t_mean, t_var, r_var


def lift_gaussian(d, t_mean, t_var, r_var, diag):
    """internal/render.py:26-46."""
    mean = d[..., None, :] * t_mean[..., None]
    d_mag_sq = torch.clamp(torch.sum(d**2, dim=-1, keepdim=True), min=1e-10)
    if diag:
        d_outer_diag = d**2
        null_outer_diag = 1 - d_outer_diag / d_mag_sq
        t_cov_diag = t_var[..., None] * d_outer_diag[..., None, :]
        xy_cov_diag = r_var[..., None] * null_outer_diag[..., None, :]
        return mean, t_cov_diag + xy_cov_diag
    d_outer = d[..., :, None] * d[..., None, :]
    eye = torch.eye(d.shape[-1], dtype=d.dtype)
    null_outer = eye - d[..., :, None] * (d / d_mag_sq)[..., None, :]
    t_cov = t_var[..., None, None] * d_outer[..., None, :, :]
    xy_cov = r_var[..., None, None] * null_outer[..., None, :, :]
    return mean, t_cov + xy_cov


def cast_rays(tdist, origins, directions, radii, ray_shape="cone", diag=True):
    """internal/render.py:106-131 (cone: :62-81, cylinder: :84-103)."""
    assert ray_shape in ("cone", "cylinder")
    t0 = tdist[..., :-1]
    t1 = tdist[..., 1:]
    if ray_shape == "cone":
        t_mean, t_var, r_var = gaussianize_frustum(t0, t1)
        r_var = r_var * radii**2
    else:
        t_mean = (t0 + t1) / 2
        r_var = radii**2 / 4 * torch.ones_like(t_mean)
        t_var = (t1 - t0) ** 2 / 12
    means, covs = lift_gaussian(directions, t_mean, t_var, r_var, diag)
    means = means + origins[..., None, :]
    return means, covs


def compute_alpha_weights(density, tdist, dirs, opaque_background=False, delta=None):
    """internal/render.py:134-169."""
    if delta is None:
        t_delta = tdist[..., 1:] - tdist[..., :-1]
        delta = t_delta * ref_math.sqrt(torch.sum(dirs[..., None, :] ** 2, dim=-1))
    density_delta = density * torch.abs(delta)
    if opaque_background:
        density_delta = torch.cat(
            [density_delta[..., :-1], torch.full_like(density_delta[..., -1:], float("inf"))], dim=-1
        )
    alpha = 1 - torch.exp(-density_delta)
    trans = torch.exp(
        -torch.cat(
            [torch.zeros_like(density_delta[..., :1]), torch.cumsum(density_delta[..., :-1], dim=-1)],
            dim=-1,
        )
    )
    weights = alpha * trans
    return weights, alpha, trans


def volumetric_rendering(
    rgbs,
    weights,
    weights_no_filter,
    tdist,
    bg_rgbs,
    compute_extras,
    extras=None,
    normalize_weights_for_extras=False,
    percentiles=(5, 50, 95),
    compute_distance=True,
):
    """internal/render.py:172-247."""
    eps = EPS
    rendering = {}
    acc = weights_no_filter.sum(dim=-1)
    acc_no_filter = weights_no_filter.sum(dim=-1)
    bg_w = torch.clamp(1 - acc[..., None], min=0)
    if rgbs is not None:
        rgb = (weights[..., None] * rgbs).sum(dim=-2) + bg_w * bg_rgbs
    else:
        rgb = None
    rendering["rgb"] = rgb
    rendering["acc"] = acc
    weights_norm = weights / torch.clamp(acc[..., None], min=eps)
    weights_norm_no_filter = weights_no_filter / torch.clamp(acc_no_filter[..., None], min=eps)
    if extras is not None:
        for k, v in extras.items():
            if v is not None:
                w = weights_norm if normalize_weights_for_extras else weights
                rendering[k] = (w[..., None] * v).sum(dim=-2)
    if compute_distance:
        t_mids = 0.5 * (tdist[..., :-1] + tdist[..., 1:])
        e = (weights_no_filter * torch.log(t_mids)).sum(dim=-1) / torch.clamp(acc_no_filter, min=eps)
        # NB: jnp.nan_to_num(x, jnp.inf) binds inf to `copy`, so NaN -> 0.0 (:235).
        dm = torch.nan_to_num(torch.exp(e), nan=0.0)
        rendering["distance_mean"] = torch.minimum(torch.maximum(dm, tdist[..., 0]), tdist[..., -1])
        pct = stepfun.weighted_percentile(tdist, weights_norm_no_filter, percentiles)
        for i, p in enumerate(percentiles):
            s = "median" if p == 50 else "percentile_" + str(p)
            rendering["distance_" + s] = pct[..., i]
    return rendering
