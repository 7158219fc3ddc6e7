"""CPU restatement (test infrastructure) of the time-resolved rendering path, BASELINE config 4:
internal/render.py volumetric_transient_rendering (:250-449, the rgb / transient outputs), shift_direct
(:452-490), shift_map_coordinates (:493-507; jax.scipy.ndimage.map_coordinates order 1, mode='constant'),
internal/inverse_render/render_utils.py zero_invalid_bins (:1699-1767) and the head post-processing of
internal/nerf.py:1660-1777 (softplus(raw + bias) * indirect_scale; tint * F * ref_rgb * indirect_scale; clip).
Pinned to the reference's source for shift_direct, shift_map_coordinates, zero_invalid_bins and the whole of
volumetric_transient_rendering (tests/test_reference_vectors.py::test_transient; map_coordinates there is SciPy's with
JAX's boundary rule); the head (get_indirect Dense stack + post-processing) is pinned through
TransientNeRFMLP._compute_indirect_lighting executed from the reference's class (test_transient_head)."""
import numpy as np
import torch


def shift_direct(dists, direct_rgbs, weights, n_bins, C):
    R, n = dists.shape
    rgb = torch.zeros((R * n_bins, C), dtype=direct_rgbs.dtype)
    lo = torch.clamp(torch.floor(dists), min=0)
    hi = torch.ceil(dists)
    w_hi = dists - lo
    w_lo = 1.0 - w_hi
    base = torch.arange(R).repeat_interleave(n) * n_bins
    vals = (weights[..., None] * direct_rgbs).reshape(-1, C)
    for idx, w in ((base + lo.reshape(-1).to(torch.int32), w_lo), (base + hi.reshape(-1).to(torch.int32), w_hi)):
        ok = (idx >= 0) & (idx < R * n_bins)              # out-of-bounds scatter updates are dropped
        rgb.index_add_(0, idx[ok].long(), (vals * w.reshape(-1, 1))[ok])
    return rgb.reshape(R, n_bins, C)


def shift_map_coordinates(x, bins_move, exposure_time, n_bins):
    """x [N, n_bins, C]: out[i, b] = linear interpolation of x[i] at b - bins_move[i]/exposure, zero outside."""
    y = torch.arange(n_bins, dtype=x.dtype)[None, :] - (bins_move / exposure_time)[:, None]
    y0 = torch.floor(y)
    t = (y - y0)[..., None]
    i0 = y0.long()

    def tap(i):
        ok = ((i >= 0) & (i < n_bins))[..., None]
        g = torch.gather(x, 1, i.clamp(0, n_bins - 1)[..., None].expand(-1, -1, x.shape[-1]))
        return torch.where(ok, g, torch.zeros_like(g))

    return (1.0 - t) * tap(i0) + t * tap(i0 + 1)


def zero_invalid_bins(diffuse, specular, light_dists, cam_dists, n_bins, exposure_time, bin_zero_threshold_light,
                      light_zero, light_near):
    """diffuse / specular [..., n_bins, C]; light_dists / cam_dists [..., 1]."""
    bins = torch.arange(n_bins, dtype=diffuse.dtype).reshape((1,) * (diffuse.dim() - 2) + (n_bins, 1))
    too_close = (bins + bin_zero_threshold_light) * exposure_time < light_dists[..., None, :]
    too_far = (bins * exposure_time + cam_dists[..., None, :]) > (n_bins - 1) * exposure_time
    bad = too_close | too_far
    if light_zero:
        bad = bad | (light_dists[..., None, :] < light_near)
    z = lambda v: torch.where(bad, torch.zeros_like(v), v)
    return z(diffuse), z(specular)


def volumetric_transient_rendering(direct_rgbs, indirect, weights, ray_dists, light_dists, n_bins, exposure_time=0.01,
                                   shift=0.0, dark_level=0.0):
    """The transient outputs of render.volumetric_transient_rendering (:322-449) without temporal filter / median filter /
    vis offsets (all off in the configs): direct splat at (light + ray distance) / exposure, indirect histograms shifted
    by the ray distance and reduced with the weights."""
    R, n, C = direct_rgbs.shape
    d = (light_dists + ray_dists) / exposure_time
    t_direct = shift_direct(d + shift / exposure_time, direct_rgbs, weights, n_bins, C)
    shifted = shift_map_coordinates(indirect.reshape(R * n, n_bins, C), ray_dists.reshape(-1) + shift, exposure_time, n_bins)
    t_indirect = (shifted.reshape(R, n, n_bins, C) * weights[..., None, None]).sum(1)
    return dict(transient_direct=t_direct, transient_indirect=t_indirect, rgb=t_direct + t_indirect + dark_level)


def transient_head(diffuse_raw, specular, spec_scale, light_dists, cam_dists, n_bins, exposure_time=0.01, diffuse_bias=-1.0,
                   indirect_scale=1.0, bin_zero_threshold_light=0.0, light_zero=False, light_near=0.0, rgb_max=10000.0):
    """Post-processing of TransientNeRFMLP._compute_indirect_lighting (nerf.py:1689-1745): diffuse = softplus(raw + bias) *
    scale; specular = (tint * F) * ref_rgb * scale; zero_invalid_bins; clip to [0, rgb_max].  diffuse_raw / specular
    [..., n_bins, C]; spec_scale [..., C]; light_dists / cam_dists [...]."""
    diffuse = torch.nn.functional.softplus(diffuse_raw + diffuse_bias) * indirect_scale
    spec = spec_scale[..., None, :] * specular * indirect_scale
    diffuse, spec = zero_invalid_bins(diffuse, spec, light_dists[..., None], cam_dists[..., None], n_bins, exposure_time,
                                      bin_zero_threshold_light, light_zero, light_near)
    return torch.clamp(diffuse, 0.0, rgb_max), torch.clamp(spec, 0.0, rgb_max)


def transient_render(direct_rgbs, diffuse_raw, specular, spec_scale, weights, ray_dists, light_dists, cam_dists, n_bins,
                     exposure_time=0.01, shift=0.0, diffuse_bias=-1.0, indirect_scale=1.0, bin_zero_threshold_light=0.0,
                     light_zero=False, light_near=0.0, rgb_max=10000.0, dark_level=0.0):
    diffuse, spec = transient_head(diffuse_raw, specular, spec_scale, light_dists, cam_dists, n_bins, exposure_time,
                                   diffuse_bias, indirect_scale, bin_zero_threshold_light, light_zero, light_near, rgb_max)
    indirect = diffuse + spec
    return volumetric_transient_rendering(direct_rgbs, indirect, weights, ray_dists, light_dists, n_bins, exposure_time, shift,
                                          dark_level)


def gaussian_tfilter(tfilter_sigma):
    """The temporal filter volumetric_transient_rendering builds from tfilter_sigma (render.py:401-404): a Gaussian over
    round(-4 sigma) .. round(4 sigma) bins minus exp(-8), normalised."""
    k = np.arange(round(-4 * tfilter_sigma), round(4 * tfilter_sigma) + 1).astype(np.float32)
    f = np.exp(-(k ** 2) / np.float32(2 * tfilter_sigma ** 2)).astype(np.float32) - np.float32(np.exp(-8))
    return torch.from_numpy((f / f.sum()).astype(np.float32))


def temporal_filter(x, filt):
    """jax.scipy.signal.convolve(x, filt[None, :, None], mode='same') (render.py:406-413): x [R, n_bins, C]."""
    R, B, C = x.shape
    taps = filt.shape[0]
    xp = torch.nn.functional.pad(x.permute(0, 2, 1).reshape(R * C, 1, B), (taps - 1, taps - 1))
    full = torch.nn.functional.conv1d(xp, filt.flip(0).reshape(1, 1, taps))[:, 0]          # full convolution, length B + taps - 1
    lo = (taps - 1) // 2
    return full[:, lo:lo + B].reshape(R, C, B).permute(0, 2, 1).contiguous()


def transient_integrate_reflect_rays(lobe, weight, pdf, local_lightdirs, radiance_in, indirect_occ=None, max_radiance=float("inf")):
    """transient_integrate_reflect_rays with direct=False (render_utils.py:1195-1302): lobe [R,S,3] (get_lobe),
    weight / pdf [R,S,1], radiance_in [R,S,n_bins,3] -> radiance_out, irradiance [R,n_bins,3], indirect_occ [R,1]."""
    den = torch.clamp(pdf, min=1e-5)
    w = torch.clamp(weight, min=0.0)
    w = torch.where(local_lightdirs[..., 2:] > 0.0, w, torch.zeros_like(w))
    dl = torch.clamp(local_lightdirs[..., 2:], min=0.0) / np.pi
    ro = (torch.clamp(radiance_in * lobe[..., None, :], 0.0, max_radiance) * w[..., None, :] / den[..., None, :]).mean(1)
    ir = (torch.clamp(radiance_in * dl[..., None, :], 0.0, max_radiance) * w[..., None, :] / den[..., None, :]).mean(1)
    return dict(radiance_out=ro, irradiance=ir, indirect_occ=indirect_occ.mean(1) if indirect_occ is not None else None)
