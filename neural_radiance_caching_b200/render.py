"""Host-side mirror of internal/render.py for the hot path (CUDA bodies, same signatures)."""
import torch

from . import _lib


def _c(t):
    return t.contiguous() if t is not None else None


class _AlphaWeightsFn(torch.autograd.Function):
    """custom_vjp analogue over nrc_ray_alpha_weights_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, density, tdist, dirs, opaque_background):
        n = density.shape[-1]
        d2, t2, r2 = _c(density.reshape(-1, n)), _c(tdist.reshape(-1, n + 1)), _c(dirs.reshape(-1, 3))
        R = d2.shape[0]
        w, a, tr = torch.empty_like(d2), torch.empty_like(d2), torch.empty_like(d2)
        _lib.call("nrc_ray_alpha_weights_fwd", _lib.stream_ptr(), _lib.ptr(d2), _lib.ptr(t2), _lib.ptr(r2), R, n,
                  int(opaque_background), _lib.ptr(w), _lib.ptr(a), _lib.ptr(tr))
        ctx.save_for_backward(d2, t2, r2)
        ctx.shape = density.shape
        ctx.opaque = opaque_background
        return w.reshape(density.shape), a.reshape(density.shape), tr.reshape(density.shape)

    @staticmethod
    def backward(ctx, gw, ga, gt):
        d2, t2, r2 = ctx.saved_tensors
        if ctx.opaque:
            raise NotImplementedError("gradient through opaque_background is outside the configs' scope")
        R, n = d2.shape
        gd = torch.empty_like(d2)
        f = lambda g: _c(g.reshape(R, n)) if g is not None else None
        _lib.call("nrc_ray_alpha_weights_bwd", _lib.stream_ptr(), _lib.ptr(d2), _lib.ptr(t2), _lib.ptr(r2),
                  _lib.ptr(f(gw)), _lib.ptr(f(ga)), _lib.ptr(f(gt)), R, n, _lib.ptr(gd))
        return gd.reshape(ctx.shape), None, None, None


def compute_alpha_weights(density, tdist, dirs, opaque_background=False, delta=None):
    """Helper function for computing alpha compositing weights (internal/render.py:134-169).

    Returns (weights, alpha, trans).  `delta` overrides are not used on the hot path.
    """
    if delta is not None:
        raise NotImplementedError("explicit `delta` is outside the CUDA path's scope")
    return _AlphaWeightsFn.apply(density, tdist, dirs, bool(opaque_background))


def cast_rays(tdist, origins, directions, radii, ray_shape, diag=True, want_covs=True):
    """Cast cone- or cylinder-shaped rays (internal/render.py:106-131): returns (means, covs) like the reference.

    Under unscented basis 'mean' only the means feed the encoding: the sampler's fused launches never
    compute covariances, and `want_covs=False` skips them here too (covs=None).  With want_covs the covariances come from
    nrc_ray_cast_covs: [...,n,3] (diag) or [...,n,3,3].  Forward only (the reference's callers on this path stop gradients
    into the Gaussians' shapes)."""
    if ray_shape not in ("cone", "cylinder"):
        raise ValueError("ray_shape must be 'cone' or 'cylinder'")
    n = tdist.shape[-1] - 1
    t2 = _c(tdist.reshape(-1, n + 1))
    R = t2.shape[0]
    o2, d2 = _c(origins.reshape(-1, 3)), _c(directions.reshape(-1, 3))
    means = torch.empty((R, n, 3), device=t2.device, dtype=torch.float32)
    cyl = ray_shape == "cylinder"
    if not cyl:
        zeros, ones = torch.zeros(R, device=t2.device), torch.ones(R, device=t2.device)
        t_out = torch.empty_like(t2)
        # identity warp: near=0, far=1 => t = s*1 + (1-s)*0
        _lib.call("nrc_ray_cast", _lib.stream_ptr(), _lib.ptr(t2), _lib.ptr(o2), _lib.ptr(d2), _lib.ptr(zeros),
                  _lib.ptr(ones), R, n, 0, 0.0, 1.0, _lib.ptr(t_out), _lib.ptr(means))
    covs = None
    if want_covs or cyl:
        r2 = _c(radii.reshape(-1).expand(R) if radii.numel() == 1 else radii.reshape(R))
        if want_covs:
            covs = torch.empty((R, n, 3) if diag else (R, n, 3, 3), device=t2.device, dtype=torch.float32)
        _lib.call("nrc_ray_cast_covs", _lib.stream_ptr(), _lib.ptr(t2), _lib.ptr(o2), _lib.ptr(d2), _lib.ptr(r2), R, n,
                  int(cyl), int(bool(diag)), _lib.ptr(covs), _lib.ptr(means) if cyl else None)
        if covs is not None:
            covs = covs.reshape(tdist.shape[:-1] + covs.shape[1:])
    return means.reshape(tdist.shape[:-1] + (n, 3)), covs


class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, weights, weights_nf, tdist, bg, has_rgb, want_dist):
        k = weights.shape[-1]
        C = values.shape[-1]
        w2 = _c(weights.reshape(-1, k))
        R = w2.shape[0]
        v2 = _c(values.reshape(R, k, C))
        same = weights_nf is None
        wn2 = None if same else _c(weights_nf.reshape(R, -1))
        n = k if same else wn2.shape[-1]
        t2 = _c(tdist.reshape(R, n + 1)) if tdist is not None else None
        b2 = _c(bg.expand(weights.shape[:-1] + (3,)).reshape(R, 3)) if bg is not None else None
        out = torch.empty((R, C), device=w2.device, dtype=torch.float32)
        acc = torch.empty((R,), device=w2.device, dtype=torch.float32)
        dist = torch.empty((R, 4), device=w2.device, dtype=torch.float32) if want_dist else None
        _lib.call("nrc_ray_composite_fwd", _lib.stream_ptr(), _lib.ptr(v2), _lib.ptr(w2), k, _lib.ptr(wn2),
                  _lib.ptr(t2), _lib.ptr(b2), R, n, C, int(has_rgb), _lib.ptr(out), _lib.ptr(acc), _lib.ptr(dist))
        ctx.save_for_backward(v2, w2, wn2, b2)
        ctx.meta = (values.shape, weights.shape, None if same else weights_nf.shape, k, n, C, has_rgb)
        lead = weights.shape[:-1]
        return out.reshape(lead + (C,)), acc.reshape(lead), (dist.reshape(lead + (4,)) if want_dist else None)

    @staticmethod
    def backward(ctx, g_out, g_acc, _g_dist):
        v2, w2, wn2, b2 = ctx.saved_tensors
        vshape, wshape, wnshape, k, n, C, has_rgb = ctx.meta
        R = w2.shape[0]
        go = _c(g_out.reshape(R, C)) if g_out is not None else torch.zeros((R, C), device=w2.device)
        ga = _c(g_acc.reshape(R)) if g_acc is not None else None
        gv, gw = torch.empty_like(v2), torch.empty_like(w2)
        gwn = torch.empty_like(wn2) if wn2 is not None else None
        _lib.call("nrc_ray_composite_bwd", _lib.stream_ptr(), _lib.ptr(v2), _lib.ptr(w2), k, _lib.ptr(wn2),
                  _lib.ptr(b2), _lib.ptr(go), _lib.ptr(ga), R, n, C, int(has_rgb), _lib.ptr(gv), _lib.ptr(gw),
                  _lib.ptr(gwn))
        return (gv.reshape(vshape), gw.reshape(wshape), gwn.reshape(wnshape) if gwn is not None else None, None,
                None, None, None)


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
    """Volumetric Rendering Function (internal/render.py:172-247).

    rgb and every entry of `extras` are composited by ONE kernel launch over the
    channel-concatenated values; distances come from the same launch.
    """
    if normalize_weights_for_extras:
        raise NotImplementedError("normalize_weights_for_extras=True is outside the configs' scope")
    if tuple(percentiles) != (5, 50, 95):
        raise NotImplementedError("the CUDA path computes percentiles (5, 50, 95)")
    names, chunks = [], []
    if rgbs is not None:
        chunks.append(rgbs)
    if extras is not None:
        for k_, v in extras.items():
            if v is not None:
                names.append((k_, v.shape[-1]))
                chunks.append(v)
    nf = None if weights_no_filter is weights else weights_no_filter
    if chunks:
        values = torch.cat(chunks, dim=-1) if len(chunks) > 1 else chunks[0]
    else:
        values = torch.zeros(weights.shape + (0,), device=weights.device)
    bg = None
    if rgbs is not None and bg_rgbs is not None:
        if isinstance(bg_rgbs, (int, float)):
            # fill on the device: no host->device copy (keeps the call CUDA-graph capturable)
            bg = torch.full(weights.shape[:-1] + (3,), float(bg_rgbs), device=weights.device, dtype=torch.float32)
        else:
            bg = torch.as_tensor(bg_rgbs, device=weights.device, dtype=torch.float32)
            bg = bg.expand(weights.shape[:-1] + (3,))
    out, acc, dist = _CompositeFn.apply(values, weights, nf, tdist, bg, rgbs is not None, bool(compute_distance))
    rendering = {}
    off = 0
    if rgbs is not None:
        rendering["rgb"] = out[..., :3]
        off = 3
    else:
        rendering["rgb"] = None
    rendering["acc"] = acc
    for k_, c in names:
        rendering[k_] = out[..., off:off + c]
        off += c
    if compute_distance:
        rendering["distance_mean"] = dist[..., 0]
        for i, p in enumerate(percentiles):
            s = "median" if p == 50 else "percentile_" + str(p)
            rendering["distance_" + s] = dist[..., 1 + i]
    return rendering


class _TransientRenderFn(torch.autograd.Function):
    """custom_vjp of the time-resolved rendering: nrc_transient_render_{fwd,bwd}.  Differentiable inputs: direct_rgbs,
    diffuse_raw, specular, spec_scale, weights; the distances are stop-gradient inputs."""

    @staticmethod
    def forward(ctx, direct_rgbs, diffuse_raw, specular, spec_scale, weights, ray_dists, light_dists, cam_dists, consts):
        R, n, C = direct_rgbs.shape
        n_bins = consts[0]
        dev = direct_rgbs.device
        c = lambda t: t.contiguous() if t is not None else None
        new = lambda: torch.empty((R, n_bins, C), device=dev, dtype=torch.float32)
        t_direct, t_indirect, rgb = new(), new(), new()
        args = [c(direct_rgbs), c(diffuse_raw), c(specular), c(spec_scale), c(weights), c(ray_dists), c(light_dists), c(cam_dists)]
        _lib.call("nrc_transient_render_fwd", _lib.stream_ptr(), *[_lib.ptr(a) for a in args], R, n, n_bins, C, *consts[1:],
                  _lib.ptr(t_direct), _lib.ptr(t_indirect), _lib.ptr(rgb))
        ctx.save_for_backward(*[a if a is not None else torch.empty(0, device=dev) for a in args])
        ctx.present = [a is not None for a in args]
        ctx.consts = consts
        return t_direct, t_indirect, rgb

    @staticmethod
    def backward(ctx, g_direct, g_indirect, g_rgb):
        args = [a if ok else None for a, ok in zip(ctx.saved_tensors, ctx.present)]
        direct_rgbs, diffuse_raw, specular, spec_scale = args[:4]
        R, n, C = direct_rgbs.shape
        consts = ctx.consts
        n_bins = consts[0]
        dev = direct_rgbs.device

        def total(g):     # rgb = direct + indirect + dark_level
            parts = [t for t in (g, g_rgb) if t is not None]
            return None if not parts else (parts[0] if len(parts) == 1 else parts[0] + parts[1]).contiguous()

        gd, gi = total(g_direct), total(g_indirect)
        need = ctx.needs_input_grad
        g_dir = torch.empty_like(direct_rgbs)
        g_w = torch.empty((R, n), device=dev, dtype=torch.float32)
        g_raw = torch.empty_like(diffuse_raw) if (diffuse_raw is not None and need[1]) else None
        g_spec = torch.empty_like(specular) if (specular is not None and need[2]) else None
        g_ss = torch.empty_like(spec_scale) if (specular is not None and need[3]) else None
        _lib.call("nrc_transient_render_bwd", _lib.stream_ptr(), *[_lib.ptr(a) for a in args], R, n, n_bins, C, *consts[1:-1],
                  _lib.ptr(gd), _lib.ptr(gi), _lib.ptr(g_dir), _lib.ptr(g_raw), _lib.ptr(g_spec), _lib.ptr(g_ss), _lib.ptr(g_w))
        return g_dir, g_raw, g_spec, g_ss, g_w, None, None, None, None


def volumetric_transient_rendering(direct_rgbs, diffuse_raw, specular, spec_scale, weights, ray_dists, light_dists,
                                   cam_dists, n_bins=700, exposure_time=0.01, shift=0.0, diffuse_bias=-1.0,
                                   indirect_scale=1.0, bin_zero_threshold_light=0.0, light_zero=False, light_near=0.0,
                                   rgb_max=10000.0, dark_level=0.0):
    """Time-resolved rendering (internal/render.py:250-449) fused with the transient heads' post-processing
    (internal/nerf.py:1660-1777) and zero_invalid_bins (render_utils.py:1699-1767): see nrc_transient_render_fwd.
    direct_rgbs [R,n,C]; diffuse_raw / specular [R,n,B,C] (either may be None); spec_scale [R,n,C];
    weights / ray_dists / light_dists / cam_dists [R,n].  Returns dict(transient_direct, transient_indirect,
    rgb [R,B,C], integrated_rgb [R,C]).  Differentiable (nrc_transient_render_bwd) with respect to direct_rgbs, diffuse_raw,
    specular, spec_scale and weights; the temporal filter is applied by temporal_filter (differentiable too)."""
    consts = (int(n_bins), float(exposure_time), float(shift), float(diffuse_bias), float(indirect_scale),
              float(bin_zero_threshold_light), int(bool(light_zero)), float(light_near), float(min(rgb_max, 3.0e38)),
              float(dark_level))
    t_direct, t_indirect, rgb = _TransientRenderFn.apply(direct_rgbs, diffuse_raw, specular, spec_scale, weights,
                                                         ray_dists.detach(), light_dists.detach(), cam_dists.detach(), consts)
    return dict(transient_direct=t_direct, transient_indirect=t_indirect, rgb=rgb, integrated_rgb=rgb.sum(-2))


_tfilter_cache = {}


def gaussian_tfilter(tfilter_sigma, device):
    """The temporal filter of volumetric_transient_rendering (internal/render.py:401-404): Gaussian over
    round(-4 sigma) .. round(4 sigma) bins minus exp(-8), normalised; as a device tensor [taps] (cached per sigma and
    device: no host-to-device copy inside a captured CUDA graph)."""
    import numpy as np
    key = (float(tfilter_sigma), str(device))
    if key in _tfilter_cache:
        return _tfilter_cache[key]
    k = np.arange(round(-4 * tfilter_sigma), round(4 * tfilter_sigma) + 1).astype(np.float32)
    f = np.exp(-(k ** 2) / np.float32(2 * tfilter_sigma ** 2)).astype(np.float32) - np.float32(np.exp(-8))
    _tfilter_cache[key] = torch.from_numpy((f / f.sum()).astype(np.float32)).to(device)
    return _tfilter_cache[key]


class _TemporalFilterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, filt):
        x = _c(x)
        R, B, Cc = x.shape
        y = torch.empty_like(x)
        filt = _c(filt)
        _lib.call("nrc_transient_filter", _lib.stream_ptr(), _lib.ptr(x), _lib.ptr(filt), int(filt.shape[0]), R, B, Cc, _lib.ptr(y))
        ctx.save_for_backward(filt)
        return y

    @staticmethod
    def backward(ctx, g):
        (filt,) = ctx.saved_tensors
        taps = int(filt.shape[0])
        if taps % 2 == 0:
            raise NotImplementedError("the adjoint of a 'same' convolution with an even number of taps is not a 'same' convolution")
        # adjoint of convolve(mode='same') with an odd filter = the same convolution with the filter reversed
        g = _c(g)
        R, B, Cc = g.shape
        gx = torch.empty_like(g)
        _lib.call("nrc_transient_filter", _lib.stream_ptr(), _lib.ptr(g), _lib.ptr(_c(filt.flip(0))), taps, R, B, Cc, _lib.ptr(gx))
        return gx, None


def temporal_filter(x, filt):
    """jax.scipy.signal.convolve(x, filt[None, :, None], mode='same') along the bin axis (internal/render.py:406-413):
    x [R, n_bins, C], filt [taps] (impulse response or gaussian_tfilter); differentiable with respect to x."""
    return _TemporalFilterFn.apply(x, filt)


def volumetric_transient_rendering_fused(direct_rgbs, h_diffuse, diffuse_layer, h_specular, specular_layer, spec_scale, weights,
                                         ray_dists, light_dists, cam_dists, n_bins=700, exposure_time=0.01, shift=0.0,
                                         diffuse_bias=-2.0, spec_premult=1.0, spec_bias=-2.0, spec_max=float("inf"),
                                         indirect_scale=1.0, bin_zero_threshold_light=0.0, light_zero=False, light_near=0.0,
                                         rgb_max=10000.0, dark_level=0.0, impulse_response=None, tfilter_sigma=0.0,
                                         filter_indirect=False, pack_cache=None):
    """volumetric_transient_rendering (internal/render.py:250-449) with the LAST LAYER of the transient heads inside the
    kernel (nrc_transient_head_render_fwd): h_diffuse [R,n,64] is the irradiance stack's last hidden activation and
    diffuse_layer = {'kernel': [64, n_bins*C], 'bias'} its transient_indirect_layer (internal/nerf.py:1757-1777);
    h_specular [R,n,128] / specular_layer the transient SurfaceLightField's last activation and output_rgba_layer
    ([128, n_bins*C + 1]; the alpha column is not used); spec_scale [R,n,C] = tint * integrated BRDF.  The per-sample
    histograms [R,n,n_bins,C] are never materialised.  Then the temporal filter (impulse_response / tfilter_sigma,
    filter_indirect) as in the reference.  Forward path."""
    R, n, Cc = direct_rgbs.shape
    dev = direct_rgbs.device
    new = lambda: torch.empty((R, n_bins, Cc), device=dev, dtype=torch.float32)
    t_direct, t_indirect, rgb = new(), new(), new()
    hd = _c(h_diffuse.reshape(R * n, -1)) if h_diffuse is not None else None
    hs = _c(h_specular.reshape(R * n, -1)) if h_specular is not None else None
    wd, bd = (diffuse_layer["kernel"], diffuse_layer["bias"]) if hd is not None else (None, None)
    ws, bs = (specular_layer["kernel"], specular_layer["bias"]) if hs is not None else (None, None)
    big = 3.0e38
    # both kernels as one bf16 operand image, repacked when a kernel tensor changed (pack_cache: a dict the caller keeps)
    key = tuple((id(t), t.data_ptr(), t._version) for t in (wd, ws) if t is not None)
    cache = pack_cache if pack_cache is not None else {}
    repack = cache.get("key") != key
    if repack:
        cache["key"], cache["keep"] = key, (wd, ws)
        cache["buf"] = torch.empty((n_bins * Cc * 192,), device=dev, dtype=torch.bfloat16)
    packed = cache["buf"]
    _lib.call("nrc_transient_head_render_fwd", _lib.stream_ptr(), _lib.ptr(_c(direct_rgbs)),
              _lib.ptr(hd), 64 if hd is not None else 0, _lib.ptr(wd), int(wd.shape[1]) if wd is not None else 0, _lib.ptr(bd),
              _lib.ptr(hs), 128 if hs is not None else 0, _lib.ptr(ws), int(ws.shape[1]) if ws is not None else 0, _lib.ptr(bs),
              _lib.ptr(_c(spec_scale)) if spec_scale is not None else None, _lib.ptr(_c(weights)), _lib.ptr(_c(ray_dists)),
              _lib.ptr(_c(light_dists)), _lib.ptr(_c(cam_dists)), R, n, n_bins, Cc, float(exposure_time), float(shift),
              float(diffuse_bias), float(spec_premult), float(spec_bias), float(min(spec_max, big)), float(indirect_scale),
              float(bin_zero_threshold_light), int(bool(light_zero)), float(light_near), float(min(rgb_max, big)), float(dark_level),
              _lib.ptr(packed), int(repack), _lib.ptr(t_direct), _lib.ptr(t_indirect), _lib.ptr(rgb))
    out = dict(transient_direct_no_filter=t_direct, transient_indirect_no_filter=t_indirect)
    if impulse_response is not None or tfilter_sigma != 0.0:
        filt = impulse_response if impulse_response is not None else gaussian_tfilter(tfilter_sigma, dev)
        t_direct = temporal_filter(t_direct, filt)
        if filter_indirect:
            t_indirect = temporal_filter(t_indirect, filt)
        rgb = t_direct + t_indirect + dark_level
    out.update(transient_direct=t_direct, transient_indirect=t_indirect, rgb=rgb, integrated_rgb=rgb.sum(-2),
               direct_rgb=t_direct.sum(-2), indirect_rgb=t_indirect.sum(-2))
    return out
