"""Host-side mirror of the surface-light-field MEMORY variant (second half): the reference's
`surface_lf_mem` = surface_light_field.SurfaceLightFieldMLP with use_distance_prediction + use_reflectance_grid
(internal/surface_light_field.py:594-1069) as internal/models.py:813-833 builds it under
configs/nerf_ngp_yobo.gin:97-165 (+ ngp_yobo.gin:232-236), queried per secondary ray by models.get_slf_results
(models.py:849-908), and the control-variate combination of material._integrate_slf_variate (material.py:2433-2513).

  distance_grid(contract(origin)) -> Dense 64 ReLU x 2          (BaseShader.run_network: the "bottleneck")
  [bottleneck | pos_enc(contract(origin), 0..2) | pos_enc(dir, 0..2)] -> 4 x Dense 128 ReLU (input re-concatenated after
      layer 2) -> distance_output_layer (8 n + 4)                 (tcgen05 chain program)
  -> nrc_slf_points_{fwd,bwd}: n points per ray, their weights, the weighted s-distance, env rgba
  -> reflectance_grid at the n points -> nrc_slf_reduce_{fwd,bwd} (weighted feature sum)
  -> Dense 64 ReLU -> Dense 128 ReLU -> output_rgba_layer (4), output_ambient_rgb_layer (3)      (chain program)

Origins and directions are stop-gradient inputs (models.py:854 utils.partial_stopgrad_rays)."""
import math

import torch

from . import _lib, coord, grid_utils, mlp_chain, nerf

DISTANCE_GRID = dict(hash_map_size=524288, max_grid_size=256, num_features=4)                           # nerf_ngp_yobo.gin:156-161
REFLECTANCE_GRID = dict(hash_map_size=524288, max_grid_size=256, num_features=4, bbox_scaling=2.0)     # nerf_ngp_yobo.gin:146-152


def _points_cfg(net, near, far):
    c = _lib.nrc_slf_points_t()
    c.num_distance_samples = net.n
    c.warp_kind = 0 if net.raydist is None else 1
    c.distance_near, c.distance_far = net.distance_near, net.distance_far
    c.near, c.far = float(near), float(far)
    c.distance_scale, c.distance_bias = net.distance_scale, net.distance_bias
    c.rgb_premultiplier, c.rgb_bias, c.alpha_bias = 1.0, net.rgb_bias, net.alpha_bias
    c.warp_p, c.warp_premult = (net.raydist if net.raydist is not None else (1.0, 1.0))
    c.ref_warp_c = net.warp_c
    return c


class _SlfPointsFn(torch.autograd.Function):
    """predict_points + the weight head of __call__ (surface_light_field.py:594-780, 899-913)."""

    @staticmethod
    def forward(ctx, raw, origins, refdirs, cfg):
        import ctypes as C
        P, n = raw.shape[0], cfg.num_distance_samples
        dev = raw.device
        raw, origins, refdirs = raw.contiguous(), origins.contiguous(), refdirs.contiguous()
        points = torch.empty((P, n, 3), device=dev, dtype=torch.float32)
        weights = torch.empty((P, n), device=dev, dtype=torch.float32)
        s_dist = torch.empty((P, 1), device=dev, dtype=torch.float32)
        distances = torch.empty((P, n), device=dev, dtype=torch.float32)
        env = torch.empty((P, 4), device=dev, dtype=torch.float32)
        _lib.call("nrc_slf_points_fwd", _lib.stream_ptr(), C.byref(cfg), _lib.ptr(raw), raw.shape[1], _lib.ptr(origins),
                  _lib.ptr(refdirs), P, _lib.ptr(points), _lib.ptr(weights), _lib.ptr(s_dist), _lib.ptr(distances),
                  _lib.ptr(env))
        ctx.save_for_backward(raw, origins, refdirs)
        ctx.cfg = cfg
        return points, weights, s_dist, distances, env

    @staticmethod
    def backward(ctx, g_points, g_weights, g_s_dist, g_distances, g_env):
        import ctypes as C
        raw, origins, refdirs = ctx.saved_tensors
        P = raw.shape[0]
        c = lambda g: g.contiguous() if g is not None else None
        g_raw = torch.empty_like(raw)
        _lib.call("nrc_slf_points_bwd", _lib.stream_ptr(), C.byref(ctx.cfg), _lib.ptr(raw), raw.shape[1], _lib.ptr(origins),
                  _lib.ptr(refdirs), P, _lib.ptr(c(g_points)), _lib.ptr(c(g_weights)), _lib.ptr(c(g_s_dist)),
                  _lib.ptr(c(g_distances)), _lib.ptr(c(g_env)), _lib.ptr(g_raw))
        return g_raw, None, None, None


class _SlfReduceFn(torch.autograd.Function):
    """(ref_grid_feat * ref_weights[..., None]).sum(-2) (surface_light_field.py:981)."""

    @staticmethod
    def forward(ctx, feat, weights):
        P, n, F = feat.shape
        feat, weights = feat.contiguous(), weights.contiguous()
        out = torch.empty((P, F), device=feat.device, dtype=torch.float32)
        _lib.call("nrc_slf_reduce_fwd", _lib.stream_ptr(), _lib.ptr(feat), _lib.ptr(weights), P, n, F, _lib.ptr(out))
        ctx.save_for_backward(feat, weights)
        return out

    @staticmethod
    def backward(ctx, g):
        feat, weights = ctx.saved_tensors
        P, n, F = feat.shape
        g_feat, g_w = torch.empty_like(feat), torch.empty_like(weights)
        _lib.call("nrc_slf_reduce_bwd", _lib.stream_ptr(), _lib.ptr(feat), _lib.ptr(weights), _lib.ptr(g.contiguous()), P, n, F,
                  _lib.ptr(g_feat), _lib.ptr(g_w))
        return g_feat, g_w


def _pos_enc(x, max_deg):
    P = x.shape[0]
    w = 3 + 2 * 3 * max_deg
    out = torch.empty((P, w), device=x.device, dtype=torch.float32)
    _lib.call("nrc_pos_enc", _lib.stream_ptr(), _lib.ptr(x.contiguous()), P, 3, 0, max_deg, 1, _lib.ptr(out), w)
    return out


class SurfaceLightFieldMemMLP:
    def __init__(self, num_distance_samples=8, distance_near=5e-2, distance_far=2.0, grid=None, reflectance_grid=None,
                 warp_c=2.0, raydist=(-1.5, 2.0), rgb_bias=-2.0, ambient_rgb_bias=-1.0, alpha_bias=2.0, distance_scale=1.0,
                 distance_bias=-2.0, bf16=True):
        self.n = int(num_distance_samples)
        self.distance_near, self.distance_far = float(distance_near), float(distance_far)
        self.grid = grid_utils.HashEncoding(**(grid or DISTANCE_GRID))                 # scale_supersample: class default
        self.reflectance_grid = grid_utils.HashEncoding(**(reflectance_grid or REFLECTANCE_GRID))
        self.warp_c = warp_c if warp_c is not None else 0.0
        self.raydist = raydist
        self.rgb_bias, self.ambient_rgb_bias, self.alpha_bias = rgb_bias, ambient_rgb_bias, alpha_bias
        self.distance_scale, self.distance_bias = distance_scale, distance_bias
        self.bf16 = bf16
        self.nf, self.nrf = self.grid.num_outputs, self.reflectance_grid.num_outputs
        self.dist_in = 64 + 15 + 15
        self.out_dim = 8 * self.n + 4
        if self.out_dim > 128:
            raise NotImplementedError("more than 15 distance samples: the output layer exceeds one 128-column head group")
        self.distance_chain = mlp_chain.ChainSpec(
            in_widths=[64, 30], hidden=[(f"distance_layer_{i}", 128, (i % 2 == 0 and i > 0)) for i in range(4)],
            heads=[[("distance_output_layer", self.out_dim)]])
        self.rgb_chain = mlp_chain.ChainSpec(
            in_widths=[self.nrf], hidden=[("layer_0", 64, False), ("layer_bottleneck", 128, False)],
            heads=[[("output_rgba_layer", 4), ("output_ambient_rgb_layer", 3)]])
        # render path: the bottleneck stack as one program too (its product is the last ACTIVATION: linear head + ReLU)
        self.bottleneck_chain = mlp_chain.ChainSpec(in_widths=[self.nf], hidden=[("layers_0", 64, False)],
                                                    heads=[[("layers_1", 64)]])
        self._pack_caches = (mlp_chain.PackCache(), mlp_chain.PackCache(), mlp_chain.PackCache())

    def layer_shapes(self):
        s = [("layers_0", self.nf, 64), ("layers_1", 64, 64)]
        d = self.dist_in
        for i in range(4):
            s.append((f"distance_layer_{i}", d, 128))
            d = 128 + (self.dist_in if (i % 2 == 0 and i > 0) else 0)
        s += [("distance_output_layer", d, self.out_dim), ("layer_0", self.nrf, 64), ("layer_bottleneck", 64, 128),
              ("output_rgba_layer", 128, 4), ("output_ambient_rgb_layer", 128, 3)]
        return s

    def init(self, device, generator=None, table_init_range=0.1):
        """Random-init parameters in the reference's layout (he_uniform kernels, zero biases).  The reference zero-initialises
        distance_output_layer (surface_light_field.py:377-379); `init` draws it like the others so that a synthetic workload
        exercises every branch of predict_points."""
        p = {}
        for key, enc in (("distance_grid", self.grid), ("reflectance_grid", self.reflectance_grid)):
            _, arena = enc.init(device, generator=generator, init_range=table_init_range)
            p[key] = dict(enc.views(arena), _arena=arena)
        for name, fi, fo in self.layer_shapes():
            lim = math.sqrt(6.0 / fi)
            p[name] = {"kernel": torch.empty((fi, fo), device=device).uniform_(-lim, lim, generator=generator),
                       "bias": torch.zeros((fo,), device=device)}
        return p

    def from_oracle(self, p, device):
        out = {}
        for key, enc in (("distance_grid", self.grid), ("reflectance_grid", self.reflectance_grid)):
            names = [n for (n, _, _, _) in enc.level_layout]
            arena = torch.cat([p[key][n].detach().reshape(-1) for n in names]).to(device)
            out[key] = dict(enc.views(arena), _arena=arena)
        for name, _, _ in self.layer_shapes():
            out[name] = {k: v.detach().to(device).contiguous() for k, v in p[name].items()}
        return out

    def _stack(self, spec, cache, p, sources):
        if self.bf16 and not torch.is_grad_enabled():
            return mlp_chain.forward_cached(spec, p, sources, cache)
        if self.bf16 and spec.supports_backward:
            return mlp_chain.apply(spec, p, sources)
        # fp32 parity variant / stacks without a data-gradient program: per-layer GEMMs
        enc = torch.cat(sources, dim=-1) if len(sources) > 1 else sources[0]
        x = enc
        for name, _, skip in spec.hidden:
            x = nerf.dense(p[name], x, relu=True, bf16=self.bf16)
            if skip:
                x = torch.cat([x, enc], dim=-1)
        return [nerf.dense(p[name], x, bf16=self.bf16) for grp in spec.heads for name, _ in grp]

    def bottleneck(self, p, origins):
        """predict_appearance_feature (shading.py:133-220) with one control point: grid features through run_network."""
        z = coord._ContractFn.apply(origins, self.warp_c)
        x = self.grid(p["distance_grid"], z).reshape(-1, self.nf)
        if self.bf16 and not torch.is_grad_enabled():
            (h,) = mlp_chain.forward_cached(self.bottleneck_chain, p, [x], self._pack_caches[2])
            return torch.relu(h), z
        x = nerf.dense(p["layers_0"], x, relu=True, bf16=self.bf16)
        return nerf.dense(p["layers_1"], x, relu=True, bf16=self.bf16), z

    def run_distances_network(self, p, bottleneck, z, refdirs):
        """surface_light_field.py:414-442; `z` = warp_fn(origins)."""
        enc = torch.cat([_pos_enc(z.detach(), 2), _pos_enc(refdirs, 2)], dim=-1)
        (raw,) = self._stack(self.distance_chain, self._pack_caches[0], p, [bottleneck, enc])
        return raw

    def __call__(self, p, origins, refdirs, near=0.0, far=float("inf")):
        """origins / refdirs [...,3] -> the reference's `incoming_*` dict, one entry per ray."""
        lead = origins.shape[:-1]
        o2, d2 = origins.detach().reshape(-1, 3).contiguous(), refdirs.detach().reshape(-1, 3).contiguous()
        P = o2.shape[0]
        bott, z = self.bottleneck(p, o2)
        raw = self.run_distances_network(p, bott, z, d2)
        points, w, s_dist, dist, env = _SlfPointsFn.apply(raw, o2, d2, _points_cfg(self, near, far))
        feat = self.reflectance_grid(p["reflectance_grid"], points.reshape(-1, 3)).reshape(P, self.n, self.nrf)
        x = _SlfReduceFn.apply(feat, w)
        rgba, amb = self._stack(self.rgb_chain, self._pack_caches[1], p, [x])
        sp = torch.nn.functional.softplus
        rgb = sp(rgba[:, :3] + self.rgb_bias)                                                    # :1045-1047 (clip at 0: no-op)
        alpha = torch.sigmoid(rgba[:, 3:4] + self.alpha_bias)                                    # :1048-1051
        out = dict(incoming_rgb=rgb, incoming_ambient_rgb=sp(amb + self.ambient_rgb_bias), incoming_alpha=alpha,
                   incoming_weights=w, incoming_s_dist=s_dist, incoming_dist=dist, incoming_env_rgba=env,
                   incoming_acc=w.sum(dim=-1))
        return {k: v.reshape(lead + v.shape[1:]) for k, v in out.items()}

    def get_slf_results(self, p, origins, viewdirs, near=0.0, far=float("inf")):
        """models.get_slf_results (models.py:849-908) as material._make_surface_lf_fn calls it (use_env_map=False: no
        environment composite), and the clamp of surface_lf_fn (material.py:2273): rgb [...,3], acc [...]."""
        res = self(p, origins, viewdirs, near, far)
        return dict(res, rgb=torch.clamp(res["incoming_rgb"], min=0.0), acc=res["incoming_acc"])


def integrate_slf_variate(integrated_cache, integrated_slf):
    """material._integrate_slf_variate (material.py:2484-2513): the cache integral minus the light-field integral for the
    radiance keys, every entry of both kept under `_cache` / `_slf` suffixes."""
    final = dict(integrated_cache)
    for k in ("radiance_out", "diffuse_radiance_out", "specular_radiance_out", "direct_radiance_out", "indirect_radiance_out",
              "irradiance"):
        if k in integrated_cache and k in integrated_slf:
            final[k] = integrated_cache[k] - integrated_slf[k]
    for k in list(final.keys()):
        final[k + "_cache"] = integrated_cache.get(k)
        final[k + "_slf"] = integrated_slf.get(k)
    return final
