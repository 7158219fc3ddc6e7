"""Host-side mirror of the material stage (BASELINE config 3):
internal/material.py MaterialMLP (predict_bottleneck_feature :1901-1926, _predict_material_and_feature
:2073-2123, _get_microfacet_material :1290-1322), the secondary-ray pass (get_secondary_rays,
render_utils.py:927-1056), the radiance-cache recursion (_make_radiance_cache_fn :2174-2231 ->
models.NeRFModel(is_secondary=True, resample=True)), the model-level environment map
(_make_env_map_fn :2283-2314 -> Model._handle_env_map, models.py:360-421, configs/nerf_ngp_yobo.gin:253-297)
and the GGX Monte-Carlo integration (integrate_reflect_rays, render_utils.py:1102-1193).

Forward (render-time) path; every body is a CUDA kernel behind the C ABI.  Random draws are inputs."""
import numpy as np
import torch

from . import _lib, coord, grid_utils, mlp_chain, nerf
from .inverse_render import render_utils

MATERIAL_GRID = dict(hash_map_size=524288, max_grid_size=2048, num_features=4)   # configs/ngp_yobo.gin:329-333


def _he_uniform(gen, device, fan_in, fan_out):
    lim = float(np.sqrt(6.0 / fan_in))
    return torch.empty((fan_in, fan_out), device=device).uniform_(-lim, lim, generator=gen)


def _layer(gen, device, fi, fo):
    return {"kernel": _he_uniform(gen, device, fi, fo), "bias": torch.zeros((fo,), device=device)}


class MaterialMLP:
    """material grid (L=8, F=4) -> bottleneck Dense 32->128 (linear) -> pred_brdf_layer 128->10 -> microfacet
    material (net_depth 0, use_density_feature False: configs/ngp_yobo.gin:315-333)."""

    def __init__(self, warp_c=2.0, bbox_scaling=1.0, min_roughness=0.01, default_F_0=0.04, bf16=True):
        self.grid = grid_utils.HashEncoding(bbox_scaling=bbox_scaling, scale_supersample=1.0, **MATERIAL_GRID)
        self.warp_c = warp_c if warp_c is not None else 0.0
        self.min_roughness, self.default_F_0 = min_roughness, default_F_0
        self.bf16 = bf16
        self.chain = mlp_chain.ChainSpec(in_widths=[self.grid.num_outputs],
                                         hidden=[("bottleneck_layer", 128, False, "linear")],
                                         heads=[[("pred_brdf_layer", 10)]])
        self._pack_cache = mlp_chain.PackCache()

    def init(self, device, generator=None, table_init_range=0.1):
        _, arena = self.grid.init(device, generator=generator, init_range=table_init_range)
        return {"material_grid": dict(self.grid.views(arena), _arena=arena),
                "bottleneck_layer": _layer(generator, device, self.grid.num_outputs, 128),
                "pred_brdf_layer": _layer(generator, device, 128, 10)}

    def from_oracle(self, p, device):
        names = [n for (n, _, _, _) in self.grid.level_layout]
        arena = torch.cat([p["material_grid"][n].detach().reshape(-1) for n in names]).to(device)
        out = {k: {kk: vv.detach().to(device).contiguous() for kk, vv in v.items()} for k, v in p.items()
               if k != "material_grid"}
        out["material_grid"] = dict(self.grid.views(arena), _arena=arena)
        return out

    def predict_material(self, p, means):
        """-> (bottleneck feature [...,128] or None, material dict) for shaded points `means` [...,3]."""
        lead = means.shape[:-1]
        z = coord._ContractFn.apply(means, self.warp_c)
        enc = self.grid(p["material_grid"], z).reshape(-1, self.grid.num_outputs)
        if self.bf16 and not torch.is_grad_enabled():
            (raw,) = mlp_chain.forward_cached(self.chain, p, [enc], self._pack_cache)
        elif self.bf16:
            (raw,) = mlp_chain.apply(self.chain, p, [enc])
        else:
            raw = nerf.dense(p["pred_brdf_layer"], nerf.dense(p["bottleneck_layer"], enc))
        mat = render_utils.microfacet_material(raw, self.min_roughness, self.default_F_0)
        return {k: v.reshape(lead + (v.shape[-1],)) for k, v in mat.items()}


class EnvMapMLP:
    """Model-level environment map (models.py:801-812 with NeRFModel.env_map_params,
    configs/nerf_ngp_yobo.gin:253-297): pos_enc(viewdirs, 0..4) (27) -> 4 x Dense 256 + ReLU with the input
    re-concatenated after layer 2 -> output_rgba_layer (4) [+ output_ambient_rgb_layer (3)];
    incoming_rgb = softplus(rgba[:3] + rgb_bias).  Under autograd (material-stage training) the parameter gradients come
    from per-layer GEMMs; the view directions are not differentiated (secondary-ray directions enter the environment map
    through utils.partial_stopgrad_rays, models.py:380)."""

    def __init__(self, deg_view=4, width=256, depth=4, skip=2, rgb_bias=-1.0, bf16=True):
        self.deg_view, self.width, self.depth, self.skip, self.rgb_bias = deg_view, width, depth, skip, rgb_bias
        self.in_dim = 3 + 2 * 3 * deg_view
        self.bf16 = bf16
        self.names = [f"layer_{i}" for i in range(depth - 1)] + ["layer_bottleneck"]
        self.chain = mlp_chain.ChainSpec(
            in_widths=[self.in_dim],
            hidden=[(n, width, (i % skip == 0 and i > 0)) for i, n in enumerate(self.names)],
            heads=[[("output_rgba_layer", 4), ("output_ambient_rgb_layer", 3)]])
        self._pack_cache = mlp_chain.PackCache()

    def init(self, device, generator=None):
        p, d = {}, self.in_dim
        for i, n in enumerate(self.names):
            p[n] = _layer(generator, device, d, self.width)
            d = self.width + (self.in_dim if (i % self.skip == 0 and i > 0) else 0)
        p["output_rgba_layer"] = _layer(generator, device, d, 4)
        p["output_ambient_rgb_layer"] = _layer(generator, device, d, 3)
        return p

    def from_oracle(self, p, device):
        return {k: {kk: vv.detach().to(device).contiguous() for kk, vv in v.items()} for k, v in p.items()}

    def __call__(self, p, viewdirs):
        lead = viewdirs.shape[:-1]
        v2 = viewdirs.reshape(-1, 3).contiguous()
        P = v2.shape[0]
        enc = torch.empty((P, self.in_dim), device=v2.device, dtype=torch.float32)
        _lib.call("nrc_pos_enc", _lib.stream_ptr(), _lib.ptr(v2), P, 3, 0, self.deg_view, 1, _lib.ptr(enc), self.in_dim)
        if self.bf16 and not torch.is_grad_enabled():
            # render path: the whole stack as one tcgen05 chain program
            rgba, amb = mlp_chain.forward_cached(self.chain, p, [enc], self._pack_cache)
        else:
            # fp32 parity variant, and TRAINING of the bf16 variant: the chain kernel's data- / weight-gradient programs
            # stop at 128-wide layers, so a 256-wide stack trains through per-layer GEMMs (nrc_dense_{fwd,bwd}: bf16
            # mma.sync operands, fp32 accumulation) with the same rounding points as the chain program
            x = enc
            for i, n in enumerate(self.names):
                x = nerf.dense(p[n], x, relu=True, bf16=self.bf16)
                if i % self.skip == 0 and i > 0:
                    x = torch.cat([x, enc], dim=-1)
            rgba = nerf.dense(p["output_rgba_layer"], x, bf16=self.bf16)
            amb = nerf.dense(p["output_ambient_rgb_layer"], x, bf16=self.bf16)
        rgb = torch.nn.functional.softplus(rgba[:, :3] + self.rgb_bias)
        return dict(incoming_rgb=rgb.reshape(lead + (3,)), incoming_alpha=rgba[:, 3:4].reshape(lead + (1,)),
                    incoming_ambient_raw=amb.reshape(lead + (3,)))


class MaterialModel:
    """One chunk of the material stage's render path (BASELINE config 3): for every shaded surface point
    32 secondary rays (16 microfacet for the specular lobe; 8 cosine + 8 vMF-mixture for the diffuse lobe,
    material.py:846-884,1593-1607) -> radiance-cache query per ray (power-ladder warp, categorical resample
    to one shaded sample) + environment map behind it -> GGX / Lambert Monte-Carlo integration."""

    def __init__(self, cache_model, bf16=True, n_specular=16, n_cosine=8, n_light=8, near_min=0.05, far=2.0,
                 normal_eps=1e-2, rgb_max=10000.0, slf_variate=False, surface_lf_mem=None):
        self.cache = cache_model
        # MaterialModel.slf_variate (nerf_ngp_yobo.gin:91) with NeRFModel.use_surface_light_field: the light field
        # `surface_lf_mem` (models.py:813-833; distance_far = Config.env_map_distance) is queried on the SAME secondary rays
        # and its integral subtracted (material._integrate_slf_variate, material.py:2433-2513)
        self.slf_variate = slf_variate
        self.surface_lf_mem = surface_lf_mem
        if slf_variate and surface_lf_mem is None:
            from . import surface_light_field
            self.surface_lf_mem = surface_light_field.SurfaceLightFieldMemMLP(distance_far=far, bf16=bf16)
        # bf16 variant: the cache recursion runs as a hand-ordered launch schedule (engine.FusedCacheQuery)
        from . import engine
        self.fused_query = engine.FusedCacheQuery(cache_model) if bf16 else None
        self.material_mlp = MaterialMLP(bf16=bf16)
        self.env_map = EnvMapMLP(bf16=bf16)
        self.n_specular, self.n_cosine, self.n_light = n_specular, n_cosine, n_light
        self.near_min, self.far, self.normal_eps, self.rgb_max = near_min, far, normal_eps, rgb_max

    @property
    def num_secondary(self):
        return self.n_specular + self.n_cosine + self.n_light

    def secondary_rays(self, means, viewdirs, normals, material, draws, light_sampler_results):
        """Both sampler sets; returns (rays dict of [R*S,...] tensors, specular samples, diffuse samples)."""
        ru = render_utils
        R = means.shape[0]
        ns, nd = self.n_specular, self.n_cosine + self.n_light
        rays_s, smp_s = ru.get_secondary_rays(
            dict(u=draws["u"][:, :ns]), None, means, viewdirs, normals, material, normal_eps=self.normal_eps,
            refdir_eps=self.near_min, samplers=[(ru.MicrofacetSampler, ns)], num_secondary_samples=ns, far=self.far)
        dsamplers = [(ru.CosineSampler, self.n_cosine)] + ([(ru.LightSampler, self.n_light)] if self.n_light else [])
        rng_d = dict(u=draws["u"][:, ns:].contiguous())
        if self.n_light:
            rng_d.update(latent=draws["latent"], normal2=draws["normal2"])
        rays_d, smp_d = ru.get_secondary_rays(
            rng_d, None, means, viewdirs, normals, material, normal_eps=self.normal_eps, refdir_eps=self.near_min,
            samplers=dsamplers, num_secondary_samples=nd, light_sampler_results=light_sampler_results, far=self.far)
        S = ns + nd
        rays = {k: torch.cat([rays_s[k], rays_d[k]], dim=1).reshape(R * S, -1).contiguous()
                for k in ("origins", "directions", "near", "far", "radii")}
        rays["viewdirs"] = rays["directions"]
        return rays, smp_s, smp_d

    def render_chunk(self, params, means, viewdirs, normals, draws, material=None, light_sampler_results=None):
        """means / viewdirs / normals [R,3].  draws: u [R,S,2]; latent [R], normal2 [R,n_light,2] (light sampler);
        u01 = 3 x [R*S,1] (proposal levels of the cache query); gumbel [R*S, 32, 1] (categorical resample).
        Returns rgb [R,3] = specular + diffuse outgoing radiance and the per-component results."""
        R = means.shape[0]
        S = self.num_secondary
        with torch.no_grad():
            if material is None:
                material = self.material_mlp.predict_material(params["Material"], means)
            rays, smp_s, smp_d = self.secondary_rays(means, viewdirs, normals, material, draws, light_sampler_results)
            if self.fused_query is not None:
                q = self.fused_query(params["Cache"], rays, draws["u01"], gumbel=draws["gumbel"], is_secondary=True,
                                     resample=True)
                rgb_raw, acc = q["rgb"], q["acc"]
            else:
                out = self.cache(params["Cache"], rays, draws["u01"], gumbel=draws["gumbel"], train=False,
                                 is_secondary=True, resample=True)
                rgb_raw, acc = out["render"]["rgb"], out["render"]["acc"]
            rgb = torch.clamp(torch.nan_to_num(rgb_raw), min=0.0)
            env = self.env_map(params["EnvMap"], rays["directions"])["incoming_rgb"]
            radiance_in = (rgb + env * (1.0 - acc)[:, None]).reshape(R, S, 3)     # models.py:423-460
            ns = self.n_specular
            occ = acc.reshape(R, S, 1)
            spec = render_utils.integrate_reflect_rays(
                "microfacet_specular", False, material,
                dict(smp_s, radiance_in=radiance_in[:, :ns].contiguous(), indirect_occ=occ[:, :ns]), max_radiance=self.rgb_max)
            diff = render_utils.integrate_reflect_rays(
                "microfacet_diffuse", False, material,
                dict(smp_d, radiance_in=radiance_in[:, ns:].contiguous(), indirect_occ=occ[:, ns:]), max_radiance=self.rgb_max)
            out = dict(rgb=spec["radiance_out"] + diff["radiance_out"], specular=spec, diffuse=diff, material=material,
                       radiance_in=radiance_in, acc=acc.reshape(R, S), rays=rays)
            if self.slf_variate:
                # second get_outgoing_radiance call with radiance_cache_fn = surface_lf_fn on the rays and samples of the
                # first (last_integrated_outputs, material.py:1378-1409): models.get_slf_results per ray, no environment
                # composite (use_env_map=False, material.py:2246-2257), clamp at zero (:2273)
                from . import surface_light_field
                slf = self.surface_lf_mem.get_slf_results(params["SurfaceLightFieldMem"], rays["origins"], rays["viewdirs"])
                rad_slf = slf["rgb"].reshape(R, S, 3)
                occ_slf = slf["acc"].reshape(R, S, 1)
                spec_l = render_utils.integrate_reflect_rays(
                    "microfacet_specular", False, material,
                    dict(smp_s, radiance_in=rad_slf[:, :ns].contiguous(), indirect_occ=occ_slf[:, :ns]), max_radiance=self.rgb_max)
                diff_l = render_utils.integrate_reflect_rays(
                    "microfacet_diffuse", False, material,
                    dict(smp_d, radiance_in=rad_slf[:, ns:].contiguous(), indirect_occ=occ_slf[:, ns:]), max_radiance=self.rgb_max)
                merged = surface_light_field.integrate_slf_variate(
                    dict(radiance_out=out["rgb"], specular_radiance_out=spec["radiance_out"], diffuse_radiance_out=diff["radiance_out"]),
                    dict(radiance_out=spec_l["radiance_out"] + diff_l["radiance_out"], specular_radiance_out=spec_l["radiance_out"],
                         diffuse_radiance_out=diff_l["radiance_out"]))
                out.update(merged, rgb=merged["radiance_out"], rgb_cache=merged["radiance_out_cache"],
                           rgb_slf=merged["radiance_out_slf"], radiance_in_slf=rad_slf, acc_slf=occ_slf.reshape(R, S))
        return out
