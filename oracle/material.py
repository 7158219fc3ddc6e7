"""CPU restatement (test infrastructure -- never imported by the product path) of the material stage's
secondary-ray machinery: internal/inverse_render/render_utils.py get_rotation_matrix (:145-168),
global_to_local / local_to_global (:698-710), CosineSampler (:417-444), MicrofacetSampler (:485-546),
eval_vmf / sample_vmf / LightSampler (:1335-1490), importance_sample_rays (:722-924, the
num_real_samples == num_secondary_samples branch with MIS power heuristic), get_secondary_rays
(:927-1056); internal/material.py _get_microfacet_material (:1290-1322) with the property table
(:957-1023) under configs/ngp_yobo.gin:256-303.  Random draws are INPUTS (uniforms, the vMF latent index,
the 2-D normal pairs), as at the C ABI: the JAX host keeps its threefry streams.

Pinned to the reference's source (tests/test_reference_vectors.py: the shading frame, both analytic samplers, the vMF
mixture pdf, eval_vmf and the whole of importance_sample_rays with 2 and 3 samplers, executed from
/root/reference under the NumPy stand-in for jax); unpinned against XLA's own rounding (JAX is not installable here).
The material / environment MLP classes are restated only."""
import numpy as np
import torch

from . import ref_math

DENOMINATOR_EPS = 1e-5
EPS = float(np.finfo(np.float32).eps)


def get_rotation_matrix(normal):
    """render_utils.py:145-168 (y_up=False): columns (new_x, new_y, new_z = normal)."""
    old_z = torch.tensor([[0.0, 0.0, 1.0]], dtype=normal.dtype)
    old_y = torch.tensor([[0.0, 1.0, 0.0]], dtype=normal.dtype)
    up = torch.where(torch.abs(normal[..., 2:3]) < 0.9, old_z, old_y)
    new_x = torch.cross(up.expand_as(normal), normal, dim=-1)
    new_x = new_x / (torch.linalg.norm(new_x, dim=-1, keepdim=True) + 1e-10)
    new_z = normal
    new_y = torch.cross(new_z, new_x, dim=-1)
    new_y = new_y / (torch.linalg.norm(new_y, dim=-1, keepdim=True) + 1e-10)
    return torch.stack([new_x, new_y, new_z], dim=-1)


def global_to_local(d, R):
    return d[..., 0:1] * R[..., 0, :] + d[..., 1:2] * R[..., 1, :] + d[..., 2:3] * R[..., 2, :]


def local_to_global(d, R):
    return d[..., 0:1] * R[..., 0] + d[..., 1:2] * R[..., 1] + d[..., 2:3] * R[..., 2]


def normalize(v):
    """internal/inverse_render/math.py:85-86."""
    return v / ref_math.sqrt(1e-10 + torch.sum(v**2, dim=-1, keepdim=True))


def reflect(w, v):
    return 2.0 * (v * w).sum(dim=-1, keepdim=True) * v - w


def GGX_D(costheta, a):
    return a**2 / torch.clamp(np.pi * (costheta**2 * (a**2 - 1.0) + 1.0) ** 2, min=EPS)


# ------------------------------------------------------------------------------------ samplers
class CosineSampler:
    global_dirs = False

    def sample_directions(self, u1, u2, wo, alpha, aux):
        r = ref_math.sqrt(u1)
        phi = u2 * 2.0 * np.pi - np.pi
        x, y = r * torch.cos(phi), r * torch.sin(phi)
        z = ref_math.sqrt(torch.clamp(1.0 - x**2 - y**2, min=DENOMINATOR_EPS))
        return torch.stack([x, y, z], dim=-1), torch.clamp(z / np.pi, min=0.0)

    def pdf(self, wo, wi, alpha, aux):
        p = wi[..., 2] / np.pi
        return torch.clamp(torch.where(wi[..., 2] < 0, torch.zeros_like(p), p), min=0.0)


class MicrofacetSampler:
    global_dirs = False

    def sample_directions(self, u1, u2, wo, alpha, aux):
        a = alpha[..., 0]
        tan2 = a**2 * u1 / torch.clamp(1.0 - u1, min=EPS)
        cost = 1.0 / ref_math.sqrt(torch.clamp(1.0 + tan2, min=EPS))
        sint = ref_math.sqrt(torch.clamp(1.0 - cost**2, min=DENOMINATOR_EPS))
        phi = u2 * 2.0 * np.pi - np.pi
        normals = torch.stack([sint * torch.cos(phi), sint * torch.sin(phi), cost], dim=-1)
        npdf = torch.clamp(GGX_D(cost, a) * torch.abs(cost), min=0.0)
        dirs = reflect(wo, normals)
        dotp = torch.sum(wo * normals, dim=-1)
        pdf = npdf * (1.0 / torch.clamp(4.0 * dotp, min=EPS))
        pdf = torch.where(dotp <= 0.0, torch.zeros_like(pdf), pdf)
        return normalize(dirs), torch.clamp(pdf, min=0.0)

    def pdf(self, wo, wi, alpha, aux):
        normals = normalize(wo + wi)
        dotp = torch.sum(wo * normals, dim=-1)
        pdf = GGX_D(normals[..., 2], alpha[..., 0]) * torch.abs(normals[..., 2]) * (1.0 / torch.clamp(4.0 * dotp, min=EPS))
        pdf = torch.where(dotp <= 0.0, torch.zeros_like(pdf), pdf)
        return torch.clamp(pdf, min=0.0)


def eval_vmf(x, means, kappa):
    """render_utils.py:1335-1346."""
    vals = kappa * ref_math.safe_exp(kappa * torch.sum(x * means, dim=-1)) / (4 * np.pi * torch.sinh(kappa))
    return torch.where(kappa <= EPS, torch.ones_like(vals) / (4.0 * np.pi), vals)


class LightSampler:
    """vMF-mixture sampler (render_utils.py:1419-1490); aux = dict(vmf_means [R,K,3], vmf_kappas [R,K,1],
    vmf_logits [R,K,1], latent [R] int64, normal2 [R,S,2], u [R,S])."""
    global_dirs = True

    def _vars(self, aux):
        means = ref_math.l2_normalize(aux["vmf_means"])
        return means, aux["vmf_kappas"][..., 0], aux["vmf_logits"][..., 0]

    def mixture_pdf(self, dirs, aux):
        means, kappas, logits = self._vars(aux)
        w = torch.softmax(logits, dim=-1)
        p = torch.sum(w[..., None, :] * eval_vmf(dirs[..., None, :], means[..., None, :, :], kappas[..., None, :]), dim=-1)
        return torch.clamp(p, min=0.0)

    def sample_directions(self, u1, u2, wo, alpha, aux):
        means, kappas, _ = self._vars(aux)
        lat = aux["latent"].long()
        mean = torch.gather(means, 1, lat[:, None, None].expand(-1, 1, 3))[:, 0]
        kappa = torch.gather(kappas, 1, lat[:, None])[:, 0]
        t_vec = ref_math.l2_normalize(torch.stack([-mean[..., 1], mean[..., 0], torch.zeros_like(mean[..., 0])], dim=-1))
        b_vec = ref_math.l2_normalize(torch.cross(mean, t_vec, dim=-1))
        v = ref_math.l2_normalize(aux["normal2"])
        tmp = aux["u"]
        w = 1.0 + (1.0 / torch.clamp(kappa[..., None], min=EPS)) * ref_math.safe_log(
            tmp + (1.0 - tmp) * torch.exp(-2.0 * kappa[..., None]))
        s = ref_math.sqrt(torch.clamp(1.0 - w**2, min=0.0))
        dirs = (t_vec[:, None, :] * (s * v[..., 0])[..., None] + b_vec[:, None, :] * (s * v[..., 1])[..., None]
                + mean[:, None, :] * w[..., None])
        return dirs, self.mixture_pdf(dirs, aux)

    def pdf(self, wo, wi, alpha, aux):
        return self.mixture_pdf(wi, aux)


def importance_sample_rays(global_viewdirs, normal, roughness, samplers, uniforms, aux=None, use_mis=True):
    """render_utils.py:722-924 without the final categorical resampling (sample counts already sum to
    num_secondary_samples).  samplers: [(sampler, count)]; uniforms: per sampler (uh, uw) [R,count]."""
    R = get_rotation_matrix(normal)
    local_viewdirs = global_to_local(global_viewdirs, R)
    num_real = sum(c for _, c in samplers)
    lightdirs, pdfs, weights = [], [], []
    for (sampler, count), (uh, uw) in zip(samplers, uniforms):
        cur_wo = local_viewdirs[:, None, :].expand(-1, count, -1)
        cur_rough = roughness[:, None, :].expand(-1, count, -1)
        cur_dirs, cur_pdf = sampler.sample_directions(uh, uw, cur_wo, cur_rough, aux)
        if sampler.global_dirs:
            cur_dirs = global_to_local(cur_dirs, R[:, None, :, :])
        cur_pdf = torch.clamp(cur_pdf, min=0.0)
        if use_mis and len(samplers) > 1:
            den = 0.0
            for sp, cp in samplers:
                if sp.global_dirs:
                    tv = local_to_global(cur_wo, R[:, None, :, :])
                    tl = local_to_global(cur_dirs, R[:, None, :, :])
                else:
                    tv, tl = cur_wo, cur_dirs
                den = den + torch.square(sp.pdf(tv, tl, cur_rough, aux) * cp)
            den = torch.clamp(den, min=DENOMINATOR_EPS)
            cur_w = torch.square(count * cur_pdf) / den * (float(num_real) / float(count))
        else:
            cur_w = torch.ones_like(cur_pdf)
        lightdirs.append(cur_dirs)
        pdfs.append(cur_pdf)
        weights.append(cur_w)
    local_lightdirs = torch.cat(lightdirs, dim=-2)
    S = local_lightdirs.shape[-2]
    return dict(
        local_lightdirs=local_lightdirs,
        local_viewdirs=local_viewdirs[:, None, :].expand(-1, S, -1),
        global_lightdirs=local_to_global(local_lightdirs, R[:, None, :, :]),
        global_viewdirs=global_viewdirs[:, None, :].expand(-1, S, -1),
        pdf=torch.cat(pdfs, dim=-1)[..., None], weight=torch.cat(weights, dim=-1)[..., None])


def get_secondary_rays(means, viewdirs, normals, roughness, samplers, uniforms, aux=None, normal_eps=1e-2,
                       refdir_eps=1e-2, far=2.0):
    """render_utils.py:927-1056 for one shaded point per ray: means/viewdirs/normals [R,3]."""
    samples = importance_sample_rays(-viewdirs, normals, roughness, samplers, uniforms, aux)
    S = samples["pdf"].shape[1]
    origins = (means + normals * normal_eps)[:, None, :].expand(-1, S, -1)
    rays = dict(origins=origins, directions=samples["global_lightdirs"], viewdirs=samples["global_lightdirs"],
                radii=torch.ones_like(origins[..., :1]), near=torch.full_like(origins[..., :1], refdir_eps),
                far=torch.full_like(origins[..., :1], far))
    return rays, samples


# ------------------------------------------------------------------------------------ material head
def microfacet_material(brdf_params, min_roughness=0.01, default_F_0=0.04):
    """material.py:1290-1322 with the property table :957-1023 under configs/ngp_yobo.gin:256-303
    (sigmoid activations; biases albedo -1, specular_albedo -1, roughness -1, metalness 0; constant
    Fresnel; diffuseness / mirrorness constant 0; reparam_roughness False)."""
    sig = torch.sigmoid
    rough = sig(brdf_params[..., 6:7] - 1.0)
    rough = rough * (1.0 - min_roughness**2) + min_roughness**2
    return dict(
        albedo=sig(brdf_params[..., 0:3] - 1.0), specular_albedo=sig(brdf_params[..., 5:6] - 1.0), roughness=rough,
        F_0=torch.full_like(brdf_params[..., 9:10], default_F_0), metalness=sig(brdf_params[..., 8:9]),
        diffuseness=torch.zeros_like(brdf_params[..., 3:4]), mirrorness=torch.zeros_like(brdf_params[..., 4:5]))


# ------------------------------------------------------------------------------------ material stage
class MaterialMLP:
    """internal/material.py:1901-1926,2073-2123 under configs/ngp_yobo.gin:315-333: material grid -> bottleneck
    Dense (linear) -> pred_brdf_layer -> microfacet material."""

    def __init__(self, warp_c=2.0, bbox_scaling=1.0):
        from . import grid_utils as og
        self.grid = og.HashEncoding(hash_map_size=524288, num_features=4, scale_supersample=1.0, max_grid_size=2048,
                                    bbox_scaling=bbox_scaling)
        self.warp_c = warp_c

    def init(self, gen, table_init_range=0.1):
        from . import geometry as ogeo
        d = 32
        return {"material_grid": self.grid.init(gen, init_range=table_init_range),
                "bottleneck_layer": {"kernel": ogeo.he_uniform(gen, d, 128), "bias": torch.zeros(128)},
                "pred_brdf_layer": {"kernel": ogeo.he_uniform(gen, 128, 10), "bias": torch.zeros(10)}}

    def predict_material(self, p, means, dense=None):
        from . import coord as ocoord, geometry as ogeo
        dense = dense or ogeo.dense
        enc = self.grid(p["material_grid"], ocoord.contract_radius(means, self.warp_c))
        raw = dense(p["pred_brdf_layer"], dense(p["bottleneck_layer"], enc))
        return microfacet_material(raw)


class EnvMapMLP:
    """models.py:801-812 / configs/nerf_ngp_yobo.gin:253-297 (directional encoding, no IDE)."""

    def __init__(self, deg_view=4, width=256, depth=4, skip=2, rgb_bias=-1.0):
        self.deg_view, self.width, self.depth, self.skip, self.rgb_bias = deg_view, width, depth, skip, rgb_bias
        self.in_dim = 3 + 6 * deg_view
        self.names = [f"layer_{i}" for i in range(depth - 1)] + ["layer_bottleneck"]

    def init(self, gen):
        from . import geometry as ogeo
        p, d = {}, self.in_dim
        for i, n in enumerate(self.names):
            p[n] = {"kernel": ogeo.he_uniform(gen, d, self.width), "bias": torch.zeros(self.width)}
            d = self.width + (self.in_dim if (i % self.skip == 0 and i > 0) else 0)
        p["output_rgba_layer"] = {"kernel": ogeo.he_uniform(gen, d, 4), "bias": torch.zeros(4)}
        p["output_ambient_rgb_layer"] = {"kernel": ogeo.he_uniform(gen, d, 3), "bias": torch.zeros(3)}
        return p

    def __call__(self, p, viewdirs, dense=None):
        from . import coord as ocoord, geometry as ogeo
        dense = dense or ogeo.dense
        enc = ocoord.pos_enc(viewdirs, 0, self.deg_view, True)
        x = enc
        for i, n in enumerate(self.names):
            x = torch.relu(dense(p[n], x))
            if i % self.skip == 0 and i > 0:
                x = torch.cat([x, enc], dim=-1)
        rgba = dense(p["output_rgba_layer"], x)
        return dict(incoming_rgb=torch.nn.functional.softplus(rgba[..., :3] + self.rgb_bias))


class MaterialModel:
    """One chunk of the material stage's render path (see neural_radiance_caching_b200/material.py)."""

    def __init__(self, cache_model, n_specular=16, n_cosine=8, n_light=8, near_min=0.05, far=2.0, normal_eps=1e-2,
                 rgb_max=10000.0):
        self.cache, self.material_mlp, self.env_map = cache_model, MaterialMLP(), EnvMapMLP()
        self.n_specular, self.n_cosine, self.n_light = n_specular, n_cosine, n_light
        self.near_min, self.far, self.normal_eps, self.rgb_max = near_min, far, normal_eps, rgb_max

    def render_chunk(self, params, means, viewdirs, normals, draws, material=None, light_aux=None, sdist_hook=None):
        from . import render_utils as oru
        R = means.shape[0]
        ns, nc, nl = self.n_specular, self.n_cosine, self.n_light
        S = ns + nc + nl
        if material is None:
            material = self.material_mlp.predict_material(params["Material"], means)
        u = draws["u"]
        rays_s, smp_s = get_secondary_rays(means, viewdirs, normals, material["roughness"], [(MicrofacetSampler(), ns)],
                                           [(u[:, :ns, 0], u[:, :ns, 1])], None, self.normal_eps, self.near_min, self.far)
        dsamplers, dunif = [(CosineSampler(), nc)], [(u[:, ns:ns + nc, 0], u[:, ns:ns + nc, 1])]
        aux = None
        if nl:
            aux = dict(light_aux, latent=draws["latent"], normal2=draws["normal2"], u=u[:, ns + nc:, 0])
            dsamplers.append((LightSampler(), nl))
            dunif.append((u[:, ns + nc:, 0], u[:, ns + nc:, 1]))
        rays_d, smp_d = get_secondary_rays(means, viewdirs, normals, material["roughness"], dsamplers, dunif, aux,
                                           self.normal_eps, self.near_min, self.far)
        rays = {k: torch.cat([rays_s[k], rays_d[k]], dim=1).reshape(R * S, -1) for k in
                ("origins", "directions", "near", "far", "radii")}
        rays["viewdirs"] = rays["directions"]
        out = self.cache(params["Cache"], rays, draws["u01"], gumbel=draws["gumbel"], is_secondary=True, resample=True)
        rgb = torch.clamp(torch.nan_to_num(out["render"]["rgb"]), min=0.0)
        acc = out["render"]["acc"]
        env = self.env_map(params["EnvMap"], rays["directions"])["incoming_rgb"]
        radiance_in = (rgb + env * (1.0 - acc)[:, None]).reshape(R, S, 3)
        occ = acc.reshape(R, S, 1)
        ones2 = torch.ones(R, ns, 2)
        spec = oru.integrate_reflect_rays("microfacet_specular", material, dict(
            smp_s, radiance_in=radiance_in[:, :ns], indirect_occ=occ[:, :ns], brdf_correction=ones2), self.rgb_max)
        diff = oru.integrate_reflect_rays("microfacet_diffuse", material, dict(
            smp_d, radiance_in=radiance_in[:, ns:], indirect_occ=occ[:, ns:], brdf_correction=torch.ones(R, S - ns, 2)),
            self.rgb_max)
        return dict(rgb=spec["radiance_out"] + diff["radiance_out"], specular=spec, diffuse=diff, material=material,
                    radiance_in=radiance_in, acc=acc.reshape(R, S), rays=rays, cache_out=out)
