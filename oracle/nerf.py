"""Oracle restatement of the cache shader (TEST INFRASTRUCTURE ONLY):
BaseNeRFMLP.predict_appearance / _predict_appearance_passive (internal/nerf.py:561-689,940-1090),
BaseShader.predict_appearance_feature (internal/shading.py:133-220), the IDE-form
SurfaceLightFieldMLP used for `SurfaceLightField` and the shader-level `EnvMap`
(internal/surface_light_field.py:782-1069) and ref_utils.generate_ide_fn
(internal/ref_utils.py:79-192).

Configuration trace (configs/ngp_yobo.gin:143-176, configs/nerf_ngp_yobo.gin:232-251,299-343,
491-506): net_depth 0 (no trunk), bottleneck 128, use_reflections, pred roughness (softplus, bias
-1), specular tint, integrated-BRDF MLP 2x64, appearance grid L=8 F=4 T=2^19; SurfaceLightField:
shader bottleneck + IDE(deg 5) -> 4x128 with skip at layer 2, use_lights False, ambient rgb softplus
bias -1; EnvMap: IDE(deg 4) only -> 4x128 with skip; irradiance / ambient-irradiance softplus bias
-2; tint sigmoid; rgb_max 1e4; no distance prediction => incoming_acc == 1.
"""
import math as pymath

import numpy as np
import torch

from . import coord, grid_utils, ref_math
from .geometry import dense, dense_bf16, he_uniform


# ----------------------------------------------------------------------------- IDE tables
def generalized_binomial_coeff(a, k):
    """internal/ref_utils.py:79-81."""
    return np.prod(a - np.arange(k)) / pymath.factorial(k)


def assoc_legendre_coeff(l, m, k):
    """internal/ref_utils.py:84-104."""
    return ((-1) ** m * 2**l * pymath.factorial(l) / pymath.factorial(k) / pymath.factorial(l - k - m)
            * generalized_binomial_coeff(0.5 * (l + k + m - 1.0), l))


def sph_harm_coeff(l, m, k):
    """internal/ref_utils.py:107-114."""
    return np.sqrt((2.0 * l + 1.0) * pymath.factorial(l - m) / (4.0 * np.pi * pymath.factorial(l + m))
                   ) * assoc_legendre_coeff(l, m, k)


def get_ml_array(deg_view):
    """internal/ref_utils.py:117-128."""
    ml_list = []
    for i in range(deg_view):
        l = 2**i
        for m in range(l + 1):
            ml_list.append((m, l))
    return np.array(ml_list).T


def ide_tables(deg_view):
    """(ml_array [2,n], mat [(l_max+1), n] float64, sigma [n] float64) -- ref_utils.py:146-158,181."""
    if deg_view > 5:
        raise ValueError("Only deg_view of at most 5 is numerically stable.")
    ml_array = get_ml_array(deg_view)
    l_max = 2 ** (deg_view - 1)
    mat = np.zeros((l_max + 1, ml_array.shape[1]))
    for i, (m, l) in enumerate(ml_array.T):
        for k in range(l - m + 1):
            mat[k, i] = sph_harm_coeff(l, m, k)
    sigma = 0.5 * ml_array[1, :] * (ml_array[1, :] + 1)
    return ml_array, mat, sigma


def generate_ide_fn(deg_view, dtype=torch.float32):
    """internal/ref_utils.py:131-192 (complex arithmetic written out in real/imaginary parts).

    dtype=float64 gives the exact value of the same expression (with the reference's fp32-rounded
    coefficient matrix): for deg_view = 5 the fp32 evaluation carries ~1e-3 relative noise from the
    l = 16 alternating sums, so fp32 implementations can only be compared through this truth."""
    ml_array, mat, sigma = ide_tables(deg_view)
    mat_t = torch.tensor(mat.astype(np.float32)).to(dtype)
    sigma_t = torch.tensor(sigma.astype(np.float32)).to(dtype)

    def integrated_dir_enc_fn(xyz, kappa_inv):
        x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
        vmz = torch.cat([z**i for i in range(mat.shape[0])], dim=-1)
        re, im = [], []
        for m in ml_array[0, :]:
            c = torch.complex(x, y) ** int(m)
            re.append(c.real)
            im.append(c.imag)
        vr, vi = torch.cat(re, dim=-1), torch.cat(im, dim=-1)
        zpart = vmz @ mat_t
        att = torch.exp(-sigma_t * kappa_inv)
        return torch.cat([vr * zpart * att, vi * zpart * att], dim=-1)

    return integrated_dir_enc_fn


def reflect(viewdirs, normals):
    """internal/ref_utils.py:25-42."""
    return 2.0 * torch.sum(normals * viewdirs, dim=-1, keepdim=True) * normals - viewdirs


# ----------------------------------------------------------------------------- SLF (IDE form)
class SurfaceLightFieldMLP:
    """IDE-form SurfaceLightFieldMLP (surface_light_field.py:782-1069): x = [shader bottleneck]? +
    IDE(refdirs, roughness) -> 4 x Dense(128)+ReLU with the input re-concatenated after layer 2
    -> output_ambient_rgb_layer (3) -> softplus(. + ambient_rgb_bias)."""

    def __init__(self, deg_view, use_shader_bottleneck, bottleneck_width=128, width=128, depth=4, skip=2,
                 ambient_rgb_bias=-1.0, ambient_rgb_max=float("inf"), bf16=False):
        self.deg_view = deg_view
        self.use_shader_bottleneck = use_shader_bottleneck
        self.width, self.depth, self.skip = width, depth, skip
        self.ambient_rgb_bias = ambient_rgb_bias
        self.ambient_rgb_max = ambient_rgb_max
        self.ide = generate_ide_fn(deg_view)
        self.n_ide = 2 * get_ml_array(deg_view).shape[1]
        self.in_dim = self.n_ide + (bottleneck_width if use_shader_bottleneck else 0)
        self.dense = dense_bf16 if bf16 else dense

    def layer_names(self):
        return [f"layer_{i}" for i in range(self.depth - 1)] + ["layer_bottleneck"]

    def init(self, gen):
        p = {}
        d = self.in_dim
        for i, name in enumerate(self.layer_names()):
            p[name] = {"kernel": he_uniform(gen, d, self.width), "bias": torch.zeros(self.width)}
            d = self.width
            if i % self.skip == 0 and i > 0:
                d += self.in_dim
        p["output_ambient_rgb_layer"] = {"kernel": he_uniform(gen, d, 3), "bias": torch.zeros(3)}
        return p

    def __call__(self, p, refdirs, roughness, shader_bottleneck):
        x = []
        if self.use_shader_bottleneck:
            x.append(shader_bottleneck)
        x.append(self.ide(refdirs, roughness))
        x = torch.cat(x, dim=-1)
        inputs = x
        for i, name in enumerate(self.layer_names()):
            x = torch.relu(self.dense(p[name], x))
            if i % self.skip == 0 and i > 0:
                x = torch.cat([x, inputs], dim=-1)
        ambient = torch.nn.functional.softplus(self.dense(p["output_ambient_rgb_layer"], x) + self.ambient_rgb_bias)
        acc = torch.ones_like(x[..., 0])  # incoming_weights == ones (no distance prediction)
        return dict(incoming_ambient_rgb=torch.clamp(ambient, 0.0, self.ambient_rgb_max), incoming_acc=acc)


# ----------------------------------------------------------------------------- cache shader
APPEARANCE_GRID = dict(hash_map_size=524288, max_grid_size=2048, num_features=4)


class NeRFMLP:
    def __init__(self, warp_c=2.0, bbox_scaling=1.0, rgb_max=10000.0, bf16=False):
        self.grid = grid_utils.HashEncoding(bbox_scaling=bbox_scaling, scale_supersample=1.0, **APPEARANCE_GRID)
        self.warp_fn = coord.make_warp(warp_c)
        self.rgb_max = rgb_max
        self.feat_dim = 64 + len(self.grid.grid_sizes) * self.grid.num_features
        self.surface_lf = SurfaceLightFieldMLP(5, True, bf16=bf16)
        self.env_map = SurfaceLightFieldMLP(4, False, bf16=bf16)
        self.dense = dense_bf16 if bf16 else dense

    def init(self, gen, table_init_range=None):
        f = self.feat_dim
        d = lambda i, o: {"kernel": he_uniform(gen, i, o), "bias": torch.zeros(o)}
        return {
            "appearance_grid": self.grid.init(gen, table_init_range),
            "bottleneck_layer": d(f, 128),
            "roughness_layer": d(f, 1),
            "ambient_irradiance_layer": d(f, 3),
            "irradiance_layer": d(f, 3),
            "tint_layer": d(f, 3),
            "integrated_brdf_layers_0": d(129, 64),
            "integrated_brdf_layers_1": d(64, 64),
            "output_integrated_brdf_layer": d(64, 1),
            "SurfaceLightField": self.surface_lf.init(gen),
            "EnvMap": self.env_map.init(gen),
        }

    def predict_appearance_feature(self, p, density_feature, means):
        """shading.py:133-220: [density feature | appearance grid(contract(means))], net_depth 0."""
        control = self.warp_fn(means[..., None, :])
        enc = self.grid(p["appearance_grid"], control, per_level_mean=True)
        return torch.cat([density_feature, enc], dim=-1)

    def get_integrated_brdf(self, p, normals, viewdirs, bottleneck):
        """nerf.py:423-434,461-482: [bottleneck | n . (-v)] -> 2 x 64 ReLU -> sigmoid(. + log 3)."""
        dotprod = torch.sum(normals * (-viewdirs[..., None, :]), dim=-1, keepdim=True)
        x = torch.cat([bottleneck, dotprod], dim=-1)
        x = torch.relu(self.dense(p["integrated_brdf_layers_0"], x))
        x = torch.relu(self.dense(p["integrated_brdf_layers_1"], x))
        return torch.sigmoid(self.dense(p["output_integrated_brdf_layer"], x) + float(np.log(3.0)))

    def heads(self, p, feature):
        """Bottleneck (nerf.py:385-408, no noise / exposure) and roughness (:633-634) of the appearance feature."""
        sp = torch.nn.functional.softplus
        return self.dense(p["bottleneck_layer"], feature), sp(self.dense(p["roughness_layer"], feature) - 1.0)

    def predict_appearance_passive(self, p, feature, bottleneck, roughness, normals, viewdirs):
        """NeRFMLP._predict_appearance_passive (nerf.py:940-1090) with use_env_map and the IDE-form sub-networks."""
        sp = torch.nn.functional.softplus
        ambient_diffuse = torch.clamp(sp(self.dense(p["ambient_irradiance_layer"], feature) - 2.0), 0.0, self.rgb_max)
        tint = torch.sigmoid(self.dense(p["tint_layer"], feature))  # :975
        F = self.get_integrated_brdf(p, normals, viewdirs, bottleneck)
        refdirs = reflect(-viewdirs[..., None, :], normals)  # :1344-1358
        env = self.env_map(p["EnvMap"], refdirs, roughness, None)  # :984-996
        env_rgb = env["incoming_ambient_rgb"]
        indirect_diffuse = torch.clamp(sp(self.dense(p["irradiance_layer"], feature) - 2.0), 0.0, self.rgb_max)
        inc = self.surface_lf(p["SurfaceLightField"], refdirs, roughness, bottleneck)  # :1020-1030
        ref_rgb = inc["incoming_ambient_rgb"]
        ref_acc = inc["incoming_acc"][..., None]
        ambient_specular = torch.clamp(tint * F * (env_rgb * (1.0 - ref_acc)), 0.0, self.rgb_max)  # :1034-1037
        indirect_specular = torch.clamp(tint * F * (ref_rgb * ref_acc), 0.0, self.rgb_max)  # :1039-1042
        ambient = ambient_diffuse + ambient_specular
        indirect = indirect_diffuse + indirect_specular
        return dict(
            rgb=ambient + indirect, diffuse_rgb=ambient_diffuse + indirect_diffuse,
            specular_rgb=ambient_specular + indirect_specular, ambient_rgb=ambient, indirect_rgb=indirect,
            albedo_rgb=tint, integrated_brdf=F, env_rgb=env_rgb, ref_rgb=ref_rgb, refdirs=refdirs,
        )

    def __call__(self, p, viewdirs, means, density_feature, normals):
        """viewdirs [R,3]; means [R,n,3]; density_feature [R,n,64]; normals [R,n,3] (normals_to_use)."""
        feature = self.predict_appearance_feature(p, density_feature, means)
        bottleneck, roughness = self.heads(p, feature)
        out = self.predict_appearance_passive(p, feature, bottleneck, roughness, normals, viewdirs)
        out.update(roughness=roughness, bottleneck=bottleneck, feature=feature)
        return out
