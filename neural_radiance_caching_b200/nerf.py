"""Host-side mirror of the cache shader: internal/nerf.py NeRFMLP (passive path),
internal/shading.py BaseShader.predict_appearance_feature, the IDE-form
internal/surface_light_field.py SurfaceLightFieldMLP (`SurfaceLightField`, shader-level `EnvMap`)
and internal/ref_utils.py generate_ide_fn / reflect.  Dense layers and the IDE run in the CUDA
library (nrc_dense_*, nrc_ide_*); the appearance grid uses the encode kernels.  The remaining
per-point scalar glue (softplus / sigmoid / clip / products of [P,1..3] tensors) is torch
elementwise code for now.
"""
import ctypes as C
import math as pymath

import numpy as np
import torch

from . import _lib, coord, grid_utils


# ----------------------------------------------------------------------------- Dense
def _pad4(t):
    """Row stride multiple of 4 floats (16-byte rows) so the GEMM tile loader can use LDG.128."""
    k = t.shape[-1]
    if k % 4 == 0:
        return t.contiguous(), k
    k4 = (k + 3) // 4 * 4
    return torch.nn.functional.pad(t, (0, k4 - k)).contiguous(), k4


class _DenseFn(torch.autograd.Function):
    """custom_vjp analogue of flax.linen.Dense (+ optional ReLU) over nrc_dense_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, x, kernel, bias, relu, bf16):
        K, N = kernel.shape
        x2, ldx = _pad4(x.reshape(-1, K))
        M = x2.shape[0]
        y = torch.empty((M, N), device=x.device, dtype=torch.float32)
        _lib.call("nrc_dense_fwd", _lib.stream_ptr(), _lib.ptr(x2), ldx, _lib.ptr(kernel), _lib.ptr(bias), M, K, N,
                  int(relu), int(bf16), _lib.ptr(y), N)
        ctx.save_for_backward(x2, kernel, y if relu else None)
        ctx.bias_ref = bias
        ctx.meta = (x.shape, relu, bf16, ldx)
        return y.reshape(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, g):
        x2, kernel, y = ctx.saved_tensors
        xshape, relu, bf16, ldx = ctx.meta
        K, N = kernel.shape
        M = x2.shape[0]
        g2 = g.reshape(M, N).contiguous()
        if relu:
            gpre = torch.empty_like(g2)
            _lib.call("nrc_relu_bwd", _lib.stream_ptr(), _lib.ptr(y), N, _lib.ptr(g2), N, M, N, _lib.ptr(gpre), N)
            g2 = gpre
        g2, ldg = _pad4(g2)
        gx = torch.empty((M, K), device=g.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        want_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gk = gb = None
        sunk = False
        if want_w:
            gk, gb = _lib.grad_sink(kernel), _lib.grad_sink(ctx.bias_ref)
            sunk = gk is not None and gb is not None
            if not sunk:
                gk = torch.zeros_like(kernel)
                gb = torch.zeros((N,), device=g.device, dtype=torch.float32)
        _lib.call("nrc_dense_bwd", _lib.stream_ptr(), _lib.ptr(x2), ldx, _lib.ptr(kernel), _lib.ptr(g2), ldg, M, K, N,
                  int(bf16), _lib.ptr(gx), K, 0, _lib.ptr(gk), _lib.ptr(gb))
        if sunk:
            gk = gb = None   # accumulated into the registered sinks
        return (gx.reshape(xshape) if gx is not None else None), gk, gb, None, None


def dense(p, x, relu=False, bf16=False):
    """flax.linen.Dense with params {'kernel': [in,out], 'bias': [out]}."""
    return _DenseFn.apply(x, p["kernel"], p["bias"], relu, bf16)


# ----------------------------------------------------------------------------- IDE
def _generalized_binomial_coeff(a, k):
    return np.prod(a - np.arange(k)) / pymath.factorial(k)


def _assoc_legendre_coeff(l, m, k):
    return ((-1) ** m * 2**l * pymath.factorial(l) / pymath.factorial(k) / pymath.factorial(l - k - m)
            * _generalized_binomial_coeff(0.5 * (l + k + m - 1.0), l))


def _sph_harm_coeff(l, m, k):
    return np.sqrt((2.0 * l + 1.0) * pymath.factorial(l - m) / (4.0 * np.pi * pymath.factorial(l + m))
                   ) * _assoc_legendre_coeff(l, m, k)


class _IdeTables:
    """Host tables of ref_utils.generate_ide_fn (internal/ref_utils.py:117-158,181)."""

    _cache = {}

    def __init__(self, deg_view):
        if deg_view > 5:
            raise ValueError("Only deg_view of at most 5 is numerically stable.")
        ml = []
        for i in range(deg_view):
            l = 2**i
            for m in range(l + 1):
                ml.append((m, l))
        self.n_sh = len(ml)
        l_max = 2 ** (deg_view - 1)
        mat = np.zeros((l_max + 1, self.n_sh))
        for i, (m, l) in enumerate(ml):
            for k in range(l - m + 1):
                mat[k, i] = _sph_harm_coeff(l, m, k)
        self.mat_host = mat.astype(np.float32)
        self.m = (C.c_int32 * self.n_sh)(*[m for m, _ in ml])
        self.l = (C.c_int32 * self.n_sh)(*[l for _, l in ml])
        self.sigma = (C.c_float * self.n_sh)(*[0.5 * l * (l + 1) for _, l in ml])
        self._mat_dev = {}

    def mat(self, device):
        key = str(device)
        if key not in self._mat_dev:
            self._mat_dev[key] = torch.from_numpy(self.mat_host).to(device).contiguous()
        return self._mat_dev[key]

    @classmethod
    def get(cls, deg_view):
        if deg_view not in cls._cache:
            cls._cache[deg_view] = cls(deg_view)
        return cls._cache[deg_view]


class _IdeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, kappa_inv, deg_view):
        t = _IdeTables.get(deg_view)
        x2 = xyz.reshape(-1, 3).contiguous()
        k2 = kappa_inv.reshape(-1).contiguous()
        P = x2.shape[0]
        out = torch.empty((P, 2 * t.n_sh), device=xyz.device, dtype=torch.float32)
        _lib.call("nrc_ide_fwd", _lib.stream_ptr(), t.n_sh, t.m, t.l, t.sigma, _lib.ptr(t.mat(xyz.device)),
                  _lib.ptr(x2), _lib.ptr(k2), P, _lib.ptr(out), 2 * t.n_sh)
        ctx.save_for_backward(x2, k2)
        ctx.meta = (deg_view, xyz.shape, kappa_inv.shape)
        return out.reshape(xyz.shape[:-1] + (2 * t.n_sh,))

    @staticmethod
    def backward(ctx, g):
        x2, k2 = ctx.saved_tensors
        deg_view, xshape, kshape = ctx.meta
        t = _IdeTables.get(deg_view)
        P = x2.shape[0]
        g2 = g.reshape(P, 2 * t.n_sh).contiguous()
        gx = torch.empty_like(x2)
        gk = torch.empty_like(k2)
        _lib.call("nrc_ide_bwd", _lib.stream_ptr(), t.n_sh, t.m, t.l, t.sigma, _lib.ptr(t.mat(g.device)),
                  _lib.ptr(x2), _lib.ptr(k2), _lib.ptr(g2), 2 * t.n_sh, P, _lib.ptr(gx), _lib.ptr(gk))
        return gx.reshape(xshape), gk.reshape(kshape), None


def generate_ide_fn(deg_view):
    """ref_utils.generate_ide_fn (internal/ref_utils.py:131-192): returns f(xyz, kappa_inv)."""
    _IdeTables.get(deg_view)
    return lambda xyz, kappa_inv: _IdeFn.apply(xyz, kappa_inv, deg_view)


def reflect(viewdirs, normals):
    """ref_utils.reflect (internal/ref_utils.py:25-42)."""
    return 2.0 * torch.sum(normals * viewdirs, dim=-1, keepdim=True) * normals - viewdirs


# ----------------------------------------------------------------------------- SLF (IDE form)
class SurfaceLightFieldMLP:
    """IDE-form SurfaceLightFieldMLP (internal/surface_light_field.py:782-1069) as configured by
    configs/nerf_ngp_yobo.gin:232-251 (`SurfaceLightField`) and :299-343 (shader `EnvMap`)."""

    def __init__(self, deg_view, use_shader_bottleneck, bottleneck_width=128, net_width_viewdirs=128,
                 net_depth_viewdirs=4, skip_layer_dir=2, ambient_rgb_bias=-1.0, ambient_rgb_max=float("inf"),
                 bf16=False):
        self.deg_view = deg_view
        self.use_shader_bottleneck = use_shader_bottleneck
        self.width, self.depth, self.skip = net_width_viewdirs, net_depth_viewdirs, skip_layer_dir
        self.ambient_rgb_bias, self.ambient_rgb_max = ambient_rgb_bias, ambient_rgb_max
        self.dir_enc_fn = generate_ide_fn(deg_view)
        self.bf16 = bf16

    def layer_names(self):
        return [f"layer_{i}" for i in range(self.depth - 1)] + ["layer_bottleneck"]

    def __call__(self, p, refdirs, roughness, shader_bottleneck):
        x = []
        if self.use_shader_bottleneck:
            x.append(shader_bottleneck)
        x.append(self.dir_enc_fn(refdirs, roughness))
        x = torch.cat(x, dim=-1) if len(x) > 1 else x[0]
        inputs = x
        for i, name in enumerate(self.layer_names()):  # run_surface_lightfield_network :480-500
            x = dense(p[name], x, relu=True, bf16=self.bf16)
            if i % self.skip == 0 and i > 0:
                x = torch.cat([x, inputs], dim=-1)
        ambient = torch.nn.functional.softplus(
            dense(p["output_ambient_rgb_layer"], x, bf16=self.bf16) + self.ambient_rgb_bias)
        acc = torch.ones_like(x[..., 0])
        return dict(incoming_ambient_rgb=torch.clamp(ambient, 0.0, self.ambient_rgb_max), incoming_acc=acc)


# ----------------------------------------------------------------------------- cache shader
APPEARANCE_GRID = dict(hash_map_size=524288, max_grid_size=2048, num_features=4)


class NeRFMLP:
    """Cache shader (internal/nerf.py:561-689,940-1090) under configs/ngp_yobo.gin:143-176 and
    configs/nerf_ngp_yobo.gin:491-506."""

    def __init__(self, warp_c=2.0, bbox_scaling=1.0, rgb_max=10000.0, bf16=False):
        self.grid = grid_utils.HashEncoding(bbox_scaling=bbox_scaling, scale_supersample=1.0, **APPEARANCE_GRID)
        self.warp_c = warp_c if warp_c is not None else 0.0
        self.rgb_max = rgb_max
        self.bf16 = bf16
        self.surface_lf = SurfaceLightFieldMLP(5, True, bf16=bf16)
        self.env_map = SurfaceLightFieldMLP(4, False, bf16=bf16)

    def from_oracle(self, p, device):
        def mv(t):
            if isinstance(t, dict):
                return {k: mv(v) for k, v in t.items()}
            return t.detach().to(device).contiguous()

        out = {k: mv(v) for k, v in p.items() if k != "appearance_grid"}
        names = [n for (n, _, _, _) in self.grid.level_layout]
        arena = torch.cat([p["appearance_grid"][n].detach().reshape(-1) for n in names]).to(device)
        out["appearance_grid"] = dict(self.grid.views(arena), _arena=arena)
        return out

    def predict_appearance_feature(self, p, density_feature, means):
        """shading.py:133-220."""
        z = coord._ContractFn.apply(means, self.warp_c)
        enc = self.grid(p["appearance_grid"], z)
        return torch.cat([density_feature, enc], dim=-1)

    def __call__(self, p, viewdirs, means, density_feature, normals):
        sp = torch.nn.functional.softplus
        b = self.bf16
        feature = self.predict_appearance_feature(p, density_feature, means)
        bottleneck = dense(p["bottleneck_layer"], feature, bf16=b)
        roughness = sp(dense(p["roughness_layer"], feature, bf16=b) - 1.0)
        ambient_diffuse = torch.clamp(sp(dense(p["ambient_irradiance_layer"], feature, bf16=b) - 2.0), 0.0, self.rgb_max)
        tint = torch.sigmoid(dense(p["tint_layer"], feature, bf16=b))
        dotprod = torch.sum(normals * (-viewdirs[..., None, :]), dim=-1, keepdim=True)
        x = torch.cat([bottleneck, dotprod], dim=-1)
        x = dense(p["integrated_brdf_layers_0"], x, relu=True, bf16=b)
        x = dense(p["integrated_brdf_layers_1"], x, relu=True, bf16=b)
        F = torch.sigmoid(dense(p["output_integrated_brdf_layer"], x, bf16=b) + float(np.log(3.0)))
        refdirs = reflect(-viewdirs[..., None, :], normals)
        env_rgb = self.env_map(p["EnvMap"], refdirs, roughness, None)["incoming_ambient_rgb"]
        indirect_diffuse = torch.clamp(sp(dense(p["irradiance_layer"], feature, bf16=b) - 2.0), 0.0, self.rgb_max)
        inc = self.surface_lf(p["SurfaceLightField"], refdirs, roughness, bottleneck)
        ref_rgb = inc["incoming_ambient_rgb"]
        ref_acc = inc["incoming_acc"][..., None]
        ambient_specular = torch.clamp(tint * F * (env_rgb * (1.0 - ref_acc)), 0.0, self.rgb_max)
        indirect_specular = torch.clamp(tint * F * (ref_rgb * ref_acc), 0.0, self.rgb_max)
        ambient = ambient_diffuse + ambient_specular
        indirect = indirect_diffuse + indirect_specular
        return dict(
            rgb=ambient + indirect, diffuse_rgb=ambient_diffuse + indirect_diffuse,
            specular_rgb=ambient_specular + indirect_specular, ambient_rgb=ambient, indirect_rgb=indirect,
            albedo_rgb=tint, roughness=roughness, integrated_brdf=F, env_rgb=env_rgb, ref_rgb=ref_rgb,
            bottleneck=bottleneck, feature=feature, refdirs=refdirs,
        )
