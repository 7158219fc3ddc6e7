"""Oracle restatement of internal/geometry.py DensityMLP (TEST INFRASTRUCTURE ONLY).

Configuration follows configs/ngp_yobo.gin:137-140,206-230 and
configs/nerf_ngp_yobo.gin:38,44,428: net_depth 2, net_width 64, ReLU, no skip
(skip_layer=4 > depth), density_activation safe_exp, density_bias -1,
unscented_mip_basis 'mean' (one control point = the mean), warp contract_radius_c.
"""
import numpy as np
import torch

from . import coord, grid_utils, ref_math


def he_uniform(gen, fan_in, fan_out):
    """jax.nn.initializers.he_uniform: U(+-sqrt(6/fan_in)), kernel [in,out] (geometry.py:127)."""
    lim = np.sqrt(6.0 / fan_in)
    return torch.from_numpy(gen.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32))


def dense(p, x):
    """flax.linen.Dense: x @ kernel + bias."""
    return x @ p["kernel"] + p["bias"]


class _RoundBf16(torch.autograd.Function):
    """Round to bf16 (value) with a straight-through gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def dense_bf16(p, x):
    """The bf16-MLP variant: operands rounded to bf16, fp32 accumulate,
    fp32 bias.  Used as the reference for the tensor-core kernels' *gradients*: bf16 rounding
    flips the ReLU mask of ~5% of the hidden units' near-zero pre-activations, so gradients are
    only comparable against a reference that sees the same rounded activations."""
    return _RoundBf16.apply(x) @ _RoundBf16.apply(p["kernel"]) + p["bias"]


class DensityMLP:
    def __init__(
        self,
        grid_params,
        net_depth=2,
        net_width=64,
        density_bias=-1.0,
        warp_c=2.0,
        bbox_scaling=1.0,
        enable_pred_normals=False,
        disable_density_normals=False,
        normals_for_filter_only=False,
        bf16=False,
    ):
        self.dense = dense_bf16 if bf16 else dense
        self.grid = grid_utils.HashEncoding(bbox_scaling=bbox_scaling, scale_supersample=1.0, **grid_params)
        self.net_depth = net_depth
        self.net_width = net_width
        self.density_bias = density_bias
        self.warp_c = warp_c
        self.warp_fn = coord.make_warp(warp_c)
        self.enable_pred_normals = enable_pred_normals
        self.disable_density_normals = disable_density_normals
        self.normals_for_filter_only = normals_for_filter_only
        self.in_dim = len(self.grid.grid_sizes) * self.grid.num_features

    def init(self, gen, table_init_range=None, bias_range=0.0):
        p = {"density_grid": self.grid.init(gen, table_init_range)}
        d = self.in_dim
        for i in range(self.net_depth):
            p[f"density_layers_{i}"] = {
                "kernel": he_uniform(gen, d, self.net_width),
                "bias": torch.from_numpy(gen.uniform(-bias_range, bias_range, self.net_width).astype(np.float32)),
            }
            d = self.net_width
        p["output_density_layer"] = {
            "kernel": he_uniform(gen, d, 1),
            "bias": torch.from_numpy(gen.uniform(-bias_range, bias_range, 1).astype(np.float32)),
        }
        if self.enable_pred_normals:
            p["pred_normals_layer"] = {
                "kernel": he_uniform(gen, d, 3),
                "bias": torch.from_numpy(gen.uniform(-bias_range, bias_range, 3).astype(np.float32)),
            }
        return p

    def run_network(self, p, x):
        """internal/geometry.py:155-168."""
        for i in range(self.net_depth):
            x = torch.relu(self.dense(p[f"density_layers_{i}"], x))
        raw_density = self.dense(p["output_density_layer"], x)[..., 0]
        return raw_density, x

    def encode(self, p, means):
        """internal/geometry.py:225-275: control = means (basis 'mean'), warp, grid."""
        control = means[..., None, :]
        if self.warp_fn is not None:
            control = self.warp_fn(control)
        return self.grid(p["density_grid"], control, per_level_mean=True)

    def predict_density(self, p, means):
        """internal/geometry.py:199-316."""
        x = self.encode(p, means)
        return self.run_network(p, x)

    def convert_raw_density(self, raw_density, means):
        """internal/geometry.py:318-341."""
        density = ref_math.safe_exp(raw_density + self.density_bias)
        warped = self.warp_fn(means) if self.warp_fn is not None else means
        bbox = self.grid.bbox
        b0 = torch.tensor(bbox[0].astype(np.float32))
        b1 = torch.tensor(bbox[1].astype(np.float32))
        valid = torch.all((warped > b0) & (warped < b1), dim=-1)
        return torch.where(valid, density, torch.zeros_like(density))

    def __call__(self, p, means, viewdirs=None, origins=None, create_graph=False):
        """internal/geometry.py:381-518 (predict_density_normals), :521-584."""
        if self.disable_density_normals:
            raw_density, x = self.predict_density(p, means)
            raw_grad = None
            normals = None
        else:
            m = means if means.requires_grad else means.detach().requires_grad_(True)
            raw_density, x = self.predict_density(p, m)
            (raw_grad,) = torch.autograd.grad(
                raw_density.sum(), m, create_graph=create_graph, retain_graph=True
            )
            normals = torch.nan_to_num(-ref_math.l2_normalize(raw_grad))
        density = self.convert_raw_density(raw_density, means)
        out = dict(feature=x, density=density, raw_density=raw_density, raw_grad_density=raw_grad, normals=normals)
        if self.enable_pred_normals:
            grad_pred = self.dense(p["pred_normals_layer"], x)
            out["grad_pred"] = grad_pred
            out["normals_pred"] = torch.nan_to_num(-ref_math.l2_normalize(grad_pred))
            out["normals_to_use"] = out["normals_pred"]
        else:
            out["normals_pred"] = None
            out["normals_to_use"] = normals
        if origins is not None:
            out["ray_dists"] = torch.linalg.norm(origins[..., None, :] - means, dim=-1, keepdim=True)
        if self.normals_for_filter_only:
            out["normals"] = None
            out["normals_to_use"] = None
            out["normals_pred"] = None
        return out
