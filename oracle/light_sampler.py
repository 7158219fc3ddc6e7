"""CPU restatement (test infrastructure -- never imported by the product path) of the light sampler:
internal/light_sampler.py LightMLP.get_vmfs / predict_lighting (:135-214) under configs/ngp_yobo.gin:335-352
(light_grid L=8 F=4, two 64-wide ReLU layers, output layer 128 lobes x 5, vmf_scale 20), and
render_utils.vmf_loss_fn (internal/inverse_render/render_utils.py:1493-1550) as called by
train_utils.light_sampling_loss (internal/train_utils.py:1985-2071).

`means_random` (jax.random.normal(PRNGKey(random_seed)) * vmf_scale / 2, light_sampler.py:141-143) is an INPUT, like
every other random draw at the C ABI: JAX's threefry stream cannot be reproduced without JAX.

vmf_loss_fn and linear_to_srgb are pinned to the reference's source (tests/test_reference_vectors.py, 1e-5); the LightMLP
class (flax parameter scoping) is restated only.  Unpinned against XLA's own rounding (JAX is not installable here)."""
import numpy as np
import torch

from . import material as omat, ref_math

NUM_COMPONENTS = 128
VMF_SCALE = 20.0
VMF_BIAS = dict(vmf_means=0.0, vmf_kappas=1.0, vmf_logits=1.0)    # light_sampler.py:73-77


def linear_to_srgb(linear):
    """internal/image.py:192-200."""
    eps = float(np.finfo(np.float32).eps)
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * torch.clamp(linear, min=eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


def get_vmfs(vmf_params, means_random, vmf_scale=VMF_SCALE):
    """LightMLP.get_vmfs (light_sampler.py:135-160): vmf_params [..., K, 5]."""
    return {
        "vmf_means": vmf_params[..., 0:3] * vmf_scale + VMF_BIAS["vmf_means"] + means_random,
        "vmf_kappas": torch.clamp(torch.nn.functional.softplus(vmf_params[..., 3:4] + VMF_BIAS["vmf_kappas"]), max=50.0),
        "vmf_logits": torch.clamp(vmf_params[..., 4:5] + VMF_BIAS["vmf_logits"], min=-50.0),
    }


class LightMLP:
    def __init__(self, warp_c=2.0, bbox_scaling=1.0, num_components=NUM_COMPONENTS):
        from . import grid_utils as og
        self.grid = og.HashEncoding(hash_map_size=524288, num_features=4, scale_supersample=1.0, max_grid_size=2048,
                                    bbox_scaling=bbox_scaling)
        self.warp_c = warp_c
        self.num_components = num_components

    def init(self, gen, table_init_range=0.1):
        from . import geometry as ogeo
        K = self.num_components
        return {"light_grid": self.grid.init(gen, init_range=table_init_range),
                "layers_0": {"kernel": ogeo.he_uniform(gen, 32, 64), "bias": torch.zeros(64)},
                "layers_1": {"kernel": ogeo.he_uniform(gen, 64, 64), "bias": torch.zeros(64)},
                "output_layer": {"kernel": ogeo.he_uniform(gen, 64, K * 5), "bias": torch.zeros(K * 5)}}

    def predict_lighting(self, p, means, means_random, normals=None, weights=None, dense=None):
        """light_sampler.py:162-214 (single illumination): grid feature -> run_network (shading.py:120-131, depth 2,
        no skip) -> output_layer -> get_vmfs; means are made relative to the (stop-gradient) position."""
        from . import coord as ocoord, geometry as ogeo
        dense = dense or ogeo.dense
        x = self.grid(p["light_grid"], ocoord.contract_radius(means, self.warp_c))
        x = torch.relu(dense(p["layers_0"], x))
        x = torch.relu(dense(p["layers_1"], x))
        raw = dense(p["output_layer"], x).reshape(means.shape[:-1] + (self.num_components, 5))
        vmfs = get_vmfs(raw, means_random)
        vmfs["vmf_means"] = vmfs["vmf_means"] - means.detach()[..., None, :]
        vmfs["vmf_origins"] = means.detach()[..., None, :]
        if normals is not None:
            vmfs["vmf_normals"] = normals.detach()[..., None, :]
        if weights is not None:
            vmfs["weights"] = weights.detach()[..., None, None]
        return vmfs


def vmf_loss_fn(vmf_vars, sample_normals, sample_dirs, pdf, weight, function_vals, lossmult, srgb=True):
    """render_utils.vmf_loss_fn (:1493-1550).  vmf_vars = (means [P,K,3], kappas [P,K,1], logits [P,K,1]);
    sample_normals [P,3]; sample_dirs [P,S,3]; pdf / weight / function_vals / lossmult [P,S]
    (function_vals_nocorr == function_vals at the call site, train_utils.py:2058-2067)."""
    means = ref_math.l2_normalize(vmf_vars[0], grad_eps=1e-5)
    kappas = vmf_vars[1][..., 0]
    weights = ref_math.safe_exp(vmf_vars[2][..., 0])
    likelihood = torch.sum(
        weights[..., None, :] * omat.eval_vmf(sample_dirs[..., None, :], means[..., None, :, :], kappas[..., None, :]), dim=-1)
    denominator = torch.clamp(pdf, min=1e-2)
    dotprod = (sample_dirs * sample_normals[..., None, :]).sum(dim=-1)
    w = torch.clamp(weight, 0.0, 10.0)
    w = torch.where(dotprod > 0.0, w, torch.zeros_like(w))
    fv = function_vals
    if srgb:
        fv = linear_to_srgb(torch.clamp(fv, min=1e-5))
        likelihood = linear_to_srgb(torch.clamp(likelihood, min=1e-5))
    return torch.mean((fv - likelihood) * (fv - likelihood).detach() * w * lossmult / denominator)


def light_sampling_loss(vmfs, sample_dirs, pdf, weight, radiance_in, srgb=True):
    """train_utils.light_sampling_loss (:1985-2071) for one suffix present (multiplier 2, the /2 inside the loop):
    function_vals = stop_gradient(|radiance_in|); lossmult = 1 / S per sample."""
    fv = torch.linalg.norm(radiance_in, dim=-1).detach()
    S = fv.shape[-1]
    lossmult = torch.full_like(fv, 1.0 / S)
    normals = vmfs["vmf_normals"].reshape(-1, 3)
    K = vmfs["vmf_means"].shape[-2]
    v = (vmfs["vmf_means"].reshape(-1, K, 3), vmfs["vmf_kappas"].reshape(-1, K, 1), vmfs["vmf_logits"].reshape(-1, K, 1))
    return vmf_loss_fn(v, normals, sample_dirs.detach(), pdf.detach(), weight.detach(), fv, lossmult, srgb) / 2.0 * 2.0
