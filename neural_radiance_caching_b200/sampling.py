"""Host-side mirror of internal/sampling.py ProposalVolumeSampler (CUDA bodies).

Per level the path is four launches with nothing else touching HBM:
  nrc_ray_sample_intervals  (anneal + safe_log + softmax + CDF + inverse + sort)
  nrc_ray_cast              (s->t warp + Gaussian means)
  nrc_density_query_fwd     (contract + hash-grid encode + fused MLP + activation [+ normals])
  nrc_ray_alpha_weights_fwd (alpha compositing weights)
The training path (`train=True`) runs the same kernels through their custom VJPs.
"""
import numpy as np
import torch

from . import _lib, coord, geometry, render, stepfun

GRID_PARAMS = (
    dict(hash_map_size=524288, max_grid_size=512, num_features=1),
    dict(hash_map_size=524288, max_grid_size=1024, num_features=1),
    dict(hash_map_size=524288, max_grid_size=2048, num_features=4),
)
MLP_PARAMS = (
    dict(disable_density_normals=False, enable_pred_normals=False, normals_for_filter_only=True),
    dict(disable_density_normals=False, enable_pred_normals=False, normals_for_filter_only=True),
    dict(disable_density_normals=False, enable_pred_normals=True, normals_for_filter_only=False),
)


class _SafeExpFn(torch.autograd.Function):
    """math.safe_exp (internal/math.py:186-192): exp(clip(x, ., 70)), grad = y * g."""

    @staticmethod
    def forward(ctx, x):
        y = torch.exp(torch.clamp(x, max=70.0))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g * y


class _NormalsFn(torch.autograd.Function):
    """nan_to_num(-ref_utils.l2_normalize(x)) (internal/ref_utils.py:45-70, geometry.py:442-479) with the
    reference's gradient override (forward clamps |x|^2 at tiny, backward at eps): nrc_normals_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, x):
        x2 = x.reshape(-1, 3).contiguous()
        n = torch.empty_like(x2)
        _lib.call("nrc_normals_fwd", _lib.stream_ptr(), _lib.ptr(x2), x2.shape[0], _lib.ptr(n))
        ctx.save_for_backward(x2)
        return n.reshape(x.shape)

    @staticmethod
    def backward(ctx, g):
        (x2,) = ctx.saved_tensors
        g2 = g.reshape(-1, 3).contiguous()
        out = torch.empty_like(x2)
        _lib.call("nrc_normals_bwd", _lib.stream_ptr(), _lib.ptr(x2), _lib.ptr(g2), x2.shape[0], _lib.ptr(out))
        return out.reshape(g.shape)


class ProposalVolumeSampler:
    """internal/sampling.py:44-649 under configs/ngp_yobo.gin:178-242 and
    configs/nerf_ngp_yobo.gin:521-562 (see oracle/sampling.py for the config trace)."""

    def __init__(self, sampling_strategy=((0, 0, 64), (1, 1, 64), (2, 2, 32)), grid_params_per_level=GRID_PARAMS,
                 mlp_params_per_level=MLP_PARAMS, anneal_slope=10.0, anneal_end=1.0, anneal_clip=0.4,
                 resample_padding=1e-5, single_jitter=True, warp_c=2.0, bbox_scaling=1.0, raydist=(-1.5, 2.0),
                 opaque_background=False, bf16=False):
        if not single_jitter:
            raise NotImplementedError("single_jitter=False is outside the configs' scope")
        self.sampling_strategy = sampling_strategy
        self.mlps = [
            geometry.DensityMLP(grid_params=g, warp_c=warp_c, bbox_scaling=bbox_scaling, bf16=bf16, **m)
            for g, m in zip(grid_params_per_level, mlp_params_per_level)
        ]
        self.anneal_slope, self.anneal_end, self.anneal_clip = anneal_slope, anneal_end, anneal_clip
        self.resample_padding = resample_padding
        self.raydist = raydist
        self.opaque_background = opaque_background

    def from_oracle(self, params, device):
        return {f"MLP_{i}": m.from_oracle(params[f"MLP_{i}"], device) for i, m in enumerate(self.mlps)}

    def anneal(self, train_frac):
        """internal/sampling.py:326-336."""
        if self.anneal_slope > 0:
            bias = lambda x, s: (s * x) / ((s - 1) * x + 1)
            return float(np.clip(bias(train_frac / self.anneal_end, self.anneal_slope), 0.0, self.anneal_clip))
        return self.anneal_clip

    def _cast(self, sdist, rays, use_raydist_fn):
        R = sdist.shape[0]
        n = sdist.shape[-1] - 1
        tdist = torch.empty_like(sdist)
        means = torch.empty((R, n, 3), device=sdist.device, dtype=torch.float32)
        p, premult = self.raydist
        _lib.call("nrc_ray_cast", _lib.stream_ptr(), _lib.ptr(sdist), _lib.ptr(rays["origins"]),
                  _lib.ptr(rays["directions"]), _lib.ptr(rays["near"]), _lib.ptr(rays["far"]), R, n,
                  1 if use_raydist_fn else 0, float(p), float(premult), _lib.ptr(tdist), _lib.ptr(means))
        return tdist, means

    def sample_and_cast(self, u01, sdist, weights, num_samples, anneal, rays, use_raydist_fn, prev=None):
        """One level of the sampler's resampling (sampling.py:340-349: annealed logits -> stepfun.sample_intervals)
        followed by render.cast_rays in ONE launch (nrc_ray_sample_cast): bit-identical to
        stepfun.sample_intervals_from_weights(...) + self._cast(...).  Returns (sdist_new, tdist, means).
        prev = (density [R,m], tdist [R,m+1]) of the level being resampled: its alpha-compositing weights are computed
        in the same launch (nrc_ray_weights_sample_cast) and written INTO `weights` (an output then)."""
        if num_samples <= 1:
            raise ValueError(f"num_samples must be > 1, is {num_samples}.")
        m = weights.shape[-1]
        if sdist.shape[-1] != m + 1:
            raise ValueError(f"Invalid shapes ({sdist.shape}, {weights.shape}) for a step function.")
        R = sdist.shape[0]
        u2 = u01.reshape(-1)
        if u2.shape[0] != R:
            raise ValueError("single_jitter=True needs one uniform per ray")
        base, max_jitter = stepfun.u_base(num_samples, sdist.device)
        dev = sdist.device
        sd = torch.empty((R, num_samples + 1), device=dev, dtype=torch.float32)
        tdist = torch.empty_like(sd)
        means = torch.empty((R, num_samples, 3), device=dev, dtype=torch.float32)
        p, premult = self.raydist
        if prev is not None:
            density, tdist_prev = prev
            if not weights.is_contiguous():
                raise ValueError("weights must be a contiguous output buffer")
            _lib.call("nrc_ray_weights_sample_cast", _lib.stream_ptr(), _lib.ptr(sdist.contiguous()), _lib.ptr(density),
                      _lib.ptr(tdist_prev), int(self.opaque_background), _lib.ptr(weights), _lib.ptr(u2.contiguous()),
                      _lib.ptr(base), R, m, num_samples, float(anneal), float(self.resample_padding), max_jitter, 0.0, 1.0,
                      _lib.ptr(rays["origins"]), _lib.ptr(rays["directions"]), _lib.ptr(rays["near"]), _lib.ptr(rays["far"]),
                      1 if use_raydist_fn else 0, float(p), float(premult), _lib.ptr(sd), _lib.ptr(tdist), _lib.ptr(means))
            return sd, tdist, means
        _lib.call("nrc_ray_sample_cast", _lib.stream_ptr(), _lib.ptr(sdist.contiguous()), _lib.ptr(weights.contiguous()),
                  _lib.ptr(u2.contiguous()), _lib.ptr(base), R, m, num_samples, float(anneal),
                  float(self.resample_padding), max_jitter, 0.0, 1.0, _lib.ptr(rays["origins"]), _lib.ptr(rays["directions"]),
                  _lib.ptr(rays["near"]), _lib.ptr(rays["far"]), 1 if use_raydist_fn else 0, float(p), float(premult),
                  _lib.ptr(sd), _lib.ptr(tdist), _lib.ptr(means))
        return sd, tdist, means

    def __call__(self, params, rays, u01_per_level, train_frac=1.0, train=False, use_raydist_fn=False,
                 normals_all_levels=False, sdist_override=None, weights_only=False, return_covs=False):
        """rays: dict of contiguous CUDA tensors origins/directions/viewdirs [R,3], near/far [R,1].

        `sdist_override` (list of per-level [R,n+1] tensors, test aid): use these fenceposts
        instead of the resampled ones, so that each level can be checked against the oracle on
        bit-identical sample positions; the resampled fenceposts are still computed and returned
        as `sdist_sampled`.  `return_covs`: also emit the reference's `covs` entry (render.cast_rays(..., diag=False),
        sampling.py:361-368) per level from nrc_ray_cast_covs; off by default because no consumer on this path reads it
        (unscented basis 'mean')."""
        near = rays["near"]
        R = near.shape[0]
        dev = near.device
        sdist = torch.cat([torch.zeros_like(near), torch.ones_like(near)], dim=-1)
        weights = torch.ones_like(near)
        anneal = self.anneal(train_frac)
        history = []
        for i_level, (i_mlp, _, num_samples) in enumerate(self.sampling_strategy):
            mlp = self.mlps[i_mlp]
            p = params[f"MLP_{i_mlp}"]
            with torch.no_grad():  # stop_level_grad (sampling.py:353-354)
                sdist = stepfun.sample_intervals_from_weights(
                    u01_per_level[i_level], sdist, weights.detach(), num_samples, anneal=anneal,
                    padding=self.resample_padding, domain=(0.0, 1.0))
                sdist_sampled = sdist
                if sdist_override is not None:
                    sdist = sdist_override[i_level].contiguous()
                tdist, means = self._cast(sdist, rays, use_raydist_fn)
            want_normals = (normals_all_levels or not mlp.normals_for_filter_only) and not mlp.disable_density_normals
            want_normals = want_normals and not weights_only   # weights_only pass (models.py:1265ff): densities only
            res = {}
            if not train:
                with torch.no_grad():
                    q = mlp.query(p, means, want_feat=True, want_normals=want_normals)
                density = q["density"]
                res.update(feature=q["feature"], density=density, raw_density=q["raw_density"],
                           raw_grad_density=q["raw_grad_density"], grad_pred=q["grad_pred"])
            else:
                last = i_level == len(self.sampling_strategy) - 1
                density, feat, gp = mlp.query_train(p, means, want_feat=(last or normals_all_levels) and not weights_only)
                res.update(feature=feat, density=density, raw_density=None, grad_pred=gp, raw_grad_density=None)
                if want_normals:   # analytic raw gradient with its parameter VJP (second-order path, 8f-1)
                    res["raw_grad_density"] = mlp.raw_grad_density(p, means)
            if res.get("raw_grad_density") is not None and want_normals:
                res["normals"] = _NormalsFn.apply(res["raw_grad_density"])
            else:
                res["normals"] = None
            if mlp.enable_pred_normals and not weights_only:
                res["normals_pred"] = _NormalsFn.apply(res["grad_pred"])
                res["normals_to_use"] = res["normals_pred"]
            else:
                res["normals_pred"] = None
                res["normals_to_use"] = res["normals"]
            if mlp.normals_for_filter_only and not normals_all_levels:
                res["normals"] = res["normals_to_use"] = res["normals_pred"] = None
            for k in list(res.keys()):
                if k.startswith("normals") and res[k] is not None:
                    pdot = torch.sum(res[k] * rays["viewdirs"][..., None, :], dim=-1, keepdim=True)
                    res[k + "_rectified"] = res[k] * torch.where(pdot > 0, -1.0, 1.0)
            weights, alphas, trans = render.compute_alpha_weights(
                density, tdist, rays["directions"], opaque_background=self.opaque_background)
            res.update(points=means, means=means, tdist=tdist, sdist=sdist, sdist_sampled=sdist_sampled,
                       weights=weights, alphas=alphas, trans=trans)
            if return_covs:
                with torch.no_grad():
                    res["covs"] = render.cast_rays(tdist, rays["origins"], rays["directions"], rays["radii"], "cone",
                                                   diag=False)[1]
            history.append(res)
        return history
