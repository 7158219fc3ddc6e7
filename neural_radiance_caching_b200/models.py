"""Host-side mirror of the cache model graph: BaseNeRFModel.__call__ (internal/models.py:656-774),
maybe_resample (:193-292), apply_shader_and_integrator (:462-614) for the cache stage
(sampler -> [categorical resample] -> NeRFMLP shader -> VolumeIntegrator)."""
import torch

from . import _lib, nerf, render, sampling


class _ResampleWeightsFn(torch.autograd.Function):
    """maybe_resample's index draw and weight rescale (nrc_ray_resample).  The denominator
    num_resample * softmax(logits)[inds] + 1e-8 is a stop_gradient in the reference, so the VJP
    only routes g to the selected weights, scaled by 1/denominator."""

    @staticmethod
    def forward(ctx, weights, gumbel, bias, mult):
        R, n = weights.shape
        k = gumbel.shape[-1]
        inds = torch.empty((R, k), device=weights.device, dtype=torch.int32)
        w_new = torch.empty((R, k), device=weights.device, dtype=torch.float32)
        w2, g2 = weights.contiguous(), gumbel.contiguous()
        _lib.call("nrc_ray_resample", _lib.stream_ptr(), _lib.ptr(w2), _lib.ptr(g2), R, n, k, float(bias),
                  float(mult), _lib.ptr(inds), _lib.ptr(w_new))
        sel = torch.gather(w2, 1, inds.long())
        scale = torch.where(sel != 0, w_new / sel, torch.zeros_like(sel))
        ctx.save_for_backward(inds, scale)
        ctx.n = n
        ctx.mark_non_differentiable(inds)
        return inds, w_new

    @staticmethod
    def backward(ctx, _gi, gw):
        inds, scale = ctx.saved_tensors
        g = torch.zeros((inds.shape[0], ctx.n), device=gw.device, dtype=torch.float32)
        g.scatter_add_(1, inds.long(), gw * scale)
        return g, None, None, None


class _GatherFn(torch.autograd.Function):
    """take_along_axis of a per-sample field at the resampled indices (nrc_ray_resample_gather)."""

    @staticmethod
    def forward(ctx, field, inds):
        R, n, Cc = field.shape
        k = inds.shape[-1]
        out = torch.empty((R, k, Cc), device=field.device, dtype=torch.float32)
        f2 = field.contiguous()
        _lib.call("nrc_ray_resample_gather", _lib.stream_ptr(), _lib.ptr(f2), _lib.ptr(inds), R, n, k, Cc,
                  _lib.ptr(out))
        ctx.save_for_backward(inds)
        ctx.shape = field.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (inds,) = ctx.saved_tensors
        R, n, Cc = ctx.shape
        gf = torch.zeros((R, n, Cc), device=g.device, dtype=torch.float32)
        gf.scatter_add_(1, inds.long()[..., None].expand(-1, -1, Cc), g)
        return gf, None


class NeRFModel:
    """Cache model (`Cache` scope of the reference): Sampler + Shader + Integrator."""

    def __init__(self, bf16=False, weights_bias=0.0, num_resample=1):
        self.sampler = sampling.ProposalVolumeSampler(bf16=bf16)
        self.shader = nerf.NeRFMLP(bf16=bf16)
        self.weights_bias = weights_bias
        self.num_resample = num_resample

    def from_oracle(self, params, device):
        return {"Sampler": self.sampler.from_oracle(params["Sampler"], device),
                "Shader": self.shader.from_oracle(params["Shader"], device)}

    def weights_only(self, params, rays, u01, train=True):
        """The weights_only pass (models.py:1265ff, passes=('cache',), weights_only=True): sampler only, returns
        the accumulation acc = sum of the final level's weights [R]."""
        hist = self.sampler(params["Sampler"], rays, u01, train=train, use_raydist_fn=False, weights_only=True)
        return hist[-1]["weights"].sum(dim=-1)

    def __call__(self, params, rays, u01, gumbel=None, train=False, is_secondary=False, resample=False,
                 extras=False, sdist_override=None):
        hist = self.sampler(params["Sampler"], rays, u01, train=train, use_raydist_fn=is_secondary,
                            sdist_override=sdist_override)
        last = hist[-1]
        if resample:
            inds, w = _ResampleWeightsFn.apply(last["weights"], gumbel, self.weights_bias, 1.0)
            means = _GatherFn.apply(last["means"], inds)
            feat = _GatherFn.apply(last["feature"], inds)
            nrm = _GatherFn.apply(last["normals_to_use"], inds)
        else:
            inds, w = None, last["weights"]
            means, feat, nrm = last["means"], last["feature"], last["normals_to_use"]
        shade = self.shader(params["Shader"], rays["viewdirs"], means, feat, nrm)
        bg = 0.0 if is_secondary else 1.0
        ex = None
        if extras:
            ex = {k: shade[k] for k in ("diffuse_rgb", "specular_rgb", "ambient_rgb", "indirect_rgb", "albedo_rgb")}
            ex["normals_to_use"] = nrm
        wnf = last["weights"] if resample else w
        rendering = render.volumetric_rendering(shade["rgb"], w, wnf, last["tdist"], bg, True, extras=ex)
        return dict(sampler=hist, shader=shade, render=rendering, inds=inds,
                    shaded=dict(means=means, normals=nrm, weights=w))
