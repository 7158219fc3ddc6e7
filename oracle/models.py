"""Oracle restatement of the cache model graph (TEST INFRASTRUCTURE ONLY):
BaseNeRFModel.__call__ (internal/models.py:656-774), maybe_resample (:193-292),
apply_shader_and_integrator (:462-614) reduced to sampler -> [resample] -> NeRFMLP shader ->
VolumeIntegrator (internal/integration.py:112-289, bg_intensity_range (1,1) for primary rays,
(0,0) for secondary rays, models.py:183-191)."""
import torch

from . import nerf, ref_math, render, sampling


def maybe_resample(weights, gumbel, num_resample, weights_bias=0.0, logits_mult=1.0):
    """models.py:193-292 with resample_argmax=False: returns (inds [R,k], new weights [R,k])."""
    logits = ref_math.safe_log(weights + weights_bias) * logits_mult
    probs = torch.softmax(logits, dim=-1)
    # jax.random.categorical(key, logits[..., None], axis=-2, shape=[..., k]) = argmax(logits + gumbel)
    inds = torch.argmax(logits[..., None] + gumbel, dim=-2)
    fp = torch.gather(probs, -1, inds)
    w = torch.gather(weights, -1, inds) / (num_resample * fp + 1e-8).detach()
    return inds, w


class NeRFModel:
    def __init__(self, bf16=False, weights_bias=0.0, num_resample=1):
        self.sampler = sampling.ProposalVolumeSampler()
        if bf16:
            for m in self.sampler.mlps:
                m.dense = __import__("oracle.geometry", fromlist=["dense_bf16"]).dense_bf16
        self.shader = nerf.NeRFMLP(bf16=bf16)
        self.weights_bias = weights_bias
        self.num_resample = num_resample

    def init(self, gen, table_init_range=None, bias_range=0.0):
        return {"Sampler": self.sampler.init(gen, table_init_range, bias_range),
                "Shader": self.shader.init(gen, table_init_range)}

    def weights_only(self, params, rays, u01):
        """weights_only pass (internal/models.py:1265ff): sampler only -> acc [R]."""
        hist = self.sampler(params["Sampler"], rays, u01, use_raydist_fn=False, weights_only=True)
        return hist[-1]["weights"].sum(dim=-1)

    def __call__(self, params, rays, u01, gumbel=None, is_secondary=False, resample=False, extras=False,
                 create_graph=False):
        hist = self.sampler(params["Sampler"], rays, u01, use_raydist_fn=is_secondary, create_graph=create_graph)
        last = hist[-1]
        take = lambda x, inds: torch.gather(x, 1, inds[..., None].expand(inds.shape + (x.shape[-1],)))
        if resample:
            inds, w = maybe_resample(last["weights"], gumbel, self.num_resample, self.weights_bias)
            means, feat, nrm = take(last["means"], inds), take(last["feature"], inds), take(last["normals_to_use"], inds)
        else:
            inds, w = None, last["weights"]
            means, feat, nrm = last["means"], last["feature"], last["normals_to_use"]
        shade = self.shader(params["Shader"], rays["viewdirs"], means, feat, nrm)
        bg = 0.0 if is_secondary else 1.0
        ex = None
        if extras:
            ex = {k: shade[k] for k in ("diffuse_rgb", "specular_rgb", "ambient_rgb", "indirect_rgb", "albedo_rgb")}
            ex["normals_to_use"] = nrm
        rendering = render.volumetric_rendering(shade["rgb"], w, last["weights"], last["tdist"],
                                                torch.full(w.shape[:-1] + (3,), bg), True, extras=ex)
        return dict(sampler=hist, shader=shade, render=rendering, inds=inds)
