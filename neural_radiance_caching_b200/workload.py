"""Synthetic workloads of BASELINE.md section 2 and the cache-stage step built from the
kernels.  Shared by bench.py, __graft_entry__.smoke() and the tests (no oracle imports here).
"""
import numpy as np
import torch

from . import _lib, sampling

SEED = 20200823  # the reference's Config.jax_rng_seed (internal/configs.py:180)
SAMPLES_PER_RAY = (64, 64, 32)  # configs/nerf_ngp_yobo.gin:521-545


def make_rays_np(g, R, near=2.0, far=6.0, radius=4.0, radii=5e-4):
    """Config 1/2 primary rays: origins on a radius-4 sphere looking inwards with jitter;
    directions are NOT unit length (the reference's camera rays are not, render.py:143-144)."""
    o = g.normal(size=(R, 3))
    o = radius * o / np.linalg.norm(o, axis=-1, keepdims=True)
    v = -o + 0.3 * g.normal(size=(R, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    d = v * g.uniform(1.0, 1.2, size=(R, 1))
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(origins=f(o), directions=f(d), viewdirs=f(v), radii=np.full((R, 1), radii, np.float32),
                near=np.full((R, 1), near, np.float32), far=np.full((R, 1), far, np.float32))


def pack_rays(rays_np, u01_np, extra=None):
    """One [R, C] fp32 host buffer per step (a single pinned H2D copy): origins 3, directions 3,
    viewdirs 3, radii 1, near 1, far 1, u01 x3, then `extra` columns (e.g. target rgb)."""
    cols = [rays_np["origins"], rays_np["directions"], rays_np["viewdirs"], rays_np["radii"], rays_np["near"],
            rays_np["far"]] + list(u01_np)
    if extra is not None:
        cols.append(extra)
    return np.ascontiguousarray(np.concatenate(cols, axis=-1), dtype=np.float32)


def unpack_rays(buf):
    """Views (made contiguous) into the packed device buffer."""
    c = lambda a: a.contiguous()
    rays = dict(origins=c(buf[:, 0:3]), directions=c(buf[:, 3:6]), viewdirs=c(buf[:, 6:9]), radii=c(buf[:, 9:10]),
                near=c(buf[:, 10:11]), far=c(buf[:, 11:12]))
    u01 = [c(buf[:, 12 + i:13 + i]) for i in range(3)]
    extra = c(buf[:, 15:]) if buf.shape[1] > 15 else None
    return rays, u01, extra


class CacheSamplerStep:
    """Cache-stage proposal sampler, forward + backward (BASELINE config 2, density path):
    3 levels x (interval resampling -> ray cast -> contract + hash-grid encode + fused MLP ->
    alpha weights), loss on the level weights, gradients for the three density tables and
    MLP weights into contiguous arenas."""

    def __init__(self, device, table_init_range=0.1, seed=SEED, bf16=False):
        self.device = device
        self.sampler = sampling.ProposalVolumeSampler(bf16=bf16)
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.params = {}
        self.leaves = []
        for i, m in enumerate(self.sampler.mlps):
            tables, arena = m.grid.init(device, generator=gen, init_range=table_init_range)
            arena.requires_grad_(True)
            # level views are taken from the detached arena: they only carry device pointers, and
            # a view of the leaf would pin its AccumulateGrad node to the init stream (breaks
            # CUDA-graph capture of the backward pass).
            p = {"density_grid": dict(m.grid.views(arena.detach()), _arena=arena)}
            dims = [(m.in_dim, 64), (64, 64), (64, 1)] + ([(64, 3)] if m.enable_pred_normals else [])
            names = ["density_layers_0", "density_layers_1", "output_density_layer", "pred_normals_layer"]
            for name, (fi, fo) in zip(names, dims):
                lim = float(np.sqrt(6.0 / fi))  # he_uniform (geometry.py:127)
                k = torch.empty((fi, fo), device=device).uniform_(-lim, lim, generator=gen).requires_grad_(True)
                b = torch.zeros((fo,), device=device).requires_grad_(True)
                p[name] = {"kernel": k, "bias": b}
                self.leaves += [k, b]
            self.leaves.append(arena)
            self.params[f"MLP_{i}"] = p

    def num_table_params(self):
        return sum(int(p["density_grid"]["_arena"].numel()) for p in self.params.values())

    def zero_grad(self):
        for t in self.leaves:
            t.grad = None

    def forward(self, rays, u01, train=True):
        return self.sampler(self.params, rays, u01, train=train)

    def step(self, rays, u01, target):
        """One training step: forward, a scalar loss over the three levels' weights,
        backward.  Returns the loss (device scalar)."""
        self.zero_grad()
        hist = self.forward(rays, u01, train=True)
        # Charbonnier-style data term on the accumulated opacity of the final level plus an
        # L2 tie between proposal and final accumulations: touches every level's weights.
        acc = [h["weights"].sum(-1) for h in hist]
        loss = torch.sqrt((acc[2] - target) ** 2 + 1e-6).mean()
        loss = loss + 0.01 * ((acc[0] - acc[2].detach()) ** 2).mean() + 0.01 * ((acc[1] - acc[2].detach()) ** 2).mean()
        loss.backward()
        return loss.detach()
