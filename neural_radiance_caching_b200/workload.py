"""Synthetic workloads of BASELINE.md section 2 and the cache-stage training step built from the
kernels.  Shared by bench.py, __graft_entry__.smoke() and the tests (no oracle imports here).
"""
import os

import numpy as np
import torch

from . import _lib, mlp_chain, models, nerf

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


SHADOW_NEAR_MAX = 0.2        # Config.shadow_near_max (internal/configs.py:635)
SECONDARY_NORMAL_EPS = 1e-2  # Config.secondary_normal_eps (internal/configs.py:643)
SECONDARY_FAR = 2.0          # Config.secondary_far (configs/nerf_ngp_yobo.gin:19)


def backward_mask_rays_np(g, rays_np):
    """Extra rays of the backward-mask loss (_compute_backward_mask_loss, internal/train_utils.py:3348-3401): one
    uniform-hemisphere direction (UniformHemisphereSampler, render_utils.py:387-414) around the normal `-look` at
    `origins + look * shadow_near_max`, origin lifted by normal_eps along the normal, near = shadow_near_max,
    far = secondary_far, radii = 1.  The synthetic batches have no camera, so the per-ray view direction stands
    in for `rays.look`.  They depend on the batch and on random draws only (no parameters): like `u01`, they are
    generated with the batch and travel in the same packed host buffer."""
    R = rays_np["origins"].shape[0]
    look = rays_np["viewdirs"].astype(np.float64)
    normal = -look
    means = rays_np["origins"].astype(np.float64) + look * SHADOW_NEAR_MAX
    u1, u2 = g.uniform(size=(R, 1)), g.uniform(size=(R, 1))
    phi = u2 * 2.0 * np.pi - np.pi
    sin_t = np.sqrt((2.0 - u1) * u1)
    local = np.concatenate([sin_t * np.cos(phi), sin_t * np.sin(phi), 1.0 - u1], axis=-1)
    # frame with new_z = normal (render_utils.py:145-168)
    up = np.where(np.abs(normal[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[0.0, 1.0, 0.0]]))
    new_x = np.cross(up, normal)
    new_x /= np.linalg.norm(new_x, axis=-1, keepdims=True) + 1e-10
    new_y = np.cross(normal, new_x)
    new_y /= np.linalg.norm(new_y, axis=-1, keepdims=True) + 1e-10
    d = local[:, 0:1] * new_x + local[:, 1:2] * new_y + local[:, 2:3] * normal
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(origins=f(means + normal * SECONDARY_NORMAL_EPS), directions=f(d), viewdirs=f(d),
                radii=np.ones((R, 1), np.float32), near=np.full((R, 1), SHADOW_NEAR_MAX, np.float32),
                far=np.full((R, 1), SECONDARY_FAR, np.float32))


_RAY_FIELDS = (("origins", 3), ("directions", 3), ("viewdirs", 3), ("radii", 1), ("near", 1), ("far", 1))


def pack_rays(rays_np, u01_np, extra=None):
    """One flat fp32 host buffer per step (a single pinned H2D copy), structure-of-arrays: the blocks
    origins [R,3] | directions [R,3] | viewdirs [R,3] | radii [R,1] | near [R,1] | far [R,1] | u01 x3 [R,1]
    | `extra` [R,E] (e.g. target rgb) lie back to back, so the device-side views are contiguous."""
    blocks = [rays_np[k] for k, _ in _RAY_FIELDS] + list(u01_np)
    if extra is not None:
        blocks.append(extra)
    return np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float32).reshape(-1) for b in blocks]))


def unpack_rays(buf, extra_cols=3):
    """Zero-copy views into the packed device buffer (see pack_rays)."""
    cols = sum(c for _, c in _RAY_FIELDS) + 3 + extra_cols
    R = buf.numel() // cols
    off, rays = 0, {}
    for k, c in _RAY_FIELDS:
        rays[k] = buf[off:off + R * c].view(R, c)
        off += R * c
    u01 = []
    for _ in range(3):
        u01.append(buf[off:off + R].view(R, 1))
        off += R
    extra = buf[off:off + R * extra_cols].view(R, extra_cols) if extra_cols else None
    return rays, u01, extra


def pack_batch(rays_np, u01_np, target, extra_rays_np, u01x_np):
    """Packed host buffer of one config-2 batch: pack_rays(main rays, u01, target) followed by
    pack_rays(backward-mask rays, their u01)."""
    return np.concatenate([pack_rays(rays_np, u01_np, target), pack_rays(extra_rays_np, u01x_np)])


def unpack_batch(buf, target_cols=3):
    """Views into a pack_batch buffer: (rays, u01, target, (extra_rays, u01x))."""
    per_ray = sum(c for _, c in _RAY_FIELDS)
    cols_main, cols_extra = per_ray + 3 + target_cols, per_ray + 3
    R = buf.numel() // (cols_main + cols_extra)
    rays, u01, target = unpack_rays(buf[:R * cols_main], target_cols)
    xrays, u01x, _ = unpack_rays(buf[R * cols_main:], 0)
    return rays, u01, target, (xrays, u01x)


def linear_to_srgb(linear):
    """image.linear_to_srgb (internal/image.py:192-200)."""
    eps = float(np.finfo(np.float32).eps)
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * torch.clamp(linear, min=eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


INTERLEVEL_MULTS = (0.01, 0.01)     # configs/ngp_yobo.gin:246
INTERLEVEL_BLURS = (0.03, 0.003)    # configs/ngp_yobo.gin:247


GEOMETRY_MULTS = (0.01, 0.001, 0.01)   # orientation, predicted normals, reverse (configs/nerf_ngp_yobo_lego.gin:7-11)
PREDICTED_NORMAL_STOPGRAD_WEIGHT = 0.1  # configs/nerf_ngp_yobo.gin:60
MASK_WEIGHTS = (1.0, 1.0)              # opaque_loss_weight, empty_loss_weight (configs/nerf_ngp_yobo.gin:367-368)
BACKWARD_MASK_WEIGHT = 0.1             # configs/nerf_ngp_yobo.gin:375
DENSITY_GRID_REGULARIZER = 1.0         # Config.param_regularizers['density_grid'] (configs/nerf_ngp_yobo.gin:47-51)
DISTORTION = (0.01, -0.25, 10000.0)    # mult (nerf_ngp_yobo_lego.gin:10), power_ladder p, premult (ngp_yobo.gin:250-253)


def density_grid_tables(params):
    """Level tables of every `density_grid` module of the cache model (the parameters the 'density_grid' regularizer
    prefix matches, internal/train_utils.py:1192-1201).  When the tables live in one arena leaf, differentiable
    views of that leaf are returned (the stored per-level views are pointer-only, see CacheTrainStep)."""
    out = []
    for name in sorted(params["Sampler"].keys()):
        grid = params["Sampler"][name]["density_grid"]
        arena = grid.get("_arena")
        for k in sorted(k for k in grid.keys() if k != "_arena"):
            t = grid[k]
            if arena is not None and arena.requires_grad and not t.requires_grad:
                off = (t.data_ptr() - arena.data_ptr()) // 4
                t = arena[off:off + t.numel()].view(t.shape)
            out.append(t)
    return out


def cache_loss(result, target_rgb, charb_padding=0.001, interlevel_fn=None, rays=None, extra_acc=None, lib=None,
               reg_tables=None):
    """Cache-stage objective of the config-2 step (internal/train_utils.py:2880-2950):
      data   Charbonnier on the sRGB-mapped render (cache_loss='charb', cache_linear_to_srgb=True,
             configs/ngp_yobo.gin:35-37; internal/configs.py:330)
      sampler  spline interlevel loss on the two proposal levels (mults (0.01, 0.01), blurs (0.03, 0.003),
             configs/ngp_yobo.gin:245-247; internal/loss_utils.py:74-108)
      distortion (when `rays` is given) mip-NeRF-360 distortion loss on the final level's metric distances through
             power_ladder(-0.25, 1e4) (loss_utils.py:108-123; mult 0.01 in the lego config)
      geometry (when `rays` is given) orientation + predicted-normal + reverse losses on the final level
             (train_utils.py:3255-3311), the middle one through the analytic normals' second-order path
      mask   compute_mask_loss on the accumulation (masks == 1 for the synthetic batches), and, when `extra_acc`
             (accumulation of the backward-mask rays' weights_only pass) is given, the backward-mask term.
      regularizer (when `reg_tables` is given) 0.5 * mean(T^2) over every density-grid level table
             (param_regularizer_loss, train_utils.py:1169-1216; nerf_ngp_yobo.gin:47-51)
    `lib` supplies spline_interlevel_loss / geometry_losses / compute_mask_loss: this package's CUDA mirrors by
    default, the oracle's restatement in the CPU legs."""
    if lib is None:
        from . import loss_utils as lib
    rgb = linear_to_srgb(result["render"]["rgb"])
    loss = torch.sqrt((rgb - target_rgb) ** 2 + charb_padding**2).mean()
    if interlevel_fn is None:
        interlevel_fn = lib.spline_interlevel_loss
    for l in interlevel_fn(result["sampler"], mults=INTERLEVEL_MULTS, blurs=INTERLEVEL_BLURS):
        loss = loss + l
    if rays is not None:
        loss = loss + lib.distortion_loss(result["sampler"], *DISTORTION)
        for l in lib.geometry_losses(rays, result["sampler"][-1], *GEOMETRY_MULTS, PREDICTED_NORMAL_STOPGRAD_WEIGHT):
            loss = loss + l
        loss = loss + lib.compute_mask_loss(result["render"]["acc"], None, charb_padding, *MASK_WEIGHTS)
    if reg_tables is not None:
        loss = loss + lib.param_regularizer_loss(reg_tables, DENSITY_GRID_REGULARIZER)
    if extra_acc is not None:
        loss = loss + lib.compute_mask_loss(extra_acc, None, charb_padding, empty_loss_weight=BACKWARD_MASK_WEIGHT,
                                            backward=True)
    return loss


def _he_uniform(gen, device, fan_in, fan_out):
    lim = float(np.sqrt(6.0 / fan_in))  # jax he_uniform (geometry.py:127)
    return torch.empty((fan_in, fan_out), device=device).uniform_(-lim, lim, generator=gen)


class CacheTrainStep:
    """BASELINE config 2: nerf_ngp_yobo_lego cache training step on a batch of rays --
    proposal sampler (64,64,32) -> cache shader on the 32 final samples -> volumetric rendering ->
    loss -> backward: gradients for the 4 hash-grid arenas (3 density grids + appearance grid) and
    every MLP weight on the path."""

    def __init__(self, device, table_init_range=0.1, seed=SEED, bf16=False, fused=True):
        self.device = device
        self.model = models.NeRFModel(bf16=bf16)
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.leaves = []

        def layer(fi, fo):
            k = _he_uniform(gen, device, fi, fo).requires_grad_(True)
            b = torch.zeros((fo,), device=device).requires_grad_(True)
            self.leaves += [k, b]
            return {"kernel": k, "bias": b}

        def grid(enc):
            _, arena = enc.init(device, generator=gen, init_range=table_init_range)
            arena.requires_grad_(True)
            self.leaves.append(arena)
            # views from the detached arena: pointers only; a view of the leaf would pin its
            # AccumulateGrad node to the init stream and break CUDA-graph capture of backward.
            return dict(enc.views(arena.detach()), _arena=arena)

        sampler = {}
        for i, m in enumerate(self.model.sampler.mlps):
            p = {"density_grid": grid(m.grid), "density_layers_0": layer(m.in_dim, 64),
                 "density_layers_1": layer(64, 64), "output_density_layer": layer(64, 1)}
            if m.enable_pred_normals:
                p["pred_normals_layer"] = layer(64, 3)
            sampler[f"MLP_{i}"] = p
        self.num_sampler_leaves = len(self.leaves)   # leaves (and the gradient arena) are ordered Sampler | Shader
        sh = self.model.shader

        def slf(in_dim):
            p, d = {}, in_dim
            for j, name in enumerate(["layer_0", "layer_1", "layer_2", "layer_bottleneck"]):
                p[name] = layer(d, 128)
                d = 128 + (in_dim if (j % 2 == 0 and j > 0) else 0)
            p["output_ambient_rgb_layer"] = layer(d, 3)
            return p

        shader = {
            "appearance_grid": grid(sh.grid), "bottleneck_layer": layer(96, 128), "roughness_layer": layer(96, 1),
            "ambient_irradiance_layer": layer(96, 3), "irradiance_layer": layer(96, 3), "tint_layer": layer(96, 3),
            "integrated_brdf_layers_0": layer(129, 64), "integrated_brdf_layers_1": layer(64, 64),
            "output_integrated_brdf_layer": layer(64, 1), "SurfaceLightField": slf(200), "EnvMap": slf(38),
        }
        self.params = {"Sampler": sampler, "Shader": shader}
        # One flat gradient arena for every parameter on the path: a single memset per step, a
        # single NCCL all-reduce under data parallelism (the reference's lax.pmean over the grad
        # pytree, internal/train_utils.py:3132-3136).  The kernels accumulate straight into it.
        from . import _lib
        align = 64  # floats: every sink starts 256-byte aligned (the kernels use 16-byte vector atomics)
        pad = lambda n: (n + align - 1) // align * align
        total = sum(pad(int(t.numel())) for t in self.leaves)
        # Under data parallelism the arena lives in symmetric memory so that the all-reduce can be the library's own
        # peer-memory kernel (dist.PeerArena); None -> plain tensor + NCCL.
        from . import dist as _ndist
        self._comm = None
        self.peer = _ndist.PeerArena.create(total, device)
        self.flat_grad = self.peer.buf[:total] if self.peer is not None else torch.zeros(total, device=device,
                                                                                         dtype=torch.float32)
        off = 0
        ranges = {}
        self.shader_offset = 0
        final_arena = sampler[f"MLP_{len(self.model.sampler.mlps) - 1}"]["density_grid"]["_arena"]
        self.final_level_offset = 0
        for i, t in enumerate(self.leaves):
            if t is final_arena:
                self.final_level_offset = off  # flat_grad[:final_level_offset] = the proposal levels' MLPs and grids
            if i == self.num_sampler_leaves:
                self.shader_offset = off       # flat_grad[:shader_offset] = sampler grads, [shader_offset:] = shader grads
                if t is not shader["appearance_grid"]["_arena"]:
                    raise RuntimeError("the appearance grid must be the shader's first leaf (gradient buckets)")
                self.shader_grid_end = off + pad(int(t.numel()))
            n = int(t.numel())
            sink = self.flat_grad[off:off + n].view(t.shape)
            _lib.register_grad_sink(t, sink)
            t.grad = sink
            ranges[id(t)] = (off, off + n)
            off += pad(n)
        # complement of the density grids' tables (see zero_grad): what a step still has to zero-fill itself
        skip = sorted(ranges[id(sampler[f"MLP_{i}"]["density_grid"]["_arena"])] for i in range(len(self.model.sampler.mlps)))
        self._zero_ranges, lo = [], 0
        for a, b in skip:
            if a > lo:
                self._zero_ranges.append((lo, a))
            lo = b
        if lo < total:
            self._zero_ranges.append((lo, total))
        from . import engine as _engine
        self.engine = _engine.FusedCacheStep(self.model, self.params) if (bf16 and fused) else None

    def num_params(self):
        return sum(int(t.numel()) for t in self.leaves)

    def zero_grad(self, skip_density_grids=False):
        """Clear the gradient arena.  skip_density_grids=True leaves out the three density grids' tables, which the
        fused step initialises with the parameter regularizer's (dense) gradient instead (nrc_grid_regularizer_init)."""
        if not skip_density_grids:
            self.flat_grad.zero_()
            return
        mode = os.environ.get("NRC_STREAM_FILLS", "1")
        rs = self._zero_ranges
        if mode == "torch" or len(rs) > 8 or any((lo | hi) & 3 for lo, hi in rs):
            for lo, hi in rs:
                self.flat_grad[lo:hi].zero_()
            return
        # one launch for all ranges, evict-first stores (NRC_STREAM_FILLS=0: plain stores)
        import ctypes as C
        lo_a = (C.c_int64 * len(rs))(*[lo for lo, _ in rs])
        hi_a = (C.c_int64 * len(rs))(*[hi for _, hi in rs])
        _lib.call("nrc_zero_ranges", _lib.stream_ptr(), _lib.ptr(self.flat_grad), lo_a, hi_a, len(rs), int(mode != "0"))

    def allreduce_grads(self, lo=0, hi=None, channel=0, num_ctas=0, mode=None, leading_barrier=True, trailing_barrier=True):
        """Mean over ranks of flat_grad[lo:hi] in place on the current stream (the reference's lax.pmean,
        internal/train_utils.py:3132-3136).  Concurrent buckets (different streams) need different channels."""
        from . import dist as _ndist
        hi = self.flat_grad.numel() if hi is None else hi
        if self.peer is not None:
            self.peer.allreduce_mean_(lo, hi - lo, channel, num_ctas, mode, leading_barrier, trailing_barrier)
        else:
            _ndist.allreduce_mean_(self.flat_grad[lo:hi])

    @property
    def allreduce_kind(self):
        return f"peer-memory kernel ({self.peer.mode})" if self.peer is not None else "nccl"

    def step(self, rays, u01, target_rgb, extra=None, fused_allreduce=False):
        """Forward + loss + backward of one ray batch; returns the loss (device scalar).  The bf16
        variant runs the hand-ordered launch schedule of engine.FusedCacheStep (same kernels, no
        elementwise glue); the fp32 parity variant goes through the autograd mirrors.
        `extra` = (backward-mask rays, their u01): see backward_mask_rays_np."""
        if self.engine is not None:
            if fused_allreduce and self.peer is not None:
                # Data parallel, all-reduce INSIDE the step (and inside its CUDA graph), in three buckets that follow the
                # order in which the backward pass finishes gradients:
                #   proposal levels (MLP_0 / MLP_1 grids + MLPs, 17 MB): final as soon as the proposal branch's backward is
                #     done - it runs beside the shader's forward - so this bucket hides behind the shader;
                #   shader (appearance grid + stacks, 47 MB): forked the moment the shader's backward is done, beside the
                #     final level's backward;
                #   final level (MLP_2 grid + MLP, 47 MB): the only one left at the end of the step.
                # The overlapped buckets run with few CTAs (they share the SMs with the backward kernels).
                if self._comm is None:
                    self._comm = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
                comm, comm_p, comm_f = self._comm
                so, fo = self.shader_offset, self.final_level_offset
                n_overlap = int(os.environ.get("NRC_AR_CTAS_OVERLAP", "32" if self.peer.mode == "multicast" else "74"))
                three = os.environ.get("NRC_AR_BUCKETS", "3") == "3" and fo > 0

                go = self.shader_grid_end      # [so, go) = appearance grid, [go, end) = the stacks' weights

                # One trailing barrier for all buckets (nothing reads the reduced gradients before the end of the step)
                # instead of one per bucket; the stacks' small range shares the final level's leading barrier.
                lean = os.environ.get("NRC_AR_LEAN", "1") == "1"
                tb = not lean
                ev = {}

                def grid_bucket():                  # the appearance grid's scatter is done (the stacks' wgrad follows)
                    comm.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(comm):   # fewer CTAs: it shares the SMs with the backward kernels
                        self.allreduce_grads(so, go, channel=1, num_ctas=n_overlap, trailing_barrier=tb)

                def shader_bucket():                # everything of the shader is final
                    if lean and split_shader and "final" in ev:
                        # the stacks' weights (0.75 MB) ride behind the final level's barrier: both are complete now
                        comm_f.wait_event(ev["final"])
                        comm_f.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(comm_f):
                            self.allreduce_grads(fo, so, channel=0, trailing_barrier=False)
                            self.allreduce_grads(go, None, channel=0, leading_barrier=False, trailing_barrier=False)
                        ev["final_done"] = True
                        return
                    comm.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(comm):
                        self.allreduce_grads(go if split_shader else so, None, channel=1, num_ctas=n_overlap, trailing_barrier=tb)

                def proposal_bucket():              # called on the proposal branch's stream, after its backward
                    comm_p.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(comm_p):
                        self.allreduce_grads(0, fo, channel=2, num_ctas=n_overlap, trailing_barrier=tb)

                def final_bucket():                 # called on the stream of the final level's backward (engine tail fork)
                    if lean and split_shader:       # reduced together with the stacks' range (shader_bucket)
                        ev["final"] = torch.cuda.Event()
                        ev["final"].record()
                        return
                    comm_f.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(comm_f):
                        self.allreduce_grads(fo, so, channel=0, trailing_barrier=tb)

                split_shader = three and os.environ.get("NRC_AR_SPLIT_SHADER", "1") == "1"
                loss = self.engine.step(rays, u01, target_rgb, extra=extra, zero_grad=self.zero_grad,
                                        on_shader_grads=shader_bucket, on_proposal_grads=proposal_bucket if three else None,
                                        on_final_grads=final_bucket if three else None,
                                        on_grid_grads=grid_bucket if split_shader else None)
                cur = torch.cuda.current_stream()
                if three and self.engine.final_grads_announced:
                    if lean and split_shader and not ev.get("final_done"):
                        raise RuntimeError("final level's bucket was announced but not reduced")
                    cur.wait_stream(comm_f)
                else:
                    self.allreduce_grads(fo if three else 0, so, channel=0, trailing_barrier=tb)
                cur.wait_stream(comm)
                if three:
                    cur.wait_stream(comm_p)
                if lean:
                    self.peer.barrier()
                return loss
            return self.engine.step(rays, u01, target_rgb, extra=extra, zero_grad=self.zero_grad)
        return self.step_autograd(rays, u01, target_rgb, extra)

    def step_front(self, rays, u01, target_rgb, extra=None):
        """First half of step() (forward, loss, shader backward); see engine.FusedCacheStep.step_front."""
        return self.engine.step_front(rays, u01, target_rgb, extra=extra, zero_grad=self.zero_grad)

    def step_back(self, state):
        self.engine.step_back(state)
        return state["loss"]

    def step_autograd(self, rays, u01, target_rgb, extra=None):
        self.zero_grad()
        res = self.model(self.params, rays, u01, train=True)
        extra_acc = self.model.weights_only(self.params, extra[0], extra[1]) if extra is not None else None
        loss = cache_loss(res, target_rgb, rays=rays, extra_acc=extra_acc, reg_tables=density_grid_tables(self.params))
        loss.backward()
        return loss.detach()

    def render(self, rays, u01):
        """Forward-only evaluation (BASELINE config 1)."""
        with torch.no_grad():
            return self.model(self.params, rays, u01, train=False)["render"]


# ----------------------------------------------------------------------------- configs 3 and 5 (render path)
def _cache_params(model, device, gen, table_init_range):
    """Random-init parameters of the cache model in the reference's layout (no gradients: render path)."""
    step = CacheTrainStep.__new__(CacheTrainStep)
    step.device, step.model, step.leaves = device, model, []

    def layer(fi, fo):
        return {"kernel": _he_uniform(gen, device, fi, fo), "bias": torch.zeros((fo,), device=device)}

    def grid(enc):
        _, arena = enc.init(device, generator=gen, init_range=table_init_range)
        return dict(enc.views(arena), _arena=arena)

    sampler = {}
    for i, m in enumerate(model.sampler.mlps):
        p = {"density_grid": grid(m.grid), "density_layers_0": layer(m.in_dim, 64), "density_layers_1": layer(64, 64),
             "output_density_layer": layer(64, 1)}
        if m.enable_pred_normals:
            p["pred_normals_layer"] = layer(64, 3)
        sampler[f"MLP_{i}"] = p

    def slf(in_dim):
        p, d = {}, in_dim
        for j, name in enumerate(["layer_0", "layer_1", "layer_2", "layer_bottleneck"]):
            p[name] = layer(d, 128)
            d = 128 + (in_dim if (j % 2 == 0 and j > 0) else 0)
        p["output_ambient_rgb_layer"] = layer(d, 3)
        return p

    shader = {"appearance_grid": grid(model.shader.grid), "bottleneck_layer": layer(96, 128), "roughness_layer": layer(96, 1),
              "ambient_irradiance_layer": layer(96, 3), "irradiance_layer": layer(96, 3), "tint_layer": layer(96, 3),
              "integrated_brdf_layers_0": layer(129, 64), "integrated_brdf_layers_1": layer(64, 64),
              "output_integrated_brdf_layer": layer(64, 1), "SurfaceLightField": slf(200), "EnvMap": slf(38)}
    return {"Sampler": sampler, "Shader": shader}


def make_surface_np(g, R):
    """Config 3 inputs: surface points p ~ U(ball r=1), unit normals, camera view directions facing them."""
    p = g.normal(size=(R, 3))
    p = p / np.linalg.norm(p, axis=-1, keepdims=True) * g.uniform(size=(R, 1)) ** (1 / 3)
    n = g.normal(size=(R, 3))
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    v = -n + 0.5 * g.normal(size=(R, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return f(p), f(v), f(n)


class MaterialRenderStep:
    """BASELINE config 3 (material_light_from_scratch_resample, chunk of 1024 shaded points): 32 secondary
    rays per point (16 microfacet + 8 cosine + 8 vMF mixture with 128 lobes), radiance-cache query along every
    secondary ray (proposal sampler (64,64,32), power-ladder warp, categorical resample to one shaded sample,
    cache shader), environment map behind it, GGX / Lambert Monte-Carlo integration."""

    NUM_LOBES = 128

    def __init__(self, device, table_init_range=0.1, seed=SEED, bf16=True, slf_variate=False):
        from . import material
        self.device = device
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.gen = gen
        self.cache = models.NeRFModel(bf16=bf16)
        # slf_variate: MaterialModel.slf_variate with NeRFModel.use_surface_light_field (nerf_ngp_yobo.gin:91): the
        # surface-light-field memory is queried on the same secondary rays and its integral subtracted
        self.model = material.MaterialModel(self.cache, bf16=bf16, slf_variate=slf_variate)
        from . import light_sampler
        self.light = light_sampler.LightMLP(bf16=bf16)      # the vMF lobes of the light sampler
        self.params = {"Cache": _cache_params(self.cache, device, gen, table_init_range),
                       "Material": self.model.material_mlp.init(device, gen, table_init_range),
                       "EnvMap": self.model.env_map.init(device, gen),
                       "Light": self.light.init(device, gen, table_init_range)}
        if slf_variate:
            self.params["SurfaceLightFieldMem"] = self.model.surface_lf_mem.init(device, gen, table_init_range)
        # LightMLP.get_vmfs adds a fixed-key normal draw * vmf_scale / 2 to the lobe means (light_sampler.py:141-143)
        self.means_random = torch.randn((self.NUM_LOBES, 3), device=device, generator=gen) * (self.light.vmf_scale / 2.0)

    def draws(self, R):
        """All random inputs of one chunk, generated on the device (the kernels take them as tensors)."""
        S, dev, gen = self.model.num_secondary, self.device, self.gen
        u = lambda *shape: torch.rand(shape, device=dev, generator=gen)
        gum = -torch.log(-torch.log(torch.clamp(u(R * S, 32, 1), min=1e-12)))
        return dict(u=u(R, S, 2), latent=torch.randint(0, self.NUM_LOBES, (R,), device=dev, generator=gen, dtype=torch.int32),
                    normal2=torch.randn((R, self.model.n_light, 2), device=dev, generator=gen),
                    u01=[u(R * S, 1) for _ in range(3)], gumbel=gum)

    def light_lobes(self, R, means=None, normals=None):
        """vMF lobes of the light sampler at the shaded points: LightMLP.predict_lighting (light grid -> tcgen05
        stack -> vMF head) when `means` is given; synthetic lobes otherwise."""
        if means is not None:
            with torch.no_grad():
                return self.light.predict_lighting(self.params["Light"], means, self.means_random, normals=normals)
        dev, gen = self.device, self.gen
        return dict(vmf_means=torch.randn((R, self.NUM_LOBES, 3), device=dev, generator=gen),
                    vmf_kappas=torch.rand((R, self.NUM_LOBES, 1), device=dev, generator=gen) * 50.0,
                    vmf_logits=torch.randn((R, self.NUM_LOBES, 1), device=dev, generator=gen))

    def render(self, means, viewdirs, normals, draws, lobes, material=None):
        return self.model.render_chunk(self.params, means, viewdirs, normals, draws, material=material,
                                       light_sampler_results=lobes)


class FrameRenderer:
    """BASELINE config 5: one 1024-ray chunk of a full-view render = cache stage on the primary rays
    (config 1, resampled to one shaded point per ray) + material stage at that point (config 3) +
    compositing with the resampled weight over a white background."""

    def __init__(self, device, bf16=True, seed=SEED, use_graph=True):
        self.stage = MaterialRenderStep(device, seed=seed, bf16=bf16)
        # random draws inside the captured chunk come from the default CUDA generator (graph-safe Philox offsets)
        self.stage.gen = None
        self.device = device
        self.use_graph = use_graph
        self._graphs = {}

    def render_chunk_graphed(self, rays, repeat=0):
        """render_chunk behind one CUDA graph per chunk size: ray buffers are static, the ~70 launches of a
        chunk replay without host work (the per-chunk host sync of models.render_image disappears)."""
        if not self.use_graph:
            return self.render_chunk(rays, repeat)
        R = rays["origins"].shape[0]
        if R not in self._graphs:
            static = {k: v.clone() for k, v in rays.items()}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.render_chunk(static, repeat)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.render_chunk(static, repeat)
            self._graphs[R] = (graph, static, out)
        graph, static, out = self._graphs[R]
        for k, v in rays.items():
            static[k].copy_(v)
        graph.replay()
        return out

    def render_chunk(self, rays, repeat=0):
        st, dev, gen = self.stage, self.device, self.stage.gen
        R = rays["origins"].shape[0]
        with torch.no_grad():
            u01 = [torch.rand((R, 1), device=dev, generator=gen) for _ in range(3)]
            gum = -torch.log(-torch.log(torch.clamp(torch.rand((R, 32, 1), device=dev, generator=gen), min=1e-12)))
            if st.model.fused_query is not None:
                q = st.model.fused_query(st.params["Cache"], rays, u01, gumbel=gum, is_secondary=False, resample=True)
                means, normals, w = q["means"].reshape(R, 3), q["normals"].reshape(R, 3), q["weights"].reshape(R, 1)
                cache_rgb, acc = q["rgb"], q["acc"].reshape(R, 1)
            else:
                prim = st.cache(st.params["Cache"], rays, u01, gumbel=gum, train=False, is_secondary=False, resample=True)
                sh = prim["shaded"]
                means, normals, w = sh["means"].reshape(R, 3), sh["normals"].reshape(R, 3), sh["weights"].reshape(R, 1)
                cache_rgb, acc = prim["render"]["rgb"], prim["render"]["acc"].reshape(R, 1)
            means, normals = means.contiguous(), normals.contiguous()
            out = st.render(means, rays["viewdirs"], normals, st.draws(R), st.light_lobes(R, means, normals))
            rgb = out["rgb"] * w + (1.0 - acc)
        return dict(rgb=rgb, cache_rgb=cache_rgb, acc=acc, albedo=out["material"]["albedo"])


class TransientRenderStep:
    """BASELINE config 4 (transient_simulation_ngp_yobo_cornell.gin, time-resolved cache): one chunk of primary rays through
    the proposal sampler (64, 64, 32), the transient cache shader on the 32 final samples and the time-resolved integrator,
    700 bins of 0.01 (Config.n_bins / exposure_time, gin:17-18), near 0.7 / far 4, light_near 0.7 with light_zero,
    bin_zero_threshold_light 100, indirect_scale 0.05, rgb_max 100, tfilter_sigma 3 (internal/configs.py:710).

    Per shaded sample: appearance feature (appearance grid) -> bottleneck, roughness, tint, albedo, integrated BRDF;
    diffuse transient head = irradiance stack on [feature | pos_enc(light position)] (internal/nerf.py:1757-1777) and
    specular transient head = transient SurfaceLightField on [bottleneck | IDE_5(reflection, roughness)]
    (internal/surface_light_field.py:782-1069 with use_indirect: 128 -> 700*3 + 1); their LAST layers run inside the
    time-resolved kernel (render.volumetric_transient_rendering_fused), so the [R, 32, 700, 3] histograms the reference
    materialises three to four times never exist.  The direct term is the un-occluded diffuse response to the point light
    (albedo n.l power / d^2 / pi, nerf.py:1141-1170,1474-1481); the shadow-ray visibility query and the light BRDF network
    of the reference's active path are not part of this workload.  render() is the fused forward path; render_unfused() /
    loss_and_grads() the differentiable one (training)."""

    def __init__(self, device, n_bins=700, table_init_range=0.1, seed=SEED, bf16=True):
        self.device, self.n_bins, self.bf16 = device, n_bins, bf16
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.cache = models.NeRFModel(bf16=bf16)
        self.params = _cache_params(self.cache, device, gen, table_init_range)
        self.head = nerf.TransientIndirectHead(n_bins=n_bins, bf16=bf16)

        def layer(fi, fo):
            return {"kernel": _he_uniform(gen, device, fi, fo), "bias": torch.zeros((fo,), device=device)}

        t = self.head.init(device, gen)
        t["albedo_layer"] = layer(96, 3)
        slf, d = {}, 200
        for j, name in enumerate(["layer_0", "layer_1", "layer_2", "layer_bottleneck"]):
            slf[name] = layer(d, 128)
            d = 128 + (200 if (j % 2 == 0 and j > 0) else 0)
        slf["output_rgba_layer"] = layer(128, n_bins * 3 + 1)
        t["TransientSurfaceLightField"] = slf
        self.tparams = t
        self.ide5 = nerf.generate_ide_fn(5)
        self.cfg = dict(exposure_time=0.01, shift=0.0, diffuse_bias=-2.0, spec_bias=-2.0, indirect_scale=0.05,
                        bin_zero_threshold_light=100.0, light_zero=True, light_near=0.7, rgb_max=100.0, dark_level=0.0,
                        tfilter_sigma=3.0, filter_indirect=False)
        self.light_power = float(np.exp(3.9))     # light_power_activation = safe_exp, light_power_bias = 3.9 (gin:111-114)
        self._head_pack = {}                      # packed bf16 image of the two head layers (repacked when they change)
        # render path of the bf16 variant: the shader-side Dense stacks as tcgen05 chain programs (spec, packed-weight cache);
        # a stack whose LAST hidden activation is the product (h_diffuse, h_specular) ends in a linear head + ReLU
        PC = mlp_chain.PackCache
        self._chains = {
            "trunk": (mlp_chain.ChainSpec(in_widths=[96], hidden=[], heads=[[("bottleneck_layer", 128)],
                      [("roughness_layer", 1), ("tint_layer", 3), ("albedo_layer", 3)]]), PC()),
            "brdf": (self.cache.shader.brdf_chain, PC()),
            "slf": (mlp_chain.ChainSpec(in_widths=[128, 72], hidden=[("layer_0", 128, False), ("layer_1", 128, False),
                    ("layer_2", 128, True)], heads=[[("layer_bottleneck", 128)]]), PC()),
            "irr": (mlp_chain.ChainSpec(in_widths=[96, 15], hidden=[("irradiance_layers_0", 64, False)],
                    heads=[[("irradiance_layers_1", 64)]]), PC()),
        } if bf16 else None

    def make_rays(self, g, R):
        """Synthetic rays of the Cornell-box scale: near 0.7 / far 4, a point light beside the camera."""
        rn = make_rays_np(g, R, near=0.7, far=4.0, radius=2.0)
        lights = (rn["origins"] + 0.05 * g.normal(size=(R, 3))).astype(np.float32)
        return dict(rn, lights=lights, cam_origins=rn["origins"].copy())

    def _shade(self, rays, u01, train):
        """Everything in front of the time-resolved integration: (direct, h_diffuse, h_specular, spec_scale, weights,
        ray / light / camera distances).  Under autograd when `train`."""
        sp = torch.nn.functional.softplus
        p, tp, b = self.params["Shader"], self.tparams, self.bf16
        shader = self.cache.shader
        with torch.enable_grad() if train else torch.no_grad():
            last = self.cache.sampler(self.params["Sampler"], rays, u01, train=train)[-1]
            means, feat, nrm, w = last["means"], last["feature"], last["normals_to_use"], last["weights"]
            R, n = w.shape
            P = R * n
            feature = shader.predict_appearance_feature(p, feat, means).reshape(P, 96)
            view = rays["viewdirs"][:, None, :].expand(R, n, 3).reshape(P, 3)
            n2 = nrm.reshape(P, 3)
            dot = torch.sum(n2 * (-view), dim=-1, keepdim=True)
            m2 = means.reshape(P, 3)
            lights = rays["lights"][:, None, :].expand(R, n, 3).reshape(P, 3)
            chains = b and not train       # render path of the bf16 variant: the Dense stacks as tcgen05 chain programs
            if chains:
                ch, fc = self._chains, mlp_chain.forward_cached
                slf = tp["TransientSurfaceLightField"]
                bott, r_raw, t_raw, a_raw = fc(ch["trunk"][0], dict(p, albedo_layer=tp["albedo_layer"]), [feature], ch["trunk"][1])
                rough, tint, albedo = sp(r_raw - 1.0), torch.sigmoid(t_raw), torch.sigmoid(a_raw - 1.0)
                (f_raw,) = fc(ch["brdf"][0], p, [bott, dot], ch["brdf"][1])
                F = torch.sigmoid(f_raw + float(np.log(3.0)))
                refdirs = nerf.reflect(-view, n2)
                (h_s,) = fc(ch["slf"][0], slf, [bott, self.ide5(refdirs, rough)], ch["slf"][1])
                h_s = torch.relu(h_s)                                                      # [P, 128]
                enc_l = torch.empty((P, 15), device=feature.device, dtype=torch.float32)
                _lib.call("nrc_pos_enc", _lib.stream_ptr(), _lib.ptr(lights.contiguous()), P, 3, 0, 2, 1, _lib.ptr(enc_l), 15)
                (h_d,) = fc(ch["irr"][0], tp, [feature, enc_l], ch["irr"][1])
                h_d = torch.relu(h_d)                                                      # [P, 64]
            else:
                bott = nerf.dense(p["bottleneck_layer"], feature, bf16=b)
                rough = sp(nerf.dense(p["roughness_layer"], feature, bf16=b) - 1.0)
                tint = torch.sigmoid(nerf.dense(p["tint_layer"], feature, bf16=b))
                albedo = torch.sigmoid(nerf.dense(tp["albedo_layer"], feature, bf16=b) - 1.0)
                x = nerf.dense(p["integrated_brdf_layers_0"], torch.cat([bott, dot], dim=-1), relu=True, bf16=b)
                x = nerf.dense(p["integrated_brdf_layers_1"], x, relu=True, bf16=b)
                F = torch.sigmoid(nerf.dense(p["output_integrated_brdf_layer"], x, bf16=b) + float(np.log(3.0)))
                refdirs = nerf.reflect(-view, n2)
                xin = torch.cat([bott, self.ide5(refdirs, rough)], dim=-1)
                x = xin
                for j, name in enumerate(["layer_0", "layer_1", "layer_2", "layer_bottleneck"]):
                    x = nerf.dense(tp["TransientSurfaceLightField"][name], x, relu=True, bf16=b)
                    if j % 2 == 0 and j > 0:
                        x = torch.cat([x, xin], dim=-1)
                h_s = x                                                                    # [P, 128]
                h_d = self.head.hidden(tp, feature, lights)                                 # [P, 64]
            off = lights - m2
            light_d = torch.linalg.norm(off, dim=-1, keepdim=True)
            n_dot_l = torch.clamp(torch.sum(n2 * (off / torch.clamp(light_d, min=1e-5)), dim=-1, keepdim=True), min=0.0)
            radiance = self.light_power / torch.clamp(light_d ** 2, min=1e-5)
            radiance = torch.where(light_d < self.cfg["light_near"], torch.zeros_like(radiance), radiance)
            direct = torch.clamp(albedo * n_dot_l * radiance / np.pi, 0.0, self.cfg["rgb_max"]).reshape(R, n, 3)
            ray_d = torch.linalg.norm(rays["origins"][:, None, :] - means, dim=-1)
            cam_d = ray_d + torch.linalg.norm(rays["origins"] - rays["cam_origins"], dim=-1)[:, None]
            return (direct, h_d.reshape(R, n, 64), h_s.reshape(R, n, 128), (tint * F).reshape(R, n, 3), w, ray_d.detach(),
                    light_d.reshape(R, n).detach(), cam_d.detach())

    def render(self, rays, u01):
        from . import render as nrender
        tp = self.tparams
        direct, h_d, h_s, spec_scale, w, ray_d, light_d, cam_d = self._shade(rays, u01, train=False)
        with torch.no_grad():
            return nrender.volumetric_transient_rendering_fused(
                direct, h_d, tp["transient_indirect_layer"], h_s, tp["TransientSurfaceLightField"]["output_rgba_layer"],
                spec_scale, w, ray_d, light_d, cam_d, n_bins=self.n_bins, pack_cache=self._head_pack, **self.cfg)

    def render_unfused(self, rays, u01, train=False):
        """The same frame through the UNFUSED kernels: both heads' last layers as GEMMs, their [R, n, n_bins, 3] histograms in
        HBM (what the reference materialises), nrc_transient_render_fwd, temporal filter.  Differentiable when `train`
        (nrc_transient_render_bwd + the Dense / encoding / sampler VJPs): the training path of config 4."""
        from . import render as nrender
        tp, c, b = self.tparams, self.cfg, self.bf16
        direct, h_d, h_s, spec_scale, w, ray_d, light_d, cam_d = self._shade(rays, u01, train=train)
        R, n = w.shape
        B = self.n_bins
        with torch.enable_grad() if train else torch.no_grad():
            diffuse_raw = nerf.dense(tp["transient_indirect_layer"], h_d.reshape(R * n, 64), bf16=b).reshape(R, n, B, 3)
            raw_s = nerf.dense(tp["TransientSurfaceLightField"]["output_rgba_layer"], h_s.reshape(R * n, 128), bf16=b)
            # rgb_activation(rgb_premultiplier * raw + rgb_bias) of the transient light field (surface_light_field.py:1045-1047)
            specular = torch.nn.functional.softplus(raw_s[:, :B * 3] + c["spec_bias"]).reshape(R, n, B, 3)
            out = nrender.volumetric_transient_rendering(
                direct, diffuse_raw, specular, spec_scale, w, ray_d, light_d, cam_d, n_bins=B, exposure_time=c["exposure_time"],
                shift=c["shift"], diffuse_bias=c["diffuse_bias"], indirect_scale=c["indirect_scale"],
                bin_zero_threshold_light=c["bin_zero_threshold_light"], light_zero=c["light_zero"], light_near=c["light_near"],
                rgb_max=c["rgb_max"], dark_level=c["dark_level"])
            t_direct, t_indirect = out["transient_direct"], out["transient_indirect"]
            if c["tfilter_sigma"] != 0.0:
                filt = nrender.gaussian_tfilter(c["tfilter_sigma"], t_direct.device)
                t_direct = nrender.temporal_filter(t_direct, filt)
                if c["filter_indirect"]:
                    t_indirect = nrender.temporal_filter(t_indirect, filt)
            rgb = t_direct + t_indirect + c["dark_level"]
        return dict(transient_direct=t_direct, transient_indirect=t_indirect, rgb=rgb, integrated_rgb=rgb.sum(-2))

    def trainable(self):
        """name -> leaf tensor of every parameter the time-resolved objective reaches (grids as their flat arenas)."""
        leaves = {}

        def walk(prefix, node):
            for k, v in node.items():
                if isinstance(v, dict):
                    if "_arena" in v:
                        leaves[prefix + k] = v["_arena"]
                    else:
                        walk(prefix + k + "/", v)
                elif isinstance(v, torch.Tensor) and v.is_floating_point():
                    leaves[prefix + k] = v
        walk("Transient/", self.tparams)
        walk("Shader/", self.params["Shader"])
        walk("Sampler/", self.params["Sampler"])
        return leaves

    def loss_and_grads(self, rays, u01, target):
        """One training evaluation of config 4 on a synthetic target histogram [R, n_bins, 3]: mean squared error of the
        time-resolved `rgb`, back-propagated through the integrator (nrc_transient_render_bwd), both heads, the shader and
        the sampler's final level.  Returns (loss, {name: gradient}); parameters nothing reaches are absent."""
        leaves = self.trainable()
        for t in leaves.values():
            t.requires_grad_(True)
            t.grad = None
        try:
            out = self.render_unfused(rays, u01, train=True)
            loss = torch.mean((out["rgb"] - target) ** 2)
            loss.backward()
            grads = {k: t.grad for k, t in leaves.items() if t.grad is not None}
        finally:
            for t in leaves.values():
                t.requires_grad_(False)
                t.grad = None
        return loss.detach(), grads
