"""The JAX side of the drop-in boundary: what a maintainer of the reference adds so that its model code
(internal/grid_utils.py, render.py, stepfun.py, models.py, geometry.py, inverse_render/render_utils.py) runs the B200
kernels of libnrc_b200.so.  The reference's function signatures stay unchanged; the dispatch seam is ResampleOpMode
(internal/grid_utils.py:651-676), patched by `install(grid_utils)` below.

Two layers:

1. Descriptor packing (pure ctypes, no jax needed; exercised by tests/test_xla_gpu.py against the compiled targets):
   mirrors of include/nrc_xla.h and `pack_*` helpers producing the `opaque` byte strings.

2. JAX primitives (need `jax` / `jaxlib`; written for the pinned jax==0.4.16 - requirements.txt:2 - whose GPU custom
   calls use the original ABI void(stream, buffers, opaque, opaque_len)): one primitive per target with abstract
   evaluation and an MLIR lowering to `stablehlo.custom_call`, and `jax.custom_vjp` wrappers with the reference's call
   signatures (cuda_hash_encode, compute_alpha_weights, sample_intervals, cast_rays_means, volumetric_rendering,
   maybe_resample, integrate_reflect_rays, contract).  JAX is not installable in the build image of this repository, so
   layer 2 is shipped untested here; the C side it calls is tested through exactly the ABI XLA uses.
"""
import ctypes as C
import functools

from .. import _lib

DESC_VERSION = 1
NRC_MAX_LEVELS = _lib.NRC_MAX_LEVELS


# ----------------------------------------------------------------------------- layer 1: descriptors
class nrc_xla_encode_desc_t(C.Structure):
    _fields_ = [("version", C.c_int32), ("reserved", C.c_int32), ("num_points", C.c_int64), ("arena_floats", C.c_int64),
                ("level_offset", C.c_int64 * NRC_MAX_LEVELS), ("enc", _lib.nrc_encoding_t)]


class nrc_xla_contract_desc_t(C.Structure):
    _fields_ = [("version", C.c_int32), ("reserved", C.c_int32), ("num_points", C.c_int64), ("c", C.c_float), ("pad", C.c_int32)]


class nrc_xla_density_query_desc_t(C.Structure):
    _fields_ = [("grid", nrc_xla_encode_desc_t), ("in_dim", C.c_int32), ("width", C.c_int32), ("has_pred_normals", C.c_int32),
                ("want_raw_grad", C.c_int32), ("bf16", C.c_int32), ("warp_c", C.c_float), ("density_bias", C.c_float),
                ("pad", C.c_int32)]


class nrc_xla_ray_desc_t(C.Structure):
    _fields_ = [("version", C.c_int32), ("n", C.c_int32), ("num_rays", C.c_int64), ("opaque_background", C.c_int32),
                ("m", C.c_int32), ("anneal", C.c_float), ("padding", C.c_float), ("max_jitter", C.c_float),
                ("dom_lo", C.c_float), ("dom_hi", C.c_float), ("warp_kind", C.c_int32), ("p", C.c_float),
                ("premult", C.c_float), ("k", C.c_int32), ("channels", C.c_int32), ("has_rgb", C.c_int32),
                ("has_bg", C.c_int32), ("has_weights_nf", C.c_int32), ("bias", C.c_float), ("mult", C.c_float),
                ("pad", C.c_int32)]


class nrc_xla_ggx_desc_t(C.Structure):
    _fields_ = [("version", C.c_int32), ("num_samples", C.c_int32), ("num_points", C.c_int64), ("lobe_kind", C.c_int32),
                ("has_occ", C.c_int32), ("rgb_max", C.c_float), ("pad", C.c_int32)]


class nrc_xla_slf_desc_t(C.Structure):
    _fields_ = [("version", C.c_int32), ("num_features", C.c_int32), ("num_points", C.c_int64), ("ld_raw", C.c_int64),
                ("cfg", _lib.nrc_slf_points_t)]


TARGETS = ("nrc_xla_encode_fwd", "nrc_xla_encode_bwd", "nrc_xla_contract_fwd", "nrc_xla_contract_bwd",
           "nrc_xla_density_query_fwd", "nrc_xla_ray_alpha_weights_fwd", "nrc_xla_ray_alpha_weights_bwd",
           "nrc_xla_ray_sample_intervals", "nrc_xla_ray_cast", "nrc_xla_ray_composite_fwd", "nrc_xla_ray_composite_bwd",
           "nrc_xla_ray_resample", "nrc_xla_ray_resample_gather", "nrc_xla_ggx_integrate_fwd", "nrc_xla_ggx_integrate_bwd",
           "nrc_xla_slf_points_fwd", "nrc_xla_slf_points_bwd", "nrc_xla_slf_reduce_fwd", "nrc_xla_slf_reduce_bwd")


def _bytes(desc):
    return bytes(memoryview(desc))


def pack_encode(encoding, num_points):
    """opaque of nrc_xla_encode_{fwd,bwd} for a grid_utils.HashEncoding (this package's mirror or any object with the
    reference's fields: grid_sizes, hash_map_size, num_features, bbox, precondition_scaling).  The arena is the level
    tables back to back in level order (HashEncoding.level_layout)."""
    import numpy as np

    d = nrc_xla_encode_desc_t()
    d.version, d.num_points = DESC_VERSION, int(num_points)
    e = d.enc
    sizes = [int(n) for n in encoding.grid_sizes]
    e.num_levels, e.num_features = len(sizes), int(encoding.num_features)
    bbox = np.asarray(encoding.bbox, dtype=np.float64)
    for a in range(3):
        e.bbox_min[a], e.bbox_max[a] = float(np.float32(bbox[0][a])), float(np.float32(bbox[1][a]))
        e.bbox_span[a] = float(np.float32(np.float32(bbox[1][a]) - np.float32(bbox[0][a])))
    e.precondition_scaling = float(encoding.precondition_scaling)
    off = 0
    for l, n in enumerate(sizes):
        is_hash = n ** 3 > int(encoding.hash_map_size)          # grid_utils.py:837 (Python ints: no int32 wrap at 2048^3)
        rows = int(encoding.hash_map_size) if is_hash else n ** 3
        e.levels[l].grid_size, e.levels[l].is_hash, e.levels[l].table_size = n, int(is_hash), rows
        d.level_offset[l] = off
        off += rows * e.num_features
    d.arena_floats = off
    return d


def pack_contract(num_points, c):
    d = nrc_xla_contract_desc_t()
    d.version, d.num_points, d.c = DESC_VERSION, int(num_points), float(c)
    return d


def pack_density_query(encoding, num_points, in_dim, width=64, has_pred_normals=False, want_raw_grad=False, bf16=False,
                       warp_c=2.0, density_bias=-1.0):
    d = nrc_xla_density_query_desc_t()
    d.grid = pack_encode(encoding, num_points)
    d.in_dim, d.width = int(in_dim), int(width)
    d.has_pred_normals, d.want_raw_grad, d.bf16 = int(has_pred_normals), int(want_raw_grad), int(bf16)
    d.warp_c, d.density_bias = float(warp_c), float(density_bias)
    return d


def pack_ray(num_rays, n, **kw):
    d = nrc_xla_ray_desc_t()
    d.version, d.num_rays, d.n = DESC_VERSION, int(num_rays), int(n)
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def pack_ggx(num_points, num_samples, lobe_kind=0, has_occ=False, rgb_max=3.4e38):
    d = nrc_xla_ggx_desc_t()
    d.version, d.num_points, d.num_samples = DESC_VERSION, int(num_points), int(num_samples)
    d.lobe_kind, d.has_occ, d.rgb_max = int(lobe_kind), int(has_occ), float(rgb_max)
    return d


def pack_slf(num_points, cfg, ld_raw=None, num_features=0):
    """cfg: a filled _lib.nrc_slf_points_t (module constants of the reference's surface_lf_mem)."""
    d = nrc_xla_slf_desc_t()
    d.version, d.num_points, d.num_features = DESC_VERSION, int(num_points), int(num_features)
    d.ld_raw = int(ld_raw if ld_raw is not None else 8 * cfg.num_distance_samples + 4)
    C.memmove(C.byref(d.cfg), C.byref(cfg), C.sizeof(_lib.nrc_slf_points_t))
    return d


# ----------------------------------------------------------------------------- layer 2: JAX primitives
def _jax():
    try:
        import jax  # noqa: F401
        from jax import core  # noqa: F401
        from jax.interpreters import mlir, xla  # noqa: F401
        from jax.lib import xla_client  # noqa: F401
        from jaxlib.hlo_helpers import custom_call  # noqa: F401
    except ImportError as e:  # pragma: no cover - jax is absent from the build image
        raise ImportError("nrc_jax's primitives need jax / jaxlib (the reference pins jax==0.4.16)") from e
    import jax
    from jax import core
    from jax.interpreters import mlir
    from jax.lib import xla_client
    from jaxlib.hlo_helpers import custom_call
    return jax, core, mlir, xla_client, custom_call


_registered = []


def register_targets():
    """xla_client.register_custom_call_target for every target of include/nrc_xla.h (platform CUDA)."""
    if _registered:
        return
    _, _, _, xla_client, _ = _jax()
    lib = _lib.load()
    PyCapsule_New = C.pythonapi.PyCapsule_New
    PyCapsule_New.restype, PyCapsule_New.argtypes = C.py_object, [C.c_void_p, C.c_char_p, C.c_void_p]
    for name in TARGETS:
        fn = C.cast(getattr(lib, name), C.c_void_p).value
        capsule = PyCapsule_New(fn, b"xla._CUSTOM_CALL_TARGET", None)
        xla_client.register_custom_call_target(name.encode(), capsule, platform="CUDA")
        _registered.append(name)


@functools.lru_cache(maxsize=None)
def _primitive(target, num_results):
    """A jax primitive whose lowering is one custom call to `target`.  Bound as
    prim.bind(*operands, opaque=bytes, out=((shape, dtype), ...))."""
    jax, core, mlir, _, custom_call = _jax()
    import numpy as np

    register_targets()
    prim = core.Primitive(target)
    prim.multiple_results = True
    prim.def_impl(functools.partial(jax.interpreters.xla.apply_primitive, prim))
    prim.def_abstract_eval(lambda *a, opaque, out: tuple(core.ShapedArray(s, np.dtype(d)) for s, d in out))

    def lowering(ctx, *operands, opaque, out):
        row_major = lambda aval: tuple(range(len(aval.shape) - 1, -1, -1))
        result_types = [mlir.aval_to_ir_type(a) for a in ctx.avals_out]
        call = custom_call(target, result_types=result_types, operands=list(operands), backend_config=opaque,
                           operand_layouts=[row_major(a) for a in ctx.avals_in],
                           result_layouts=[row_major(a) for a in ctx.avals_out])
        return call.results if hasattr(call, "results") else call

    mlir.register_lowering(prim, lowering, platform="gpu")
    return prim


def _call(target, operands, opaque_desc, out):
    prim = _primitive(target, len(out))
    return prim.bind(*operands, opaque=_bytes(opaque_desc), out=tuple((tuple(s), str(d)) for s, d in out))


def cuda_hash_encode(encoding):
    """HashEncoding.__call__'s per-level loop (internal/grid_utils.py:807-905) as one custom call with a custom VJP:
    returns f(x [...,3], arena [arena_floats]) -> features [..., L*F].  x already mapped by the caller exactly as the
    reference does (the bbox map is part of the kernel: pass the UNMAPPED x, as HashEncoding.__call__ receives it)."""
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp

    @jax.custom_vjp
    def f(x, arena):
        P = int(x.size // 3)
        d = pack_encode(encoding, P)
        (out,) = _call("nrc_xla_encode_fwd", (x.reshape(P, 3), arena), d,
                       [((P, d.enc.num_levels * d.enc.num_features), jnp.float32)])
        return out.reshape(x.shape[:-1] + (out.shape[-1],))

    def fwd(x, arena):
        return f(x, arena), (x, arena)

    def bwd(res, g):
        x, arena = res
        P = int(x.size // 3)
        d = pack_encode(encoding, P)
        g_arena, g_x = _call("nrc_xla_encode_bwd", (x.reshape(P, 3), arena, g.reshape(P, -1)), d,
                             [(arena.shape, jnp.float32), ((P, 3), jnp.float32)])
        return g_x.reshape(x.shape), g_arena

    f.defvjp(fwd, bwd)
    return f


def contract(x, c=1.0):
    """coord.contract(x / c) (internal/coord.py:33-69)."""
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp

    @jax.custom_vjp
    def f(x_):
        P = int(x_.size // 3)
        (z,) = _call("nrc_xla_contract_fwd", (x_.reshape(P, 3),), pack_contract(P, c), [((P, 3), jnp.float32)])
        return z.reshape(x_.shape)

    def bwd(x_, g):
        P = int(x_.size // 3)
        (gx,) = _call("nrc_xla_contract_bwd", (x_.reshape(P, 3), g.reshape(P, 3)), pack_contract(P, c), [((P, 3), jnp.float32)])
        return (gx.reshape(x_.shape),)

    f.defvjp(lambda x_: (f(x_), x_), bwd)
    return f(x)


def compute_alpha_weights(density, tdist, dirs, opaque_background=False):
    """render.compute_alpha_weights (internal/render.py:134-169): (weights, alpha, trans), VJP w.r.t. density."""
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp
    R, n = density.shape

    @jax.custom_vjp
    def f(dens):
        d = pack_ray(R, n, opaque_background=int(opaque_background))
        return tuple(_call("nrc_xla_ray_alpha_weights_fwd", (dens, tdist, dirs), d, [((R, n), jnp.float32)] * 3))

    def bwd(dens, gs):
        if opaque_background:
            raise NotImplementedError("opaque_background=True has no backward pass")
        d = pack_ray(R, n)
        (g,) = _call("nrc_xla_ray_alpha_weights_bwd", (dens, tdist, dirs) + tuple(gs), d, [((R, n), jnp.float32)])
        return (g,)

    f.defvjp(lambda dens: (f(dens), dens), bwd)
    return f(density)


def sample_intervals(u01, t, w, num_samples, anneal, padding, domain=(0.0, 1.0)):
    """The sampler's annealed resampling (internal/sampling.py:340-349 over stepfun.sample_intervals,
    internal/stepfun.py:207-250, single_jitter=True): u01 [R] = jax.random.uniform(key, [R]) stays in JAX."""
    _jax()
    import jax.numpy as jnp
    from ..stepfun import _u_base_host     # the fp32 linspace part of stepfun.sample's u (stepfun.py:158-166), host side
    R, m = w.shape
    base, max_jitter = _u_base_host(num_samples)
    u_base = jnp.asarray(base)
    d = pack_ray(R, num_samples, m=m, anneal=float(anneal), padding=float(padding), max_jitter=max_jitter,
                 dom_lo=float(domain[0]), dom_hi=float(domain[1]))
    (t_new,) = _call("nrc_xla_ray_sample_intervals", (t, w, u01.reshape(R), u_base), d, [((R, num_samples + 1), jnp.float32)])
    return t_new


def cast_rays_means(sdist, origins, directions, near, far, raydist=None):
    """coord s_to_t (internal/coord.py:223-260) + render.cast_rays means (internal/render.py:26-131):
    (tdist [R,n+1], means [R,n,3]).  raydist = (p, premult) selects the power-ladder warp."""
    _jax()
    import jax.numpy as jnp
    R, n1 = sdist.shape
    d = pack_ray(R, n1 - 1, warp_kind=0 if raydist is None else 1, p=0.0 if raydist is None else float(raydist[0]),
                 premult=1.0 if raydist is None else float(raydist[1]))
    return tuple(_call("nrc_xla_ray_cast", (sdist, origins, directions, near, far), d,
                       [((R, n1), jnp.float32), ((R, n1 - 1, 3), jnp.float32)]))


def volumetric_rendering(values, weights, tdist, bg=None, weights_no_filter=None, has_rgb=True):
    """render.volumetric_rendering (internal/render.py:172-247): (out [R,C], acc [R], dist [R,4]) with a VJP w.r.t.
    values and weights."""
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp
    R, k, Cc = values.shape
    n = tdist.shape[-1] - 1
    flags = dict(k=k, channels=Cc, has_rgb=int(has_rgb), has_bg=int(bg is not None), has_weights_nf=int(weights_no_filter is not None))
    extra = (() if weights_no_filter is None else (weights_no_filter,))
    bgs = (() if bg is None else (bg,))

    @jax.custom_vjp
    def f(v, w):
        return tuple(_call("nrc_xla_ray_composite_fwd", (v, w) + extra + (tdist,) + bgs, pack_ray(R, n, **flags),
                           [((R, Cc), jnp.float32), ((R,), jnp.float32), ((R, 4), jnp.float32)]))

    def bwd(res, gs):
        v, w = res
        outs = [((R, k, Cc), jnp.float32), ((R, k), jnp.float32)] + ([((R, n), jnp.float32)] if extra else [])
        g = _call("nrc_xla_ray_composite_bwd", (v, w) + extra + bgs + (gs[0], gs[1]), pack_ray(R, n, **flags), outs)
        return g[0], g[1]

    f.defvjp(lambda v, w: (f(v, w), (v, w)), bwd)
    return f(values, weights)


def maybe_resample(weights, gumbel, num_resample, weights_bias=0.0, logits_mult=1.0):
    """Model.maybe_resample (internal/models.py:193-292): gumbel [R,n,k] = jax.random.gumbel(key, ...) stays in JAX;
    returns (inds [R,k] int32, new weights [R,k]); gather fields with resample_gather."""
    _jax()
    import jax.numpy as jnp
    R, n = weights.shape
    d = pack_ray(R, n, k=int(num_resample), bias=float(weights_bias), mult=float(logits_mult))
    return tuple(_call("nrc_xla_ray_resample", (weights, gumbel), d,
                       [((R, num_resample), jnp.int32), ((R, num_resample), jnp.float32)]))


def resample_gather(field, inds):
    _jax()
    import jax.numpy as jnp
    R, n, Cc = field.shape
    k = inds.shape[-1]
    (out,) = _call("nrc_xla_ray_resample_gather", (field, inds), pack_ray(R, n, k=k, channels=Cc), [((R, k, Cc), jnp.float32)])
    return out


def integrate_reflect_rays(lobe_kind, wi, wo, radiance, weight, pdf, albedo, roughness, metalness, f0, occ=None,
                           rgb_max=3.4e38):
    """render_utils.integrate_reflect_rays (internal/inverse_render/render_utils.py:1102-1193) with a VJP w.r.t. the
    incoming cache radiance."""
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp
    R, S = weight.shape
    has_occ = occ is not None

    @jax.custom_vjp
    def f(rad):
        ops = (wi, wo, rad, weight, pdf) + ((occ,) if has_occ else ()) + (albedo, roughness, metalness, f0)
        outs = [((R, 3), jnp.float32), ((R, 3), jnp.float32)] + ([((R,), jnp.float32)] if has_occ else [])
        return tuple(_call("nrc_xla_ggx_integrate_fwd", ops, pack_ggx(R, S, lobe_kind, has_occ, rgb_max), outs))

    def bwd(rad, gs):
        ops = (wi, wo, rad, weight, pdf, albedo, roughness, metalness, f0, gs[0], gs[1])
        (g,) = _call("nrc_xla_ggx_integrate_bwd", ops, pack_ggx(R, S, lobe_kind, False, rgb_max), [((R, S, 3), jnp.float32)])
        return (g,)

    f.defvjp(lambda rad: (f(rad), rad), bwd)
    return f(radiance)


def slf_points(raw, origins, refdirs, cfg):
    """BaseSurfaceLightFieldMLP.predict_points + ref_warp_fn + the weight head of __call__
    (internal/surface_light_field.py:594-780,899-913) for the `surface_lf_mem` configuration, with its VJP w.r.t. the
    distance-network outputs: -> (points [P,n,3], ref_weights [P,n], s_dist [P], distances [P,n], env_rgba [P,4])."""
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp
    P, n = raw.shape[0], cfg.num_distance_samples
    outs = [((P, n, 3), jnp.float32), ((P, n), jnp.float32), ((P,), jnp.float32), ((P, n), jnp.float32), ((P, 4), jnp.float32)]

    @jax.custom_vjp
    def f(r):
        return tuple(_call("nrc_xla_slf_points_fwd", (r, origins, refdirs), pack_slf(P, cfg, r.shape[1]), outs))

    def bwd(r, gs):
        (g,) = _call("nrc_xla_slf_points_bwd", (r, origins, refdirs) + tuple(gs), pack_slf(P, cfg, r.shape[1]),
                     [((P, 8 * n + 4), jnp.float32)])
        return (g,)

    f.defvjp(lambda r: (f(r), r), bwd)
    return f(raw)


def slf_reduce(feat, weights):
    """(ref_grid_feat * ref_weights[..., None]).sum(axis=-2) (internal/surface_light_field.py:981) with its VJP."""
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp
    P, n, F = feat.shape
    cfg = _lib.nrc_slf_points_t()
    cfg.num_distance_samples = n

    @jax.custom_vjp
    def f(x, w):
        (out,) = _call("nrc_xla_slf_reduce_fwd", (x, w), pack_slf(P, cfg, num_features=F), [((P, F), jnp.float32)])
        return out

    def bwd(res, g):
        x, w = res
        return tuple(_call("nrc_xla_slf_reduce_bwd", (x, w, g), pack_slf(P, cfg, num_features=F),
                           [((P, n, F), jnp.float32), ((P, n), jnp.float32)]))

    f.defvjp(lambda x, w: (f(x, w), (x, w)), bwd)
    return f(feat, weights)


def install(grid_utils):
    """Patch the reference's internal/grid_utils.py in place: add ResampleOpMode.CUDA and make HashEncoding.__call__ take
    the custom call when it is selected (gin: HashEncoding.resample_op_mode = %ResampleOpMode.CUDA) and neither x_scale
    nor feature_filter is in use (the configurations of BASELINE.md).  Parameters keep their checkpoint names: the level
    tables are concatenated into the arena in level order at call time."""
    import enum
    jax, _, _, _, _ = _jax()
    import jax.numpy as jnp

    members = {m.name: m.value for m in grid_utils.ResampleOpMode}
    if "CUDA" not in members:
        members["CUDA"] = max(members.values()) + 1
        grid_utils.ResampleOpMode = enum.Enum("ResampleOpMode", members)
    reference_call = grid_utils.HashEncoding.__call__

    def __call__(self, x, *, x_scale=None, per_level_fn=None, **kw):
        mode = getattr(self.resample_op_mode, "name", None)
        if mode != "CUDA" or x_scale is not None or kw.get("feature_filter") is not None:
            return reference_call(self, x, x_scale=x_scale, per_level_fn=per_level_fn, **kw)
        tables = []
        for n in self.grid_sizes:
            ds = "grid" if int(n) ** 3 <= self.hash_map_size else "hash"
            tables.append(self.get_variable("params", f"{ds}_{self.get_grid_size_str(n)}").reshape(-1))
        feats = cuda_hash_encode(self)(x, jnp.concatenate(tables))          # [..., L*F], level-major like the reference
        if per_level_fn is not None:                                        # e.g. the multisample mean (shading.py:199)
            L, Fn = len(self.grid_sizes), self.num_features
            per = feats.reshape(feats.shape[:-1] + (L, Fn))
            feats = jnp.concatenate([per_level_fn(per[..., l, :]) for l in range(L)], axis=-1)
        return feats

    grid_utils.HashEncoding.__call__ = __call__
    return grid_utils
