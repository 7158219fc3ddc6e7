"""Host-side mirror of the reference's internal/grid_utils.py for the hot path.

Same names, argument meaning and error behaviour as the reference
(`HashEncoding`, `trilerp`, `ResampleOpMode`), with one new op mode,
`ResampleOpMode.CUDA`, whose bodies are the sm_100a kernels behind the C ABI
(include/nrc_b200.h).  The reference's JAX modes are not re-implemented here:
there is no CPU or non-CUDA fallback.

Parameters keep the reference's checkpoint layout and names
(`grid_0016` [N,N,N,F] ... `hash_2048` [T,F], internal/grid_utils.py:835-852);
`HashEncoding.init` lays them out back to back in ONE fp32 arena so the
gradient all-reduce (dist.py) is a single NCCL call.
"""
import ctypes as C
import enum

import numpy as np
import torch

from . import _lib


class ResampleOpMode(enum.Enum):
    """internal/grid_utils.py:651-660, plus the CUDA mode this package adds."""

    DEFAULT_JAX = enum.auto()
    SIMPLEX_JAX = enum.auto()
    CUDA = enum.auto()


def _require_cuda_mode(op_mode):
    if op_mode != ResampleOpMode.CUDA:
        raise NotImplementedError(
            f"{op_mode} lives in the reference's JAX code; this package implements ResampleOpMode.CUDA only."
        )


class _EncodeFn(torch.autograd.Function):
    """custom_vjp analogue: forward = nrc_encode_fwd, backward = nrc_encode_bwd.

    `tables` is either the per-level tensors or ONE flat arena holding all levels back to
    back (then the table gradient is one arena-shaped buffer: no per-level zero-fill/add)."""

    @staticmethod
    def forward(ctx, enc, x, use_arena, *tables):
        x2 = x.reshape(-1, 3).contiguous()
        out = torch.empty((x2.shape[0], enc.num_outputs), device=x.device, dtype=torch.float32)
        levels = enc.tables(enc.views(tables[0])) if use_arena else tables
        desc = enc._descriptor(levels, None)
        _lib.call("nrc_encode_fwd", _lib.stream_ptr(), C.byref(desc), _lib.ptr(x2), x2.shape[0], _lib.ptr(out))
        ctx.enc = enc
        ctx.use_arena = use_arena
        ctx.save_for_backward(x2, *tables)
        ctx.x_shape = x.shape
        return out.reshape(x.shape[:-1] + (enc.num_outputs,))

    @staticmethod
    def backward(ctx, g):
        enc = ctx.enc
        x2, *tables = ctx.saved_tensors
        g2 = g.reshape(-1, enc.num_outputs).contiguous()
        need_x = ctx.needs_input_grad[1]
        need_t = any(ctx.needs_input_grad[3:])
        sunk = False
        if need_t and ctx.use_arena and _lib.grad_sink(tables[0]) is not None:
            grads, sunk = [_lib.grad_sink(tables[0])], True
        else:
            grads = [torch.zeros_like(t) for t in tables] if need_t else None
        if ctx.use_arena:
            levels = enc.tables(enc.views(tables[0]))
            glevels = enc.tables(enc.views(grads[0])) if need_t else None
        else:
            levels, glevels = tables, grads
        g_x = torch.empty_like(x2) if need_x else None
        desc = enc._descriptor(levels, glevels)
        _lib.call("nrc_encode_bwd", _lib.stream_ptr(), C.byref(desc), _lib.ptr(x2), _lib.ptr(g2), x2.shape[0],
                  _lib.ptr(g_x))
        gx = g_x.reshape(ctx.x_shape) if need_x else None
        if sunk:
            grads = None   # accumulated into the registered sink
        return (None, gx, None) + (tuple(grads) if grads is not None else (None,) * len(tables))


class HashEncoding:
    """Multiresolution grid/hash encoding (internal/grid_utils.py:738-905).

    Field names and defaults follow the reference's dataclass fields (:741-758).
    """

    def __init__(
        self,
        hash_map_size=2**19,
        num_features=2,
        scale_supersample=2.0,
        min_grid_size=16,
        max_grid_size=2048,
        hash_init_range=1e-4,
        precondition_scaling=10.0,
        bbox_scaling=2.0,
        resample_op_mode=ResampleOpMode.CUDA,
        feature_aggregator="concatenate",
        append_scale=False,
    ):
        self.hash_map_size = int(hash_map_size)
        self.num_features = int(num_features)
        self.scale_supersample = scale_supersample
        self.min_grid_size = min_grid_size
        self.max_grid_size = max_grid_size
        self.hash_init_range = hash_init_range
        self.precondition_scaling = precondition_scaling
        self.bbox_scaling = bbox_scaling
        self.resample_op_mode = resample_op_mode
        self.feature_aggregator = feature_aggregator
        self.append_scale = append_scale
        if self.num_features not in (1, 2, 4, 8):
            raise ValueError(f"num_features must be 1, 2, 4 or 8 for the CUDA op mode, got {num_features}.")
        if len(self.grid_sizes) > _lib.NRC_MAX_LEVELS:
            raise ValueError(f"at most {_lib.NRC_MAX_LEVELS} levels are supported.")

    # -- reference properties -------------------------------------------------
    @property
    def grid_sizes(self):
        """internal/grid_utils.py:772-794."""
        desired_num_scales = 1 + self.scale_supersample * np.log2(self.max_grid_size / self.min_grid_size)
        num_scales = int(np.round(desired_num_scales))
        if np.abs(desired_num_scales - num_scales) > 1e-4:
            raise ValueError(
                "grid scale parameters are ("
                + f"min_grid_size={self.min_grid_size}, "
                + f"max_grid_size={self.max_grid_size}, "
                + f"scale_supersample={self.scale_supersample}), "
                + f"which yields a non-integer number of scales {desired_num_scales}."
            )
        return np.round(np.geomspace(self.min_grid_size, self.max_grid_size, num_scales)).astype(np.int32)

    def get_grid_size_str(self, grid_size):
        """internal/grid_utils.py:796-798."""
        return str(grid_size).zfill(len(str(np.max(self.grid_sizes))))

    @property
    def bbox(self):
        """internal/grid_utils.py:800-805."""
        bbox = self.bbox_scaling
        if isinstance(bbox, float):
            bbox = ((-bbox,) * 3, (bbox,) * 3)
        return np.array(bbox)

    # -- layout ---------------------------------------------------------------
    @property
    def level_layout(self):
        """[(param_name, datastructure, N, shape)] (internal/grid_utils.py:834-852)."""
        out = []
        for n in self.grid_sizes:
            n = int(n)
            if n**3 <= self.hash_map_size:
                ds, shape = "grid", (n, n, n, self.num_features)
            else:
                ds, shape = "hash", (self.hash_map_size, self.num_features)
            out.append((f"{ds}_{self.get_grid_size_str(n)}", ds, n, shape))
        return out

    @property
    def num_outputs(self):
        return len(self.grid_sizes) * self.num_features

    @property
    def num_params(self):
        return sum(int(np.prod(s)) for (_, _, _, s) in self.level_layout)

    def init(self, device, generator=None, init_range=None, arena=None):
        """Allocate the level tables as views of one contiguous arena.

        uniform(+-hash_init_range/precondition_scaling) like :844-850 unless
        `init_range` is given.  Returns (params: dict name -> tensor, arena).
        """
        maxval = self.hash_init_range / self.precondition_scaling if init_range is None else init_range
        if arena is None:
            arena = torch.empty(self.num_params, device=device, dtype=torch.float32)
            arena.uniform_(-maxval, maxval, generator=generator)
        return self.views(arena), arena

    def views(self, arena):
        """Per-level checkpoint-layout views into a flat fp32 arena."""
        params, off = {}, 0
        for name, _, _, shape in self.level_layout:
            n = int(np.prod(shape))
            params[name] = arena[off:off + n].view(shape)
            off += n
        return params

    def _descriptor(self, tables, grads):
        layout = self.level_layout
        if len(tables) != len(layout):
            raise ValueError(f"expected {len(layout)} level tables, got {len(tables)}")
        d = _lib.nrc_encoding_t()
        d.num_levels = len(layout)
        d.num_features = self.num_features
        bbox = self.bbox
        b0 = bbox[0].astype(np.float32)
        b1 = bbox[1].astype(np.float32)
        span = (bbox[1] - bbox[0]).astype(np.float32)
        for a in range(3):
            d.bbox_min[a] = float(b0[a])
            d.bbox_max[a] = float(b1[a])
            d.bbox_span[a] = float(span[a])
        d.precondition_scaling = float(self.precondition_scaling)
        for l, ((name, ds, n, shape), t) in enumerate(zip(layout, tables)):
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"level {name}: expected shape {shape}, got {tuple(t.shape)}")
            lv = d.levels[l]
            lv.d_table = _lib.ptr(t).value
            g = grads[l] if grads is not None else None
            lv.d_grad = _lib.ptr(g).value if g is not None else None
            lv.grid_size = n
            lv.is_hash = 1 if ds == "hash" else 0
            lv.table_size = int(np.prod(shape[:-1]))
        return d

    def tables(self, params):
        return [params[name] for (name, _, _, _) in self.level_layout]

    def apply(self, params, x):
        """Encode through the custom VJP; `params["_arena"]` (flat buffer the level tensors are
        views of) selects the single-gradient-buffer path."""
        arena = params.get("_arena") if isinstance(params, dict) else None
        if arena is not None:
            return _EncodeFn.apply(self, x, True, arena)
        return _EncodeFn.apply(self, x, False, *self.tables(params))

    # -- reference call signature --------------------------------------------
    def __call__(
        self,
        params,
        x,
        *,
        x_scale=None,
        per_level_fn=None,
        train=True,
        train_frac=1.0,
        feature_filter=None,
        feature_filter_size=None,
    ):
        """HashEncoding.__call__ (internal/grid_utils.py:807-905).

        `params` is the module's parameter dict (Flax `self.param` in the reference).
        Supported: x_scale=None, feature_filter=None, 'concatenate' aggregation and
        per_level_fn in {None, average_across_multisamples}; the other settings are
        inactive under every BASELINE config and raise.
        """
        _require_cuda_mode(self.resample_op_mode)
        if x_scale is not None or feature_filter is not None or self.append_scale:
            raise NotImplementedError("x_scale / feature_filter / append_scale are outside the CUDA path's scope.")
        if self.feature_aggregator != "concatenate":
            raise ValueError(f"Aggregator {self.feature_aggregator} not implemented.")
        if per_level_fn is not None:
            # math.average_across_multisamples over the multisample axis (-2).  The
            # configs use unscented basis 'mean' => exactly one control point.
            if x.shape[-2] != 1:
                raise NotImplementedError("per_level_fn with more than one multisample is outside the path's scope.")
            x = x[..., 0, :]
        return self.apply(params, x)

    def corner_indices(self, params, x, level):
        """Parity aid: integer corner indices of one level (nrc_encode_indices)."""
        x2 = x.reshape(-1, 3).contiguous()
        idx = torch.empty((x2.shape[0], 8), device=x.device, dtype=torch.int32)
        desc = self._descriptor(self.tables(params), None)
        _lib.call("nrc_encode_indices", _lib.stream_ptr(), C.byref(desc), level, _lib.ptr(x2), x2.shape[0],
                  _lib.ptr(idx))
        return idx


def trilerp(values, coordinates, datastructure, op_mode=ResampleOpMode.CUDA):
    """Sample from a hash or 3D voxel grid (internal/grid_utils.py:679-726).

    `coordinates` are in voxel units (x * N), exactly like the reference.  Implemented
    as a one-level encoding whose bbox maps voxel units back to [0,1]: exact for
    power-of-two N (every N in the configs).
    """
    _require_cuda_mode(op_mode)
    if datastructure == "hash":
        raise NotImplementedError("stand-alone hash trilerp needs the level's N; use HashEncoding.")
    if datastructure != "grid":
        raise ValueError(f"datastructure must be either `grid` or `hash` but `{datastructure}` was given.")
    n = values.shape[0]
    if n & (n - 1):
        raise NotImplementedError("stand-alone trilerp supports power-of-two grids only.")
    enc = HashEncoding(hash_map_size=n**3, num_features=values.shape[-1], scale_supersample=1.0,
                       min_grid_size=n, max_grid_size=n, precondition_scaling=1.0,
                       bbox_scaling=((0.0, 0.0, 0.0), (float(n),) * 3))
    return _EncodeFn.apply(enc, coordinates, False, values)
