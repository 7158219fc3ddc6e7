"""Oracle restatement of internal/grid_utils.py (TEST INFRASTRUCTURE ONLY).

Integer paths (corner coordinates, spatial hash, dense indices) are restated
in NumPy uint32/int32; the float paths in PyTorch-CPU fp32 with the reference's
operation order, so torch.autograd supplies d/dtable and d/dx.
"""
import numpy as np
import torch

PI_2 = 19349663  # internal/grid_utils.py:102
PI_3 = 83492791  # internal/grid_utils.py:103


def grid_sizes(min_grid_size=16, max_grid_size=2048, scale_supersample=1.0):
    """internal/grid_utils.py:772-794 (HashEncoding.grid_sizes)."""
    desired = 1 + scale_supersample * np.log2(max_grid_size / min_grid_size)
    num_scales = int(np.round(desired))
    if np.abs(desired - num_scales) > 1e-4:
        raise ValueError(
            f"grid scale parameters (min_grid_size={min_grid_size}, max_grid_size={max_grid_size}, "
            f"scale_supersample={scale_supersample}) yield a non-integer number of scales {desired}."
        )
    return np.round(np.geomspace(min_grid_size, max_grid_size, num_scales)).astype(np.int32)


def level_layout(hash_map_size, num_features, sizes):
    """internal/grid_utils.py:834-852: dense [N,N,N,F] if N^3 <= T else hash [T,F]."""
    out = []
    for n in sizes:
        n = int(n)
        if n**3 <= hash_map_size:
            out.append(("grid", n, (n, n, n, num_features)))
        else:
            out.append(("hash", n, (hash_map_size, num_features)))
    return out


def param_name(kind, n, sizes):
    """internal/grid_utils.py:796-798,851: f'{datastructure}_{zero-padded N}'."""
    width = len(str(int(np.max(sizes))))
    return f"{kind}_{str(int(n)).zfill(width)}"


# --------------------------------------------------------------------------
# Integer oracle (NumPy).
# --------------------------------------------------------------------------
def hash_corner_indices_np(locations, table_size):
    """Bit-exact corner hash indices; internal/grid_utils.py:61-77,98-111.

    locations: [P,3] float32 (already multiplied by N).  Returns int32 [P,8]
    in the reference corner order (fff, ffc, fcf, fcc, cff, cfc, ccf, ccc).
    """
    loc = locations.astype(np.float32) - np.float32(0.5)
    fl = np.floor(loc)
    ce = fl + np.float32(1.0)
    out = np.empty((loc.shape[0], 8), np.int32)
    k = 0
    for cx in (fl[:, 0], ce[:, 0]):
        for cy in (fl[:, 1], ce[:, 1]):
            for cz in (fl[:, 2], ce[:, 2]):
                ux = cx.astype(np.int32).astype(np.uint32)
                uy = cy.astype(np.int32).astype(np.uint32)
                uz = cz.astype(np.int32).astype(np.uint32)
                h = ux ^ ((uy * np.uint32(PI_2)) ^ (uz * np.uint32(PI_3)))
                out[:, k] = (h % np.uint32(table_size)).astype(np.int32)
                k += 1
    return out


def dense_corner_indices_np(coords, n):
    """Padded-grid corner indices of the dense path; grid_utils.py:703-715,384-438.

    coords: [P,3] float32 = x*N (before trilerp's -0.5).  Returns int32 [P,8,3]
    of (ix,iy,iz) indices into the zero-padded (N+2)^3 grid, reference corner
    order (which iterates the *flipped* axes: z outermost ... x innermost).
    """
    c = coords.astype(np.float32) - np.float32(0.5)
    loc = c[:, ::-1] + np.float32(1.0)  # flipped: (z, y, x)
    fl = np.floor(loc)
    ce = fl + np.float32(1.0)
    out = np.empty((c.shape[0], 8, 3), np.int32)
    k = 0
    for a0 in (fl[:, 0], ce[:, 0]):  # z
        for a1 in (fl[:, 1], ce[:, 1]):  # y
            for a2 in (fl[:, 2], ce[:, 2]):  # x
                iz = np.clip(a0.astype(np.int32), 0, n + 1)
                iy = np.clip(a1.astype(np.int32), 0, n + 1)
                ix = np.clip(a2.astype(np.int32), 0, n + 1)
                out[:, k, 0], out[:, k, 1], out[:, k, 2] = ix, iy, iz
                k += 1
    return out


# --------------------------------------------------------------------------
# Float oracle (torch CPU fp32).
# --------------------------------------------------------------------------
def _u32(t):
    return t.to(torch.int64) & 0xFFFFFFFF


def hash_resample_3d(data, locations):
    """internal/grid_utils.py:41-121 (TRILINEAR, half_pixel_center=True)."""
    assert data.dim() == 2
    loc = locations - 0.5
    fl = torch.floor(loc)
    ce = fl + 1.0
    cw = loc - fl
    fw = 1.0 - cw
    T = data.shape[0]
    out = None
    for bx in (0, 1):
        for by in (0, 1):
            for bz in (0, 1):
                px = (ce if bx else fl)[..., 0]
                py = (ce if by else fl)[..., 1]
                pz = (ce if bz else fl)[..., 2]
                w = ((cw if bx else fw)[..., 0] * (cw if by else fw)[..., 1]) * (cw if bz else fw)[..., 2]
                ux = _u32(px.detach().to(torch.int32))
                uy = _u32(py.detach().to(torch.int32))
                uz = _u32(pz.detach().to(torch.int32))
                h = ux ^ (((uy * PI_2) & 0xFFFFFFFF) ^ ((uz * PI_3) & 0xFFFFFFFF))
                idx = h % T
                g = data[idx] * w[..., None]
                out = g if out is None else out + g
    return out


def resample_3d(data, locations):
    """internal/grid_utils.py:352-445 with CONSTANT_OUTSIDE, coordinate_order='xyz',
    half_pixel_center=False.  `locations` are the flipped (z,y,x) coordinates."""
    n0, n1, n2 = data.shape[:3]
    padded = torch.nn.functional.pad(data, (0, 0, 1, 1, 1, 1, 1, 1))  # :384-389
    loc = locations + 1.0  # :390
    fl = torch.floor(loc)
    ce = fl + 1.0
    cw = loc - fl
    fw = 1.0 - cw
    # max_indices = flip(shape[:3]) - 1 for 'xyz' (:433-435)
    max_idx = [padded.shape[2] - 1, padded.shape[1] - 1, padded.shape[0] - 1]
    out = torch.zeros(locations.shape[:-1] + (data.shape[-1],), dtype=data.dtype)
    for b0 in (0, 1):
        for b1 in (0, 1):
            for b2 in (0, 1):
                p0 = (ce if b0 else fl)[..., 0].detach().to(torch.int32).long().clamp(0, max_idx[0])
                p1 = (ce if b1 else fl)[..., 1].detach().to(torch.int32).long().clamp(0, max_idx[1])
                p2 = (ce if b2 else fl)[..., 2].detach().to(torch.int32).long().clamp(0, max_idx[2])
                w = ((cw if b0 else fw)[..., 0] * (cw if b1 else fw)[..., 1]) * (cw if b2 else fw)[..., 2]
                # gather_volume 'xyz': data[z_coord, y_coord, x_coord] with
                # x_coord = loc[...,0], z_coord = loc[...,2]   (:328-349)
                g = padded[p2, p1, p0]
                out = out + g * w[..., None]
    return out


def trilerp(values, coordinates, datastructure):
    """internal/grid_utils.py:679-726 (op_mode DEFAULT_JAX)."""
    if datastructure == "hash":
        flat = coordinates.reshape(-1, coordinates.shape[-1])
        res = hash_resample_3d(values, flat)
    elif datastructure == "grid":
        c = torch.flip(coordinates - 0.5, dims=(-1,))
        flat = c.reshape(-1, c.shape[-1])
        res = resample_3d(values, flat)
    else:
        raise ValueError(f"datastructure must be either `grid` or `hash` but `{datastructure}` was given.")
    return res.reshape(coordinates.shape[:-1] + (values.shape[-1],))


class HashEncoding:
    """internal/grid_utils.py:738-905 with x_scale=None, feature_filter=None,
    feature_aggregator='concatenate' (the BASELINE configs)."""

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
    ):
        self.hash_map_size = hash_map_size
        self.num_features = num_features
        self.scale_supersample = scale_supersample
        self.min_grid_size = min_grid_size
        self.max_grid_size = max_grid_size
        self.hash_init_range = hash_init_range
        self.precondition_scaling = precondition_scaling
        self.bbox_scaling = bbox_scaling
        self.grid_sizes = grid_sizes(min_grid_size, max_grid_size, scale_supersample)
        self.layout = level_layout(hash_map_size, num_features, self.grid_sizes)

    @property
    def bbox(self):
        b = self.bbox_scaling
        if isinstance(b, float):
            b = ((-b,) * 3, (b,) * 3)
        return np.array(b)  # float64, as in the reference (:800-805)

    def param_names(self):
        return [param_name(k, n, self.grid_sizes) for (k, n, _) in self.layout]

    def init(self, gen, init_range=None):
        """Random tables: uniform(+-hash_init_range/precondition_scaling) (:844-850)
        unless `init_range` overrides (the "trained-like" stress distribution)."""
        maxval = self.hash_init_range / self.precondition_scaling if init_range is None else init_range
        params = {}
        for name, (_, _, shape) in zip(self.param_names(), self.layout):
            params[name] = torch.from_numpy(gen.uniform(-maxval, maxval, size=shape).astype(np.float32))
        return params

    def __call__(self, params, x, per_level_mean=False):
        """x: [...,3] (or [...,S,3] with per_level_mean: mean over the multisample
        axis, math.average_across_multisamples, internal/math.py:471-473)."""
        bbox = self.bbox
        b0 = torch.tensor(bbox[0].astype(np.float32))
        span = torch.tensor((bbox[1] - bbox[0]).astype(np.float32))
        x = (x - b0) / span  # :820
        feats = []
        for name, (kind, n, _) in zip(self.param_names(), self.layout):
            f = trilerp(params[name], x * float(n), kind)  # :863
            if per_level_mean:
                f = torch.mean(f, dim=-2)
            feats.append(f)
        features = torch.cat(feats, dim=-1)
        return features * self.precondition_scaling  # :903
