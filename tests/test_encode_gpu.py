"""K1/K2 parity: CUDA hash-grid encoding vs the CPU oracle (tests call through the C ABI).

Bar: corner indices bit-exact (int32), features within fp32 rel 1e-5 (they are in
fact bit-identical because both sides use the same individually-rounded op order),
gradients within 1e-5 relative to the gradient's scale (atomics reorder the sums).
"""
import numpy as np
import pytest
import torch

from oracle import grid_utils as og
from neural_radiance_caching_b200 import grid_utils as ng
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu

CONFIGS = [
    # (max_grid_size, F, T, bbox_scaling)           proposal MLP_0 / MLP_1 / MLP_2 grids
    (512, 1, 524288, 1.0),
    (1024, 1, 524288, 1.0),
    (2048, 4, 524288, 1.0),
    (256, 2, 4096 * 3 + 5, 2.0),   # non power-of-two T (mod path), F=2, bbox +-2 (cornell)
    (128, 8, 2**14, ((-1.0, -2.0, -0.5), (1.0, 2.0, 1.5))),  # F=8, anisotropic bbox
]


def _points(g, n, lim):
    x = g.uniform(-lim, lim, size=(n, 3)).astype(np.float32)
    # edge cases: exactly on the bbox faces, far outside (negative voxel coordinates ->
    # int32->uint32 wrap-around in the hash), voxel centres and cell boundaries.
    extra = np.array(
        [[-1, -1, -1], [1, 1, 1], [0, 0, 0], [-1.9, 1.7, 0.3], [1.999, -1.999, 1.5],
         [1 / 1024, 3 / 1024, 5 / 1024], [0.5, 0.25, 0.125], [-2.5, 2.5, -2.5]], dtype=np.float32)
    return np.concatenate([x, extra], 0)


def _make(cfg, device, g, init_range):
    nmax, F, T, bbox = cfg
    kw = dict(hash_map_size=T, num_features=F, scale_supersample=1.0, max_grid_size=nmax, bbox_scaling=bbox)
    o = og.HashEncoding(**kw)
    n = ng.HashEncoding(**kw)
    po = o.init(g, init_range=init_range)
    assert o.param_names() == [name for (name, _, _, _) in n.level_layout]
    pn = {k: v.to(device) for k, v in po.items()}
    return o, n, po, pn


@pytest.mark.parametrize("cfg", CONFIGS)
def test_corner_indices_bit_exact(cuda_device, cfg):
    g = gen(1)
    o, n, po, pn = _make(cfg, cuda_device, g, 0.1)
    x = _points(g, 4096, 2.2)
    bbox = o.bbox
    xn = ((f32(x) - f32(bbox[0])) / f32(bbox[1] - bbox[0])).numpy()
    xd = f32(x).to(cuda_device)
    for l, (kind, N, _) in enumerate(o.layout):
        got = n.corner_indices(pn, xd, l).cpu().numpy()
        pos = xn * np.float32(N)
        if kind == "hash":
            want = og.hash_corner_indices_np(pos, o.hash_map_size)
        else:
            c = og.dense_corner_indices_np(pos, N)
            want = (c[..., 0] * (N + 2) + c[..., 1]) * (N + 2) + c[..., 2]
        assert got.dtype == np.int32
        assert np.array_equal(got, want), f"level {l} ({kind} N={N}): {np.sum(got != want)} mismatches"
        if kind == "hash":
            assert got.min() >= 0 and got.max() < o.hash_map_size


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("init_range", [None, 0.1])
def test_encode_forward(cuda_device, cfg, init_range):
    g = gen(2)
    o, n, po, pn = _make(cfg, cuda_device, g, init_range)
    x = f32(_points(g, 8192, 2.2))
    want = o(po, x)
    got = n(pn, x.to(cuda_device)).cpu()
    assert got.shape == want.shape
    assert rel_err(got, want) <= 1e-5
    # stronger: identical op order => bit-identical features
    assert torch.equal(got, want)


def test_encode_multisample_axis_and_batch_shape(cuda_device):
    g = gen(3)
    o, n, po, pn = _make(CONFIGS[2], cuda_device, g, 0.1)
    x = f32(g.uniform(-1, 1, size=(7, 5, 1, 3)))
    want = o(po, x, per_level_mean=True)
    got = n(pn, x.to(cuda_device), per_level_fn="mean").cpu()
    assert got.shape == (7, 5, 32)
    assert torch.equal(got, want)


def test_encode_empty(cuda_device):
    g = gen(4)
    o, n, po, pn = _make(CONFIGS[0], cuda_device, g, 0.1)
    got = n(pn, torch.zeros((0, 3), device=cuda_device))
    assert got.shape == (0, 6)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_encode_backward(cuda_device, cfg):
    g = gen(5)
    o, n, po, pn = _make(cfg, cuda_device, g, 0.1)
    x = f32(_points(g, 4096, 1.3))
    go = f32(g.normal(size=(x.shape[0], n.num_outputs)))
    # oracle autograd
    xo = x.clone().requires_grad_(True)
    po_r = {k: v.clone().requires_grad_(True) for k, v in po.items()}
    (o(po_r, xo) * go).sum().backward()
    # CUDA custom VJP
    xn_ = x.to(cuda_device).requires_grad_(True)
    pn_r = {k: v.clone().requires_grad_(True) for k, v in pn.items()}
    (n(pn_r, xn_) * go.to(cuda_device)).sum().backward()
    assert rel_err(xn_.grad, xo.grad) <= 1e-5
    for k in po_r:
        assert rel_err(pn_r[k].grad, po_r[k].grad) <= 1e-5, k


def test_dense_level_matches_explicit_trilinear(cuda_device):
    """a dense level equals half-pixel-centred trilinear interpolation of
    an explicit N^3 grid with zero padding (checked against torch grid_sample-free math)."""
    g = gen(6)
    N, F = 16, 4
    grid = f32(g.normal(size=(N, N, N, F)))
    coords = f32(g.uniform(-1.0, N + 1.0, size=(2048, 3)))
    want = og.trilerp(grid, coords, "grid")
    got = ng.trilerp(grid.to(cuda_device), coords.to(cuda_device), "grid").cpu()
    assert torch.equal(got, want)
    # voxel centres reproduce the stored values exactly
    centres = f32(np.stack(np.meshgrid(*[np.arange(N) + 0.5] * 3, indexing="ij"), -1).reshape(-1, 3))
    got_c = ng.trilerp(grid.to(cuda_device), centres.to(cuda_device), "grid").cpu()
    assert torch.equal(got_c, grid.reshape(-1, F))


def test_zero_ranges(cuda_device):
    """nrc_zero_ranges: several ranges of one buffer cleared in one launch, plain and evict-first stores; everything
    outside the ranges untouched; argument checks."""
    import ctypes as C
    from neural_radiance_caching_b200 import _lib
    lib = _lib.load()
    n = 1 << 20
    ranges = [(0, 4096), (8192, 8192), (100000, 700000), (n - 8, n)]
    for streaming in (0, 1):
        buf = torch.ones(n, device=cuda_device)
        lo = (C.c_int64 * len(ranges))(*[a for a, _ in ranges])
        hi = (C.c_int64 * len(ranges))(*[b for _, b in ranges])
        _lib.call("nrc_zero_ranges", _lib.stream_ptr(), _lib.ptr(buf), lo, hi, len(ranges), streaming)
        want = torch.ones(n)
        for a, b in ranges:
            want[a:b] = 0.0
        assert torch.equal(buf.cpu(), want)
    buf = torch.ones(64, device=cuda_device)
    one = lambda a, b: ((C.c_int64 * 1)(a), (C.c_int64 * 1)(b))
    assert lib.nrc_zero_ranges(_lib.stream_ptr(), _lib.ptr(buf), *one(2, 8), 1, 0) == -1      # not a multiple of 4 floats
    assert lib.nrc_zero_ranges(_lib.stream_ptr(), _lib.ptr(buf), *one(8, 4), 1, 0) == -1      # hi < lo
    assert lib.nrc_zero_ranges(_lib.stream_ptr(), _lib.ptr(buf), None, None, 0, 0) == 0
    assert lib.nrc_zero_ranges(_lib.stream_ptr(), _lib.ptr(buf), *one(0, 4), 9, 0) == -1
    assert float(buf.sum()) == 64.0
