"""Size-independent properties at BASELINE.json's full sizes, and the edge cases of the C ABI (empty and
ragged inputs).  The oracle finishes in seconds only at small sizes, so at 1024 rays x (64,64,32) samples
the CUDA path is checked through invariants of the domain instead."""
import ctypes as C

import numpy as np
import pytest
import torch

from neural_radiance_caching_b200 import _lib, engine, grid_utils as ng, mlp_chain as mc, render as nrender, workload
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu


def _batch(dev, R):
    g = np.random.Generator(np.random.PCG64(workload.SEED + 3))
    rn = workload.make_rays_np(g, R)
    u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
    tgt = g.uniform(size=(R, 3)).astype(np.float32)
    return workload.unpack_rays(torch.from_numpy(workload.pack_rays(rn, u, tgt)).to(dev))


def test_config2_full_size_invariants(cuda_device):
    """One full config-2 step (1024 rays): fenceposts sorted inside [0,1], weights form a sub-probability,
    rendering bounded, gradients finite and reproducible (checksum of the gradient arena over two runs)."""
    R = 1024
    step = workload.CacheTrainStep(cuda_device, bf16=True)
    rays, u01, tgt = _batch(cuda_device, R)
    loss1 = float(step.step(rays, u01, tgt))
    sum1 = step.flat_grad.double().sum().item()
    abs1 = step.flat_grad.double().abs().sum().item()
    for lv in step.engine.last["levels"]:
        sd = lv["sdist"]
        assert bool((sd[:, 1:] >= sd[:, :-1]).all()) and float(sd.min()) >= 0.0 and float(sd.max()) <= 1.0
        w = lv["weights"]
        assert float(w.min()) >= 0.0 and float(w.sum(-1).max()) <= 1.0 + 1e-5
        assert bool((lv["density"] >= 0).all())
    rgb, acc = step.engine.last["rgb"], step.engine.last["acc"]
    assert bool(torch.isfinite(rgb).all()) and float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert np.isfinite(loss1) and loss1 > 0 and bool(torch.isfinite(step.flat_grad).all()) and abs1 > 0
    loss2 = float(step.step(rays, u01, tgt))
    assert abs(loss2 - loss1) <= 1e-6 * abs(loss1)
    assert abs(step.flat_grad.double().sum().item() - sum1) <= 1e-4 * abs1     # atomics reorder the sums only


def test_compositing_is_linear_and_encode_is_additive_in_tables(cuda_device):
    g = gen(800)
    R, n = 1024, 32
    dev = cuda_device
    w = torch.rand((R, n), device=dev) / n
    t = torch.sort(torch.rand((R, n + 1), device=dev) * 4 + 2, dim=-1).values
    a, b = torch.rand((R, n, 3), device=dev), torch.rand((R, n, 3), device=dev)
    ra = nrender.volumetric_rendering(a, w, w, t, 0.0, True)["rgb"]
    rb = nrender.volumetric_rendering(b, w, w, t, 0.0, True)["rgb"]
    rab = nrender.volumetric_rendering(2.0 * a + 3.0 * b, w, w, t, 0.0, True)["rgb"]
    assert rel_err(rab, 2.0 * ra + 3.0 * rb) <= 1e-5
    enc = ng.HashEncoding(hash_map_size=524288, num_features=4, scale_supersample=1.0, max_grid_size=2048, bbox_scaling=1.0)
    p1, a1 = enc.init(dev, init_range=0.1)
    p2, a2 = enc.init(dev, init_range=0.1)
    x = torch.rand((32768, 3), device=dev) * 2.4 - 1.2
    f12 = enc(enc.views(a1 + a2), x)
    assert rel_err(f12, enc(p1, x) + enc(p2, x)) <= 1e-5       # trilinear interpolation is linear in the table


def test_empty_and_ragged_inputs(cuda_device):
    dev = cuda_device
    enc = ng.HashEncoding(hash_map_size=524288, num_features=4, scale_supersample=1.0, max_grid_size=2048, bbox_scaling=1.0)
    p, _ = enc.init(dev, init_range=0.1)
    assert enc(p, torch.empty((0, 3), device=dev)).shape == (0, 32)                     # empty batch
    assert enc(p, torch.rand((1, 3), device=dev)).shape == (1, 32)                      # single point
    w, a, t = nrender.compute_alpha_weights(torch.empty((0, 8), device=dev), torch.empty((0, 9), device=dev),
                                            torch.empty((0, 3), device=dev))
    assert w.shape == (0, 8)
    # ragged tile counts on the tensor-core chain: P = 1, 127, 129, 257 (partial and odd numbers of 128-row tiles)
    spec = mc.ChainSpec(in_widths=[64, 32], hidden=[("l0", 64, False)], heads=[[("out", 3)]])
    g = gen(810)
    lim = lambda fi: float(np.sqrt(6.0 / fi))
    params = {"l0": {"kernel": (torch.rand((96, 64), device=dev) * 2 - 1) * lim(96), "bias": torch.zeros(64, device=dev)},
              "out": {"kernel": (torch.rand((64, 3), device=dev) * 2 - 1) * lim(64), "bias": torch.zeros(3, device=dev)}}
    big = [torch.randn((257, 64), device=dev), torch.randn((257, 32), device=dev)]
    with torch.no_grad():
        (ref,) = mc.apply(spec, params, big)
        for P in (1, 127, 129, 257):
            (got,) = mc.apply(spec, params, [s[:P].contiguous() for s in big])
            assert got.shape == (P, 3)
            assert torch.equal(got, ref[:P])           # a row's result does not depend on the batch around it
    # invalid arguments are reported as status codes, not crashes
    st = _lib.load().nrc_encode_fwd(None, None, None, 5, None)
    assert st == -1


def test_peer_allreduce_kernel_single_rank(cuda_device):
    """nrc_allreduce_mean_peer with world = 1 (the arena is its own only peer): the mean over one rank is the identity on
    the bucket and nothing outside the bucket may change; bad arguments come back as status codes.  The N >= 2 check
    against NCCL (bit-identical at N = 2) is tools/test_allreduce.py (needs several GPUs)."""
    import ctypes as C
    from neural_radiance_caching_b200 import _lib

    n = 1 << 20
    g = torch.Generator(device=cuda_device).manual_seed(5)
    buf = torch.randn(n, device=cuda_device, generator=g)
    ref = buf.clone()
    peers = (C.c_void_p * 1)(buf.data_ptr())
    lo, cnt = 4096, n - 8192
    _lib.call("nrc_allreduce_mean_peer", _lib.stream_ptr(), C.cast(peers, C.c_void_p), lo, cnt, 0, 1, 0)
    torch.cuda.synchronize()
    assert torch.equal(buf, ref)
    lib = _lib.load()
    assert lib.nrc_allreduce_mean_peer(None, C.cast(peers, C.c_void_p), 2, cnt, 0, 1, 0) == -1      # offset % 4
    assert lib.nrc_allreduce_mean_peer(None, C.cast(peers, C.c_void_p), 0, cnt, 1, 1, 0) == -1      # rank >= world
    assert lib.nrc_allreduce_mean_multicast(None, None, 0, cnt, 0, 1, 0) == -1
