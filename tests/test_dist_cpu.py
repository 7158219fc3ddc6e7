"""World-size-2 gloo tests of the data-parallel host logic (no GPU): ray sharding, the single
flat-arena gradient all-reduce (mean), and the row-band tile gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from neural_radiance_caching_b200 import dist as ndist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # gradient arena: each rank holds its shard's gradient; the mean must equal the full-batch mean
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        ndist.allreduce_mean_(g)
        ok_grad = torch.allclose(g, torch.arange(1000, dtype=torch.float32) * (1 + world) / 2)
        # ray sharding covers the batch exactly once
        a, b = ndist.shard_rays(1024, rank, world)
        ok_shard = (b - a) == 1024 // world and a == rank * (1024 // world)
        # tile gather: 7 rows over 2 ranks (4 + 3), every rank ends with the full image
        H, W = 7, 5
        full = torch.arange(H * W * 3, dtype=torch.float32).reshape(H, W, 3)
        r0, r1 = ndist.row_bands(H, world)[rank]
        img = ndist.gather_tiles(full[r0:r1].clone(), H)
        ok_tiles = torch.equal(img, full)
        q.put((rank, bool(ok_grad), bool(ok_shard), bool(ok_tiles)))
    finally:
        tdist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(all(r[1:]) for r in res), res


def test_row_bands_cover_image():
    for H in (800, 7, 1):
        for ws in (1, 2, 4, 8):
            bands = ndist.row_bands(H, ws)
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(b0[1] == b1[0] for b0, b1 in zip(bands, bands[1:]))
            sizes = [b - a for a, b in bands]
            assert max(sizes) - min(sizes) <= 1


def test_welford_matches_batch_statistics():
    from neural_radiance_caching_b200.render_image import Welford
    g = np.random.Generator(np.random.PCG64(3))
    xs = [torch.from_numpy(g.normal(size=(50, 3)).astype(np.float32)) for _ in range(7)]
    w = Welford()
    for x in xs:
        w.update(x)
    st = torch.stack(xs)
    assert torch.allclose(w.mean, st.mean(0), atol=1e-6)
    assert torch.allclose(w.variance(), st.var(0, unbiased=True), atol=1e-5)


def test_render_image_chunking_and_bands_cpu():
    """Host logic of render_image on CPU tensors with a fake per-chunk renderer: edge padding, chunk
    scatter, and equality with a single-pass evaluation."""
    from neural_radiance_caching_b200 import render_image as ri
    H, W, focal = 13, 9, 20.0
    c2w = ri.orbit_camera()
    calls = []

    def fake(rays, rep):
        calls.append(rays["origins"].shape[0])
        return dict(rgb=rays["viewdirs"] * 0.5 + 0.5, depth=rays["directions"][:, :1])

    img = ri.render_image(fake, H, W, focal, c2w, torch.device("cpu"), chunk=32)
    assert all(c == 32 for c in calls) and len(calls) == -(-H * W // 32)
    full = ri.pinhole_rays(H, W, focal, c2w, torch.device("cpu"))
    assert torch.allclose(img["rgb"].reshape(-1, 3), full["viewdirs"] * 0.5 + 0.5)
    assert img["rgb"].shape == (H, W, 3) and img["depth"].shape == (H, W, 1)
    # unit view directions, non-unit ray directions whose z-component in camera space is -1
    assert torch.allclose(torch.linalg.norm(full["viewdirs"], dim=-1), torch.ones(H * W), atol=1e-6)
