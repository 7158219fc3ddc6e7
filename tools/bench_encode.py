"""Micro-benchmark of the encode kernels (development aid, not the contract bench)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_radiance_caching_b200 import grid_utils as ng

dev = torch.device("cuda:0")
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (nmax, F, P) in [(512, 1, 65536), (1024, 1, 65536), (2048, 4, 32768), (512, 1, 2097152), (1024, 1, 2097152),
                     (2048, 4, 1048576)]:
    enc = ng.HashEncoding(hash_map_size=524288, num_features=F, scale_supersample=1.0, max_grid_size=nmax,
                          bbox_scaling=1.0)
    params, arena = enc.init(dev, init_range=0.1)
    x = (torch.rand(P, 3, device=dev) * 2 - 1)
    L = len(enc.grid_sizes)
    with torch.no_grad():
        ms = timeit(lambda: enc(params, x))
    alg = P * (12 + 8 * F * 4 * L + 4 * L * F)
    print(f"fwd nmax={nmax} F={F} P={P}: {ms*1e3:.1f} us  {P/ms/1e6:.2f} Gpts/s  alg {alg/ms/1e6:.1f} GB/s "
          f"({alg/ms/1e6/PEAK*100:.1f}% of measured HBM)")
    xg = x.clone().requires_grad_(True)
    pg = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = enc(pg, xg)
    go = torch.randn_like(out)
    def bwd():
        torch.autograd.grad(out, list(pg.values()), go, retain_graph=True)
    ms = timeit(bwd, iters=10, warm=3)
    print(f"bwd(tables, incl. zero-fill) nmax={nmax} F={F} P={P}: {ms*1e3:.1f} us  {P/ms/1e6:.2f} Gpts/s")
