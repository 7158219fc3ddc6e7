"""A/B of the two density-query kernels (mma.sync fused query vs nrc_chain_query) on ray-coherent points.
python tests/tools/bench_query.py [--points 2097152] [--reps 10] [--only tc|mma]"""
import argparse, os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neural_radiance_caching_b200 import geometry, mlp_chain, sampling  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 21)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = np.random.default_rng(0)
    grids = [dict(hash_map_size=524288, max_grid_size=512, num_features=1),
             dict(hash_map_size=524288, max_grid_size=2048, num_features=4)]
    R, n = a.points // 64, 64
    o = g.uniform(-1, 1, size=(R, 1, 3)); d = g.normal(size=(R, 1, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    t = np.sort(g.uniform(0.05, 3.0, size=(R, n, 1)), axis=1)
    means = torch.tensor((o + t * d).reshape(-1, 3).astype(np.float32), device=dev)
    P = means.shape[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for gi, grid in enumerate(grids):
        import oracle.geometry as ogeo   # parameter initialiser only
        pred = gi == 1
        om = ogeo.DensityMLP(grid, enable_pred_normals=pred)
        m = geometry.DensityMLP(grid, enable_pred_normals=pred, bf16=True)
        p = m.from_oracle(om.init(g, table_init_range=0.5, bias_range=0.1), dev)
        density = torch.empty((P,), device=dev); feat = torch.empty((P, 64), device=dev) if pred else None
        gp = torch.empty((P, 3), device=dev) if pred else None
        cache = mlp_chain.PackCache()
        runs = {"tc": lambda: m.query_tc(p, means, density, feat, gp, cache=cache),
                "mma": lambda: m.query(p, means, want_feat=pred)}
        for name, fn in runs.items():
            if a.only and a.only != name:
                continue
            ts = []
            for i in range(a.reps + 3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            print(f"grid{gi} L*F={m.in_dim} {name}: {np.median(ts)*1e3:.1f} us  ({P/np.median(ts)/1e6:.2f} Gpoints/s)", flush=True)


if __name__ == "__main__":
    main()
