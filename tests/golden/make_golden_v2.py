"""Freezes the oracle's restatement of the SURVEY 8f-1 pieces as tests/golden/oracle_v2.npz: geometry losses
(internal/loss_utils.py:127-199), mask loss (internal/train_utils.py:785-836) and the second-order gradients of the
analytic normals (double back-propagation through internal/geometry.py:442-460).  Same role and caveats as
make_golden.py (the reference ships no vectors for this path; JAX is not installable here).
Regenerate only on purpose:  python tests/golden/make_golden_v2.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import geometry as ogeo, loss_utils as oloss  # noqa: E402

SEED = 20200823 + 81
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v2.npz")
GRID = dict(hash_map_size=4096, max_grid_size=128, num_features=2)
MULTS = (0.01, 0.001, 0.01)


def f32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def build():
    torch.set_num_threads(1)
    g = np.random.Generator(np.random.PCG64(SEED))
    out = {}
    unit = lambda x: x / np.linalg.norm(x, axis=-1, keepdims=True)
    # ---- geometry losses: values and gradients
    R, n = 6, 8
    w = g.uniform(0, 0.3, size=(R, n)).astype(np.float32)
    nrm = unit(g.normal(size=(R, n, 3))).astype(np.float32)
    npred = unit(g.normal(size=(R, n, 3))).astype(np.float32)
    vd = unit(g.normal(size=(R, 3))).astype(np.float32)
    out.update(gl_w=w, gl_n=nrm, gl_np=npred, gl_vd=vd)
    W, N, NP = (f32(a).requires_grad_(True) for a in (w, nrm, npred))
    terms = oloss.geometry_losses({"viewdirs": f32(vd)}, {"weights": W, "normals": N, "normals_pred": NP}, *MULTS, 0.1)
    sum(terms).backward()
    out["gl_terms"] = np.array([float(t) for t in terms], np.float32)
    out["gl_gw"], out["gl_gn"], out["gl_gnp"] = W.grad.numpy(), N.grad.numpy(), NP.grad.numpy()
    # ---- mask loss (plain and the backward-mask call)
    acc = g.uniform(0, 1.2, size=(16,)).astype(np.float32)
    out["ml_acc"] = acc
    A = f32(acc).requires_grad_(True)
    l = oloss.compute_mask_loss(A, None, 0.001, 1.0, 1.0)
    l.backward()
    out["ml_loss"], out["ml_g"] = np.float32(float(l)), A.grad.numpy().copy()
    A.grad = None
    l = oloss.compute_mask_loss(A, torch.zeros(16, 1), 0.001, empty_loss_weight=0.1, backward=True)
    l.backward()
    out["mlb_loss"], out["mlb_g"] = np.float32(float(l)), A.grad.numpy().copy()
    # ---- distortion loss on metric distances through power_ladder(-0.25, 1e4)
    Rd, nd = 5, 32
    t = np.sort(g.uniform(2.0, 6.0, size=(Rd, nd + 1)).astype(np.float32), axis=-1)
    wd = (g.uniform(0, 1, size=(Rd, nd)).astype(np.float32) ** 3) * 0.2
    out["dl_t"], out["dl_w"] = t, wd
    Wd = f32(wd).requires_grad_(True)
    l = oloss.distortion_loss([{"tdist": f32(t), "weights": Wd}], 0.01, -0.25, 10000.0)
    l.backward()
    out["dl_loss"], out["dl_g"] = np.float32(float(l.detach())), Wd.grad.numpy().copy()
    # ---- second-order path: d/d theta <G, d raw / d means>
    mlp = ogeo.DensityMLP(grid_params=GRID, enable_pred_normals=True)
    gp = np.random.Generator(np.random.PCG64(SEED + 1))
    p = mlp.init(gp, table_init_range=0.5, bias_range=0.1)
    for v in p.values():
        for t in v.values():
            t.requires_grad_(True)
    P = 24
    means = g.uniform(-2.5, 2.5, size=(P, 3)).astype(np.float32)
    G = g.normal(size=(P, 3)).astype(np.float32)
    out["so_means"], out["so_G"] = means, G
    res = mlp(p, f32(means), create_graph=True)
    out["so_raw_grad"] = res["raw_grad_density"].detach().numpy()
    (res["raw_grad_density"] * f32(G)).sum().backward()
    for k in ("density_layers_0", "density_layers_1", "output_density_layer"):
        out[f"so_d_{k}"] = p[k]["kernel"].grad.numpy()
    names = sorted(p["density_grid"].keys())
    out["so_d_tables"] = np.concatenate([p["density_grid"][k].grad.numpy().reshape(-1) for k in names])
    return out


if __name__ == "__main__":
    np.savez_compressed(OUT, **build())
    print("wrote", OUT)
