"""Freezes outputs of the CPU oracle (oracle/) for fixed seeds as tests/golden/oracle_v1.npz.

The reference ships no golden vectors for this path and JAX cannot be installed here, so
these fixtures pin the RESTATEMENT: any later edit of oracle/ that changes a value is caught by
tests/test_golden.py on CPU, and the CUDA kernels are compared with the same frozen numbers in
tests/test_golden_gpu.py.  Regenerate only on purpose:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import coord as ocoord, geometry as ogeo, grid_utils as og, models as omodels  # noqa: E402
from oracle import nerf as onerf, render as orender, render_utils as oru, stepfun as ostep  # noqa: E402

SEED = 20200823
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz")


def f32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def build():
    torch.set_num_threads(1)
    g = np.random.Generator(np.random.PCG64(SEED))
    out = {}
    # ---- hash / dense corner indices and encoded features (rows 1-5)
    x = g.uniform(-1.3, 1.3, size=(48, 3)).astype(np.float32)
    x[:4] = [[-1, -1, -1], [1, 1, 1], [0, 0, 0], [-2.5, 2.5, -2.5]]
    out["enc_x"] = x
    enc = og.HashEncoding(hash_map_size=524288, num_features=4, scale_supersample=1.0, max_grid_size=2048,
                          bbox_scaling=1.0)
    bbox = enc.bbox
    xn = ((f32(x) - f32(bbox[0])) / f32(bbox[1] - bbox[0])).numpy()
    for l, (kind, N, _) in enumerate(enc.layout):
        pos = xn * np.float32(N)
        if kind == "hash":
            idx = og.hash_corner_indices_np(pos, enc.hash_map_size)
        else:
            c = og.dense_corner_indices_np(pos, N)
            idx = (c[..., 0] * (N + 2) + c[..., 1]) * (N + 2) + c[..., 2]
        out[f"enc_idx_l{l}"] = idx.astype(np.int32)
    ge = np.random.Generator(np.random.PCG64(SEED + 1))
    p = enc.init(ge, init_range=0.1)
    out["enc_feat"] = enc(p, f32(x)).numpy()
    # ---- contraction (row 7)
    out["contract_c2"] = ocoord.contract_radius(f32(x * 3.0), 2.0).numpy()
    # ---- step-function resampling (row 12): fenceposts and bin indices
    R, m, n = 6, 16, 12
    t = np.sort(g.uniform(0, 1, size=(R, m + 1)).astype(np.float32), axis=-1)
    t[:, 0], t[:, -1] = 0.0, 1.0
    w = g.uniform(0, 1, size=(R, m)).astype(np.float32) ** 3
    u01 = g.uniform(size=(R, 1)).astype(np.float32)
    out["si_t"], out["si_w"], out["si_u01"] = t, w, u01
    logits = 0.4 * torch.log(torch.clamp(f32(w) + 1e-5, min=float(np.finfo(np.float32).tiny)))
    sd, idx = ostep.sample_intervals(f32(u01), f32(t), logits, n, single_jitter=True, domain=(0.0, 1.0), return_idx=True)
    out["si_sdist"] = sd.numpy()
    out["si_bins"] = idx.numpy().astype(np.int32)
    # ---- alpha compositing weights and volumetric rendering (rows 14, 17)
    dens = g.gamma(0.5, 4.0, size=(R, n)).astype(np.float32)
    tdist = np.sort(g.uniform(2, 6, size=(R, n + 1)).astype(np.float32), axis=-1)
    dirs = g.normal(size=(R, 3)).astype(np.float32)
    out["aw_density"], out["aw_tdist"], out["aw_dirs"] = dens, tdist, dirs
    wts, alpha, trans = orender.compute_alpha_weights(f32(dens), f32(tdist), f32(dirs))
    out["aw_weights"], out["aw_alpha"], out["aw_trans"] = wts.numpy(), alpha.numpy(), trans.numpy()
    rgbs = g.uniform(size=(R, n, 3)).astype(np.float32)
    out["vr_rgbs"] = rgbs
    ren = orender.volumetric_rendering(f32(rgbs), wts, wts, f32(tdist), 1.0, True)
    for k in ("rgb", "acc", "distance_mean", "distance_median", "distance_percentile_5", "distance_percentile_95"):
        out["vr_" + k] = ren[k].numpy()
    # ---- categorical resampling (row 15)
    gum = g.gumbel(size=(R, n, 2)).astype(np.float32)
    out["rs_gumbel"] = gum
    inds, w_new = omodels.maybe_resample(wts, f32(gum), 2)
    out["rs_inds"], out["rs_w"] = inds.numpy().astype(np.int32), w_new.numpy()
    # ---- density MLP (row 8)
    gm = np.random.Generator(np.random.PCG64(SEED + 2))
    mlp = ogeo.DensityMLP(grid_params=dict(hash_map_size=524288, max_grid_size=512, num_features=1))
    pm = mlp.init(gm, table_init_range=0.1, bias_range=0.1)
    raw, feat = mlp.predict_density(pm, f32(x))
    out["mlp_raw"], out["mlp_feat"] = raw.detach().numpy(), feat.detach().numpy()
    out["mlp_density"] = mlp.convert_raw_density(raw, f32(x)).detach().numpy()
    # ---- integrated directional encoding (row 16)
    d = g.normal(size=(16, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    kinv = g.gamma(1.0, 0.3, size=(16, 1)).astype(np.float32)
    out["ide_dirs"], out["ide_kinv"] = d.astype(np.float32), kinv
    out["ide4"] = onerf.generate_ide_fn(4)(f32(d), f32(kinv)).numpy()
    out["ide5_f64"] = onerf.generate_ide_fn(5, dtype=torch.float64)(f32(d).double(), f32(kinv).double()).numpy()
    # ---- GGX lobe integration (row 21)
    S = 8
    def unit(a):
        return a / np.linalg.norm(a, axis=-1, keepdims=True)
    wi = unit(g.normal(size=(R, S, 3)))
    wo = unit(np.abs(g.normal(size=(R, 1, 3))) + [0, 0, 0.05]) * np.ones((1, S, 1))
    samples = dict(
        local_lightdirs=f32(wi), local_viewdirs=f32(wo), radiance_in=f32(g.gamma(1.0, 1.0, size=(R, S, 3))),
        pdf=f32(g.gamma(1.0, 0.5, size=(R, S, 1))), weight=f32(g.uniform(-0.1, 2.0, size=(R, S, 1))),
        indirect_occ=f32(g.uniform(size=(R, S, 1))), brdf_correction=torch.ones(R, S, 2))
    material = dict(albedo=f32(g.uniform(size=(R, 3))), roughness=f32(g.uniform(0.01, 1.0, size=(R, 1))),
                    metalness=f32(g.uniform(size=(R, 1))), F_0=torch.full((R, 1), 0.04))
    for k, v in samples.items():
        out["ggx_s_" + k] = v.numpy()
    for k, v in material.items():
        out["ggx_m_" + k] = v.numpy()
    res = oru.integrate_reflect_rays("microfacet", material, samples, max_radiance=10000.0)
    for k in ("radiance_out", "irradiance", "indirect_occ"):
        out["ggx_" + k] = res[k].numpy()
    return out


if __name__ == "__main__":
    data = build()
    np.savez_compressed(OUT, **data)
    print(OUT, {k: v.shape for k, v in data.items()})
