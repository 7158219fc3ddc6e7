import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import nerf as onerf
from neural_radiance_caching_b200 import nerf as nnerf
from tests.util import f32, gen, rel_err, rel_l2
from tests.test_shader_gpu import _shader_inputs
dev = torch.device("cuda:0")
for R in (48, 512):
  for trial in range(2):
    g = gen(341)
    o = onerf.NeRFMLP(); n = nnerf.NeRFMLP(bf16=True)
    po = o.init(g, table_init_range=0.1)
    pn = n.from_oracle(po, dev)
    v, means, feat, nrm = _shader_inputs(g, R, 32)
    G = f32(g.normal(size=(R, 32, 3)))
    def leaves(p, prefix=""):
        out = []
        for k in sorted(p.keys()):
            if k == "_arena": continue
            if isinstance(p[k], dict): out += leaves(p[k], prefix + k + "/")
            else: out.append((prefix + k, p, k))
        return out
    lo = leaves(po)
    for _, d, k in lo: d[k] = d[k].clone().requires_grad_(True)
    fo, no = feat.clone().requires_grad_(True), nrm.clone().requires_grad_(True)
    (o(po, v, means, fo, no)["rgb"] * G).sum().backward()
    arena = pn["appearance_grid"]["_arena"].clone().requires_grad_(True)
    pn["appearance_grid"] = dict(n.grid.views(arena.detach()), _arena=arena)
    ln = leaves(pn)
    for name, d, k in ln:
        if "appearance_grid" not in name: d[k] = d[k].clone().requires_grad_(True)
    fn_, nn_ = feat.to(dev).requires_grad_(True), nrm.to(dev).requires_grad_(True)
    out = n(pn, v.to(dev), means.to(dev), fn_, nn_)
    (out["rgb"] * G.to(dev)).sum().backward()
    print(f"R={R} trial {trial}: d_feat l2 {rel_l2(fn_.grad, fo.grad):.4f} max {rel_err(fn_.grad, fo.grad):.4f} | d_nrm l2 {rel_l2(nn_.grad, no.grad):.4f}")
    msg = []
    for (name, dn, kn), (_, do, ko) in zip(ln, lo):
        if "appearance_grid" in name: continue
        ref = do[ko].grad
        if ref is None or float(ref.abs().max()) == 0.0 or dn[kn].grad is None: continue
        msg.append(f"{name.split('/')[-2][-12:]}.{name[-1]} {rel_l2(dn[kn].grad, ref):.3f}")
    print("   ", " ".join(msg))
