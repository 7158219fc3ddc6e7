import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import sampling as osamp
from neural_radiance_caching_b200 import sampling as nsamp
from tests.util import f32, gen, make_rays, rel_err, to_dev
dev = torch.device("cuda:0")
g = gen(101)
R = 128
o = osamp.ProposalVolumeSampler(); n = nsamp.ProposalVolumeSampler()
po = o.init(g, table_init_range=0.1, bias_range=0.05)
pn = n.from_oracle(po, dev)
rays = make_rays(g, R)
u = [f32(g.uniform(size=(R, 1))) for _ in range(3)]
G = [f32(g.normal(size=(R, ns))) for (_, _, ns) in o.sampling_strategy]
for i in range(3):
    for k in po[f"MLP_{i}"]["density_grid"]:
        po[f"MLP_{i}"]["density_grid"][k].requires_grad_(True)
ho = o(po, rays, u)
for h in ho:
    h["density"].retain_grad(); h["raw_density"].retain_grad()
sum((h["weights"] * Gl).sum() for h, Gl in zip(ho, G)).backward()
for i, m in enumerate(n.mlps):
    p = pn[f"MLP_{i}"]
    arena = p["density_grid"]["_arena"].clone().requires_grad_(True)
    p["density_grid"] = dict(m.grid.views(arena), _arena=arena)
hn = n(pn, to_dev(rays, dev), to_dev(u, dev), train=True)
for h in hn:
    h["density"].retain_grad(); h["raw_density"].retain_grad()
sum((h["weights"] * Gl.to(dev)).sum() for h, Gl in zip(hn, G)).backward()
for l in range(3):
    for k in ("sdist", "means", "density", "weights"):
        print(l, k, rel_err(hn[l][k], ho[l][k]))
    print(l, "g_density", rel_err(hn[l]["density"].grad, ho[l]["density"].grad))
    print(l, "g_raw", rel_err(hn[l]["raw_density"].grad, ho[l]["raw_density"].grad))
    d = (hn[l]["density"].grad.cpu() - ho[l]["density"].grad).abs()
    print("   worst ray/sample", np.unravel_index(int(d.argmax()), d.shape), float(d.max()),
          "valid flips:", int(((hn[l]["density"].cpu() == 0) != (ho[l]["density"] == 0)).sum()))
    gv = n.mlps[l].grid.views(pn[f"MLP_{l}"]["density_grid"]["_arena"].grad)
    for name in gv:
        print("   ", name, rel_err(gv[name], po[f"MLP_{l}"]["density_grid"][name].grad))
