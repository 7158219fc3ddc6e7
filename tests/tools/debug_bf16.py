import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import geometry as ogeo
from neural_radiance_caching_b200 import geometry as ngeo
from tests.util import f32, gen, rel_err
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
g = gen(1)
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
GR = dict(hash_map_size=524288, max_grid_size=2048, num_features=4)
o = ogeo.DensityMLP(grid_params=GR, enable_pred_normals=False)
n = ngeo.DensityMLP(grid_params=GR, enable_pred_normals=False, bf16=True)
po = o.init(g, table_init_range=0.1, bias_range=0.1)
pn = n.from_oracle(po, dev)
P = 1000
x = f32(g.normal(size=(P, 32))).to(dev)
g_raw = f32(g.normal(size=(P,))).to(dev)
W0, b0 = pn["density_layers_0"]["kernel"], pn["density_layers_0"]["bias"]
W1, b1 = pn["density_layers_1"]["kernel"], pn["density_layers_1"]["bias"]
wd = pn["output_density_layer"]["kernel"]
# emulation (double accumulate to take summation order out of the picture)
D = torch.float64
a1 = (bf(x).to(D) @ bf(W0).to(D) + b0.to(D)); h1 = torch.relu(a1)
a2 = (bf(h1.float()).to(D) @ bf(W1).to(D) + b1.to(D)); h2 = torch.relu(a2)
g2 = (g_raw[:, None].to(D) * wd[:, 0][None].to(D)) * (a2 > 0)
g1 = (bf(g2.float()).to(D) @ bf(W1).to(D).T) * (a1 > 0)
ge = bf(g1.float()).to(D) @ bf(W0).to(D).T
xn = x.clone().requires_grad_(True)
for k in ("density_layers_0", "density_layers_1", "output_density_layer"):
    for kk in pn[k]:
        pn[k][kk] = pn[k][kk].detach().clone().requires_grad_(True)
outs = n.run_network(pn, xn)
(outs[0] * g_raw).sum().backward()
d = (xn.grad.double() - ge).abs().max(-1).values / ge.abs().max()
print("vs bf16 emulation: x.grad max", float(d.max()), "median", float(d.median()), "p99", float(d.quantile(0.99)))
dW1 = bf(h1.float()).to(D).T @ bf(g2.float()).to(D)
dW0 = bf(x).to(D).T @ bf(g1.float()).to(D)
print("dW1", rel_err(pn["density_layers_1"]["kernel"].grad, dW1), "dW0", rel_err(pn["density_layers_0"]["kernel"].grad, dW0))
print("db1", rel_err(pn["density_layers_1"]["bias"].grad, g2.sum(0)), "db0", rel_err(pn["density_layers_0"]["bias"].grad, g1.sum(0)))
# fp32 oracle masks vs bf16 masks
a1f = x @ W0 + b0; a2f = torch.relu(a1f) @ W1 + b1
print("mask flips per point: layer1", float(((a1f > 0) != (a1 > 0)).float().sum(1).mean()), "layer2", float(((a2f > 0) != (a2 > 0)).float().sum(1).mean()))
