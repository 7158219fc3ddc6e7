"""Second-order path of the analytic normals: nrc_density_normals_bwd against double
back-propagation through the oracle (internal/geometry.py:442-460)."""
import pytest
import torch

from oracle import geometry as ogeo
from neural_radiance_caching_b200 import geometry as ngeo
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu

GRIDS = [
    dict(hash_map_size=524288, max_grid_size=512, num_features=1),
    dict(hash_map_size=524288, max_grid_size=2048, num_features=4),
    dict(hash_map_size=4096, max_grid_size=256, num_features=2),
]


@pytest.mark.parametrize("gi,P", [(0, 1000), (1, 4099), (2, 257)])
def test_normals_second_order_matches_oracle(cuda_device, gi, P):
    g = gen(900 + gi)
    kw = dict(grid_params=GRIDS[gi], enable_pred_normals=True)
    o, n = ogeo.DensityMLP(**kw), ngeo.DensityMLP(**kw)
    po = o.init(g, table_init_range=0.5, bias_range=0.1)
    pn = n.from_oracle(po, cuda_device)
    means = f32(g.uniform(-2.5, 2.5, size=(P, 3)))
    G = f32(g.normal(size=(P, 3)))
    # oracle: L = <G, d raw / d means>, back-propagated to every parameter (create_graph => second order)
    leaves = {}
    for k, v in po.items():
        for kk, t in v.items():
            t.requires_grad_(True)
            leaves[(k, kk)] = t
    out = o(po, means, create_graph=True)
    (out["raw_grad_density"] * G).sum().backward()
    # CUDA: forward value and the VJP through the autograd wrapper
    arena = pn["density_grid"]["_arena"].requires_grad_(True)
    for k in ("density_layers_0", "density_layers_1", "output_density_layer", "pred_normals_layer"):
        for kk in pn[k]:
            pn[k][kk].requires_grad_(True)
    rg = n.raw_grad_density(pn, means.to(cuda_device))
    assert rel_err(rg, out["raw_grad_density"].detach()) <= 1e-5
    (rg * G.to(cuda_device)).sum().backward()
    torch.cuda.synchronize()
    # fp32, different summation order (per-CTA register partial sums + atomics): 1e-5 of the gradient's scale
    for k in ("density_layers_0", "density_layers_1", "output_density_layer"):
        assert rel_err(pn[k]["kernel"].grad, leaves[(k, "kernel")].grad) <= 2e-5, k
        gb = pn[k]["bias"].grad
        assert gb is None or float(gb.abs().max()) == 0.0          # the tangent map has no bias term
        ob = leaves[(k, "bias")].grad
        assert ob is None or float(ob.abs().max()) == 0.0
    names = [nm for (nm, _, _, _) in n.grid.level_layout]
    want = torch.cat([leaves[("density_grid", nm)].grad.reshape(-1) for nm in names])
    assert rel_err(arena.grad, want) <= 2e-5
    assert float(want.abs().max()) > 0
