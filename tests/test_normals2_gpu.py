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


def _tc_setup(cuda_device, gi, P, seed):
    import ctypes as C
    from neural_radiance_caching_b200 import _lib

    g = gen(seed)
    kw = dict(grid_params=GRIDS[gi], enable_pred_normals=True)
    o = ogeo.DensityMLP(**kw)
    n32, n16 = ngeo.DensityMLP(**kw), ngeo.DensityMLP(bf16=True, **kw)
    po = o.init(g, table_init_range=0.5, bias_range=0.1)
    pn = n32.from_oracle(po, cuda_device)
    means_c = f32(g.uniform(-2.5, 2.5, size=(P, 3)))
    G_c = f32(g.normal(size=(P, 3)))
    means, G = means_c.to(cuda_device), G_c.to(cuda_device)
    arena = pn["density_grid"]["_arena"]
    # primal features as the fused query saves them
    enc_out = torch.empty((P, n16.in_dim), device=cuda_device)
    density = torch.empty((P,), device=cuda_device)
    enc = n16.grid._descriptor(n16.grid.tables(n16.grid.views(arena)), None)
    desc = ngeo._mlp_desc(pn, n16.in_dim, True)
    _lib.call("nrc_density_query_fwd", _lib.stream_ptr(), C.byref(enc), C.byref(desc), _lib.ptr(means), P, float(n16.warp_c),
              float(n16.density_bias), 1, _lib.ptr(density), None, None, None, None, _lib.ptr(enc_out))
    return dict(o=o, po=po, n32=n32, n16=n16, pn=pn, means=means, G=G, means_c=means_c, G_c=G_c, arena=arena,
                enc_out=enc_out, desc=desc)


@pytest.mark.parametrize("gi,P", [(0, 1000), (1, 2049), (2, 257)])
def test_tangent_gather_and_scatter(cuda_device, gi, P):
    """nrc_encode_tangent_fwd against the oracle's forward-mode derivative of the encoding along G (fp32, 1e-5), and
    nrc_encode_tangent_bwd as its exact adjoint in the tables: <K_c(ge), dT> == <ge, K_a(T = dT)>."""
    import ctypes as C
    from neural_radiance_caching_b200 import _lib

    t = _tc_setup(cuda_device, gi, P, 960 + gi)
    n, arena, means, G = t["n16"], t["arena"], t["means"], t["G"]
    LF = n.in_dim
    enc = n.grid._descriptor(n.grid.tables(n.grid.views(arena)), None)
    edot = torch.empty((P, LF), device=cuda_device)
    _lib.call("nrc_encode_tangent_fwd", _lib.stream_ptr(), C.byref(enc), _lib.ptr(means), _lib.ptr(G), P, float(n.warp_c),
              _lib.ptr(edot))
    _, want = torch.autograd.functional.jvp(lambda m: t["o"].encode(t["po"], m), t["means_c"], t["G_c"])
    assert rel_err(edot, want.reshape(P, LF)) <= 1e-5
    # adjoint identity
    gq = gen(77)
    ge = f32(gq.normal(size=(P, LF))).to(cuda_device)
    d_arena = f32(gq.normal(size=tuple(arena.shape))).to(cuda_device)
    ga = torch.zeros_like(arena)
    enc_g = n.grid._descriptor(n.grid.tables(n.grid.views(arena)), n.grid.tables(n.grid.views(ga)))
    _lib.call("nrc_encode_tangent_bwd", _lib.stream_ptr(), C.byref(enc_g), _lib.ptr(means), _lib.ptr(G), _lib.ptr(ge), P,
              float(n.warp_c))
    enc_d = n.grid._descriptor(n.grid.tables(n.grid.views(d_arena)), None)
    edot_d = torch.empty((P, LF), device=cuda_device)
    _lib.call("nrc_encode_tangent_fwd", _lib.stream_ptr(), C.byref(enc_d), _lib.ptr(means), _lib.ptr(G), P, float(n.warp_c),
              _lib.ptr(edot_d))
    torch.cuda.synchronize()
    lhs, rhs = float((ga.double() * d_arena.double()).sum()), float((ge.double() * edot_d.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), float(ge.norm() * edot_d.norm()) * 1e-3)


@pytest.mark.parametrize("gi,P", [(0, 1000), (1, 4099), (1, 32768)])
def test_normals_second_order_tensor_core_path(cuda_device, gi, P):
    """bf16-MLP variant of the second-order term.  nrc_density_mlp_bwd_tangent is held to a torch emulation that rounds
    operands to bf16 at exactly the kernel's rounding points (fp32 accumulation): like the first-order gradients
    (tests/test_mlp_gpu.py) the result is only comparable with a reference that sees the same ReLU masks.  The whole
    three-launch path is then bounded against the fp32 kernel nrc_density_normals_bwd (oracle-checked above) in the L2
    norm; bf16 mask flips on white-noise tables allow up to 0.2 there."""
    import ctypes as C
    from neural_radiance_caching_b200 import _lib

    t = _tc_setup(cuda_device, gi, P, 950 + gi)
    n32, n16, pn, arena, means, G, enc_out = t["n32"], t["n16"], t["pn"], t["arena"], t["means"], t["G"], t["enc_out"]
    LF = n16.in_dim
    enc = n16.grid._descriptor(n16.grid.tables(n16.grid.views(arena)), None)
    edot = torch.empty((P, LF), device=cuda_device)
    _lib.call("nrc_encode_tangent_fwd", _lib.stream_ptr(), C.byref(enc), _lib.ptr(means), _lib.ptr(G), P, float(n16.warp_c),
              _lib.ptr(edot))
    zeros = lambda: {k: {kk: torch.zeros_like(x) for kk, x in v.items()} for k, v in pn.items() if k != "density_grid"}
    gp = zeros()
    ge = torch.empty((P, LF), device=cuda_device)
    gd = ngeo._grad_desc(n16, gp)
    _lib.call("nrc_density_mlp_bwd_tangent", _lib.stream_ptr(), C.byref(t["desc"]), _lib.ptr(enc_out), _lib.ptr(edot), P,
              _lib.ptr(ge), C.byref(gd))
    torch.cuda.synchronize()
    # emulation with the kernel's rounding points
    r = lambda x: x.to(torch.bfloat16).float()
    W0, b0 = pn["density_layers_0"]["kernel"], pn["density_layers_0"]["bias"]
    W1, b1 = pn["density_layers_1"]["kernel"], pn["density_layers_1"]["bias"]
    wd = pn["output_density_layer"]["kernel"][:, 0]
    h1pre = r(enc_out) @ r(W0) + b0
    M1 = (h1pre > 0).float()
    h2pre = r(torch.relu(h1pre)) @ r(W1) + b1
    M2 = (h2pre > 0).float()
    ed16 = r(edot)
    h1d = r(M1 * (ed16 @ r(W0)))
    h2d = r(M2 * (h1d @ r(W1)))
    a2 = r(M2 * wd[None, :])
    a1 = r(M1 * (a2 @ r(W1).T))
    l2 = lambda a, b: float((a - b).norm() / b.norm())
    assert l2(ge, a1 @ r(W0).T) <= 1e-3
    assert l2(gp["density_layers_1"]["kernel"], h1d.T @ a2) <= 1e-3
    assert l2(gp["density_layers_0"]["kernel"], ed16.T @ a1) <= 1e-3
    assert l2(gp["output_density_layer"]["kernel"][:, 0], h2d.sum(0)) <= 1e-3
    for k in ("density_layers_0", "density_layers_1", "output_density_layer", "pred_normals_layer"):
        assert float(gp[k]["bias"].abs().max()) == 0.0
    assert float(gp["pred_normals_layer"]["kernel"].abs().max()) == 0.0

    # whole path against the fp32 kernel
    def run(mlp, eo):
        g_p, g_a = zeros(), torch.zeros_like(arena)
        ngeo.density_normals_bwd(mlp, pn, arena, means, G, g_p, g_a, enc_out=eo)
        torch.cuda.synchronize()
        return g_p, g_a

    want_p, want_a = run(n32, None)
    got_p, got_a = run(n16, enc_out)
    for k in ("density_layers_0", "density_layers_1", "output_density_layer"):
        assert l2(got_p[k]["kernel"], want_p[k]["kernel"]) <= 0.2, k
    assert l2(got_a, want_a) <= 0.2
    assert float(want_a.abs().max()) > 0
