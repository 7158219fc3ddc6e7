"""The XLA custom-call targets of libnrc_b200.so (include/nrc_xla.h) driven through ctypes with EXACTLY the arguments XLA's
GPU runtime passes for the jaxlib the reference pins (jax==0.4.16: void(cudaStream_t, void** buffers, const char* opaque,
size_t opaque_len), operands first, results after): every target must write the same bits as the direct C-ABI call the
Python mirrors make (which the other GPU tests hold to the oracle and to the reference's own vectors)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from neural_radiance_caching_b200 import _lib, coord as ncoord, geometry as ngeo, grid_utils as ng, models as nmodels
from neural_radiance_caching_b200 import render as nrender, stepfun as nstep
from neural_radiance_caching_b200.inverse_render import render_utils as nru
from neural_radiance_caching_b200.jax_binding import nrc_jax as J
from tests.util import ENC_CONFIGS, dense_params, f32, gen, level_table

pytestmark = pytest.mark.gpu


def xla_call(name, operands, results, desc, expect_status=0):
    """One custom call: buffers = operand pointers then result pointers; opaque = the packed descriptor."""
    lib = _lib.load()
    fn = getattr(lib, name)
    fn.restype, fn.argtypes = None, [C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    bufs = (C.c_void_p * (len(operands) + len(results)))(*[t.data_ptr() for t in list(operands) + list(results)])
    opaque = bytes(memoryview(desc))
    fn(C.c_void_p(torch.cuda.current_stream().cuda_stream), bufs, opaque, len(opaque))
    lib.nrc_xla_last_status.restype = C.c_int32
    assert lib.nrc_xla_last_status() == expect_status, name
    torch.cuda.synchronize()


def same(a, b):
    return np.array_equal(a.detach().cpu().numpy(), b.detach().cpu().numpy())


def test_encode_fwd_bwd_and_contract(cuda_device):
    dev = cuda_device
    g = gen(900)
    enc = ng.HashEncoding(**ENC_CONFIGS["b"])
    tables = {name: torch.from_numpy(level_table(shape, i + 1)).to(dev) for i, (name, _, _, shape) in enumerate(enc.level_layout)}
    arena = torch.cat([tables[name].reshape(-1) for (name, _, _, _) in enc.level_layout]).contiguous()
    P = 777
    x = f32(g.normal(size=(P, 3)) * 1.5).to(dev).requires_grad_(True)
    a2 = arena.clone().requires_grad_(True)
    want = enc(dict(enc.views(a2.detach()), _arena=a2), x)     # single-arena path of the custom VJP
    gout = f32(g.normal(size=tuple(want.shape))).to(dev)
    want.backward(gout)
    d = J.pack_encode(enc, P)
    assert d.arena_floats == arena.numel()
    out = torch.empty_like(want)
    xla_call("nrc_xla_encode_fwd", [x.detach(), arena], [out], d)
    assert same(out, want)
    g_arena, g_x = torch.full_like(arena, 7.0), torch.empty((P, 3), device=dev)     # XLA result buffers are uninitialised
    xla_call("nrc_xla_encode_bwd", [x.detach(), arena, gout], [g_arena, g_x], d)
    assert float((g_x - x.grad).abs().max()) <= 1e-6 * float(x.grad.abs().max())    # the mirror's VJP sums the levels in another order
    # scatter-add order differs between two launches (atomics): equal up to summation order
    assert float((g_arena - a2.grad).abs().max()) <= 1e-6 * float(a2.grad.abs().max())
    # contraction
    z = torch.empty_like(x)
    xla_call("nrc_xla_contract_fwd", [x.detach()], [z], J.pack_contract(P, 2.0))
    x2 = x.detach().clone().requires_grad_(True)
    zw = ncoord._ContractFn.apply(x2, 2.0)
    assert same(z, zw)
    zw.backward(gout[:, :3].contiguous())
    gx = torch.empty_like(x)
    xla_call("nrc_xla_contract_bwd", [x.detach(), gout[:, :3].contiguous()], [gx], J.pack_contract(P, 2.0))
    assert same(gx, x2.grad)


def test_density_query(cuda_device):
    dev = cuda_device
    g = gen(901)
    cfg = {k: v for k, v in ENC_CONFIGS["a"].items() if k != "scale_supersample"}
    mlp = ngeo.DensityMLP(cfg, net_depth=2, net_width=64, density_bias=-1.0, warp_c=2.0, bbox_scaling=2.0, enable_pred_normals=True)
    p = {"density_grid": {name: torch.from_numpy(level_table(shape, i + 1)) for i, (name, _, _, shape) in enumerate(mlp.grid.level_layout)}}
    for name, d_in, d_out, salt in (("density_layers_0", mlp.in_dim, 64, 100), ("density_layers_1", 64, 64, 101),
                                    ("output_density_layer", 64, 1, 102), ("pred_normals_layer", 64, 3, 110)):
        k, b = dense_params(d_in, d_out, salt)
        p[name] = {"kernel": torch.from_numpy(k), "bias": torch.from_numpy(b)}
    pd = mlp.from_oracle(p, dev)
    P = 515
    means = f32(g.normal(size=(P, 3)) * 1.5).to(dev)
    want = mlp.query(pd, means, want_feat=True, want_normals=True)
    new = lambda *s: torch.empty(s, device=dev)
    outs = [new(P), new(P), new(P, 64), new(P, 3), new(P, 3)]
    ops = [means, pd["density_grid"]["_arena"]] + [pd[n][k] for n in ("density_layers_0", "density_layers_1", "output_density_layer",
                                                                       "pred_normals_layer") for k in ("kernel", "bias")]
    d = J.pack_density_query(mlp.grid, P, mlp.in_dim, has_pred_normals=True, want_raw_grad=True, warp_c=2.0, density_bias=-1.0)
    xla_call("nrc_xla_density_query_fwd", ops, outs, d)
    for got, k in zip(outs, ("density", "raw_density", "feature", "grad_pred", "raw_grad_density")):
        assert same(got, want[k]), k


def test_ray_targets(cuda_device):
    dev = cuda_device
    g = gen(902)
    R, m, n = 300, 64, 32
    new = lambda *s: torch.empty(s, device=dev)
    t = torch.sort(f32(g.uniform(size=(R, m + 1))), dim=-1).values.to(dev)
    w = f32(g.uniform(size=(R, m)) ** 3).to(dev)
    u01 = f32(g.uniform(size=(R,))).to(dev)
    # sample_intervals
    want = nstep.sample_intervals_from_weights(u01, t, w, n, anneal=0.4, padding=1e-5, domain=(0.0, 1.0))
    base, mj = nstep.u_base(n, dev)
    t_new = new(R, n + 1)
    xla_call("nrc_xla_ray_sample_intervals", [t, w, u01, base], [t_new],
             J.pack_ray(R, n, m=m, anneal=0.4, padding=1e-5, max_jitter=mj, dom_lo=0.0, dom_hi=1.0))
    assert same(t_new, want)
    # cast (power-ladder warp)
    from neural_radiance_caching_b200.sampling import ProposalVolumeSampler
    o, dr = f32(g.normal(size=(R, 3))).to(dev), f32(g.normal(size=(R, 3))).to(dev)
    near, far = torch.full((R, 1), 0.05, device=dev), torch.full((R, 1), 6.0, device=dev)
    s = ProposalVolumeSampler()
    td_w, mu_w = s._cast(t_new, dict(origins=o, directions=dr, near=near, far=far), True)
    td, mu = new(R, n + 1), new(R, n, 3)
    xla_call("nrc_xla_ray_cast", [t_new, o, dr, near, far], [td, mu], J.pack_ray(R, n, warp_kind=1, p=s.raydist[0], premult=s.raydist[1]))
    assert same(td, td_w) and same(mu, mu_w)
    # alpha weights fwd / bwd
    dens = f32(g.gamma(1.0, 2.0, size=(R, n))).to(dev).requires_grad_(True)
    ww, aw, tw = nrender.compute_alpha_weights(dens, td, dr, opaque_background=False)
    gw = f32(g.normal(size=(R, n))).to(dev)
    ww.backward(gw)
    wo_, ao_, to_ = new(R, n), new(R, n), new(R, n)
    xla_call("nrc_xla_ray_alpha_weights_fwd", [dens.detach(), td, dr], [wo_, ao_, to_], J.pack_ray(R, n))
    assert same(wo_, ww) and same(ao_, aw) and same(to_, tw)
    gd = new(R, n)
    zeros = torch.zeros((R, n), device=dev)
    xla_call("nrc_xla_ray_alpha_weights_bwd", [dens.detach(), td, dr, gw, zeros, zeros], [gd], J.pack_ray(R, n))
    assert float((gd - dens.grad).abs().max()) <= 1e-6 * float(dens.grad.abs().max())
    # composite fwd / bwd (rgb + background)
    vals = f32(g.uniform(size=(R, n, 3))).to(dev).requires_grad_(True)
    wts = wo_.clone().requires_grad_(True)
    bg = f32(g.uniform(size=(R, 3))).to(dev)
    vr = nrender.volumetric_rendering(vals, wts, wts, td, bg, True)
    out, acc, dist = new(R, 3), new(R), new(R, 4)
    flags = dict(k=n, channels=3, has_rgb=1, has_bg=1, has_weights_nf=0)
    xla_call("nrc_xla_ray_composite_fwd", [vals.detach(), wts.detach(), td, bg], [out, acc, dist], J.pack_ray(R, n, **flags))
    assert same(out, vr["rgb"]) and same(acc, vr["acc"]) and same(dist[:, 0], vr["distance_mean"])
    g_out, g_acc = f32(g.normal(size=(R, 3))).to(dev), f32(g.normal(size=(R,))).to(dev)
    (vr["rgb"] * g_out).sum().add((vr["acc"] * g_acc).sum()).backward()
    gv, gwt = new(R, n, 3), new(R, n)
    xla_call("nrc_xla_ray_composite_bwd", [vals.detach(), wts.detach(), bg, g_out, g_acc], [gv, gwt], J.pack_ray(R, n, **flags))
    assert float((gv - vals.grad).abs().max()) <= 1e-6 * float(vals.grad.abs().max())
    assert float((gwt - wts.grad).abs().max()) <= 1e-6 * float(wts.grad.abs().max())
    # resample + gather
    gum = f32(-np.log(-np.log(g.uniform(1e-12, 1.0, size=(R, n, 4))))).to(dev)
    inds_w, nw_w = nmodels._ResampleWeightsFn.apply(wo_, gum, 1e-3, 1.0)
    inds, nw = torch.empty((R, 4), device=dev, dtype=torch.int32), new(R, 4)
    xla_call("nrc_xla_ray_resample", [wo_, gum], [inds, nw], J.pack_ray(R, n, k=4, bias=1e-3, mult=1.0))
    assert same(inds, inds_w) and same(nw, nw_w)
    got = new(R, 4, 3)
    xla_call("nrc_xla_ray_resample_gather", [mu, inds], [got], J.pack_ray(R, n, k=4, channels=3))
    assert same(got, nmodels._GatherFn.apply(mu, inds_w))


def test_ggx_targets(cuda_device):
    dev = cuda_device
    V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_np.npz"))
    D = lambda k: torch.from_numpy(V[k]).to(dev).contiguous()
    material = {k: D("ggx_mat_" + k) for k in ("albedo", "roughness", "F_0", "metalness")}
    samples = {k: D("ggx_smp_" + k) for k in ("local_lightdirs", "local_viewdirs", "pdf", "weight", "radiance_in", "indirect_occ")}
    samples["radiance_in"].requires_grad_(True)
    res = nru.integrate_reflect_rays("microfacet", False, material, samples)
    R, S = samples["pdf"].shape[0], samples["pdf"].shape[1]
    g = gen(903)
    g_out, g_irr = f32(g.normal(size=(R, 3))).to(dev), f32(g.normal(size=(R, 3))).to(dev)
    ((res["radiance_out"] * g_out).sum() + (res["irradiance"] * g_irr).sum()).backward()
    flat = lambda k: samples[k].detach().reshape(R, S).contiguous()
    vec = lambda k: material[k].reshape(R).contiguous()
    wi, wo = samples["local_lightdirs"], samples["local_viewdirs"].expand(R, S, 3).contiguous()
    rad = samples["radiance_in"].detach()
    new = lambda *s: torch.empty(s, device=dev)
    out, irr, occ = new(R, 3), new(R, 3), new(R)
    ops = [wi, wo, rad, flat("weight"), flat("pdf"), flat("indirect_occ"), material["albedo"], vec("roughness"), vec("metalness"), vec("F_0")]
    xla_call("nrc_xla_ggx_integrate_fwd", ops, [out, irr, occ], J.pack_ggx(R, S, 0, True))
    assert same(out, res["radiance_out"]) and same(irr, res["irradiance"]) and same(occ[:, None], res["indirect_occ"])
    g_rad = new(R, S, 3)
    ops = [wi, wo, rad, flat("weight"), flat("pdf"), material["albedo"], vec("roughness"), vec("metalness"), vec("F_0"), g_out, g_irr]
    xla_call("nrc_xla_ggx_integrate_bwd", ops, [g_rad], J.pack_ggx(R, S, 0, False))
    assert same(g_rad, samples["radiance_in"].grad)


def test_slf_targets(cuda_device):
    """The surface-light-field point stage and feature reduction through the XLA ABI == the C-ABI entry points."""
    from neural_radiance_caching_b200 import surface_light_field as nslf
    dev = cuda_device
    g = gen(905)
    P, n, Fq = 301, 8, 36
    net = nslf.SurfaceLightFieldMemMLP(num_distance_samples=n, grid=dict(hash_map_size=2 ** 12, max_grid_size=64, num_features=4),
                                       reflectance_grid=dict(hash_map_size=2 ** 12, max_grid_size=64, num_features=4))
    cfg = nslf._points_cfg(net, 0.1, 1.7)
    raw = f32(g.normal(size=(P, 8 * n + 4))).to(dev).requires_grad_(True)
    o = f32(g.normal(size=(P, 3))).to(dev)
    v = torch.nn.functional.normalize(f32(g.normal(size=(P, 3))), dim=-1).to(dev)
    want = nslf._SlfPointsFn.apply(raw, o, v, cfg)
    ups = [f32(g.normal(size=tuple(w.shape))).to(dev) for w in want]
    sum((a * b).sum() for a, b in zip(want, ups)).backward()
    new = lambda *s: torch.empty(s, device=dev)
    outs = [new(P, n, 3), new(P, n), new(P), new(P, n), new(P, 4)]
    xla_call("nrc_xla_slf_points_fwd", [raw.detach(), o, v], outs, J.pack_slf(P, cfg))
    for a, b in zip(outs, want):
        assert same(a.reshape(b.shape), b)
    g_raw = new(P, 8 * n + 4)
    xla_call("nrc_xla_slf_points_bwd", [raw.detach(), o, v, ups[0], ups[1], ups[2].reshape(P), ups[3], ups[4]], [g_raw], J.pack_slf(P, cfg))
    assert same(g_raw, raw.grad)
    feat = f32(g.normal(size=(P, n, Fq))).to(dev).requires_grad_(True)
    w = f32(g.uniform(size=(P, n))).to(dev).requires_grad_(True)
    red = nslf._SlfReduceFn.apply(feat, w)
    up = f32(g.normal(size=(P, Fq))).to(dev)
    (red * up).sum().backward()
    out = new(P, Fq)
    xla_call("nrc_xla_slf_reduce_fwd", [feat.detach(), w.detach()], [out], J.pack_slf(P, cfg, num_features=Fq))
    assert same(out, red)
    gf, gw = new(P, n, Fq), new(P, n)
    xla_call("nrc_xla_slf_reduce_bwd", [feat.detach(), w.detach(), up], [gf, gw], J.pack_slf(P, cfg, num_features=Fq))
    assert same(gf, feat.grad) and same(gw, w.grad)
    # a descriptor of the wrong size is refused and reported
    xla_call("nrc_xla_slf_points_fwd", [raw.detach(), o, v], outs, J.pack_ggx(1, 1), expect_status=-1)
