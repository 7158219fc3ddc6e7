"""K4 parity: per-ray kernels vs the oracle (alpha weights, interval resampling, ray cast,
compositing + distance statistics, categorical resampling)."""
import numpy as np
import pytest
import torch

from oracle import coord as ocoord, ref_math, render as orender, stepfun as ostep
from neural_radiance_caching_b200 import render as nrender, stepfun as nstep, _lib
from tests.util import f32, gen, make_rays, rel_err, to_dev

pytestmark = pytest.mark.gpu


def _tdist(g, R, n, lo=2.0, hi=6.0):
    t = np.sort(g.uniform(lo, hi, size=(R, n + 1)), axis=-1)
    return f32(t)


@pytest.mark.parametrize("n", [1, 32, 64, 100])
def test_alpha_weights_fwd_bwd(cuda_device, n):
    g = gen(40 + n)
    R = 257
    dens = f32(g.gamma(0.5, 4.0, size=(R, n)))
    dens[0] = 0.0  # empty ray
    dens[1] = 1e6  # opaque ray
    t = _tdist(g, R, n)
    dirs = f32(g.normal(size=(R, 3)))
    gw, ga, gt = (f32(g.normal(size=(R, n))) for _ in range(3))
    do = dens.clone().requires_grad_(True)
    wo, ao, to = orender.compute_alpha_weights(do, t, dirs)
    ((wo * gw).sum() + (ao * ga).sum() + (to * gt).sum()).backward()
    dn = dens.to(cuda_device).requires_grad_(True)
    wn, an, tn = nrender.compute_alpha_weights(dn, t.to(cuda_device), dirs.to(cuda_device))
    ((wn * gw.to(cuda_device)).sum() + (an * ga.to(cuda_device)).sum() + (tn * gt.to(cuda_device)).sum()).backward()
    assert rel_err(wn, wo) <= 1e-5 and rel_err(an, ao) <= 1e-5 and rel_err(tn, to) <= 1e-5
    assert rel_err(dn.grad, do.grad) <= 1e-5


def test_alpha_weights_opaque_background(cuda_device):
    g = gen(45)
    R, n = 33, 32
    dens, t, dirs = f32(g.gamma(0.5, 1.0, size=(R, n))), _tdist(g, R, n), f32(g.normal(size=(R, 3)))
    wo, ao, to = orender.compute_alpha_weights(dens, t, dirs, opaque_background=True)
    wn, an, tn = nrender.compute_alpha_weights(dens.to(cuda_device), t.to(cuda_device), dirs.to(cuda_device),
                                               opaque_background=True)
    assert rel_err(wn, wo) <= 1e-5
    assert torch.allclose(wn.sum(-1).cpu(), torch.ones(R), atol=1e-5)


@pytest.mark.parametrize("m,n", [(1, 64), (64, 64), (64, 32), (32, 2), (100, 128)])
def test_sample_intervals(cuda_device, m, n):
    g = gen(50 + m + n)
    R = 300
    t = f32(np.sort(g.uniform(0, 1, size=(R, m + 1)), axis=-1))
    t[:, 0], t[:, -1] = 0.0, 1.0
    w = f32(g.gamma(0.3, 1.0, size=(R, m)))
    w[0] = 0.0            # all-zero weights -> uniform after padding
    w[1, 1:] = 0.0        # a single spike
    u01 = f32(g.uniform(size=(R, 1)))
    u01[2] = 0.0
    anneal, pad = 0.4, 1e-5
    logits = anneal * ref_math.safe_log(w + pad)
    want, idx_o = ostep.sample_intervals(u01, t, logits, n, single_jitter=True, domain=(0.0, 1.0), return_idx=True)
    got, idx_n = nstep.sample_intervals_from_weights(u01.to(cuda_device), t.to(cuda_device), w.to(cuda_device), n,
                                                     anneal=anneal, padding=pad, domain=(0.0, 1.0), return_bins=True)
    got, idx_n = got.cpu(), idx_n.cpu()
    assert got.shape == (R, n + 1)
    assert torch.all(got[:, 1:] >= got[:, :-1]), "output must be sorted"
    assert got.min() >= 0.0 and got.max() <= 1.0
    # (100,128) is a conditioning stress case: gamma(0.3) weights give near-empty bins where the
    # inverse CDF's slope amplifies last-ulp differences of the softmax/cumsum.
    assert float((got - want).abs().max()) <= (1e-5 if m <= 64 else 5e-5)
    # CDF bin indices: identical except where u sits within float rounding of a CDF knot
    # (the softmax/cumsum of oracle and kernel differ in the last ulp; DESIGN.md "Parity").
    mism = (idx_n.long() != idx_o).float().mean().item()
    assert mism <= 2e-3, f"{mism:.2%} bin mismatches"


def test_sample_intervals_generic_entry_and_errors(cuda_device):
    g = gen(60)
    R, m, n = 64, 16, 8
    t = f32(np.sort(g.uniform(0, 1, size=(R, m + 1)), axis=-1))
    logits = f32(g.normal(size=(R, m)))
    u01 = f32(g.uniform(size=(R, 1)))
    want = ostep.sample_intervals(u01, t, logits, n, single_jitter=True)
    got = nstep.sample_intervals(u01.to(cuda_device), t.to(cuda_device), logits.to(cuda_device), n,
                                 single_jitter=True).cpu()
    assert float((got - want).abs().max()) <= 1e-5
    with pytest.raises(ValueError):
        nstep.sample_intervals(u01.to(cuda_device), t.to(cuda_device), logits.to(cuda_device), 1, single_jitter=True)
    # the C ABI reports the same condition as a status code
    lib = _lib.load()
    assert lib.nrc_ray_sample_intervals(None, None, None, None, None, 4, 16, 1, 1.0, 0.0, 0.0, 0.0, 1.0, None,
                                        None) == -1


@pytest.mark.parametrize("use_raydist", [False, True])
def test_ray_cast(cuda_device, use_raydist):
    from neural_radiance_caching_b200.sampling import ProposalVolumeSampler

    g = gen(70)
    R, n = 200, 64
    rays = make_rays(g, R, near=0.05 if use_raydist else 2.0, far=2.0 if use_raydist else 6.0)
    s = f32(np.sort(g.uniform(0, 1, size=(R, n + 1)), axis=-1))
    s[:, 0], s[:, -1] = 0.0, 1.0
    if use_raydist:
        _, s_to_t = ocoord.power_ladder_warps(rays["near"], rays["far"], -1.5, 2.0)
    else:
        _, s_to_t = ocoord.construct_ray_warps(None, rays["near"], rays["far"])
    t_o = s_to_t(s)
    means_o, _ = orender.cast_rays(t_o, rays["origins"], rays["directions"], rays["radii"], "cone", diag=False)
    sampler = ProposalVolumeSampler()
    rd = to_dev(rays, cuda_device)
    t_n, means_n = sampler._cast(s.to(cuda_device), rd, use_raydist)
    assert rel_err(t_n, t_o) <= 1e-5
    assert rel_err(means_n, means_o) <= 1e-5
    # end points map back to near / far
    assert torch.allclose(t_n[:, 0].cpu(), rays["near"][:, 0], rtol=1e-5)
    assert torch.allclose(t_n[:, -1].cpu(), rays["far"][:, 0], rtol=1e-5)


@pytest.mark.parametrize("use_raydist", [False, True])
@pytest.mark.parametrize("m,n", [(1, 64), (64, 64), (64, 32), (100, 128)])
def test_sample_cast_fused_is_bit_identical(cuda_device, use_raydist, m, n):
    """nrc_ray_sample_cast (one launch per sampler level) against nrc_ray_sample_intervals + nrc_ray_cast, which the
    tests above hold to the oracle: resampled fenceposts, metric distances and means must agree bit for bit - also on
    rays whose fenceposts need the rank sort (near-duplicate CDF knots) and on the already-sorted fast path."""
    from neural_radiance_caching_b200.sampling import ProposalVolumeSampler

    g = gen(90 + m + n)
    R = 257
    rays = make_rays(g, R, near=0.05 if use_raydist else 2.0, far=2.0 if use_raydist else 6.0)
    t = f32(np.sort(g.uniform(0, 1, size=(R, m + 1)), axis=-1))
    t[:, 0], t[:, -1] = 0.0, 1.0
    w = f32(g.gamma(0.3, 1.0, size=(R, m)))
    w[0] = 0.0
    if m > 1:
        w[1, 1:] = 0.0                         # one spike: many samples land in one bin
        t[2, 1:-1] = t[2, 1:2]                 # degenerate bins: equal fenceposts -> ties in the sort
    u01 = f32(g.uniform(size=(R, 1)))
    sampler = ProposalVolumeSampler()
    rd = to_dev(rays, cuda_device)
    td, wd, ud = t.to(cuda_device), w.to(cuda_device), u01.to(cuda_device)
    anneal = 0.4
    s_ref = nstep.sample_intervals_from_weights(ud, td, wd, n, anneal=anneal, padding=sampler.resample_padding,
                                                domain=(0.0, 1.0))
    t_ref, means_ref = sampler._cast(s_ref, rd, use_raydist)
    s_new, t_new, means_new = sampler.sample_and_cast(ud, td, wd, n, anneal, rd, use_raydist)
    assert torch.all(s_new[:, 1:] >= s_new[:, :-1])
    assert torch.equal(s_new, s_ref)
    assert torch.equal(t_new, t_ref)
    assert torch.equal(means_new, means_ref)
    with pytest.raises(ValueError):
        sampler.sample_and_cast(ud, td, wd, 1, anneal, rd, use_raydist)


@pytest.mark.parametrize("n,C", [(32, 3), (64, 10), (32, 0)])
def test_volumetric_rendering(cuda_device, n, C):
    g = gen(80 + n + C)
    R = 129
    w = f32(g.dirichlet(np.ones(n) * 0.2, size=R) * g.uniform(0, 1, size=(R, 1)))
    w[0] = 0.0
    t = _tdist(g, R, n)
    rgbs = f32(g.uniform(size=(R, n, 3))) if C else None
    extras = {"normals": f32(g.normal(size=(R, n, 3))), "feat": f32(g.normal(size=(R, n, 4)))} if C > 3 else None
    bg = f32(g.uniform(size=(R, 3)))
    wo = w.clone().requires_grad_(True)
    ro = rgbs.clone().requires_grad_(True) if C else None
    want = orender.volumetric_rendering(ro, wo, wo, t, bg, True, extras=extras)
    wn = w.to(cuda_device).requires_grad_(True)
    rn = rgbs.to(cuda_device).requires_grad_(True) if C else None
    got = nrender.volumetric_rendering(rn, wn, wn, t.to(cuda_device), bg.to(cuda_device), True,
                                       extras=to_dev(extras, cuda_device) if extras else None)
    assert set(got.keys()) == set(want.keys())
    for k, v in want.items():
        if v is None:
            assert got[k] is None
            continue
        tol = 1e-5 if not k.startswith("distance") else 2e-5
        assert rel_err(got[k], v) <= tol, k
    if C:
        g_rgb, g_acc = f32(g.normal(size=(R, 3))), f32(g.normal(size=(R,)))
        ((want["rgb"] * g_rgb).sum() + (want["acc"] * g_acc).sum()).backward()
        ((got["rgb"] * g_rgb.to(cuda_device)).sum() + (got["acc"] * g_acc.to(cuda_device)).sum()).backward()
        assert rel_err(wn.grad, wo.grad) <= 1e-5
        assert rel_err(rn.grad, ro.grad) <= 1e-5


@pytest.mark.parametrize("use_mask", [False, True])
def test_render_loss_fused(cuda_device, use_mask):
    """nrc_render_loss (rendering + data term + mask loss + compositing VJP in one launch) against the four entry
    points it replaces in the training step, each of which is held to the oracle elsewhere: rgb and acc bit-identical,
    loss and gradients to rounding."""
    g = gen(333)
    R, n = 257, 32
    w = f32(g.dirichlet(np.ones(n) * 0.2, size=R) * g.uniform(0, 1.2, size=(R, 1)))   # some rays with acc > 1
    w[0] = 0.0
    vals = f32(g.uniform(size=(R, n, 3)))
    vals[1] = 0.0                                                                      # linear branch of the sRGB curve
    bg, target = f32(g.uniform(size=(R, 3))), f32(g.uniform(size=(R, 3)))
    t = _tdist(g, R, n)
    d = lambda a: a.to(cuda_device).contiguous()
    wd, vd, bd, td, tt = d(w), d(vals), d(bg), d(target), d(t)
    new = lambda *shape: torch.empty(shape, device=cuda_device, dtype=torch.float32)
    pad, ow, ew = 1e-3, 0.7, 0.3
    # reference sequence
    out_r, acc_r, dist_r = new(R, 3), new(R), new(R, 4)
    loss_r = torch.zeros((), device=cuda_device)
    _lib.call("nrc_ray_composite_fwd", _lib.stream_ptr(), _lib.ptr(vd), _lib.ptr(wd), n, None, _lib.ptr(tt), _lib.ptr(bd), R, n,
              3, 1, _lib.ptr(out_r), _lib.ptr(acc_r), _lib.ptr(dist_r))
    g_rgb = new(R, 3)
    _lib.call("nrc_charb_srgb_loss", _lib.stream_ptr(), _lib.ptr(out_r), _lib.ptr(td), R, pad, _lib.ptr(loss_r), _lib.ptr(g_rgb))
    g_acc = None
    if use_mask:
        g_acc = new(R)
        _lib.call("nrc_mask_loss", _lib.stream_ptr(), _lib.ptr(acc_r), 0, None, R, pad, ow, ew, _lib.ptr(loss_r), _lib.ptr(g_acc))
    gv_r, gw_r = new(R, n, 3), new(R, n)
    _lib.call("nrc_ray_composite_bwd", _lib.stream_ptr(), _lib.ptr(vd), _lib.ptr(wd), n, None, _lib.ptr(bd), _lib.ptr(g_rgb),
              _lib.ptr(g_acc), R, n, 3, 1, _lib.ptr(gv_r), _lib.ptr(gw_r), None)
    # fused
    out_f, acc_f, gv_f, gw_f = new(R, 3), new(R), new(R, n, 3), new(R, n)
    loss_f = torch.zeros((), device=cuda_device)
    _lib.call("nrc_render_loss", _lib.stream_ptr(), _lib.ptr(vd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(td), None, R, n, pad,
              1 if use_mask else 0, ow, ew, _lib.ptr(loss_f), _lib.ptr(out_f), _lib.ptr(acc_f), _lib.ptr(gv_f), _lib.ptr(gw_f))
    torch.cuda.synchronize()
    assert torch.equal(out_f, out_r) and torch.equal(acc_f, acc_r)
    assert abs(float(loss_f) - float(loss_r)) <= 1e-6 * abs(float(loss_r))
    assert rel_err(gv_f, gv_r) <= 1e-6
    assert rel_err(gw_f, gw_r) <= 1e-6
    lib = _lib.load()
    assert lib.nrc_render_loss(None, None, None, None, None, None, 4, 8, pad, 0, 0.0, 0.0, None, None, None, None, None) == -1


@pytest.mark.parametrize("n,use_mask", [(32, True), (100, False)])
def test_shade_render_loss_fused(cuda_device, n, use_mask):
    """nrc_shade_render_loss (the shader's `out` stage + rendering + losses + both VJPs in one launch) against the three
    launches it replaces in the training step (nrc_shader_out_fwd -> nrc_render_loss -> nrc_shader_out_bwd), which are
    held to the oracle elsewhere: per-sample colours, rgb and acc bit-identical, loss and gradients to rounding."""
    g = gen(334 + n)
    R = 129
    P = R * n
    w = f32(g.dirichlet(np.ones(n) * 0.2, size=R) * g.uniform(0, 1.2, size=(R, 1)))
    w[0] = 0.0
    heads, fraw, slf, env = (f32(g.normal(size=(P, 16)) * 2.0) for _ in range(4))
    heads[:n, 1:7] = 40.0                    # clipped diffuse branch (rgb_max)
    bg, target = f32(g.uniform(size=(R, 3))), f32(g.uniform(size=(R, 3)))
    d = lambda a: a.to(cuda_device).contiguous()
    wd, hd, fd, sd, ed, bd, td = d(w), d(heads), d(fraw), d(slf), d(env), d(bg), d(target)
    new = lambda *shape: torch.zeros(shape, device=cuda_device, dtype=torch.float32)
    pad, ow, ew = 1e-3, 0.7, 0.3
    consts = (8.0, -2.0, 0.3, float(np.log(3.0)))
    st = _lib.stream_ptr
    # reference sequence
    rgb_r, loss_r = new(P, 3), new()
    _lib.call("nrc_shader_out_fwd", st(), _lib.ptr(hd), 16, _lib.ptr(fd), 16, _lib.ptr(sd), 16, _lib.ptr(ed), 16, P, *consts,
              _lib.ptr(rgb_r), None)
    out_r, acc_r, gv_r, gw_r = new(R, 3), new(R), new(R, n, 3), new(R, n)
    _lib.call("nrc_render_loss", st(), _lib.ptr(rgb_r), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(td), None, R, n, pad,
              1 if use_mask else 0, ow, ew, _lib.ptr(loss_r), _lib.ptr(out_r), _lib.ptr(acc_r), _lib.ptr(gv_r), _lib.ptr(gw_r))
    gh_r, gf_r, gs_r = new(P, 16), new(P, 16), new(P, 16)
    _lib.call("nrc_shader_out_bwd", st(), _lib.ptr(hd), 16, _lib.ptr(fd), 16, _lib.ptr(sd), 16, P, *consts, _lib.ptr(gv_r),
              _lib.ptr(gh_r), 16, _lib.ptr(gf_r), 16, _lib.ptr(gs_r), 16)
    # fused
    rgb_f, loss_f, out_f, acc_f, gw_f = new(P, 3), new(), new(R, 3), new(R), new(R, n)
    gh_f, gf_f, gs_f = new(P, 16), new(P, 16), new(P, 16)
    _lib.call("nrc_shade_render_loss", st(), _lib.ptr(hd), 16, _lib.ptr(fd), 16, _lib.ptr(sd), 16, _lib.ptr(ed), 16, *consts,
              _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(td), None, R, n, pad, 1 if use_mask else 0, ow, ew, _lib.ptr(loss_f),
              _lib.ptr(rgb_f), _lib.ptr(out_f), _lib.ptr(acc_f), _lib.ptr(gw_f), _lib.ptr(gh_f), 16, _lib.ptr(gf_f), 16,
              _lib.ptr(gs_f), 16)
    torch.cuda.synchronize()
    assert torch.equal(rgb_f, rgb_r) and torch.equal(out_f, out_r) and torch.equal(acc_f, acc_r)
    assert abs(float(loss_f) - float(loss_r)) <= 1e-6 * abs(float(loss_r))
    assert rel_err(gw_f, gw_r) <= 1e-6
    for a, b in ((gh_f, gh_r), (gf_f, gf_r), (gs_f, gs_r)):
        assert rel_err(a, b) <= 1e-6
    lib = _lib.load()
    assert lib.nrc_shade_render_loss(None, None, 16, None, 16, None, 16, None, 16, *consts, None, None, None, None, 4, 8, pad, 0,
                                     0.0, 0.0, None, None, None, None, None, None, 16, None, 16, None, 16) == -1


@pytest.mark.parametrize("m,n,opaque", [(64, 64, 0), (64, 32, 0), (100, 128, 1)])
def test_weights_sample_cast_fused(cuda_device, m, n, opaque):
    """nrc_ray_weights_sample_cast == nrc_ray_alpha_weights_fwd followed by nrc_ray_sample_cast, bit for bit (weights,
    resampled fenceposts, metric distances, means)."""
    g = gen(900 + m + n)
    R = 259
    rays = to_dev(make_rays(g, R), cuda_device)
    dens = f32(g.gamma(0.5, 4.0, size=(R, m)))
    dens[0] = 0.0
    dens[1] = 1e6
    sd = f32(np.sort(g.uniform(size=(R, m + 1)), axis=-1))
    sd[:, 0], sd[:, -1] = 0.0, 1.0
    tprev = _tdist(g, R, m)
    u = f32(g.uniform(size=(R,)))
    d = lambda a: a.to(cuda_device).contiguous()
    dd, sdd, tp, ud = d(dens), d(sd), d(tprev), d(u)
    base, max_jitter = nstep.u_base(n, cuda_device)
    new = lambda *shape: torch.empty(shape, device=cuda_device, dtype=torch.float32)
    st = _lib.stream_ptr
    common = (R, m, n, 0.7, 0.01, max_jitter, 0.0, 1.0, _lib.ptr(rays["origins"]), _lib.ptr(rays["directions"]),
              _lib.ptr(rays["near"]), _lib.ptr(rays["far"]), 1, -1.5, 2.0)
    w_r, s_r, t_r, m_r = new(R, m), new(R, n + 1), new(R, n + 1), new(R, n, 3)
    _lib.call("nrc_ray_alpha_weights_fwd", st(), _lib.ptr(dd), _lib.ptr(tp), _lib.ptr(rays["directions"]), R, m, opaque,
              _lib.ptr(w_r), None, None)
    _lib.call("nrc_ray_sample_cast", st(), _lib.ptr(sdd), _lib.ptr(w_r), _lib.ptr(ud), _lib.ptr(base), *common,
              _lib.ptr(s_r), _lib.ptr(t_r), _lib.ptr(m_r))
    w_f, s_f, t_f, m_f = new(R, m), new(R, n + 1), new(R, n + 1), new(R, n, 3)
    _lib.call("nrc_ray_weights_sample_cast", st(), _lib.ptr(sdd), _lib.ptr(dd), _lib.ptr(tp), opaque, _lib.ptr(w_f), _lib.ptr(ud),
              _lib.ptr(base), *common, _lib.ptr(s_f), _lib.ptr(t_f), _lib.ptr(m_f))
    torch.cuda.synchronize()
    assert torch.equal(w_f, w_r) and torch.equal(s_f, s_r) and torch.equal(t_f, t_r) and torch.equal(m_f, m_r)
