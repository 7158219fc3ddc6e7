"""Golden vectors from the REFERENCE'S OWN SOURCE (/root/reference/internal/*.py), executed in this container under the
NumPy stand-in for jax (tests/golden/jax_numpy_shim.py - JAX itself is not installable here).  Run from the repo root:

    python tests/golden/make_reference_vectors.py          # writes tests/golden/reference_np.npz

The file pins the oracle (tests/test_reference_vectors.py, CPU) and, through the oracle-independent GPU test, the
kernels to what the reference's code computes on the same float32 inputs.  /root/reference exists only in the build
container; the .npz travels."""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jax_numpy_shim as shim  # noqa: E402

REF = os.environ.get("NRC_REFERENCE", "/root/reference")


def load_reference():
    shim.install()
    sys.path.insert(0, REF)
    importlib.import_module("internal")
    # internal/utils.py itself is the reference's (its tensorflow / cv2 / flax / PIL imports are pass-through stubs);
    # utils.device_is_tpu() evaluates to False under the shim (no device is a "tpu").
    names = ["utils", "math", "linspline", "stepfun", "coord", "render", "ref_utils", "grid_utils", "loss_utils", "image",
             "geometry", "shading", "nerf", "sampling", "models", "material", "light_sampler", "train_utils"]
    mods = {n: importlib.import_module("internal." + n) for n in names}
    mods["render_utils"] = importlib.import_module("internal.inverse_render.render_utils")
    return mods


def main():
    R = load_reference()
    rmath, rstep, rcoord, rrender, rgrid, rref, rlin = (R[k] for k in ("math", "stepfun", "coord", "render", "grid_utils",
                                                                        "ref_utils", "linspline"))
    g = np.random.Generator(np.random.PCG64(20200823))
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    out = {}

    # ---- internal/math.py ------------------------------------------------------------------------------------------
    x = f(np.concatenate([g.normal(size=200) * 3, [0.0, -0.0, 1e-30, -1e-30, 50.0, 100.0, -100.0, 1e6]]))
    out["math_x"] = x
    out["math_safe_exp"] = rmath.safe_exp(x)
    out["math_safe_log"] = rmath.safe_log(np.abs(x))
    out["math_safe_sign"] = rmath.safe_sign(x)
    for p, pre in ((-1.5, 2.0), (-0.25, 1e4)):
        xs = f(np.abs(g.normal(size=256)) * 3)
        p32, pre32 = np.float32(p), np.float32(pre)        # JAX weak typing keeps everything float32
        y = rmath.power_ladder(xs, p32, premult=pre32)
        out[f"math_pl_x_{p}"] = xs
        out[f"math_pl_y_{p}"] = y
        out[f"math_ipl_{p}"] = rmath.inv_power_ladder(y, p32, premult=pre32)

    # ---- internal/coord.py -----------------------------------------------------------------------------------------
    pts = f(g.normal(size=(512, 3)) * 2.5)
    pts[:4] = [[0, 0, 0], [1, 0, 0], [0.3, 0.4, 0.5], [10, -20, 30]]
    out["coord_x"] = pts
    out["coord_contract"] = rcoord.contract(pts)
    out["coord_contract_radius_2"] = rcoord.contract_radius_2(pts)
    out["coord_contract_radius_5"] = rcoord.contract_radius_5(pts)

    # ---- internal/stepfun.py ---------------------------------------------------------------------------------------
    nr, m, n = 64, 64, 32
    t = f(np.sort(g.uniform(0, 1, size=(nr, m + 1)), axis=-1)); t[:, 0], t[:, -1] = 0.0, 1.0
    w = f(g.gamma(0.3, 1.0, size=(nr, m))); w[0] = 0.0; w[1, 1:] = 0.0
    logits = f(0.4 * np.asarray(rmath.safe_log(w + np.float32(1e-5))))
    u01 = f(g.uniform(size=(nr, 1)))
    out.update(step_t=t, step_w=w, step_logits=logits, step_u01=u01)
    wsm = np.asarray(shim.nn_mod.softmax(logits, axis=-1), np.float32)
    out["step_integrate_weights"] = rstep.integrate_weights(wsm)
    out["step_sample_intervals"] = rstep.sample_intervals(u01, t, logits, n, single_jitter=True, domain=(0.0, 1.0))
    out["step_sample_centres"] = rstep.sample(u01, t, logits, n, single_jitter=True)
    tm = f(np.sort(g.uniform(2, 6, size=(nr, n + 1)), axis=-1))
    wn = f(g.dirichlet(np.ones(n) * 0.3, size=nr) * g.uniform(0.2, 1, size=(nr, 1)))
    out.update(dist_t=tm, dist_w=wn)
    out["step_lossfun_distortion"] = rstep.lossfun_distortion(tm, wn)
    out["step_weighted_percentile"] = rstep.weighted_percentile(tm, wn, [5, 50, 95])
    tq = f(np.sort(g.uniform(0, 1, size=(nr, 48 + 1)), axis=-1)); tq[:, 0], tq[:, -1] = 0.0, 1.0
    tb = f(np.sort(g.uniform(0, 1, size=(nr, n + 1)), axis=-1)); tb[:, 0], tb[:, -1] = 0.0, 1.0
    out.update(blur_tq=tq, blur_t=tb)
    for hw in (0.03, 0.003):
        out[f"step_blur_and_resample_{hw}"] = rstep.blur_and_resample_weights(tq, tb, wn, np.float32(hw))

    # ---- internal/render.py ----------------------------------------------------------------------------------------
    dens = f(np.exp(g.normal(size=(nr, n)) * 2))
    dirs = f(g.normal(size=(nr, 3)))
    out.update(render_density=dens, render_dirs=dirs)
    for opaque in (False, True):
        wts, alpha, trans = rrender.compute_alpha_weights(dens, tm, dirs, opaque_background=opaque)[:3]
        out[f"render_weights_{int(opaque)}"] = wts
        out[f"render_alpha_{int(opaque)}"] = alpha
        out[f"render_trans_{int(opaque)}"] = trans
    origins = f(g.normal(size=(nr, 3)) * 4)
    radii = f(np.full((nr, 1), 5e-4))
    means, covs = rrender.cast_rays(tm, origins, dirs, radii, "cone", diag=False)
    out.update(render_origins=origins, render_means=means, render_radii=radii, render_covs=covs)
    # covariances of the other three variants (diagonal; cylinder rays, render.py:84-103), with per-ray radii
    radii_v = (radii * (1.0 + np.arange(nr, dtype=np.float32)[:, None] * np.float32(37.0))).astype(np.float32)
    out["render_radii_v"] = radii_v
    out["render_covs_diag"] = rrender.cast_rays(tm, origins, dirs, radii_v, "cone", diag=True)[1]
    cyl_m, cyl_c = rrender.cast_rays(tm, origins, dirs, radii_v, "cylinder", diag=False)
    out.update(render_cyl_means=cyl_m, render_cyl_covs=cyl_c,
               render_cyl_covs_diag=rrender.cast_rays(tm, origins, dirs, radii_v, "cylinder", diag=True)[1])
    rgbs = f(g.uniform(size=(nr, n, 3)))
    bg = f(g.uniform(size=(nr, 3)))
    vr = rrender.volumetric_rendering(rgbs, wn, wn, tm, bg, True)
    out.update(render_rgbs=rgbs, render_bg=bg)
    for k in ("rgb", "acc", "distance_mean", "distance_median", "distance_percentile_5", "distance_percentile_95"):
        if k in vr:
            out["render_vr_" + k] = vr[k]

    # ---- internal/grid_utils.py: trilerp on a hash level and on a dense level --------------------------------------
    T, F, N = 4096, 4, 16
    table = f(g.normal(size=(T, F)))
    grid = f(g.normal(size=(N, N, N, F)))
    loc = f(g.uniform(-0.2, 1.2, size=(777, 3)))          # in units of the unit cube; includes out-of-range points
    out.update(grid_table=table, grid_dense=grid, grid_loc=loc)
    for res in (64, 256):
        out[f"grid_hash_{res}"] = rgrid.trilerp(table, loc * np.float32(res), "hash", rgrid.ResampleOpMode.DEFAULT_JAX)
    out["grid_dense_16"] = rgrid.trilerp(grid, loc * np.float32(N), "grid", rgrid.ResampleOpMode.DEFAULT_JAX)

    # ---- internal/ref_utils.py -------------------------------------------------------------------------------------
    v = f(g.normal(size=(300, 3))); v[0] = 0.0; v[1] = [1e-20, 0, 0]
    out["ref_l2n_x"] = v
    out["ref_l2n"] = rref.l2_normalize(v)
    d = v[2:] / np.linalg.norm(v[2:], axis=-1, keepdims=True)
    kappa_inv = f(g.uniform(0.01, 1.0, size=(d.shape[0], 1)))
    out.update(ide_dirs=f(d), ide_kappa_inv=kappa_inv)
    for deg in (4, 5):
        try:
            out[f"ref_ide_{deg}"] = rref.generate_ide_fn(deg)(f(d), kappa_inv)
        except Exception as e:   # recorded, not hidden
            print(f"generate_ide_fn({deg}) not runnable under the shim: {type(e).__name__}: {e}")

    # ---- internal/image.py, internal/loss_utils.py ----------------------------------------------------------------
    rimage, rloss, rru = R["image"], R["loss_utils"], R["render_utils"]
    lin = f(np.concatenate([g.uniform(0, 1, size=500), [0.0, 0.0031308, 0.0031309, 1e-9, 1.0, 2.5]]))
    out["image_linear"] = lin
    out["image_linear_to_srgb"] = rimage.linear_to_srgb(lin)
    one = np.float32(1.0)
    w0 = f(g.dirichlet(np.ones(m) * 0.3, size=nr) * g.uniform(0.2, 1, size=(nr, 1)))
    w1 = f(g.dirichlet(np.ones(48) * 0.3, size=nr) * g.uniform(0.2, 1, size=(nr, 1)))
    hist = [dict(sdist=t, weights=w0, lossmult=one), dict(sdist=tq, weights=w1, lossmult=one),
            dict(sdist=tb, weights=wn, lossmult=one, tdist=tm)]
    out.update(loss_w0=w0, loss_w1=w1)
    out["loss_spline_interlevel"] = np.stack(rloss.spline_interlevel_loss(hist, mults=(0.01, 0.01), blurs=(0.03, 0.003)))
    out["loss_distortion"] = np.asarray(rloss.distortion_loss(
        hist, target="tdist", mult=np.float32(0.01),
        curve_fn=lambda x: rmath.power_ladder(x, np.float32(-0.25), premult=np.float32(1e4))))

    # ---- internal/inverse_render/render_utils.py: GGX lobe, Monte Carlo integration, frames, vMF ---------------------
    P_, S_ = 96, 16
    unit = lambda a: a / np.linalg.norm(a, axis=-1, keepdims=True)
    wi = f(unit(g.normal(size=(P_, S_, 3)))); wi[:, :, 2] = np.abs(wi[:, :, 2]) * np.where(g.uniform(size=(P_, S_)) < 0.9, 1, -1)
    wo = f(unit(g.normal(size=(P_, 1, 3)))); wo[..., 2] = np.abs(wo[..., 2])
    material = dict(albedo=f(g.uniform(size=(P_, 3))), roughness=f(g.uniform(0.01, 1.0, size=(P_, 1))),
                    F_0=f(np.full((P_, 1), 0.04)), metalness=f(g.uniform(size=(P_, 1))))
    samples = dict(local_lightdirs=wi, local_viewdirs=np.broadcast_to(wo, wi.shape).copy(),
                   brdf_correction=f(np.ones((P_, S_, 2))), pdf=f(g.uniform(0.0, 3.0, size=(P_, S_, 1))),
                   weight=f(g.uniform(-0.1, 1.0, size=(P_, S_, 1))), radiance_in=f(g.uniform(0, 4, size=(P_, S_, 3))),
                   indirect_occ=f(g.uniform(size=(P_, S_, 1))))
    for k_, v_ in material.items():
        out["ggx_mat_" + k_] = v_
    for k_, v_ in samples.items():
        out["ggx_smp_" + k_] = v_
    cth, rough = f(g.uniform(0, 1, size=512)), f(g.uniform(0.01, 1, size=512))
    out.update(ggx_costheta=cth, ggx_a=rough)
    out["ggx_D"] = rru.GGX_D(cth, rough)
    res = rru.integrate_reflect_rays("microfacet", False, dict(material), dict(samples))
    for k_ in ("radiance_out", "indirect_occ", "irradiance"):
        out["ggx_int_" + k_] = res[k_]
    nrm = f(unit(g.normal(size=(256, 3)))); nrm[0] = [0, 0, 1]; nrm[1] = [0, 1, 0]; nrm[2] = [0.1, 0.2, 0.97]
    nrm = f(unit(nrm))
    out["rot_normal"] = nrm
    out["rot_matrix"] = rru.get_rotation_matrix(nrm)
    vx = f(unit(g.normal(size=(200, 8, 3)))); vmu = f(unit(g.normal(size=(200, 8, 3))))
    vk = f(g.uniform(0, 50, size=(200, 8))); vk[0] = 0.0
    out.update(vmf_x=vx, vmf_means=vmu, vmf_kappa=vk)
    out["vmf_eval"] = rru.eval_vmf(vx, vmu, vk)

    # ---- importance samplers (render_utils.py:417-546, 1431-1490) and the light-sampling loss (:1493-1550) -------------
    Ps, Ss = 200, 16
    su1, su2 = f(g.uniform(size=(Ps, Ss))), f(g.uniform(size=(Ps, Ss)))
    su1[0, :4] = [0.0, 1.0 - 2.0 ** -24, 0.5, 1e-7]
    swo = f(unit(g.normal(size=(Ps, 1, 3)))); swo[..., 2] = np.abs(swo[..., 2]) + 0.02
    swo = f(np.broadcast_to(unit(swo), (Ps, Ss, 3)).copy())
    salpha = f(np.broadcast_to(g.uniform(0.01, 1.0, size=(Ps, 1, 1)) ** 2, (Ps, Ss, 1)).copy())
    swi = f(unit(g.normal(size=(Ps, Ss, 3))))
    out.update(smp_u1=su1, smp_u2=su2, smp_wo=swo, smp_alpha=salpha, smp_wi=swi)
    for name, cls in (("cosine", rru.CosineSampler), ("microfacet", rru.MicrofacetSampler)):
        smp = cls()
        dirs, pdf = smp.sample_directions(None, su1, su2, swo, salpha, None, {})
        out[f"smp_{name}_dirs"], out[f"smp_{name}_pdf"] = dirs, pdf
        out[f"smp_{name}_pdf_of_wi"] = smp.pdf(swo, swi, salpha, {})
    K_ = 8
    lmeans = f(g.normal(size=(Ps, K_, 3)) * g.uniform(0.2, 3.0, size=(Ps, K_, 1)))
    lkappas = f(g.uniform(0.0, 40.0, size=(Ps, K_, 1))); lkappas[0, 0] = 0.0
    llogits = f(g.normal(size=(Ps, K_, 1)))
    out.update(light_means=lmeans, light_kappas=lkappas, light_logits=llogits)
    out["smp_light_pdf_of_wi"] = rru.LightSampler().pdf(swo, swi, salpha, dict(vmf_means=lmeans, vmf_kappas=lkappas,
                                                                                  vmf_logits=llogits))
    lnormals = f(unit(g.normal(size=(Ps, 3))))
    lpdf = f(g.uniform(0.0, 2.0, size=(Ps, Ss, 1))); lwt = f(g.uniform(-0.5, 12.0, size=(Ps, Ss, 1)))
    lfv = f(g.gamma(1.0, 1.0, size=(Ps, Ss))); lmult = f(np.full((Ps, Ss), 1.0 / Ss))
    out.update(light_normals=lnormals, light_pdf=lpdf, light_weight=lwt, light_fv=lfv, light_lossmult=lmult)
    for srgb in (True, False):
        out[f"light_vmf_loss_{int(srgb)}"] = np.asarray(rru.vmf_loss_fn(
            (lmeans, lkappas, llogits), lnormals, swi, dict(pdf=lpdf, weight=lwt), lfv, lfv, lmult, linear_to_srgb=srgb))

    # ---- time-resolved rendering: render.py:250-507, render_utils.zero_invalid_bins (:1699-1767) ----------------------
    import types as _types
    Rt, nt, Bt, Ct, expo = 24, 8, 96, 3, np.float32(0.01)
    tw = f(g.dirichlet(np.ones(nt) * 0.4, size=Rt) * g.uniform(0.3, 1.0, size=(Rt, 1)))
    ttd = f(np.sort(g.uniform(0.05, 1.0, size=(Rt, nt + 1)), axis=-1))
    tdirect = f(g.uniform(0, 2, size=(Rt, nt, Ct)))
    tind = f(g.gamma(1.0, 1.0, size=(Rt, nt, Bt, Ct)))
    tray = f(g.uniform(0.0, 0.55, size=(Rt, nt, 1))); tlight = f(g.uniform(0.0, 0.55, size=(Rt, nt, 1)))
    tray[0, 0], tlight[0, 0] = 0.25, 0.13              # integer bin: floor == ceil
    tray[-1, -1], tlight[-1, -1] = 0.55, 0.47          # past the last bin of the last ray: dropped, not wrapped
    tray[1, 1], tlight[1, 1] = 0.5, 0.49               # past the last bin of an inner ray: lands in the NEXT ray's bins
    out.update(tr_weights=tw, tr_tdist=ttd, tr_direct=tdirect, tr_indirect=tind, tr_ray_dists=tray, tr_light_dists=tlight)
    cfg = _types.SimpleNamespace(no_shift_direct=False, vis_only=False)
    for shift_ in (0.0, 0.0137):
        res = rrender.volumetric_transient_rendering(
            tdirect, tind, tw, tw, ttd, None, False, extras=dict(light_dists=tlight, ray_dists=tray, transient_indirect=None),
            n_bins=Bt, shift=np.float32(shift_), dark_level=np.float32(0.001), exposure_time=expo, config=cfg)
        tag = "tr_s%d_" % int(shift_ > 0)
        for k_ in ("transient_direct", "transient_indirect", "rgb"):
            out[tag + k_] = res[k_]
    out["tr_shift_direct"] = rrender.shift_direct((tray + tlight)[..., 0] / expo, tdirect, tw, Bt, Ct, None)
    out["tr_shift_map"] = rrender.shift_map_coordinates(tind.reshape(-1, Bt, Ct), tray.reshape(-1), expo, Bt, Ct)
    tmeans = f(g.uniform(-0.12, 0.12, size=(Rt, nt, 3)))
    trays = _types.SimpleNamespace(lights=f(g.uniform(-0.12, 0.12, size=(Rt, 3))), origins=f(g.uniform(-0.12, 0.12, size=(Rt, 3))),
                                   cam_origins=f(g.uniform(-0.12, 0.12, size=(Rt, 3))))
    tspec = f(g.gamma(1.0, 1.0, size=(Rt, nt, Bt, Ct)))
    out.update(tr_means=tmeans, tr_lights=trays.lights, tr_origins=trays.origins, tr_cam_origins=trays.cam_origins, tr_spec=tspec)
    for lz in (False, True):
        zc = _types.SimpleNamespace(n_bins=Bt, bin_zero_threshold_light=np.float32(2.0), exposure_time=expo, light_zero=lz,
                                    light_near=np.float32(0.12))
        zd, zs = rru.zero_invalid_bins(tind, tspec, trays, tmeans, zc)
        out[f"tr_zero_diffuse_{int(lz)}"], out[f"tr_zero_specular_{int(lz)}"] = zd, zs

    # ---- Model.maybe_resample (models.py:193-292), resample_argmax off, as configured (ngp_yobo.gin:369) ----------------
    Rr, nr_, kr = 128, 32, 4
    rw = f(g.dirichlet(np.ones(nr_) * 0.2, size=Rr) * g.uniform(0.0, 1.0, size=(Rr, 1))); rw[0] = 0.0
    gum = f(g.gumbel(size=(Rr, nr_, kr)))
    sres = dict(points=f(g.normal(size=(Rr, nr_, 3))), weights=rw, tdist=f(np.sort(g.uniform(size=(Rr, nr_ + 1)), -1)),
                sdist=f(np.sort(g.uniform(size=(Rr, nr_ + 1)), -1)), feature=f(g.normal(size=(Rr, nr_, 5))))
    out.update(rs_weights=rw, rs_gumbel=gum, rs_points=sres["points"], rs_feature=sres["feature"])
    for kk, bias in ((1, 0.0), (kr, 1e-3)):
        self_ = _types.SimpleNamespace(weights_bias=np.float32(bias), resample_argmax=False)
        fres, inds = R["models"].Model.maybe_resample(self_, gum[..., :kk], True, dict(sres), kk)
        out[f"rs_inds_{kk}"], out[f"rs_new_weights_{kk}"] = inds, fres["weights"]
        out[f"rs_new_points_{kk}"], out[f"rs_new_feature_{kk}"] = fres["points"], fres["feature"]
        assert fres["weights_no_filter"] is rw and fres["tdist"] is sres["tdist"]

    # ---- geometry losses (loss_utils.py:127-199) and the mask loss (train_utils.py:785-836) ----------------------------
    Rg, ng = 96, 32
    gw = f(g.dirichlet(np.ones(ng) * 0.3, size=Rg) * g.uniform(0.2, 1, size=(Rg, 1)))
    gn = f(unit(g.normal(size=(Rg, ng, 3)))); gnp = f(unit(gn + 0.4 * g.normal(size=(Rg, ng, 3))))
    gn[0, 0] = np.nan                                      # nan_to_num'd by both losses
    gview = f(unit(g.normal(size=(Rg, 3))))
    out.update(gl_weights=gw, gl_normals=gn, gl_normals_pred=gnp, gl_viewdirs=gview)
    rr = dict(weights=gw, lossmult=one, normals=gn, normals_pred=gnp)
    grays = _types.SimpleNamespace(viewdirs=gview, lossmult=f(np.ones((Rg, 1))))
    beta = f(np.ones((Rg, ng, 1)))
    out["gl_orientation"] = rloss.orientation_loss(grays, rr, target="normals_pred", mult=np.float32(0.01))
    out["gl_predicted_normal"] = rloss.predicted_normal_loss(rr, beta, mult=np.float32(0.001), gt="normals_pred", pred="normals",
                                                             stopgrad=False, stopgrad_weight=0.1)
    out["gl_predicted_normal_reverse"] = rloss.predicted_normal_loss(rr, beta, mult=np.float32(0.01), gt="normals",
                                                                     pred="normals_pred", stopgrad=True)
    gacc = f(g.uniform(0, 1, size=(Rg,))); gmask = f((g.uniform(size=(Rg, 1)) > 0.4))
    out.update(gl_acc=gacc, gl_masks=gmask)
    mcfg = _types.SimpleNamespace(charb_padding=np.float32(0.001), opaque_loss_weight=np.float32(1.0), empty_loss_weight=np.float32(10.0),
                                  use_mask_weight_decay=False, mask_weight_decay_start=0.0, mask_weight_decay_frac=0.1,
                                  mask_weight_decay_min=0.0, use_mask_weight_ease=False, mask_weight_ease_start=0.0,
                                  mask_weight_ease_frac=0.1, mask_weight_ease_min=0.0)
    rtrain = R["train_utils"]
    out["gl_mask_loss"] = rtrain.compute_mask_loss(_types.SimpleNamespace(masks=gmask), dict(acc=gacc), grays, mcfg)
    out["gl_mask_loss_none"] = rtrain.compute_mask_loss(_types.SimpleNamespace(masks=None), dict(acc=gacc), grays, mcfg)
    out["gl_mask_loss_backward"] = rtrain.compute_mask_loss(_types.SimpleNamespace(masks=f(np.zeros((Rg, 1)))), dict(acc=gacc),
                                                            grays, mcfg, empty_loss_weight=np.float32(0.5))

    # ---- HashEncoding.__call__ (grid_utils.py:738-905): level schedule, dense / hash choice, parameter names, bbox map,
    #      per-level multisample mean, precondition scaling.  Tables are a closed form of the entry index (level_table
    #      below, repeated in the test) so that only inputs and outputs are stored. ------------------------------------
    def level_table(shape, salt):
        idx = np.arange(int(np.prod(shape)), dtype=np.uint64)
        h = (idx * np.uint64(2654435761) + np.uint64(salt) * np.uint64(40503)) % np.uint64(1 << 32)
        return ((h.astype(np.float64) / float(1 << 32) - 0.5) * 2e-2).astype(np.float32).reshape(shape)

    class _Enc(rgrid.HashEncoding):
        def param(self, name, init_fn):
            self.seen.append(name)
            shape = init_fn.keywords["shape"]
            return level_table(shape, len(self.seen))

    hx = f(g.normal(size=(600, 1, 3)) * 1.2); hx[0, 0] = [-2.0, 2.0, 0.0]; hx[1, 0] = [2.5, -2.5, 0.3]   # corners, outside
    out["enc_x"] = hx
    ENC_A = dict(hash_map_size=2 ** 15, num_features=2, scale_supersample=1.0, max_grid_size=256)
    for tag, kw in (("a", ENC_A),
                    ("b", dict(hash_map_size=2 ** 12, num_features=4, scale_supersample=1.0, max_grid_size=128,
                               precondition_scaling=1.0, bbox_scaling=((-1.0, -2.0, -3.0), (1.5, 2.0, 2.5))))):
        enc = _Enc()
        for k_, v_ in kw.items():
            setattr(enc, k_, v_)
        enc.seen = []
        feats = enc(hx.view(shim.F32Array), per_level_fn=rmath.average_across_multisamples)
        assert feats.dtype == np.float32
        out[f"enc_{tag}_features"] = np.asarray(feats)
        out[f"enc_{tag}_names"] = np.array(enc.seen)
        out[f"enc_{tag}_grid_sizes"] = np.asarray(enc.grid_sizes)

    # ---- importance_sample_rays (render_utils.py:722-924): shading frame, per-sampler draws, MIS power heuristic with
    #      energy correction, for the two-sampler and the three-sampler (vMF light sampler) set-ups -----------------------
    class _Pre2D:                                          # random_generator_2d: hands out the pre-drawn (uh, uw)
        def __init__(self, draws):
            self.draws = list(draws)

        def sample(self, rng, n_, stratified):
            uh, uw = self.draws.pop(0)
            assert uh.size == n_
            return uh.reshape(-1), uw.reshape(-1)

    Pi, counts = 160, (16, 8, 8)
    iv = f(unit(g.normal(size=(Pi, 3)))); inrm = f(unit(iv + 0.8 * g.normal(size=(Pi, 3))))   # mostly front facing
    irough = f(g.uniform(0.05, 1.0, size=(Pi, 1)) ** 2)
    iu = [(f(g.uniform(size=(Pi, c_))), f(g.uniform(size=(Pi, c_)))) for c_ in counts]
    ikey = dict(gumbel=f(g.gumbel(size=(Pi, K_))), normal=f(g.normal(size=(Pi, counts[2], 2))),
                uniform=f(g.uniform(size=(Pi, counts[2]))))
    laux = dict(vmf_means=lmeans[:Pi], vmf_kappas=lkappas[:Pi], vmf_logits=llogits[:Pi])
    out.update(is_viewdirs=iv, is_normals=inrm, is_roughness=irough, is_light_normal2=ikey["normal"], is_light_u=ikey["uniform"],
               is_light_latent=np.argmax(ikey["gumbel"] + llogits[:Pi, :, 0], axis=-1).astype(np.int32))
    for j_, (uh_, uw_) in enumerate(iu):
        out[f"is_uh_{j_}"], out[f"is_uw_{j_}"] = uh_, uw_
    smps = [(rru.MicrofacetSampler(), counts[0]), (rru.CosineSampler(), counts[1]), (rru.LightSampler(), counts[2])]
    for ns_ in (2, 3):
        res = rru.importance_sample_rays(ikey, iv, inrm, dict(roughness=irough), random_generator_2d=_Pre2D(iu[:ns_]),
                                         use_mis=True, samplers=smps[:ns_], num_secondary_samples=sum(counts[:ns_]),
                                         light_sampler_results=laux if ns_ == 3 else None)
        for k_ in ("local_lightdirs", "local_viewdirs", "global_lightdirs", "pdf", "weight"):
            out[f"is{ns_}_{k_}"] = res[k_]

    # ---- DensityMLP.predict_density / run_network / convert_raw_density (geometry.py:155-341) as configured
    #      (ngp_yobo.gin:137-140,206-230: depth 2, width 64, ReLU, safe_exp, bias -1, 'mean' basis, contract_radius_2) ------
    def dense_params(n_in, n_out, salt):
        k = level_table((n_in, n_out), salt) * np.float32(100.0 * np.sqrt(6.0 / n_in))
        return k.astype(np.float32), (level_table((n_out,), salt + 50) * np.float32(10.0)).astype(np.float32)

    mlp = R["geometry"].DensityMLP(config=_types.SimpleNamespace(num_rgb_channels=3, n_bins=1), net_depth=2, net_width=64,
                                    net_activation=shim.nn_mod.relu,   # the class default (an instance field in flax)
                                    density_activation=rmath.safe_exp, density_bias=-1.0, warp_fn=rcoord.contract_radius_2,
                                    grid_params=dict(ENC_A, bbox_scaling=2.0))
    mlp.setup()
    mlp.grid.seen = []
    mlp.grid.param = lambda name, init_fn: (mlp.grid.seen.append(name), level_table(init_fn.keywords["shape"], len(mlp.grid.seen)))[1]
    d_in = 10
    for i_, layer in enumerate(mlp.density_layers + [mlp.output_density_layer]):
        layer.kernel, layer.bias = dense_params(d_in, layer.features, 100 + i_)
        d_in = layer.features
    dmeans = f(g.normal(size=(700, 3)) * 1.5); dmeans[0] = 0.0; dmeans[1] = [30.0, -2.0, 5.0]
    raw, feat = mlp.predict_density(dmeans.view(shim.F32Array), None, control_offsets=f(np.zeros((1, 3))), perp_mag=None)
    dens = mlp.convert_raw_density(raw, dmeans.view(shim.F32Array))
    assert raw.dtype == np.float32 and dens.dtype == np.float32
    out.update(dmlp_means=dmeans, dmlp_raw_density=np.asarray(raw), dmlp_feature=np.asarray(feat), dmlp_density=np.asarray(dens))
    # the module call itself (geometry.py:381-584) without the jax.value_and_grad branch: control points of the 'mean' basis
    # (coord.compute_control_points), predicted normals head, ray distances
    mlp.disable_density_normals, mlp.enable_pred_normals = True, True
    mlp.grid.seen = []                                     # the table salts count parameter requests from 1 again
    mlp.pred_normals_layer.kernel, mlp.pred_normals_layer.bias = dense_params(64, 3, 110)
    cm = dmeans[:696].reshape(58, 12, 3)
    g2 = np.random.Generator(np.random.PCG64(58))          # own stream: blocks added later leave earlier vectors unchanged
    crays = _types.SimpleNamespace(origins=f(g2.normal(size=(58, 3))), directions=f(g2.normal(size=(58, 3))),
                                   viewdirs=f(unit(g2.normal(size=(58, 3)))), radii=f(np.full((58, 1), 1e-3)),
                                   lights=f(g2.normal(size=(58, 3))))
    cres = mlp(None, crays, (cm.view(shim.F32Array), f(np.zeros((58, 12, 3, 3)))), tdist=None)
    out.update(dmlp_call_origins=crays.origins, dmlp_call_viewdirs=crays.viewdirs)
    for k_ in ("feature", "density", "grad_pred", "normals_pred", "normals_to_use", "ray_dists"):
        out["dmlp_call_" + k_] = np.asarray(cres[k_])
    assert cres["normals"] is None and cres["raw_grad_density"] is None

    # ---- ProposalVolumeSampler.__call__ (sampling.py:142-649), the whole level loop as configured (ngp_yobo.gin:178-242):
    #      ray warps (identity, and the power ladder of the secondary pass), annealed resampling logits, sample_intervals,
    #      s_to_t, cast_rays, compute_alpha_weights.  The three density MLPs are replaced by closed-form rational fields
    #      (IEEE +, *, / only - the same values in NumPy and PyTorch) so that the LOOP is what is pinned. ------------------
    def field(scale, k, c):
        scale, k, c = np.float32(scale), np.float32(k), [np.float32(v_) for v_ in c]

        def mlp(rng=None, rays=None, gaussians=None, tdist=None, **kw):
            m_ = np.asarray(gaussians[0])
            dx, dy, dz = m_[..., 0] - c[0], m_[..., 1] - c[1], m_[..., 2] - c[2]
            return dict(density=scale / (np.float32(1.0) + k * (dx * dx + dy * dy + dz * dz)))
        return mlp

    FIELDS = ((4.0, 3.0, (0.1, -0.2, 0.3)), (9.0, 6.0, (0.0, -0.1, 0.2)), (40.0, 14.0, (0.05, -0.15, 0.25)))
    Rp = 48
    po = f(g.normal(size=(Rp, 3)) * 0.3 + [0.0, 0.0, -2.5]); pd = f(g.normal(size=(Rp, 3)) * 0.25 + [0.0, 0.0, 1.3])
    pv = f(pd / np.linalg.norm(pd, axis=-1, keepdims=True))
    prays = R["utils"].Rays(origins=po, lights=None, directions=pd, viewdirs=pv, radii=f(g.uniform(2e-4, 2e-3, size=(Rp, 1))),
                            imageplane=None, look=None, up=None, cam_origins=None, vcam_look=None, vcam_up=None, vcam_origins=None,
                            lossmult=f(np.ones((Rp, 1))), near=f(g.uniform(0.1, 0.4, size=(Rp, 1))),
                            far=f(g.uniform(4.0, 7.0, size=(Rp, 1))), cam_idx=None, light_idx=None)
    pu = f(g.uniform(size=(Rp, 1)))
    out.update(pvs_origins=po, pvs_directions=pd, pvs_viewdirs=pv, pvs_radii=prays.radii, pvs_near=prays.near, pvs_far=prays.far,
               pvs_u01=pu)
    pvs = R["sampling"].ProposalVolumeSampler(
        config=None, sampling_strategy=((0, 0, 64), (1, 1, 64), (2, 2, 32)), anneal_slope=10.0, anneal_end=1.0, anneal_clip=0.4,
        resample_padding=1e-5, dilation_bias=0.0, dilation_multiplier=0.0,
        raydist_fn=(rmath.power_ladder, rmath.inv_power_ladder, dict(p=np.float32(-1.5), premult=np.float32(2.0))))
    pvs.mlps = [field(*a_) for a_ in FIELDS]
    for tag, use_rd in (("id", False), ("pl", True)):
        hist = pvs(pu, prays, train_frac=1.0, train=True, use_raydist_fn=use_rd)
        for lvl, h_ in enumerate(hist):
            for k_ in ("sdist", "tdist", "means", "weights"):
                out[f"pvs_{tag}_{lvl}_{k_}"] = h_[k_]

    # ---- cache shader pieces: BaseShader.predict_appearance_feature with net_depth 0 (shading.py:133-220),
    #      NeRFMLP.get_integrated_brdf (nerf.py:423-434,461-482) and _get_refdirs (:1344-1358, ref_utils.reflect) ------------
    ENC_S = dict(hash_map_size=2 ** 15, num_features=4, scale_supersample=1.0, max_grid_size=256, bbox_scaling=2.0)
    sh = R["shading"].BaseShader(net_depth=0, use_density_feature=True, warp_fn=rcoord.contract_radius_2)
    sh.grid = rgrid.HashEncoding(**ENC_S)
    sh.grid.seen = []
    sh.grid.param = lambda name, init_fn: (sh.grid.seen.append(name), level_table(init_fn.keywords["shape"], len(sh.grid.seen)))[1]
    Rs, ns = 40, 6
    smeans = f(g.normal(size=(Rs, ns, 3)) * 1.5); sfeat = f(g.normal(size=(Rs, ns, 64)))
    afeat = sh.predict_appearance_feature(dict(means=smeans.view(shim.F32Array), covs=None, feature=sfeat),
                                          control_offsets=f(np.zeros((1, 3))), perp_mag=None)
    out.update(shd_means=smeans, shd_density_feature=sfeat, shd_appearance_feature=np.asarray(afeat))
    nm = R["nerf"].NeRFMLP(net_activation=shim.nn_mod.relu, net_depth_integrated_brdf=2, skip_layer_integrated_brdf=4,
                           use_reflections=True)
    nm.integrated_brdf_layers = [shim.nn_mod_linen.Dense(64), shim.nn_mod_linen.Dense(64)]
    nm.output_integrated_brdf_layer = shim.nn_mod_linen.Dense(1)
    d_in = 129
    for i_, layer in enumerate(nm.integrated_brdf_layers + [nm.output_integrated_brdf_layer]):
        layer.kernel, layer.bias = dense_params(d_in, layer.features, 200 + i_)
        d_in = layer.features
    snrm = f(unit(g.normal(size=(Rs, ns, 3)))); sview = f(unit(g.normal(size=(Rs, 3)))); sbott = f(g.normal(size=(Rs, ns, 128)))
    out.update(shd_normals=snrm, shd_viewdirs=sview, shd_bottleneck=sbott)
    out["shd_integrated_brdf"] = nm.get_integrated_brdf(snrm, sview, sbott)
    out["shd_refdirs"] = nm._get_refdirs(sview, snrm, {})

    # ---- LightMLP.get_vmfs (light_sampler.py:135-160) with the class's own bias / activation tables ---------------------
    vraw = f(g.normal(size=(64, 16, 5)) * 3.0); vraw[0, 0, 3:] = [60.0, -80.0]     # both clamps
    vnorm = f(g.normal(size=(64, 16, 3)))
    lm = R["light_sampler"].LightMLP(num_components=16, random_seed={"normal": vnorm})   # PRNGKey(seed) = the pre-drawn normals
    vm = lm.get_vmfs(vraw)
    out.update(vmfs_raw=vraw, vmfs_normal=vnorm)
    for k_ in ("vmf_means", "vmf_kappas", "vmf_logits"):
        out["vmfs_" + k_] = vm[k_]

    # ---- MaterialMLP._get_microfacet_material (material.py:1276-1322) over its own property table (:957-1023), fields
    #      set as configs/ngp_yobo.gin:256-303 sets them ------------------------------------------------------------------
    sig = shim.nn_mod.sigmoid
    props = ("albedo", "specular_albedo", "roughness", "F_0", "metalness", "diffuseness", "mirrorness")
    mm = R["material"].MaterialMLP(
        num_rgb_channels=3, brdf_activation={k_: sig for k_ in props},
        brdf_bias=dict(albedo=-1.0, specular_albedo=-1.0, roughness=-1.0, F_0=-3.078, metalness=0.0, diffuseness=0.0, mirrorness=2.0),
        brdf_stopgrad=dict(albedo=1.0, specular_albedo=1.0, roughness=0.25, F_0=1.0, metalness=1.0, diffuseness=1.0, mirrorness=1.0),
        use_diffuseness=False, use_mirrorness=False, use_constant_metalness=False, use_constant_fresnel=True,
        min_roughness=0.01, default_F_0=0.04, max_F_0=1.0, reparam_roughness=False)
    mm._initialize_microfacet_properties()
    braw = f(g.normal(size=(300, 10)) * 3.0)
    mat = mm._get_microfacet_material(braw)
    out["mat_brdf_params"] = braw
    for k_ in props:
        out["mat_" + k_] = mat[k_]

    # ---- TransientNeRFMLP._compute_indirect_lighting / get_indirect (nerf.py:1660-1777) as configured
    #      (transient_ngp_yobo.gin:176-179: irradiance stack 2 x 64, skip 2; deg_lights 2) --------------------------------
    Dn = shim.nn_mod_linen.Dense
    Rh, nh, Bh = 12, 4, 48
    hcfg = _types.SimpleNamespace(n_bins=Bh, num_rgb_channels=3, light_intensity_conditioning=False,
                                  bin_zero_threshold_light=np.float32(2.0), exposure_time=expo, light_zero=False,
                                  light_near=np.float32(0.0))
    tn = R["nerf"].TransientNeRFMLP(config=hcfg, use_indirect=True, net_activation=shim.nn_mod.relu, deg_lights=2,
                                    net_depth_irradiance=2, net_width_irradiance=64, bottleneck_irradiance=64,
                                    skip_layer_irradiance=2, irradiance_activation=shim.nn_mod.softplus, irradiance_bias=-2.0,
                                    indirect_scale=np.float32(0.7), rgb_max=np.float32(1.5), net_depth_integrated_brdf=2,
                                    skip_layer_integrated_brdf=4, dense_layer=Dn)
    tn._initialize_irradiance_layers()
    tn.integrated_brdf_layers, tn.output_integrated_brdf_layer = [Dn(64), Dn(64)], Dn(1)
    tn.tint_layer, tn.transient_indirect_layer = Dn(3), Dn(Bh * 3)
    for layers, d_in, salt in ((tn.irradiance_layers + [tn.transient_indirect_layer], 96 + 15, 300),
                               (tn.integrated_brdf_layers + [tn.output_integrated_brdf_layer], 129, 200),
                               ([tn.tint_layer], 96, 320)):
        for i_, layer in enumerate(layers):
            layer.kernel, layer.bias = dense_params(d_in, layer.features, salt + i_)
            d_in = layer.features
    hfeat = f(g.normal(size=(Rh, nh, 96))); hmeans = f(g.uniform(-0.05, 0.05, size=(Rh, nh, 3)))
    hnrm = f(unit(g.normal(size=(Rh, nh, 3)))); href = f(g.gamma(1.0, 1.0, size=(Rh, nh, Bh * 3)))
    hbott = f(g.normal(size=(Rh, nh, 128))); hview = f(unit(g.normal(size=(Rh, 3))))
    hrays = _types.SimpleNamespace(lights=f(g.uniform(-0.05, 0.05, size=(Rh, 3))), origins=f(g.uniform(-0.05, 0.05, size=(Rh, 3))),
                                   cam_origins=f(g.uniform(-0.05, 0.05, size=(Rh, 3))))
    out.update(th_feature=hfeat, th_means=hmeans, th_normals=hnrm, th_ref_rgb=href, th_bottleneck=hbott, th_viewdirs=hview,
               th_lights=hrays.lights, th_origins=hrays.origins, th_cam_origins=hrays.cam_origins)
    res = tn._compute_indirect_lighting(hfeat, hmeans, hnrm, hnrm, href, hbott, hview, None, hrays, None, None)
    for k_, v_ in zip(("indirect_diffuse", "indirect_specular", "transient_indirect", "transient_indirect_diffuse",
                       "transient_indirect_specular"), res):
        out["th_" + k_] = v_

    # ---- train_utils.light_sampling_loss (:1985-2071): the call site of vmf_loss_fn (function values = |radiance_in|,
    #      lossmult / S, one suffix present -> multiplier 2 and the / 2 inside the loop) ---------------------------------
    lrad = f(g.gamma(1.0, 1.0, size=(Ps, Ss, 3)))
    out["light_radiance_in"] = lrad
    mres = dict(light_sampler=dict(vmf_means=lmeans[:, None], vmf_kappas=lkappas[:, None], vmf_logits=llogits[:, None],
                                   vmf_normals=lnormals[:, None, None, :]),
                shader=dict(ref_rays_indirect_diffuse=None, ref_samples_indirect_diffuse=None,
                            ref_rays_indirect_specular=_types.SimpleNamespace(viewdirs=swi.reshape(-1, 3)),
                            ref_samples_indirect_specular=dict(radiance_in=lrad, pdf=lpdf, weight=lwt)))
    for srgb in (True, False):
        out[f"light_sampling_loss_{int(srgb)}"] = np.asarray(rtrain.light_sampling_loss(
            None, None, None, _types.SimpleNamespace(lossmult=f(np.ones((Ps, 1)))),
            _types.SimpleNamespace(light_sampling_linear_to_srgb=srgb), None, mres))

    # ---- SurfaceLightFieldMLP.__call__ (surface_light_field.py:782-1069) in the two configured shapes
    #      (nerf_ngp_yobo.gin:232-251 `SurfaceLightField`: IDE_5 + shader bottleneck; :253-297 shader `EnvMap`: IDE_4), and
    #      the cache shader's composition around them: NeRFMLP.get_bottleneck_feature (nerf.py:385-408), the roughness head
    #      (:633-634), _predict_appearance_passive (:940-1090) with both sub-networks attached.  setup() is the class's own:
    #      it builds the encodings and every Dense (the flax stand-in's Dense = x @ kernel + bias); parameters are the closed
    #      form dense_params.  Callable CLASS defaults (activations) are passed as constructor fields: a plain class attribute
    #      would bind as a method here, flax stores them per instance. ------------------------------------------------------
    rslf = importlib.import_module("internal.surface_light_field")
    sp_, sg_, relu_ = shim.nn_mod.softplus, shim.nn_mod.sigmoid, shim.nn_mod.relu
    scfg = _types.SimpleNamespace(num_rgb_channels=3, n_bins=1, multi_illumination=False, rotate_illumination=False,
                                  num_illuminations=1, multiple_illumination_outputs=False, env_map_distance=float("inf"))

    def make_slf(deg_view, use_shader_bottleneck, salt):
        m_ = rslf.SurfaceLightFieldMLP(
            config=scfg, net_depth=2, net_width=64, skip_layer=2, bottleneck_width=128, use_directional_enc=True, use_ide=True,
            deg_view=deg_view, net_depth_viewdirs=4, net_width_viewdirs=128, bottleneck_viewdirs=128, skip_layer_dir=2,
            use_grid=False, use_bottleneck=False, use_density_feature=False, use_shader_bottleneck=use_shader_bottleneck,
            use_lights=False, net_activation=relu_, rgb_activation=sp_, rgb_bias=-2.0, ambient_rgb_activation=sp_,
            ambient_rgb_bias=-1.0, alpha_activation=sg_)
        m_.setup()
        n_in = (128 if use_shader_bottleneck else 0) + {4: 38, 5: 72}[deg_view]
        d_in_ = n_in
        for i_, layer in enumerate(m_.view_dependent_layers):
            layer.kernel, layer.bias = dense_params(d_in_, layer.features, salt + i_)
            d_in_ = layer.features + (n_in if (i_ % 2 == 0 and i_ > 0) else 0)
        for j_, layer in enumerate((m_.output_ambient_rgb_layer, m_.output_rgba_layer)):
            layer.kernel, layer.bias = dense_params(d_in_, layer.features, salt + 10 + j_)
        return m_

    Rq, nq = 24, 8
    qdirs = f(unit(g.normal(size=(Rq, nq, 3)))); qrough = f(g.uniform(0.02, 1.5, size=(Rq, nq, 1)))
    qbott = f(g.normal(size=(Rq, nq, 128))); qmeans = f(g.normal(size=(Rq, nq, 3)) * 1.2)
    qrays = _types.SimpleNamespace(viewdirs=f(unit(g.normal(size=(Rq, 3)))), origins=f(g.normal(size=(Rq, 3)) * 3.0),
                                   light_idx=None, lights=None)
    out.update(slf_refdirs=qdirs, slf_roughness=qrough, slf_bottleneck=qbott)
    slf5, slf4 = make_slf(5, True, 400), make_slf(4, False, 420)
    for tag, net, bott_ in (("slf5", slf5, qbott), ("slf4", slf4, None)):
        res = net(None, qrays, dict(means=qmeans), qmeans, qdirs, roughness=qrough, shader_bottleneck=bott_, train=False)
        out[tag + "_incoming_ambient_rgb"] = res["incoming_ambient_rgb"]
        out[tag + "_incoming_acc"] = res["incoming_acc"]

    ENC_P = dict(hash_map_size=2 ** 15, num_features=4, scale_supersample=1.0, max_grid_size=2048, bbox_scaling=2.0)
    shp = R["shading"].BaseShader(net_depth=0, use_density_feature=True, warp_fn=rcoord.contract_radius_2)
    class _Enc64(rgrid.HashEncoding):
        # `grid_size**3 <= hash_map_size` (grid_utils.py:837) is evaluated on NumPy int32 scalars: NumPy 1.x (the reference's
        # pin) promotes int32 ** python-int to int64, NumPy 2 keeps int32 and 2048**3 wraps to 0.  Same values, wider type.
        grid_sizes = property(lambda self: rgrid.HashEncoding.grid_sizes.fget(self).astype(np.int64))

    shp.grid = _Enc64(**ENC_P)
    shp.grid.seen = []
    shp.grid.param = lambda name, init_fn: (shp.grid.seen.append(name), level_table(init_fn.keywords["shape"], len(shp.grid.seen)))[1]
    qdf = f(g.normal(size=(Rq, nq, 64))); qnrm = f(unit(g.normal(size=(Rq, nq, 3))))
    qfeat = np.asarray(shp.predict_appearance_feature(dict(means=qmeans.view(shim.F32Array), covs=None, feature=qdf),
                                                      control_offsets=f(np.zeros((1, 3))), perp_mag=None)).astype(np.float32)
    assert qfeat.shape == (Rq, nq, 96)
    Dn_ = shim.nn_mod_linen.Dense
    pm = R["nerf"].NeRFMLP(config=scfg, net_activation=relu_, net_depth_integrated_brdf=2, skip_layer_integrated_brdf=2,
                           use_reflections=True, bottleneck_width=128, bottleneck_noise=0.0, use_exposure_at_bottleneck=False,
                           roughness_activation=sp_, roughness_bias=-1.0, irradiance_activation=sp_, irradiance_bias=-2.0,
                           ambient_irradiance_activation=sp_, ambient_irradiance_bias=-2.0, rgb_max=np.float32(10000.0),
                           use_env_map=True, stopgrad_ambient_weight=1.0, stopgrad_indirect_weight=1.0)
    pm.bottleneck_layer, pm.roughness_layer = Dn_(128), Dn_(1)
    pm.ambient_irradiance_layer, pm.irradiance_layer, pm.tint_layer = Dn_(3), Dn_(3), Dn_(3)
    pm.integrated_brdf_layers, pm.output_integrated_brdf_layer = [Dn_(64), Dn_(64)], Dn_(1)
    for layers, d_in, salt in (([pm.bottleneck_layer], 96, 440), ([pm.roughness_layer], 96, 441),
                               ([pm.ambient_irradiance_layer], 96, 442), ([pm.irradiance_layer], 96, 443),
                               ([pm.tint_layer], 96, 444),
                               (pm.integrated_brdf_layers + [pm.output_integrated_brdf_layer], 129, 445)):
        for i_, layer in enumerate(layers):
            layer.kernel, layer.bias = dense_params(d_in, layer.features, salt + 10 * i_ if len(layers) > 1 else salt)
            d_in = layer.features
    pm.surface_lf, pm.env_map = slf5, slf4
    qb = pm.get_bottleneck_feature(None, qfeat, None)                                            # nerf.py:385-408
    qr = pm.roughness_activation(pm.roughness_layer(qfeat) + pm.roughness_bias)                  # nerf.py:633-634
    res = pm._predict_appearance_passive(None, qrays, dict(means=qmeans), qfeat, qb, qr, qnrm, qnrm, train=False)
    out.update(shp_means=qmeans, shp_density_feature=qdf, shp_normals=qnrm, shp_viewdirs=qrays.viewdirs, shp_origins=qrays.origins,
               shp_feature=qfeat, shp_bottleneck=qb, shp_roughness=qr)
    for k_ in ("rgb", "diffuse_rgb", "specular_rgb", "ambient_rgb", "indirect_rgb", "albedo_rgb", "indirect_occ", "ray_dists"):
        out["shp_" + k_] = res[k_]

    # ---- analytic normals (geometry.py:442-460): jax.value_and_grad(predict_density) cannot be executed without JAX; the
    #      VALUE it returns - d raw_density / d mean - is pinned by CENTRAL DIFFERENCES of the reference's own predict_density
    #      (h = 2^-11 per axis, the realised step taken from the rounded fp32 arguments).  The field is piecewise trilinear
    #      through a ReLU MLP: a difference that straddles a cell face or a ReLU kink averages two one-sided slopes, so the
    #      consumers compare quantiles, not the maximum. ------------------------------------------------------------------
    hfd = np.float32(2.0 ** -11)

    def raw_at(x_):
        mlp.grid.seen = []                                 # the table salts count parameter requests from 1 again
        r_, _ = mlp.predict_density(x_.view(shim.F32Array), None, control_offsets=f(np.zeros((1, 3))), perp_mag=None)
        return np.asarray(r_).astype(np.float64)

    npts = np.ascontiguousarray(dmeans[2:402])
    fd = np.zeros((npts.shape[0], 3), np.float64)
    for a_ in range(3):
        e_ = np.zeros(3, np.float32); e_[a_] = hfd
        xp, xm = (npts + e_).astype(np.float32), (npts - e_).astype(np.float32)
        fd[:, a_] = (raw_at(xp) - raw_at(xm)) / (xp[:, a_].astype(np.float64) - xm[:, a_].astype(np.float64))
    out.update(dnrm_means=npts, dnrm_fd_raw_grad=fd.astype(np.float32))

    # ---- camera_utils.pixels_to_rays (camera_utils.py:896-1073), perspective camera without distortion / NDC / jitter, and
    #      get_pixtocam (:749-763): a 20 x 12 image seen from an orbit pose.  Pixel coordinates are passed as float32 (jnp
    #      promotes int32 + 0.5 to float32; NumPy would make it float64). -------------------------------------------------
    rcam = importlib.import_module("internal.camera_utils")
    Wc, Hc, fc = 20, 12, 17.3
    az, el = np.deg2rad(30.0), np.deg2rad(25.0)
    cpos = 4.0 * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    cfwd = -cpos / np.linalg.norm(cpos)
    cright = np.cross(cfwd, [0.0, 0.0, 1.0]); cright /= np.linalg.norm(cright)
    cup = np.cross(cright, cfwd)
    c2w_ = f(np.concatenate([np.stack([cright, cup, -cfwd], axis=1), cpos[:, None]], axis=1))
    k_ = f(rcam.get_pixtocam(fc, Wc, Hc))
    cys, cxs = np.meshgrid(np.arange(Hc, dtype=np.float32), np.arange(Wc, dtype=np.float32), indexing="ij")
    cres = rcam.pixels_to_rays(cxs, cys, k_, c2w_, xnp=shim.jnp)
    out.update(cam_pixtocam=k_, cam_camtoworld=c2w_, cam_size=np.array([Wc, Hc], np.int32))
    for k_name, v_ in zip(("origins", "directions", "viewdirs", "radii", "imageplane"), cres[:5]):
        out["cam_" + k_name] = np.asarray(v_)

    # ---- the temporal filter of volumetric_transient_rendering (render.py:397-415): Gaussian of tfilter_sigma bins on the
    #      direct histogram, and on the indirect one with filter_indirect; jax.scipy.signal.convolve(mode='same') ----------
    import scipy.signal as _sps
    shim.jscipy.signal = _types.SimpleNamespace(
        convolve=lambda a_, b_, mode="full": _sps.convolve(np.asarray(a_, np.float32), np.asarray(b_, np.float32), mode=mode, method="direct").astype(np.float32))
    for fi in (False, True):
        res = rrender.volumetric_transient_rendering(
            tdirect, tind, tw, tw, ttd, None, False, extras=dict(light_dists=tlight, ray_dists=tray, transient_indirect=None),
            n_bins=Bt, shift=np.float32(0.0137), dark_level=np.float32(0.001), exposure_time=expo, config=cfg,
            tfilter_sigma=np.float32(1.5), filter_indirect=fi)
        for k_ in ("transient_direct", "transient_indirect", "rgb"):
            out["tr_f%d_%s" % (int(fi), k_)] = res[k_]

    # ---- transient_integrate_reflect_rays (render_utils.py:1195-1302), direct=False: time-resolved incoming radiance ----
    trad = f(g.gamma(1.0, 1.0, size=(P_, S_, 24, 3)))
    tsamples = dict(samples, radiance_in=trad)
    out["ggxt_radiance_in"] = trad
    tres = rru.transient_integrate_reflect_rays("microfacet", False, dict(material), tsamples, direct=False)
    for k_ in ("radiance_out", "irradiance", "indirect_occ"):
        out["ggxt_" + k_] = tres[k_]

    out = {k: np.asarray(v_) for k, v_ in out.items()}
    out = {k: (v_.astype(np.float32) if v_.dtype == np.float64 else v_) for k, v_ in out.items()}   # see the shim's header
    path = os.path.join(HERE, "reference_np.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KB")
    for k, v_ in sorted(out.items()):
        print(f"  {k:34s} {str(v_.dtype):8s} {v_.shape}")


if __name__ == "__main__":
    main()
