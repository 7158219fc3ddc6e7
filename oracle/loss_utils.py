"""CPU restatement (test infrastructure) of the proposal supervision of the cache stage:
internal/loss_utils.py spline_interlevel_loss (:74-108), internal/stepfun.py weight_to_pdf (:75-79),
blur_and_resample_weights (:463-483), internal/linspline.py blur_stepfun (:187-221), compute_integral (:95-109),
interpolate_integral (:124-141), internal/math.py plus_eps / minus_eps (:56-66); configs/ngp_yobo.gin:245-247
(mults (0.01, 0.01), blurs (0.03, 0.003)).  Pinned to the reference's source (tests/test_reference_vectors.py:
blur_and_resample_weights, spline_interlevel_loss, distortion_loss, orientation / predicted-normal losses, compute_mask_loss);
unpinned against XLA's own rounding (JAX not installable here).  Note: torch.cumsum on the host carries its
running sum in float64 and rounds per element, so this restatement is MORE accurate than XLA's fp32 cumsum in the
ill-conditioned double running sum of blur_stepfun; the CUDA body does the same."""
import numpy as np
import torch

from . import ref_math

TINY = float(np.finfo(np.float32).tiny)
EPS = float(np.finfo(np.float32).eps)


def plus_eps(x):
    return torch.where(torch.abs(x) < TINY, torch.full_like(x, TINY), torch.nextafter(x, torch.full_like(x, float("inf"))))


def minus_eps(x):
    return torch.where(torch.abs(x) < TINY, torch.full_like(x, -TINY), torch.nextafter(x, torch.full_like(x, -float("inf"))))


def weight_to_pdf(t, w):
    td = torch.diff(t, dim=-1)
    return torch.where(td < TINY, torch.zeros_like(w), ref_math.safe_div(w, td))


def blur_stepfun(ts, ys, halfwidth):
    ts_lo = torch.minimum(minus_eps(ts), ts - halfwidth)
    ts_hi = torch.maximum(plus_eps(ts), ts + halfwidth)
    z = torch.zeros_like(ys[..., :1])
    ys0 = torch.cat([z, ys, z], dim=-1)
    dy = torch.diff(ys0, dim=-1) / (ts_hi - ts_lo)
    tp = torch.cat([ts_lo, ts_hi], dim=-1)
    dyp = torch.cat([dy, -dy], dim=-1)
    idx = torch.argsort(tp, dim=-1, stable=True)
    tp = torch.gather(tp, -1, idx)
    dyp = torch.gather(dyp, -1, idx[..., :-2])
    yp = torch.cumsum(torch.diff(tp, dim=-1)[..., :-1] * torch.cumsum(dyp, dim=-1), dim=-1)
    return tp, torch.cat([torch.zeros_like(yp[..., :1]), yp, torch.zeros_like(yp[..., -1:])], dim=-1)


def compute_integral(t, y):
    dt = torch.diff(t, dim=-1)
    a = torch.diff(y, dim=-1) / torch.clamp(2 * dt, min=EPS**2)
    b = y[..., :-1]
    c1 = 0.5 * torch.cumsum(dt[..., :-1] * (y[..., :-2] + y[..., 1:-1]), dim=-1)
    return a, b, torch.cat([torch.zeros_like(y[..., :1]), c1], dim=-1)


def interpolate_integral(tq, t, a, b, c):
    tq = torch.maximum(torch.minimum(tq, minus_eps(t[..., -1:])), t[..., :1])
    idx = torch.searchsorted(t.contiguous(), tq.contiguous(), right=True)
    idx0 = torch.clamp(idx - 1, min=0)
    t0, a0, b0, c0 = (torch.gather(v, -1, idx0) for v in (t, a, b, c))
    td = tq - t0
    return a0 * td**2 + b0 * td + c0


def blur_and_resample_weights(tq, t, w, blur_halfwidth):
    p = weight_to_pdf(t, w)
    tl, pl = blur_stepfun(t, p, blur_halfwidth)
    acc = interpolate_integral(tq, tl, *compute_integral(tl, pl))
    return torch.clamp(torch.diff(acc, dim=-1), min=0)


def spline_interlevel_loss(ray_history, mults=(0.01, 0.01), blurs=(0.03, 0.003), eps=1e-5):
    """-> list of per-level losses (gradient flows to the proposal weights only: w_blur is stop_gradient)."""
    c, w = ray_history[-1]["sdist"], ray_history[-1]["weights"]
    out = []
    for mult, blur, h in zip(mults, blurs, ray_history[:-1]):
        w_blur = blur_and_resample_weights(h["sdist"], c, w, blur).detach()
        wp = h["weights"]
        out.append(mult * torch.mean(torch.clamp(w_blur - wp, min=0) ** 2 / (wp + eps)))
    return out


# ------------------------------------------------------------------ geometry / mask losses
def stopgrad_with_weight(x, weight):
    """internal/utils.py:87-95."""
    if weight is None or weight == 1.0:
        return x
    if weight == 0.0:
        return x.detach()
    return (x - x.detach()) * weight + x.detach()


def orientation_loss(rays, ray_results, target="normals", mult=1.0, normalize=False, stopgrad=False):
    """internal/loss_utils.py:127-166 (lossmult == 1)."""
    w = ray_results["weights"]
    if normalize:
        w = w / torch.sum(w, dim=-1, keepdim=True)
    if stopgrad:
        w = w.detach()
    n = torch.nan_to_num(ray_results[target])
    v = -rays["viewdirs"]
    n_dot_v = (n * v[..., None, :]).sum(dim=-1)
    loss = torch.mean(torch.abs(torch.abs(w * torch.clamp(n_dot_v, max=0.0) ** 2).sum(dim=-1) + 1e-5))
    return loss * mult


def predicted_normal_loss(ray_results, beta, mult=1.0, gt="normals", pred="normals_pred", normalize=False,
                          stopgrad=False, stopgrad_weight=1.0):
    """internal/loss_utils.py:169-199 (lossmult == 1)."""
    w = ray_results["weights"]
    if normalize:
        w = w / (torch.sum(w, dim=-1, keepdim=True) + 1e-8)
    w = w.detach() if stopgrad else stopgrad_with_weight(w, stopgrad_weight)
    n = torch.nan_to_num(ray_results[gt]).detach()
    n_pred = torch.nan_to_num(ray_results[pred])
    loss = torch.mean(torch.abs(
        (torch.abs(w * (1.0 - torch.sum(n * n_pred, dim=-1))) * beta[..., 0]).sum(dim=-1, keepdim=True) + 1e-5))
    return loss * mult


def geometry_losses(rays, geometry, orientation_mult=0.01, predicted_normal_mult=0.001,
                    predicted_normal_reverse_mult=0.01, stopgrad_weight=0.1):
    """_compute_geometry_losses (internal/train_utils.py:3255-3311) as configured by
    configs/nerf_ngp_yobo_lego.gin:7-11 + configs/nerf_ngp_yobo.gin:59-72 at train_frac = 1 (ease / decay
    factors 1): orientation on 'normals_pred'; predicted normals with gt='normals_pred', pred='normals'
    (train_utils.py:1049-1070); reverse with gt='normals', pred='normals_pred', stopgrad (:1073-1093)."""
    beta = torch.ones_like(geometry["normals"][..., :1])
    return [
        orientation_loss(rays, geometry, target="normals_pred", mult=orientation_mult),
        predicted_normal_loss(geometry, beta, mult=predicted_normal_mult, gt="normals_pred", pred="normals",
                              stopgrad=False, stopgrad_weight=stopgrad_weight),
        predicted_normal_loss(geometry, beta, mult=predicted_normal_reverse_mult, gt="normals", pred="normals_pred",
                              stopgrad=True),
    ]


def compute_mask_loss(acc, masks=None, charb_padding=0.001, opaque_loss_weight=1.0, empty_loss_weight=1.0,
                      backward=False):
    """compute_mask_loss (internal/train_utils.py:785-836), lossmult == 1, schedule factors 1.  `backward=True` is
    the backward-mask call (:2929-2945): zero masks, opaque part off, empty part weighted by `empty_loss_weight`."""
    if masks is None:
        masks = torch.ones_like(acc)[..., None]
    data_loss = torch.sqrt((acc[..., None] - masks) ** 2 + charb_padding**2)
    if backward:
        data_loss = torch.where(masks > 0.5, data_loss * 0.0, data_loss * empty_loss_weight)
    else:
        data_loss = torch.where(masks > 0.5, data_loss * opaque_loss_weight, data_loss * empty_loss_weight)
    return torch.mean(data_loss)


def lossfun_distortion(t, w):
    """stepfun.lossfun_distortion (internal/stepfun.py:253-269), normalize=False."""
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    loss_intra = torch.sum(w**2 * torch.diff(t, dim=-1), dim=-1) / 3
    return loss_inter + loss_intra


def distortion_loss(ray_history, mult=0.01, p=-0.25, premult=10000.0, target="tdist"):
    """loss_utils.distortion_loss (internal/loss_utils.py:108-123) as configured by configs/ngp_yobo.gin:250-253
    (target 'tdist', curve_fn power_ladder(p=-0.25, premult=10000)) and nerf_ngp_yobo_lego.gin:10 (mult 0.01)."""
    from . import ref_math
    last = ray_history[-1]
    c = ref_math.power_ladder(last[target], p, premult=premult)
    return mult * torch.mean(lossfun_distortion(c, last["weights"]))


def param_regularizer_loss(tables, mult=1.0):
    """param_regularizer_loss (internal/train_utils.py:1169-1216) for the (mult, jnp.mean, alpha=2, scale=1) setting of
    Config.param_regularizers (configs/nerf_ngp_yobo.gin:47-51): sum over the given parameter tensors of
    mult * 0.5 * mean(param^2).  `tables`: the level tables of every module whose name matches the prefix."""
    loss = 0.0
    for t in tables:
        loss = loss + mult * 0.5 * torch.mean(t**2)
    return loss
