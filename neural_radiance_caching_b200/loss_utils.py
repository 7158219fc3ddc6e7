"""Host-side mirror of the cache stage's proposal supervision (internal/loss_utils.py:74-108,
configs/ngp_yobo.gin:245-247): spline_interlevel_loss over the CUDA body nrc_interlevel_loss."""
import torch

from . import _lib


class _InterlevelLossFn(torch.autograd.Function):
    """One proposal level: mult * mean(max(0, stop_grad(w_blur) - wp)^2 / (wp + eps)); VJP wrt wp only."""

    @staticmethod
    def forward(ctx, c, w, cp, wp, blur, mult, eps):
        m, nq = w.shape[-1], wp.shape[-1]
        c2, w2 = c.reshape(-1, m + 1).contiguous(), w.reshape(-1, m).contiguous()
        cp2, wp2 = cp.reshape(-1, nq + 1).contiguous(), wp.reshape(-1, nq).contiguous()
        R = w2.shape[0]
        loss = torch.zeros((), device=w.device, dtype=torch.float32)
        g_wp = torch.empty_like(wp2)
        _lib.call("nrc_interlevel_loss", _lib.stream_ptr(), _lib.ptr(c2), _lib.ptr(w2), m, _lib.ptr(cp2), _lib.ptr(wp2), nq, R,
                  float(blur), float(mult), float(eps), _lib.ptr(loss), _lib.ptr(g_wp), None)
        ctx.save_for_backward(g_wp)
        ctx.shape = wp.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (g_wp,) = ctx.saved_tensors
        return None, None, None, (g_wp * g).reshape(ctx.shape), None, None, None


def blur_and_resample_weights(tq, t, w, blur_halfwidth):
    """stepfun.blur_and_resample_weights (internal/stepfun.py:463-483); forward value only."""
    m, nq = w.shape[-1], tq.shape[-1] - 1
    t2, w2, tq2 = t.reshape(-1, m + 1).contiguous(), w.reshape(-1, m).contiguous(), tq.reshape(-1, nq + 1).contiguous()
    R = w2.shape[0]
    dev = w.device
    wb = torch.empty((R, nq), device=dev, dtype=torch.float32)
    scratch = torch.zeros((), device=dev, dtype=torch.float32)
    zeros = torch.zeros((R, nq), device=dev, dtype=torch.float32)
    _lib.call("nrc_interlevel_loss", _lib.stream_ptr(), _lib.ptr(t2), _lib.ptr(w2), m, _lib.ptr(tq2), _lib.ptr(zeros), nq, R,
              float(blur_halfwidth), 0.0, 1e-5, _lib.ptr(scratch), _lib.ptr(torch.empty_like(zeros)), _lib.ptr(wb))
    return wb.reshape(tq.shape[:-1] + (nq,))


def spline_interlevel_loss(ray_history, *, mults=(0.01, 0.01), blurs=(0.03, 0.003), eps=1e-5):
    """A spline-based alternative to interlevel_loss that lets us blur stuff (internal/loss_utils.py:74-108).
    Returns the list of per-level losses (lossmult == 1)."""
    num_rounds = len(ray_history[:-1])
    if not isinstance(mults, tuple):
        mults = (mults,) * num_rounds
    if len(mults) < num_rounds or len(blurs) < num_rounds:
        return []
    c = ray_history[-1]["sdist"].detach()
    w = ray_history[-1]["weights"].detach()
    return [_InterlevelLossFn.apply(c, w, h["sdist"].detach(), h["weights"], blur, mult, eps)
            for mult, blur, h in zip(mults, blurs, ray_history[:-1])]


# ------------------------------------------------------------------ geometry / mask losses
class _GeometryLossFn(torch.autograd.Function):
    """orientation + predicted-normal (+ reverse) losses of the final sampler level in one pass
    (internal/train_utils.py:3255-3311, internal/loss_utils.py:127-199): nrc_geometry_losses."""

    @staticmethod
    def forward(ctx, weights, normals, normals_pred, viewdirs, mult_o, mult_p, mult_r, sg_w):
        R, n = weights.shape
        dev = weights.device
        loss = torch.zeros((), device=dev, dtype=torch.float32)
        g_w = torch.zeros((R, n), device=dev, dtype=torch.float32)
        g_np = torch.zeros((R, n, 3), device=dev, dtype=torch.float32)
        g_n = torch.empty((R, n, 3), device=dev, dtype=torch.float32)
        _lib.call("nrc_geometry_losses", _lib.stream_ptr(), _lib.ptr(weights.contiguous()), _lib.ptr(normals.contiguous()),
                  _lib.ptr(normals_pred.contiguous()), _lib.ptr(viewdirs.contiguous()), R, n, float(mult_o), float(mult_p),
                  float(mult_r), float(sg_w), _lib.ptr(loss), _lib.ptr(g_w), _lib.ptr(g_np), _lib.ptr(g_n))
        ctx.save_for_backward(g_w, g_n, g_np)
        return loss

    @staticmethod
    def backward(ctx, g):
        g_w, g_n, g_np = ctx.saved_tensors
        return g_w * g, g_n * g, g_np * g, None, None, None, None, None


def geometry_losses(rays, geometry, orientation_mult=0.01, predicted_normal_mult=0.001,
                    predicted_normal_reverse_mult=0.01, stopgrad_weight=0.1):
    """_compute_geometry_losses as configured by configs/nerf_ngp_yobo_lego.gin:7-11 at train_frac = 1: the sum of
    the orientation loss (target 'normals_pred'), the predicted-normal loss (pred = the analytic normals: second-
    order path) and its reverse."""
    return [_GeometryLossFn.apply(geometry["weights"], geometry["normals"], geometry["normals_pred"], rays["viewdirs"],
                                  orientation_mult, predicted_normal_mult, predicted_normal_reverse_mult, stopgrad_weight)]


class _MaskLossFn(torch.autograd.Function):
    """compute_mask_loss (internal/train_utils.py:785-836): nrc_mask_loss."""

    @staticmethod
    def forward(ctx, acc, masks, charb_padding, opaque_w, empty_w):
        R = acc.shape[0]
        loss = torch.zeros((), device=acc.device, dtype=torch.float32)
        g_acc = torch.empty((R,), device=acc.device, dtype=torch.float32)
        _lib.call("nrc_mask_loss", _lib.stream_ptr(), _lib.ptr(acc.contiguous()), 0,
                  _lib.ptr(masks.reshape(R).contiguous()) if masks is not None else None, R, float(charb_padding),
                  float(opaque_w), float(empty_w), _lib.ptr(loss), _lib.ptr(g_acc))
        ctx.save_for_backward(g_acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        (g_acc,) = ctx.saved_tensors
        return g_acc * g, None, None, None, None


def compute_mask_loss(acc, masks=None, charb_padding=0.001, opaque_loss_weight=1.0, empty_loss_weight=1.0,
                      backward=False):
    """`backward=True`: the backward-mask call (train_utils.py:2929-2945): zero masks, opaque part off."""
    if backward:
        masks = torch.zeros_like(acc) if masks is None else masks
        return _MaskLossFn.apply(acc, masks, charb_padding, 0.0, empty_loss_weight)
    return _MaskLossFn.apply(acc, masks, charb_padding, opaque_loss_weight, empty_loss_weight)


class _DistortionLossFn(torch.autograd.Function):
    """loss_utils.distortion_loss over stepfun.lossfun_distortion: nrc_distortion_loss (VJP w.r.t. the weights)."""

    @staticmethod
    def forward(ctx, t, w, p, premult, mult):
        R, n = w.shape
        loss = torch.zeros((), device=w.device, dtype=torch.float32)
        g_w = torch.zeros((R, n), device=w.device, dtype=torch.float32)
        _lib.call("nrc_distortion_loss", _lib.stream_ptr(), _lib.ptr(t.contiguous()), _lib.ptr(w.contiguous()), n, R,
                  float(p), float(premult), float(mult), _lib.ptr(loss), _lib.ptr(g_w))
        ctx.save_for_backward(g_w)
        return loss

    @staticmethod
    def backward(ctx, g):
        (g_w,) = ctx.saved_tensors
        return None, g_w * g, None, None, None


def distortion_loss(ray_history, mult=0.01, p=-0.25, premult=10000.0, target="tdist"):
    """internal/loss_utils.py:108-123 with configs/ngp_yobo.gin:250-253 and nerf_ngp_yobo_lego.gin:10."""
    last = ray_history[-1]
    return _DistortionLossFn.apply(last[target].detach(), last["weights"], p, premult, mult)


def param_regularizer_loss(tables, mult=1.0):
    """param_regularizer_loss (internal/train_utils.py:1169-1216), (mult, jnp.mean, 2, 1) setting: autograd mirror in
    torch ops over the level tables (the fused step uses nrc_grid_regularizer)."""
    loss = 0.0
    for t in tables:
        loss = loss + mult * 0.5 * torch.mean(t**2)
    return loss
