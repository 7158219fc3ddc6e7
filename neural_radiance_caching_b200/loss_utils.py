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
