"""Oracle restatement of internal/math.py (TEST INFRASTRUCTURE ONLY).

PyTorch-CPU fp32; custom-JVP clip semantics are reproduced with
torch.autograd.Function so that oracle gradients follow the reference's.
"""
import numpy as np
import torch

tiny_val = float(np.finfo(np.float32).tiny)  # internal/math.py:24
min_val = float(np.finfo(np.float32).min)  # internal/math.py:25
max_val = float(np.finfo(np.float32).max)  # internal/math.py:26
EPS = float(np.finfo(np.float32).eps)


def sqrt(x):
    """Correctly rounded fp32 sqrt.  torch.sqrt on CPU dispatches large tensors to MKL VML,
    which is NOT correctly rounded (0.4% of inputs are 1 ulp off); XLA:CPU and CUDA's
    sqrtf are IEEE-exact, so the oracle goes through float64."""
    return torch.sqrt(x.double()).to(x.dtype)


class _SafeExp(torch.autograd.Function):
    """internal/math.py:186-192: exp(clip(x, min, 70)); grad = y * x_dot."""

    @staticmethod
    def forward(ctx, x):
        y = torch.exp(torch.clamp(x, min=min_val, max=70.0))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g * y


def safe_exp(x):
    return _SafeExp.apply(x)


class _SafeLog(torch.autograd.Function):
    """internal/math.py:177-183: log(clip(x, tiny, max)); grad = x_dot / clip(x)."""

    @staticmethod
    def forward(ctx, x):
        xc = torch.clamp(x, min=tiny_val, max=max_val)
        ctx.save_for_backward(xc)
        return torch.log(xc)

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        return g / xc


def safe_log(x):
    return _SafeLog.apply(x)


def safe_sign(x):
    """internal/math.py:128-130."""
    return torch.where(x < 0, -torch.ones_like(x), torch.ones_like(x))


def remove_zero(x):
    """internal/math.py:133-135."""
    return torch.where(torch.abs(x) < tiny_val, torch.full_like(x, tiny_val), x)


def safe_div(n, d):
    """internal/math.py:138-158 (forward value; autograd of the clipped form)."""
    r = torch.clamp(n / remove_zero(d), min_val, max_val)
    return torch.where(torch.abs(d) < tiny_val, torch.zeros_like(r), r)


def power_ladder(x, p, premult=None, postmult=None):
    """internal/math.py:295-316, general branch (p not in {1, 0, +-inf})."""
    assert p not in (1.0, 0.0, float("inf"), float("-inf"))
    if premult is not None:
        x = x * premult
    xp = torch.abs(x)
    xs = xp / max(tiny_val, abs(p - 1))
    y = safe_sign(x) * (abs(p - 1) / p * ((xs + 1) ** p - 1))
    if postmult is not None:
        y = y * postmult
    return y


def power_ladder_max_output(p):
    """internal/math.py:284-292."""
    if p == float("-inf"):
        return 1.0
    if p >= 0:
        return float("inf")
    return (p - 1) / p


def inv_power_ladder(y, p, premult=None, postmult=None):
    """internal/math.py:319-341, general branch."""
    assert p not in (1.0, 0.0, float("inf"), float("-inf"))
    if postmult is not None:
        y = y / postmult
    yp = torch.abs(y)
    y_max = float(np.nextafter(np.float32(power_ladder_max_output(p)), np.float32(-np.inf)))
    yp = torch.clamp(yp, -y_max, y_max)
    # safe_div(p, |p-1|) is evaluated in fp32 in the reference (p is a weak
    # python float promoted against yp's dtype).
    ratio = float(np.float32(p) / np.float32(abs(p - 1)))
    x = safe_sign(y) * (abs(p - 1) * ((ratio * yp + 1) ** (1 / p) - 1))
    if premult is not None:
        x = x / premult
    return x


def sorted_lookup_idx(x, xp):
    """internal/math.py:433-439: searchsorted(xp, x, side='right') then clamp."""
    idx = torch.searchsorted(xp.contiguous(), x.contiguous(), right=True)
    idx1 = torch.clamp(idx, max=xp.shape[-1] - 1)
    idx0 = torch.clamp(idx - 1, min=0)
    return idx0, idx1


def sorted_interp(x, xp, fp, eps=EPS**2):
    """internal/math.py:447-457."""
    idx0, idx1 = sorted_lookup_idx(x, xp)
    xp0 = torch.gather(xp, -1, idx0)
    xp1 = torch.gather(xp, -1, idx1)
    fp0 = torch.gather(fp, -1, idx0)
    fp1 = torch.gather(fp, -1, idx1)
    offset = torch.clamp((x - xp0) / torch.clamp(xp1 - xp0, min=eps), 0, 1)
    return fp0 + offset * (fp1 - fp0), idx0


def approx_erf(x):
    """internal/math.py:365-367."""
    return torch.sign(x) * sqrt(1 - torch.exp(-(4 / np.pi) * x**2))


def l2_normalize(x, grad_eps=EPS, tiny=tiny_val):
    """internal/ref_utils.py:45-70 (forward uses tiny, backward uses grad_eps)."""
    grad_eps = max(tiny, grad_eps)
    denom_sq = torch.sum(x**2, dim=-1, keepdim=True)
    normal_val = x / sqrt(torch.clamp(denom_sq, min=tiny))
    normal_grad = x / sqrt(torch.clamp(denom_sq, min=grad_eps))
    normal = normal_val.detach() + (normal_grad - normal_grad.detach())
    return torch.where(denom_sq < tiny, torch.zeros_like(normal), normal)
