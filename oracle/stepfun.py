"""Oracle restatement of internal/stepfun.py (TEST INFRASTRUCTURE ONLY).

Randomness: the reference draws the per-ray jitter with jax.random.uniform
(threefry).  RNG is not part of the kernel contract: the oracle and the CUDA
kernels both take the uniform draw `u01` in [0,1) as an input tensor and scale
it by max_jitter exactly like jax.random.uniform(maxval=max_jitter) does
(u01 * (maxval - 0) + 0, jax/_src/random.py uniform: floats * (maxval-minval) + minval,
then max(minval, .)).
"""
import numpy as np
import torch

from . import ref_math

EPS = float(np.finfo(np.float32).eps)


def linspace_f32(start, stop, num):
    """jnp.linspace in fp32 (jax 0.4.16 jax/_src/numpy/lax_numpy.py `linspace`):
    out[i] = start*(1 - i/div) + stop*(i/div) for i < div, out[div] = stop."""
    start = np.float32(start)
    stop = np.float32(stop)
    div = num - 1
    step = (np.arange(div, dtype=np.float32) / np.float32(div)).astype(np.float32)
    out = start * (np.float32(1) - step) + stop * step
    return np.concatenate([out, np.array([stop], np.float32)]).astype(np.float32)


def sample_u_base(num_samples):
    """The deterministic part of `u` in stepfun.sample (:196-201) and the jitter scale."""
    eps = np.float32(EPS)
    u_max = eps + (np.float32(1) - eps) / np.float32(num_samples)
    max_jitter = (np.float32(1) - u_max) / np.float32(num_samples - 1) - eps
    base = linspace_f32(0.0, np.float32(1) - u_max, num_samples)
    return base, np.float32(max_jitter)


def integrate_weights(w):
    """internal/stepfun.py:125-144."""
    cw = torch.clamp(torch.cumsum(w[..., :-1], dim=-1), max=1.0)
    shape = cw.shape[:-1] + (1,)
    return torch.cat([torch.zeros(shape), cw, torch.ones(shape)], dim=-1)


def invert_cdf(u, t, w_logits, return_idx=False):
    """internal/stepfun.py:147-155."""
    w = torch.softmax(w_logits, dim=-1)
    cw = integrate_weights(w)
    t_new, idx0 = ref_math.sorted_interp(u, cw, t)
    if return_idx:
        return t_new, idx0
    return t_new


def sample(u01, t, w_logits, num_samples, single_jitter=False, deterministic_center=False, return_idx=False):
    """internal/stepfun.py:158-204.  u01 None -> linspace sampling (rng=None)."""
    eps = EPS
    if u01 is None:
        if deterministic_center:
            pad = 1 / (2 * num_samples)
            u = torch.from_numpy(linspace_f32(pad, 1.0 - pad - eps, num_samples))
        else:
            u = torch.from_numpy(linspace_f32(0, 1.0 - eps, num_samples))
        u = u.expand(t.shape[:-1] + (num_samples,))
    else:
        base, max_jitter = sample_u_base(num_samples)
        d = 1 if single_jitter else num_samples
        assert u01.shape[-1] == d
        jitter = torch.clamp(u01 * float(max_jitter), min=0.0)
        u = torch.from_numpy(base) + jitter
    return invert_cdf(u, t, w_logits, return_idx=return_idx)


def sample_intervals(u01, t, w_logits, num_samples, single_jitter=False, domain=(-float("inf"), float("inf")),
                     return_idx=False):
    """internal/stepfun.py:207-250."""
    if num_samples <= 1:
        raise ValueError(f"num_samples must be > 1, is {num_samples}.")
    res = sample(u01, t, w_logits, num_samples, single_jitter, deterministic_center=True, return_idx=return_idx)
    centers, idx0 = res if return_idx else (res, None)
    mid = (centers[..., 1:] + centers[..., :-1]) / 2
    first = 2 * centers[..., :1] - mid[..., :1]
    last = 2 * centers[..., -1:] - mid[..., -1:]
    samples = torch.cat([first, mid, last], dim=-1)
    samples = torch.sort(torch.clamp(samples, domain[0], domain[1]), dim=-1).values
    if return_idx:
        return samples, idx0
    return samples


def _interp(x, xp, fp):
    """jnp.interp (jax 0.4.16 lax_numpy._interp), left/right = None, no period."""
    i = torch.clamp(torch.searchsorted(xp.contiguous(), x.contiguous(), right=True), 1, xp.shape[-1] - 1)
    fp_hi = torch.gather(fp, -1, i)
    fp_lo = torch.gather(fp, -1, i - 1)
    xp_hi = torch.gather(xp, -1, i)
    xp_lo = torch.gather(xp, -1, i - 1)
    df = fp_hi - fp_lo
    dx = xp_hi - xp_lo
    delta = x - xp_lo
    epsilon = float(np.spacing(np.finfo(np.float32).eps))
    dx0 = torch.abs(dx) <= epsilon
    f = torch.where(dx0, fp_lo, fp_lo + (delta / torch.where(dx0, torch.ones_like(dx), dx)) * df)
    f = torch.where(x < xp[..., :1], fp[..., :1].expand_as(f), f)
    f = torch.where(x > xp[..., -1:], fp[..., -1:].expand_as(f), f)
    return f


def weighted_percentile(t, w, ps):
    """internal/stepfun.py:306-314."""
    cw = integrate_weights(w)
    x = (torch.tensor(ps, dtype=torch.float32) / 100).expand(t.shape[:-1] + (len(ps),))
    return _interp(x, cw, t)
