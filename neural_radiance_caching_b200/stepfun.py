"""Host-side mirror of internal/stepfun.py for the hot path (CUDA bodies, same signatures).

Randomness: the reference passes a jax PRNG key; RNG is not part of the kernel contract
(SURVEY section 7 "RNG parity"), so `rng` here is either None (not used on the hot path) or a
tensor of uniforms in [0,1) with shape t.shape[:-1] + (1,) -- what
jax.random.uniform(rng, shape) would have produced before scaling by max_jitter.
"""
import functools

import numpy as np
import torch

from . import _lib

_EPS = np.float32(np.finfo(np.float32).eps)


def _linspace_f32(start, stop, num):
    """jnp.linspace semantics in fp32 (start*(1-i/div) + stop*(i/div), last = stop)."""
    start, stop = np.float32(start), np.float32(stop)
    div = num - 1
    step = (np.arange(div, dtype=np.float32) / np.float32(div)).astype(np.float32)
    out = start * (np.float32(1) - step) + stop * step
    return np.concatenate([out, np.array([stop], np.float32)]).astype(np.float32)


@functools.lru_cache(maxsize=None)
def _u_base_host(num_samples):
    """Deterministic part of `u` (internal/stepfun.py:196-201) and the jitter scale."""
    u_max = _EPS + (np.float32(1) - _EPS) / np.float32(num_samples)
    max_jitter = (np.float32(1) - u_max) / np.float32(num_samples - 1) - _EPS
    return _linspace_f32(0.0, np.float32(1) - u_max, num_samples), float(max_jitter)


_u_base_dev = {}


def u_base(num_samples, device):
    key = (num_samples, str(device))
    if key not in _u_base_dev:
        base, mj = _u_base_host(num_samples)
        _u_base_dev[key] = (torch.from_numpy(base).to(device), mj)
    return _u_base_dev[key]


def sample_intervals_from_weights(u01, t, weights, num_samples, anneal=1.0, padding=0.0,
                                  domain=(-float("inf"), float("inf")), return_bins=False):
    """sampling.py:340 + stepfun.sample_intervals in one launch:
    logits = anneal * safe_log(weights + padding); single_jitter=True."""
    if num_samples <= 1:
        raise ValueError(f"num_samples must be > 1, is {num_samples}.")
    m = weights.shape[-1]
    if t.shape[-1] != m + 1:
        raise ValueError(f"Invalid shapes ({t.shape}, {weights.shape}) for a step function.")
    t2 = t.reshape(-1, m + 1).contiguous()
    w2 = weights.reshape(-1, m).contiguous()
    R = t2.shape[0]
    u2 = u01.reshape(-1).contiguous()
    if u2.shape[0] != R:
        raise ValueError("single_jitter=True needs one uniform per ray")
    base, max_jitter = u_base(num_samples, t.device)
    out = torch.empty((R, num_samples + 1), device=t.device, dtype=torch.float32)
    bins = torch.empty((R, num_samples), device=t.device, dtype=torch.int32) if return_bins else None
    _lib.call("nrc_ray_sample_intervals", _lib.stream_ptr(), _lib.ptr(t2), _lib.ptr(w2), _lib.ptr(u2),
              _lib.ptr(base), R, m, num_samples, float(anneal), float(padding), max_jitter, float(domain[0]),
              float(domain[1]), _lib.ptr(out), _lib.ptr(bins))
    out = out.reshape(t.shape[:-1] + (num_samples + 1,))
    if return_bins:
        return out, bins.reshape(t.shape[:-1] + (num_samples,))
    return out


def sample_intervals(rng, t, w_logits, num_samples, single_jitter=False, domain=(-float("inf"), float("inf"))):
    """Sample *intervals* from a step function (internal/stepfun.py:207-250).

    The CUDA body takes weights, not logits (it fuses the sampler's annealed safe_log);
    logits are mapped back with weights = exp(logits), anneal = 1, padding = 0, which
    leaves softmax(logits) unchanged.
    """
    if num_samples <= 1:
        raise ValueError(f"num_samples must be > 1, is {num_samples}.")
    if rng is None or not single_jitter:
        raise NotImplementedError("the hot path uses single_jitter=True with supplied uniforms")
    w = torch.exp(w_logits - w_logits.max(dim=-1, keepdim=True).values)
    return sample_intervals_from_weights(rng, t, w, num_samples, 1.0, 0.0, domain)
