"""Host-side mirror of internal/coord.py pieces on the hot path (CUDA bodies)."""
import torch

from . import _lib


class _ContractFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c):
        x2 = x.reshape(-1, 3).contiguous()
        z = torch.empty_like(x2)
        _lib.call("nrc_contract_fwd", _lib.stream_ptr(), _lib.ptr(x2), x2.shape[0], float(c), _lib.ptr(z))
        ctx.save_for_backward(x2)
        ctx.c = c
        ctx.shape = x.shape
        return z.reshape(x.shape)

    @staticmethod
    def backward(ctx, gz):
        (x2,) = ctx.saved_tensors
        g2 = gz.reshape(-1, 3).contiguous()
        gx = torch.empty_like(x2)
        _lib.call("nrc_contract_bwd", _lib.stream_ptr(), _lib.ptr(x2), _lib.ptr(g2), x2.shape[0], float(ctx.c),
                  _lib.ptr(gx))
        return gx.reshape(ctx.shape), None


def contract(x):
    """Contracts points towards the origin (internal/coord.py:63-69)."""
    return _ContractFn.apply(x, 1.0)


def contract_radius_5(x, c=5.0):
    """internal/coord.py:33-34."""
    return _ContractFn.apply(x, c)


def contract_radius_2(x, c=2.0):
    """internal/coord.py:37-38."""
    return _ContractFn.apply(x, c)
