"""engine.FusedCacheStep (hand-ordered launch schedule of the config-2 training step) against the
autograd mirrors running the same kernels: loss and every gradient must agree up to the summation
order of the atomic reductions."""
import numpy as np
import pytest
import torch

from neural_radiance_caching_b200 import workload
from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _batch(dev, R, seed_offset=0):
    g = np.random.Generator(np.random.PCG64(workload.SEED + 7 + seed_offset))
    rn = workload.make_rays_np(g, R)
    u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
    tgt = g.uniform(size=(R, 3)).astype(np.float32)
    buf = torch.from_numpy(workload.pack_rays(rn, u, tgt)).to(dev)
    return workload.unpack_rays(buf)


def _full_batch(dev, R, seed_offset=0):
    """Main rays + the backward-mask rays, through the packed batch buffer the bench uses."""
    g = np.random.Generator(np.random.PCG64(workload.SEED + 7 + seed_offset))
    rn = workload.make_rays_np(g, R)
    u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
    tgt = g.uniform(size=(R, 3)).astype(np.float32)
    xr = workload.backward_mask_rays_np(g, rn)
    ux = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
    buf = torch.from_numpy(workload.pack_batch(rn, u, tgt, xr, ux)).to(dev)
    return workload.unpack_batch(buf)


@pytest.mark.parametrize("R,with_extra", [(256, True), (1000, True), (512, False)])
def test_fused_step_matches_autograd(cuda_device, R, with_extra):
    """The whole config-2 objective: data + interlevel + geometry (incl. the second-order path) + mask
    (+ backward-mask pass on the extra rays)."""
    step = workload.CacheTrainStep(cuda_device, bf16=True)
    assert step.engine is not None
    rays, u01, tgt, extra = _full_batch(cuda_device, R)
    if not with_extra:
        extra = None
    loss_a = float(step.step_autograd(rays, u01, tgt, extra))
    grad_a = step.flat_grad.clone()
    loss_f = float(step.step(rays, u01, tgt, extra))
    grad_f = step.flat_grad.clone()
    assert abs(loss_f - loss_a) <= 1e-5 * max(1.0, abs(loss_a))
    assert float(grad_a.abs().max()) > 0
    off = 0
    for t in step.leaves:   # per parameter, so that small tensors are not hidden by large ones
        n = t.numel()
        pad = (n + 63) // 64 * 64
        a, f = grad_a[off:off + n], grad_f[off:off + n]
        if float(a.abs().max()) > 0:
            assert rel_l2(f, a) <= 1e-3, (tuple(t.shape), rel_l2(f, a))
        else:
            assert float(f.abs().max()) == 0.0
        off += pad
    # the forward state the engine exposes matches the autograd model's rendering
    with torch.no_grad():
        res = step.model(step.params, rays, u01, train=False)
    assert rel_err(step.engine.last["rgb"], res["render"]["rgb"]) <= 1e-5


@pytest.mark.parametrize("secondary", [False, True])
def test_fused_query_matches_model(cuda_device, secondary):
    """engine.FusedCacheQuery (render path schedule) == models.NeRFModel.__call__ on the same kernels."""
    from neural_radiance_caching_b200 import engine
    step = workload.CacheTrainStep(cuda_device, bf16=True)
    R = 500
    g = np.random.Generator(np.random.PCG64(workload.SEED + 11 + int(secondary)))
    rn = workload.make_rays_np(g, R, near=0.05, far=2.0, radius=0.7) if secondary else workload.make_rays_np(g, R)
    rays = {k: torch.from_numpy(v).to(cuda_device) for k, v in rn.items()}
    u01 = [torch.from_numpy(g.uniform(size=(R, 1)).astype(np.float32)).to(cuda_device) for _ in range(3)]
    gum = torch.from_numpy(g.gumbel(size=(R, 32, 1)).astype(np.float32)).to(cuda_device)
    for resample in (False, True):
        with torch.no_grad():
            want = step.model(step.params, rays, u01, gumbel=gum if resample else None, train=False,
                              is_secondary=secondary, resample=resample)
            got = engine.FusedCacheQuery(step.model)(step.params, rays, u01, gumbel=gum if resample else None,
                                                     is_secondary=secondary, resample=resample)
        assert rel_err(got["rgb"], want["render"]["rgb"]) <= 1e-5
        assert rel_err(got["acc"], want["render"]["acc"]) <= 1e-5
        assert rel_err(got["distance"][:, 0], want["render"]["distance_mean"]) <= 1e-5
        if resample:
            assert torch.equal(got["inds"], want["inds"])
            assert rel_err(got["means"], want["shaded"]["means"]) <= 1e-6
            assert rel_err(got["normals"], want["shaded"]["normals"]) <= 1e-6
