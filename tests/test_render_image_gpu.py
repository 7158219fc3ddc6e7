"""Row 23 / 8f-3 on the device: nrc_camera_rays against the reference's own pixels_to_rays (tests/golden/reference_np.npz),
and the whole-image chunk loop (models.render_image, internal/models.py:2361-2525): the device-side schedule (one CUDA
graph replayed per chunk, chunk counter on the device) against the Python loop, and row bands against the whole image."""
import os

import numpy as np
import pytest
import torch

from neural_radiance_caching_b200 import camera_utils as ncam, models as nmodels, render_image as ri, workload
from tests.util import rel_err

pytestmark = pytest.mark.gpu
V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_np.npz"))


def test_pixels_to_rays_against_reference(cuda_device):
    Wc, Hc = (int(v) for v in V["cam_size"])
    rays = ncam.pixels_to_rays(0, Wc * Hc, Wc, Hc, V["cam_pixtocam"], V["cam_camtoworld"], cuda_device, near=2.0, far=6.0)
    for k in ("origins", "directions", "viewdirs", "radii", "imageplane"):
        want = torch.from_numpy(V["cam_" + k]).reshape(Wc * Hc, -1)
        # the cone radius is a norm of DIFFERENCES of neighbouring pixel directions (three digits cancel)
        assert rel_err(rays[k], want) <= (5e-6 if k == "radii" else 1e-6), (k, rel_err(rays[k], want))
    assert float(rays["near"].min()) == 2.0 and float(rays["far"].max()) == 6.0
    # a chunk that starts at a device counter and runs past the image repeats the last pixel (edge padding)
    cnt = torch.tensor([Wc * Hc - 5], device=cuda_device, dtype=torch.int64)
    tail = ncam.pixels_to_rays(0, 16, Wc, Hc, V["cam_pixtocam"], V["cam_camtoworld"], cuda_device, d_first_pixel=cnt)
    assert torch.equal(tail["directions"][:5], rays["directions"][-5:])
    assert torch.equal(tail["directions"][5:], rays["directions"][-1:].expand(11, 3))
    assert np.allclose(ncam.get_pixtocam(17.3, Wc, Hc).astype(np.float32), V["cam_pixtocam"], rtol=0, atol=0)


class _Renderer:
    """Deterministic chunk renderer: the cache model (bf16 tensor-core path) with a fixed jitter, no random draws."""

    def __init__(self, dev):
        gen = torch.Generator(device=dev)
        gen.manual_seed(11)
        self.model = nmodels.NeRFModel(bf16=True)
        self.params = workload._cache_params(self.model, dev, gen, 0.1)
        self.dev = dev

    def __call__(self, rays, repeat=0):
        R = rays["origins"].shape[0]
        with torch.no_grad():
            u = [torch.full((R, 1), 0.5, device=self.dev) for _ in range(3)]
            res = self.model(self.params, rays, u, train=False)["render"]
        return dict(rgb=res["rgb"], acc=res["acc"].reshape(R, 1))


def test_device_schedule_and_bands(cuda_device):
    dev = cuda_device
    H, W, focal = 40, 52, 60.0           # 2080 pixels: three chunks of 1024 with an edge-padded tail
    c2w = ri.orbit_camera()
    fn = _Renderer(dev)
    loop = ri.render_image(fn, H, W, focal, c2w, dev, chunk=1024, device_scheduler=False)
    sched = ri.render_image(fn, H, W, focal, c2w, dev, chunk=1024, device_scheduler=True)
    assert loop["rgb"].shape == (H, W, 3) and sched["acc"].shape == (H, W, 1)
    for k in ("rgb", "acc"):
        assert torch.equal(loop[k], sched[k]), k
    assert float(sched["acc"].max()) > 0.0
    # two row bands (what two ranks render) put together == the whole image, bit for bit: chunk composition differs,
    # per-ray results do not
    top = ri.render_image(fn, H, W, focal, c2w, dev, chunk=1024, band=(0, 17))
    bottom = ri.render_image(fn, H, W, focal, c2w, dev, chunk=1024, band=(17, H))
    for k in ("rgb", "acc"):
        assert torch.equal(torch.cat([top[k], bottom[k]], dim=0), sched[k]), k
    # replaying a cached graph renders the same frame again
    cache = {}
    a = ri.render_image(fn, H, W, focal, c2w, dev, chunk=1024, _graphs=cache)["rgb"].clone()
    b = ri.render_image(fn, H, W, focal, c2w, dev, chunk=1024, _graphs=cache)["rgb"]
    assert len(cache) == 1 and torch.equal(a, b) and torch.equal(a, sched["rgb"])
