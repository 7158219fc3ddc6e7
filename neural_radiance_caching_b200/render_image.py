"""Row 23 of SURVEY 8a: whole-image rendering.  models.render_image (internal/models.py:2361-2525) chunk
loop + utils.shard/unshard (internal/utils.py:333-343) + the all_gather of rendered chunks
(internal/train_utils.py:3795-3815), re-designed for one process per GPU:

  * the image is split into contiguous ROW BANDS, one per rank (dist.row_bands); every rank runs full
    1024-ray chunks (a 1024-ray chunk is never split across ranks: 128 rays per GPU is below launch-latency
    scale);
  * chunk outputs are written into a preallocated device band [rows*W, C] (no host sync per chunk);
  * `render_repeats` averaging uses Welford's running mean / M2 on the device;
  * the bands are assembled with ONE all_gather (dist.gather_tiles).

Camera rays: camera_utils.pixels_to_rays (internal/camera_utils.py:896-1073) for the perspective / no-distortion case
(pixel centres at +0.5, OpenGL camera: x right, y up, looking down -z), on the device (nrc_camera_rays).

On a CUDA device the chunk loop itself leaves Python: the chunk's first pixel is a DEVICE counter, ONE CUDA graph -
camera rays of the chunk, the renderer, the stores of its results into the band, the counter's increment - is captured
and replayed once per chunk (`device_scheduler=True`, the default for capturable renderers)."""
import ctypes as C

import numpy as np
import torch

from . import dist as ndist


def orbit_camera(radius=4.0, azimuth_deg=30.0, elevation_deg=30.0):
    """cam-to-world [3,4] of a camera on a sphere looking at the origin (TensoIR-style synthetic view)."""
    az, el = np.deg2rad(azimuth_deg), np.deg2rad(elevation_deg)
    pos = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    return np.concatenate([np.stack([right, up, -fwd], axis=1), pos[:, None]], axis=1).astype(np.float32)


def pinhole_rays(height, width, focal, camtoworld, device, rows=None, near=2.0, far=6.0, radii_scale=None):
    """Rays of the pixel rows [rows[0], rows[1]) as a dict of [N,·] device tensors (N = rows*width):
    origins, directions (NOT unit length, like the reference), viewdirs (unit), radii, near, far - pixels_to_rays
    of the reference (camera_utils.py:896-1073) through nrc_camera_rays; on a CPU device (host-logic tests) the same
    expressions in torch."""
    r0, r1 = rows if rows is not None else (0, height)
    n = (r1 - r0) * width
    if torch.device(device).type == "cuda":
        from . import camera_utils
        rays = camera_utils.pixels_to_rays(r0 * width, n, width, height, camera_utils.get_pixtocam(focal, width, height),
                                           camtoworld, device, near=near, far=far)
        if radii_scale is not None:
            rays["radii"].fill_(float(radii_scale))
        rays.pop("imageplane")
        return rays
    c2w = torch.as_tensor(camtoworld, device=device, dtype=torch.float32)
    ys, xs = torch.meshgrid(torch.arange(r0, r1, device=device, dtype=torch.float32),
                            torch.arange(width, device=device, dtype=torch.float32), indexing="ij")
    cam = torch.stack([(xs + 0.5 - 0.5 * width) / focal, -(ys + 0.5 - 0.5 * height) / focal, -torch.ones_like(xs)], dim=-1)
    dirs = (cam.reshape(-1, 3) @ c2w[:, :3].T).contiguous()
    viewdirs = dirs / torch.linalg.norm(dirs, dim=-1, keepdim=True)
    # pixel footprint radius: distance between neighbouring pixel directions * 2 / sqrt(12) (camera_utils.py:1046-1062)
    rad = (1.0 / focal) * 2.0 / np.sqrt(12.0) if radii_scale is None else radii_scale
    full = lambda v: torch.full((n, 1), float(v), device=device, dtype=torch.float32)
    return dict(origins=c2w[:, 3].expand(n, 3).contiguous(), directions=dirs, viewdirs=viewdirs.contiguous(),
                radii=full(rad), near=full(near), far=full(far))


class ChunkGraph:
    """One captured chunk of a band render: camera rays of the pixels [counter, counter + chunk) -> render_chunk_fn ->
    results stored at their band position -> counter += chunk.  Replaying it n times renders n consecutive chunks with no
    host work in between (models.render_image's Python loop, internal/models.py:2361-2525, as a device-side schedule)."""

    def __init__(self, render_chunk_fn, height, width, focal, camtoworld, device, chunk, band, near, far, repeat=0):
        from . import _lib, camera_utils
        r0, r1 = band
        self.n = (r1 - r0) * width
        self.chunk = chunk
        self.counter = torch.zeros((1,), device=device, dtype=torch.int64)
        self.first = r0 * width
        cnt = C.c_void_p(self.counter.data_ptr())
        pixtocam = camera_utils.get_pixtocam(focal, width, height)

        def body():
            rays = camera_utils.pixels_to_rays(0, chunk, width, height, pixtocam, camtoworld, device, near=near, far=far,
                                               d_first_pixel=self.counter, last_pixel=r1 * width - 1)
            rays.pop("imageplane")
            res = render_chunk_fn(rays, repeat)
            for k, v in res.items():
                v2 = v.reshape(chunk, -1).contiguous()
                if k not in self.bands:
                    self.bands[k] = torch.empty((self.n, v2.shape[1]), device=device, dtype=torch.float32)
                _lib.call("nrc_band_store", _lib.stream_ptr(), _lib.ptr(v2), v2.shape[1], cnt, self.first, self.n, chunk,
                          _lib.ptr(self.bands[k]))
            _lib.call("nrc_chunk_advance", _lib.stream_ptr(), cnt, chunk)

        self.bands = {}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):       # warm-up (allocations, lazily built state of the renderer)
                self.counter.fill_(self.first)
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()

    def render_band(self):
        self.counter.fill_(self.first)
        for _ in range((self.n + self.chunk - 1) // self.chunk):
            self.graph.replay()
        return self.bands


class Welford:
    """Running mean / variance over render repeats (models.py:2455-2500), on the device."""

    def __init__(self):
        self.n, self.mean, self.m2 = 0, None, None

    def update(self, x):
        self.n += 1
        if self.mean is None:
            self.mean, self.m2 = x.clone(), torch.zeros_like(x)
            return
        delta = x - self.mean
        self.mean += delta / self.n
        self.m2 += delta * (x - self.mean)

    def variance(self):
        return self.m2 / max(self.n - 1, 1)


def render_image(render_chunk_fn, height, width, focal, camtoworld, device, chunk=1024, num_repeats=1,
                 compute_variance=False, near=2.0, far=6.0, gather=True, band=None, device_scheduler=None, _graphs=None):
    """render_chunk_fn(rays_chunk: dict of [chunk,·] tensors, repeat: int) -> dict of [chunk, C] tensors.
    Every rank renders its row band in fixed-size chunks (the last chunk is edge-padded like
    models.py:2434-2445) and the bands are gathered.  Returns dict of [H, W, C] tensors (+ '<key>_var').
    `band` = (row0, row1) overrides the rank's band (tests).  `device_scheduler` (default: on for CUDA devices and
    a single repeat) runs the chunk loop as replays of one CUDA graph (ChunkGraph); `_graphs` (a dict the caller keeps)
    caches the captured graph across frames of the same camera and size."""
    rank, world = ndist.world()
    r0, r1 = band if band is not None else ndist.row_bands(height, world)[rank]
    if device_scheduler is None:
        device_scheduler = torch.device(device).type == "cuda" and num_repeats == 1
    if device_scheduler:
        key = (height, width, float(focal), np.asarray(camtoworld, dtype=np.float32).tobytes(), chunk, r0, r1, near, far)
        cache = _graphs if _graphs is not None else {}
        if key not in cache:
            cache[key] = ChunkGraph(render_chunk_fn, height, width, focal, camtoworld, device, chunk, (r0, r1), near, far)
        bands = cache[key].render_band()
        image = {}
        for k, v in bands.items():
            b3 = v.reshape(r1 - r0, width, -1)
            image[k] = ndist.gather_tiles(b3, height) if (gather and band is None) else b3
        return image
    rays = pinhole_rays(height, width, focal, camtoworld, device, rows=(r0, r1), near=near, far=far)
    n = (r1 - r0) * width
    bands, stats = {}, {}
    for rep in range(num_repeats):
        outs = {}
        for c0 in range(0, n, chunk):
            c1 = min(c0 + chunk, n)
            idx = None
            if c1 - c0 < chunk:   # edge padding: repeat the last ray so that every launch sees a full chunk
                idx = torch.cat([torch.arange(c0, c1, device=device),
                                 torch.full((chunk - (c1 - c0),), c1 - 1, device=device, dtype=torch.long)])
            rc = {k: (v[c0:c1] if idx is None else v[idx].contiguous()) for k, v in rays.items()}
            res = render_chunk_fn(rc, rep)
            for k, v in res.items():
                if k not in outs:
                    outs[k] = torch.empty((n,) + tuple(v.shape[1:]), device=device, dtype=v.dtype)
                outs[k][c0:c1] = v[:c1 - c0]
        for k, v in outs.items():
            stats.setdefault(k, Welford()).update(v)
    for k, w in stats.items():
        bands[k] = w.mean
        if compute_variance and num_repeats > 1:
            bands[k + "_var"] = w.variance()
    image = {}
    for k, v in bands.items():
        band_k = v.reshape(r1 - r0, width, -1)
        image[k] = ndist.gather_tiles(band_k, height) if (gather and band is None) else band_k
    return image
