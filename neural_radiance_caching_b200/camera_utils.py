"""Host-side mirror of internal/camera_utils.py for the render path: intrinsic_matrix / get_pixtocam (:749-763) and
pixels_to_rays (:896-1073) + the near / far broadcast of cast_ray_batch (:1225-1330) for the perspective camera without
distortion, NDC or jitter - the body is nrc_camera_rays (csrc/camera.cu)."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def intrinsic_matrix(fx, fy, cx, cy):
    """camera_utils.intrinsic_matrix (:749-757): pinhole intrinsics, OpenCV convention."""
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])


def get_pixtocam(focal, width, height):
    """camera_utils.get_pixtocam (:760-763): inverse intrinsics of a perfect pinhole camera."""
    return np.linalg.inv(intrinsic_matrix(focal, focal, width * 0.5, height * 0.5))


def _f32_array(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    if a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return (C.c_float * n)(*a.tolist())


def pixels_to_rays(first_pixel, num_rays, width, height, pixtocam, camtoworld, device, near=None, far=None,
                   d_first_pixel=None, last_pixel=None, out=None, distortion_params=None, pixtocam_ndc=None, jitter=0):
    """Rays of `num_rays` consecutive pixels (row-major flat index, starting at `first_pixel` or at the device counter
    `d_first_pixel` [1] int64) as a dict of device tensors: origins / directions / viewdirs [N,3], radii [N,1],
    imageplane [N,2] and, when near / far are given, near / far [N,1].  `out` reuses the tensors of an earlier call
    (static buffers of a captured graph).  Unsupported settings of the reference function raise."""
    if distortion_params is not None or pixtocam_ndc is not None or jitter:
        raise NotImplementedError("the CUDA path covers the perspective camera without distortion, NDC or jitter")
    new = lambda *s: torch.empty(s, device=device, dtype=torch.float32)
    if out is None:
        out = dict(origins=new(num_rays, 3), directions=new(num_rays, 3), viewdirs=new(num_rays, 3), radii=new(num_rays, 1),
                   imageplane=new(num_rays, 2))
        if near is not None:
            out["near"], out["far"] = new(num_rays, 1), new(num_rays, 1)
    last = width * height - 1 if last_pixel is None else int(last_pixel)
    cnt = C.c_void_p(d_first_pixel.data_ptr()) if d_first_pixel is not None else None
    _lib.call("nrc_camera_rays", _lib.stream_ptr(), _f32_array(pixtocam, 9), _f32_array(np.asarray(camtoworld)[:3, :4], 12),
              int(width), int(height), int(first_pixel), cnt, last, int(num_rays), float(near or 0.0), float(far or 0.0),
              _lib.ptr(out["origins"]), _lib.ptr(out["directions"]), _lib.ptr(out["viewdirs"]), _lib.ptr(out["radii"]),
              _lib.ptr(out.get("imageplane")), _lib.ptr(out.get("near")), _lib.ptr(out.get("far")))
    return out
