"""Test infrastructure (CPU oracle): restatement of internal/camera_utils.py pixels_to_rays (:896-1073) for the
perspective camera without distortion, NDC or jitter, and get_pixtocam (:749-763), in NumPy float32 with the
reference's op order.  Only tests/ and bench.py's CPU legs may import it."""
import numpy as np


def get_pixtocam(focal, width, height):
    camtopix = np.array([[focal, 0, width * 0.5], [0, focal, height * 0.5], [0, 0, 1.0]])
    return np.linalg.inv(camtopix)


def pixels_to_rays(pix_x, pix_y, pixtocam, camtoworld):
    """pix_x / pix_y float32 arrays of any shape SH -> origins, directions, viewdirs [SH,3], radii [SH,1], imageplane [SH,2]."""
    f = np.float32
    pix_x, pix_y = np.asarray(pix_x, f), np.asarray(pix_y, f)
    k, c2w = np.asarray(pixtocam, f), np.asarray(camtoworld, f)

    def pix_to_dir(x, y):
        return np.stack([x + f(0.5), y + f(0.5), np.ones_like(x)], axis=-1)

    stacked = np.stack([pix_to_dir(pix_x, pix_y), pix_to_dir(pix_x + f(1), pix_y), pix_to_dir(pix_x, pix_y + f(1))], axis=0)
    mat_vec = lambda A, b: np.matmul(A, b[..., None])[..., 0].astype(f)
    cam = mat_vec(k, stacked)
    cam = np.matmul(cam, np.diag(np.array([1.0, -1.0, -1.0], f))).astype(f)      # OpenCV -> OpenGL
    imageplane = cam[0, ..., :2]
    dirs = mat_vec(c2w[:3, :3], cam)
    directions, dx, dy = dirs
    origins = np.broadcast_to(c2w[:3, -1], directions.shape)
    viewdirs = directions / np.linalg.norm(directions, axis=-1, keepdims=True)
    dx_norm = np.linalg.norm(dx - directions, axis=-1)
    dy_norm = np.linalg.norm(dy - directions, axis=-1)
    radii = (f(0.5) * (dx_norm + dy_norm))[..., None] * f(2) / np.sqrt(f(12))
    return origins.astype(f), directions.astype(f), viewdirs.astype(f), radii.astype(f), imageplane.astype(f)
