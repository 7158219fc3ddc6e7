"""Shared helpers for the parity tests (synthetic workloads of BASELINE.md section 2)."""
import numpy as np
import torch

SEED = 20200823  # the reference's Config.jax_rng_seed (internal/configs.py:180)


def gen(seed_offset=0):
    return np.random.Generator(np.random.PCG64(SEED + seed_offset))


def f32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def make_rays(g, R, near=2.0, far=6.0, radius=4.0, radii=5e-4):
    """BASELINE.md config 1 primary rays: origins on a sphere, non-unit directions."""
    o = g.normal(size=(R, 3))
    o = radius * o / np.linalg.norm(o, axis=-1, keepdims=True)
    v = -o + 0.3 * g.normal(size=(R, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    d = v * g.uniform(1.0, 1.2, size=(R, 1))
    return dict(
        origins=f32(o), directions=f32(d), viewdirs=f32(v), radii=torch.full((R, 1), radii),
        near=torch.full((R, 1), near), far=torch.full((R, 1), far),
    )


def to_dev(tree, device):
    if isinstance(tree, dict):
        return {k: to_dev(v, device) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(to_dev(v, device) for v in tree)
    if isinstance(tree, torch.Tensor):
        return tree.to(device)
    return tree


def rel_err(a, b, floor=1e-30):
    """max |a-b| / max(|b|) -- scale-relative error used with the 1e-5 fp32 tolerance."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), floor))


def rel_l2(a, b, floor=1e-30):
    """||a-b||_2 / ||b||_2 -- used for bf16 gradients, where a flipped ReLU mask moves single entries."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), floor))


def level_table(shape, salt):
    """The closed-form table of tests/golden/make_reference_vectors.py (kept in step with it)."""
    idx = np.arange(int(np.prod(shape)), dtype=np.uint64)
    h = (idx * np.uint64(2654435761) + np.uint64(salt) * np.uint64(40503)) % np.uint64(1 << 32)
    return ((h.astype(np.float64) / float(1 << 32) - 0.5) * 2e-2).astype(np.float32).reshape(shape)


ENC_CONFIGS = {
    "a": dict(hash_map_size=2 ** 15, num_features=2, scale_supersample=1.0, max_grid_size=256),
    "b": dict(hash_map_size=2 ** 12, num_features=4, scale_supersample=1.0, max_grid_size=128, precondition_scaling=1.0,
              bbox_scaling=((-1.0, -2.0, -3.0), (1.5, 2.0, 2.5))),
}


def dense_params(n_in, n_out, salt):
    """Closed-form Dense kernel [in,out] / bias [out] of tests/golden/make_reference_vectors.py (kept in step with it)."""
    k = level_table((n_in, n_out), salt) * np.float32(100.0 * np.sqrt(6.0 / n_in))
    return k.astype(np.float32), (level_table((n_out,), salt + 50) * np.float32(10.0)).astype(np.float32)


def decode_image(img, img_atoms, atom0, ncols, P):
    """Columns [0, ncols) of the atoms starting at `atom0` of a bf16 tile image [tiles][img_atoms][16 KB]
    (neural_radiance_caching_b200/csrc/tc05.cuh: 128 rows x 64 bf16 per atom, row r at (r/8)*1024 + (r%8)*128 bytes,
    16-byte chunks XOR-swizzled with r%8) as an fp32 [P, ncols] CPU tensor."""
    import torch
    flat = img.detach().cpu().view(torch.int16).reshape(-1)
    p = torch.arange(P).reshape(-1, 1)
    c = torch.arange(ncols).reshape(1, -1)
    tile, r = p // 128, p % 128
    atom = atom0 + c // 64
    chunk = (c % 64) // 8
    byte = (tile * img_atoms + atom) * 16384 + (r // 8) * 1024 + (r % 8) * 128 + ((chunk ^ (r % 8)) << 4) + (c % 8) * 2
    vals = flat[(byte // 2).reshape(-1)].reshape(P, ncols)
    return vals.view(torch.bfloat16).to(torch.float32)


# ---- surface-light-field memory variant: the closed-form parameters of tests/golden/make_reference_vectors_slf.py
SLF_GRID = dict(hash_map_size=2 ** 14, max_grid_size=128, num_features=4)
SLF_TABLE_GAIN = 40.0


def slf_dense_params(n_in, n_out, salt, gain=100.0):
    k = level_table((n_in, n_out), salt) * np.float32(gain * np.sqrt(6.0 / n_in))
    return {"kernel": torch.from_numpy(k.astype(np.float32)),
            "bias": torch.from_numpy((level_table((n_out,), salt + 50) * np.float32(10.0)).astype(np.float32))}


def slf_mem_params(net):
    """Parameters of an oracle SurfaceLightFieldMemMLP `net` as the generator assigned them to the reference's class."""
    p = {}
    for key, enc, salt0 in (("distance_grid", net.grid, 700), ("reflectance_grid", net.reflectance_grid, 720)):
        p[key] = {name: torch.from_numpy(level_table(shape, salt0 + i + 1) * np.float32(SLF_TABLE_GAIN))
                  for i, (name, (_, _, shape)) in enumerate(zip(enc.param_names(), enc.layout))}
    salts = {"layers_0": 740, "layers_1": 741, "distance_layer_0": 750, "distance_layer_1": 751, "distance_layer_2": 752,
             "distance_layer_3": 753, "distance_output_layer": 760, "layer_0": 770, "layer_bottleneck": 771,
             "output_rgba_layer": 780, "output_ambient_rgb_layer": 781}
    for name, fi, fo in net.layer_shapes():
        p[name] = slf_dense_params(fi, fo, salts[name], gain=300.0 if name == "distance_output_layer" else 100.0)
    return p
