"""Golden vectors for the SURFACE-LIGHT-FIELD MEMORY variant (second half) from the REFERENCE'S OWN CLASS:
internal/surface_light_field.py SurfaceLightFieldMLP as models.py:813-833 builds `surface_lf_mem` (use_env_alpha=True,
distance_near / distance_far from the model) under configs/nerf_ngp_yobo.gin:97-165 + ngp_yobo.gin:232-236
(power-ladder ray distances), called the way models.get_slf_results does (models.py:849-876: means = origins, one
direction per ray, roughness 0, no shader bottleneck).  Executed under tests/golden/jax_numpy_shim.py like
make_reference_vectors.py (same closed-form parameters); run from the repo root:

    python tests/golden/make_reference_vectors_slf.py      # writes tests/golden/reference_slf.npz

The hash tables are scaled down (hash_map_size 2**14, max_grid_size 128) so that the tables need not be stored: they are
the closed form `level_table` of the entry index, repeated in tests/util.py."""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jax_numpy_shim as shim  # noqa: E402
import make_reference_vectors as mrv  # noqa: E402

GRID = dict(hash_map_size=2 ** 14, max_grid_size=128, num_features=4)
TABLE_GAIN = 40.0


def level_table(shape, salt):
    idx = np.arange(int(np.prod(shape)), dtype=np.uint64)
    h = (idx * np.uint64(2654435761) + np.uint64(salt) * np.uint64(40503)) % np.uint64(1 << 32)
    return ((h.astype(np.float64) / float(1 << 32) - 0.5) * 2e-2).astype(np.float32).reshape(shape)


def dense_params(n_in, n_out, salt, gain=100.0):
    k = level_table((n_in, n_out), salt) * np.float32(gain * np.sqrt(6.0 / n_in))
    return k.astype(np.float32), (level_table((n_out,), salt + 50) * np.float32(10.0)).astype(np.float32)


def main():
    R = mrv.load_reference()
    rmath, rcoord, rgrid = R["math"], R["coord"], R["grid_utils"]
    rslf = importlib.import_module("internal.surface_light_field")
    g = np.random.Generator(np.random.PCG64(20240611))
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    unit = lambda v: v / np.linalg.norm(v, axis=-1, keepdims=True)
    sp_, sg_, relu_ = shim.nn_mod.softplus, shim.nn_mod.sigmoid, shim.nn_mod.relu
    cfg = types.SimpleNamespace(num_rgb_channels=3, n_bins=1, multi_illumination=False, rotate_illumination=False,
                                num_illuminations=1, multiple_illumination_outputs=False, env_map_distance=2.0)
    out = {}

    class _Enc64(rgrid.HashEncoding):
        # NumPy 2 keeps int32 ** int in int32 (see make_reference_vectors.py): same values, wider type
        grid_sizes = property(lambda self: rgrid.HashEncoding.grid_sizes.fget(self).astype(np.int64))

    def make_grid(salt0, **kw):
        # entries in +-0.4 (TABLE_GAIN x level_table): features large enough for every output to vary across rays
        enc = _Enc64(**kw)
        enc.seen = []
        enc.param = lambda name, init_fn: (enc.seen.append(name),
                                           level_table(init_fn.keywords["shape"], salt0 + len(enc.seen)) * np.float32(TABLE_GAIN))[1]
        return enc

    for tag, nsamp, far_kw in (("slfm", 8, {}), ("slfm1", 1, dict(near=np.float32(0.07), far=np.float32(0.13)))):
        m = rslf.SurfaceLightFieldMLP(
            config=cfg, use_env_alpha=True, distance_near=5e-2, distance_far=2.0,                     # models.py:813-833
            net_depth=2, net_width=64, skip_layer=2, bottleneck_width=128, use_directional_enc=False, use_ide=False, deg_view=2,
            net_depth_viewdirs=2, net_width_viewdirs=64, skip_layer_dir=2, use_distance_prediction=True, use_distance_ide=False,
            deg_view_distance=2, net_depth_distance=4, net_width_distance=128, skip_layer_distance=2, use_origins=False,
            deg_origins=2, num_distance_samples=nsamp, distance_scale=1.0, use_voxel_grid=False, use_point_offsets=False,
            use_far_field_points=False, use_points=False, use_reflectance_grid=True, reflectance_grid_representation="ngp",
            reflectance_grid_params=dict(GRID, bbox_scaling=2.0), per_ref_feature_output=False, use_grid=True,
            grid_representation="ngp", grid_params=dict(GRID), use_roughness=False, use_bottleneck=False,
            use_density_feature=False, use_shader_bottleneck=False, use_lights=False,
            warp_fn=rcoord.contract_radius_2, ref_warp_fn=rcoord.contract_radius_2,
            raydist_fn=(rmath.power_ladder, rmath.inv_power_ladder, dict(p=np.float32(-1.5), premult=np.float32(2.0))),
            net_activation=relu_, rgb_activation=sp_, rgb_bias=-2.0, ambient_rgb_activation=sp_, ambient_rgb_bias=-1.0,
            alpha_activation=sg_)
        m.setup()
        m.grid = make_grid(700, **GRID)
        m.reflectance_grid = make_grid(720, **dict(GRID, bbox_scaling=2.0))
        nf = len(m.grid.grid_sizes) * GRID["num_features"]
        d_in = nf
        for i_, layer in enumerate(m.layers):                                         # BaseShader.run_network (shading.py:116-130)
            layer.kernel, layer.bias = dense_params(d_in, layer.features, 740 + i_)
            d_in = layer.features
        n_in = 64 + 15 + 15
        d_in = n_in
        for i_, layer in enumerate(m.distance_layers):
            layer.kernel, layer.bias = dense_params(d_in, layer.features, 750 + i_)
            d_in = layer.features + (n_in if (i_ % 2 == 0 and i_ > 0) else 0)
        # the reference initialises this layer with zeros (distance_output_layer, zeros_layer); trained values are not zero
        m.output_distance_layer.kernel, m.output_distance_layer.bias = dense_params(d_in, 8 * nsamp + 4, 760, gain=300.0)
        d_in = nf
        for i_, layer in enumerate(m.view_dependent_layers):
            layer.kernel, layer.bias = dense_params(d_in, layer.features, 770 + i_)
            d_in = layer.features
        m.output_rgba_layer.kernel, m.output_rgba_layer.bias = dense_params(d_in, 4, 780)
        m.output_ambient_rgb_layer.kernel, m.output_ambient_rgb_layer.bias = dense_params(d_in, 3, 781)

        Rq = 96
        origins = f(g.normal(size=(Rq, 3)) * 1.1); origins[:3] = [[0, 0, 0], [3.0, -2.5, 0.4], [0.2, 0.1, -0.3]]
        dirs = f(unit(g.normal(size=(Rq, 3))))
        rays = types.SimpleNamespace(origins=origins, viewdirs=dirs, near=f(np.full((Rq, 1), 0.05)), far=f(np.full((Rq, 1), 2.0)),
                                     light_idx=None, lights=None)
        o1 = origins[:, None, :].view(shim.F32Array)
        res = m(None, rays, dict(means=o1, covs=np.ones_like(origins[:, None, :])), origins[:, None, :], dirs[:, None, :],
                roughness=np.zeros_like(origins[:, None, :1]), shader_bottleneck=None, train=False, **far_kw)
        out.update({tag + "_origins": origins, tag + "_viewdirs": dirs})
        out[tag + "_grid_names"] = np.array(m.grid.seen)
        out[tag + "_ref_grid_names"] = np.array(m.reflectance_grid.seen)
        for k_ in ("incoming_rgb", "incoming_ambient_rgb", "incoming_alpha", "incoming_weights", "incoming_s_dist",
                   "incoming_dist", "incoming_env_rgba", "incoming_acc"):
            out[tag + "_" + k_] = np.asarray(res[k_])
        # the intermediate the points kernel is held to: predict_points on the same inputs
        m.grid.seen = []                                   # the table salts count parameter requests from 1 again
        bott = m.predict_appearance_feature(dict(means=o1, covs=None), train=False, control_offsets=f(np.zeros((1, 3))),
                                            perp_mag=None) * np.ones_like(dirs[:, None, :1])
        pts, raw_w, mask, s_d, dist, env_rgb, env_a = m.predict_points(None, rays, origins[:, None, :] * np.ones_like(dirs[:, None, :]),
                                                                       dirs[:, None, :], bott, np.zeros_like(origins[:, None, :1]),
                                                                       **far_kw)
        raw = m.run_distances_network(bott, origins[:, None, :], dirs[:, None, :], None)
        out.update({tag + "_bottleneck": bott, tag + "_dist_net_outputs": raw, tag + "_points": pts, tag + "_raw_weights": raw_w,
                    tag + "_ref_mask": mask, tag + "_s_distances": s_d, tag + "_distances": dist})

    # ---- material._integrate_slf_variate (material.py:2433-2513): the reference's own method on a stand-in `self` whose
    #      get_outgoing_radiance returns the cache integral on the first call and the light-field integral on the second
    #      (it must be handed the first call's outputs as last_integrated_outputs: the SAME secondary rays) ----------------
    keys = ("radiance_out", "diffuse_radiance_out", "specular_radiance_out", "irradiance", "indirect_occ")
    cache_o = {k_: f(g.uniform(size=(12, 3))) for k_ in keys}
    slf_o = {k_: f(g.uniform(size=(12, 3))) for k_ in keys if k_ != "indirect_occ"}
    slf_o["only_slf"] = f(g.uniform(size=(12, 1)))
    calls = []

    def fake_get_outgoing_radiance(**kw):
        calls.append(kw)
        return dict(cache_o) if len(calls) == 1 else dict(slf_o)

    me = types.SimpleNamespace(get_outgoing_radiance=fake_get_outgoing_radiance, get_num_secondary_samples_diff=lambda train: 4)
    merged = R["material"].BaseMaterialMLP._integrate_slf_variate(
        me, np.zeros(2, np.uint32), None, None, None, None, "cache_fn", "slf_fn", None, None, 1.0, False)
    assert calls[0]["radiance_cache_fn"] == "cache_fn" and calls[1]["radiance_cache_fn"] == "slf_fn"
    assert calls[1]["last_integrated_outputs"]["radiance_out"] is cache_o["radiance_out"]
    for k_, v_ in cache_o.items():
        out["variate_cache_" + k_] = v_
    for k_, v_ in slf_o.items():
        out["variate_slf_" + k_] = v_
    out["variate_keys"] = np.array(sorted(merged.keys()))
    for k_, v_ in merged.items():
        if v_ is not None:
            out["variate_out_" + k_] = np.asarray(v_)

    out = {k: np.asarray(v_) for k, v_ in out.items()}
    out = {k: (v_.astype(np.float32) if v_.dtype == np.float64 else v_) for k, v_ in out.items()}
    path = os.path.join(HERE, "reference_slf.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KB")
    for k, v_ in sorted(out.items()):
        print(f"  {k:34s} {str(v_.dtype):8s} {v_.shape}")


if __name__ == "__main__":
    main()
