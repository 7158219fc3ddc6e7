"""Oracle restatement (TEST INFRASTRUCTURE ONLY - never imported by the product path) of the surface-light-field
MEMORY variant: internal/surface_light_field.py BaseSurfaceLightFieldMLP with use_distance_prediction +
use_reflectance_grid as models.py:813-833 builds `surface_lf_mem` under configs/nerf_ngp_yobo.gin:97-165 and
ngp_yobo.gin:232-236, called like models.get_slf_results (models.py:849-908).

Pinned: tests/test_reference_vectors.py::test_slf_memory holds every output and the predict_points intermediates to
tests/golden/reference_slf.npz, which tests/golden/make_reference_vectors_slf.py produced by executing the reference's
own class."""
import numpy as np
import torch

from . import coord as ocoord
from . import geometry as ogeo
from . import grid_utils as ogrid

DISTANCE_GRID = dict(hash_map_size=524288, max_grid_size=256, num_features=4)          # nerf_ngp_yobo.gin:156-161
REFLECTANCE_GRID = dict(hash_map_size=524288, max_grid_size=256, num_features=4, bbox_scaling=2.0)   # :146-152


def predict_points(raw, origins, refdirs, num_samples, distance_near, distance_far, near=0.0, far=float("inf"),
                   distance_scale=1.0, distance_bias=-2.0, rgb_premultiplier=1.0, rgb_bias=-2.0, alpha_bias=2.0,
                   raydist=(-1.5, 2.0), warp_c=2.0):
    """surface_light_field.py:594-780 for use_voxel_grid=False, num_far_samples=0, use_sorted_distances=False,
    use_point_offsets=False, use_env_alpha=True, followed by the head of __call__ that turns the raw weights into
    `ref_weights` and the weighted s-distance (:899-913).

    raw [P, 8*n+4] distance-network outputs, origins / refdirs [P,3] ->
    points (already through ref_warp_fn) [P,n,3], ref_weights [P,n], s_dist [P,1], distances [P,n], env_rgba [P,4]."""
    P, n = raw.shape[0], num_samples
    env_rgb = torch.nn.functional.softplus(rgb_premultiplier * raw[..., -4:-1] + rgb_bias)            # :626-629
    env_alpha = torch.sigmoid(raw[..., -1:] + alpha_bias)                                               # :630-633
    o = raw[..., :-4].reshape(P, n, -1)                                                                 # :636-642
    distance_offsets, distance_sigma, raw_weights = o[..., 0], o[..., 1], o[..., 4]
    distance_offsets = distance_offsets * distance_scale / n * torch.sigmoid(distance_sigma + distance_bias)   # :652-657
    start = torch.linspace(1e-8, 1.0 - 1e-8, n, dtype=torch.float32).reshape(1, n)                     # :715-718
    s = distance_offsets + start
    s_floor = torch.floor(s).to(torch.int32)
    s_frac = s - s_floor.to(torch.float32)
    s = torch.where((s_floor % 2) == 0, s_frac, 1.0 - s_frac)                                           # :721-728
    tn = torch.full((P, 1), distance_near, dtype=torch.float32)
    tf = torch.full((P, 1), distance_far, dtype=torch.float32)
    if raydist is None:
        _, s_to_t = ocoord.construct_ray_warps(None, tn, tf)
    else:
        _, s_to_t = ocoord.power_ladder_warps(tn, tf, p=raydist[0], premult=raydist[1])                 # :532-545
    distances = s_to_t(s)                                                                               # :731
    mask = ((distances > distance_near).float() * (distances < distance_far).float()
            * (distances > near).float() * (distances < far).float())                                   # :747-750
    distances = torch.clamp(distances, distance_near, distance_far)                                     # :753
    points = origins[:, None, :] + distances[..., None] * refdirs[:, None, :]                           # :771
    points_raw = points
    points = ocoord.contract_radius(points, warp_c)                                                     # :905 ref_warp_fn
    w = torch.softmax(raw_weights, dim=-1)                                                              # :908
    s_dist = (s * w).sum(dim=-1, keepdim=True)                                                          # :909
    w = w * mask * env_alpha                                                                            # :910
    return dict(points=points, ref_weights=w, s_dist=s_dist, distances=distances, env_rgba=torch.cat([env_rgb, env_alpha], -1),
                raw_weights=raw_weights, ref_mask=mask, s_distances=s, points_raw=points_raw)


class SurfaceLightFieldMemMLP:
    """`surface_lf_mem`: distance_grid(contract(origins)) -> 2 x Dense 64 ReLU (BaseShader.run_network) = bottleneck;
    [bottleneck, pos_enc(contract(origins), 0, 2), pos_enc(refdirs, 0, 2)] -> 4 x Dense 128 ReLU (input re-concatenated
    after layer 2) -> distance_output_layer (8 n + 4) -> predict_points -> reflectance_grid at the n points -> weighted
    feature sum -> Dense 64 ReLU -> Dense 128 ReLU -> output_rgba_layer (4), output_ambient_rgb_layer (3)."""

    def __init__(self, num_distance_samples=8, distance_near=5e-2, distance_far=2.0, grid=None, reflectance_grid=None,
                 warp_c=2.0, raydist=(-1.5, 2.0), rgb_bias=-2.0, ambient_rgb_bias=-1.0, alpha_bias=2.0, dense=None):
        self.n = num_distance_samples
        self.distance_near, self.distance_far = distance_near, distance_far
        self.grid = ogrid.HashEncoding(**(grid or DISTANCE_GRID))            # scale_supersample: the class default (2.0)
        self.reflectance_grid = ogrid.HashEncoding(**(reflectance_grid or REFLECTANCE_GRID))
        self.warp_c, self.raydist = warp_c, raydist
        self.rgb_bias, self.ambient_rgb_bias, self.alpha_bias = rgb_bias, ambient_rgb_bias, alpha_bias
        self.nf = len(self.grid.grid_sizes) * self.grid.num_features
        self.nrf = len(self.reflectance_grid.grid_sizes) * self.reflectance_grid.num_features
        self.dist_in = 64 + 15 + 15
        self.dense = dense or ogeo.dense

    def layer_shapes(self):
        """(name, fan_in, fan_out) of every Dense in the reference's module tree."""
        s = [("layers_0", self.nf, 64), ("layers_1", 64, 64)]
        d = self.dist_in
        for i in range(4):
            s.append((f"distance_layer_{i}", d, 128))
            d = 128 + (self.dist_in if (i % 2 == 0 and i > 0) else 0)
        s += [("distance_output_layer", d, 8 * self.n + 4), ("layer_0", self.nrf, 64), ("layer_bottleneck", 64, 128),
              ("output_rgba_layer", 128, 4), ("output_ambient_rgb_layer", 128, 3)]
        return s

    def init(self, gen, table_init_range=0.1):
        p = {"distance_grid": self.grid.init(gen, init_range=table_init_range),
             "reflectance_grid": self.reflectance_grid.init(gen, init_range=table_init_range)}
        for name, fi, fo in self.layer_shapes():
            p[name] = {"kernel": ogeo.he_uniform(gen, fi, fo), "bias": torch.from_numpy(gen.normal(size=fo).astype(np.float32) * 0.1)}
        return p

    def bottleneck(self, p, origins):
        x = self.grid(p["distance_grid"], ocoord.contract_radius(origins, self.warp_c))     # shading.py:162-199 (one control point)
        for name in ("layers_0", "layers_1"):                                               # shading.py:116-130
            x = torch.relu(self.dense(p[name], x))
        return x

    def run_distances_network(self, p, bottleneck, origins, refdirs):
        """surface_light_field.py:414-442."""
        enc = torch.cat([bottleneck, ocoord.pos_enc(ocoord.contract_radius(origins, self.warp_c), 0, 2, True),
                         ocoord.pos_enc(refdirs, 0, 2, True)], dim=-1)
        x = enc
        for i in range(4):
            x = torch.relu(self.dense(p[f"distance_layer_{i}"], x))
            if i % 2 == 0 and i > 0:
                x = torch.cat([x, enc], dim=-1)
        return self.dense(p["distance_output_layer"], x)

    def __call__(self, p, origins, refdirs, near=0.0, far=float("inf")):
        """origins / refdirs [P,3] -> the reference's `incoming_*` dict (surface_light_field.py:782-1069)."""
        bott = self.bottleneck(p, origins)
        raw = self.run_distances_network(p, bott, origins, refdirs)
        pp = predict_points(raw, origins, refdirs, self.n, self.distance_near, self.distance_far, near, far,
                            rgb_bias=self.rgb_bias, alpha_bias=self.alpha_bias, raydist=self.raydist, warp_c=self.warp_c)
        feat = self.reflectance_grid(p["reflectance_grid"], pp["points"])                   # [P,n,nrf] (:925-931, per_level_fn = id)
        x = (feat * pp["ref_weights"][..., None]).sum(dim=-2)                               # :981
        x = torch.relu(self.dense(p["layer_0"], x))                                         # :1037 (no skip: depth 2)
        x = torch.relu(self.dense(p["layer_bottleneck"], x))
        rgba = self.dense(p["output_rgba_layer"], x)
        rgb = torch.nn.functional.softplus(rgba[..., :-1] + self.rgb_bias)                  # :1045-1047
        alpha = torch.clamp(torch.sigmoid(rgba[..., -1:] + self.alpha_bias), 0.0, 1.0)      # :1048-1051
        amb = torch.nn.functional.softplus(self.dense(p["output_ambient_rgb_layer"], x) + self.ambient_rgb_bias)
        return dict(incoming_rgb=torch.clamp(rgb, min=0.0), incoming_ambient_rgb=torch.clamp(amb, min=0.0), incoming_alpha=alpha,
                    incoming_weights=pp["ref_weights"], incoming_s_dist=pp["s_dist"], incoming_dist=pp["distances"],
                    incoming_env_rgba=pp["env_rgba"], incoming_acc=pp["ref_weights"].sum(dim=-1),
                    bottleneck=bott, dist_net_outputs=raw, points=pp["points"])


def slf_results(res):
    """models.get_slf_results (models.py:881-908) with use_env_map=False in the call (material.py:2246-2257): no environment
    composite, `rgb` / `acc` are the light field's own; surface_lf_fn then clamps at zero (material.py:2273)."""
    return dict(rgb=torch.clamp(res["incoming_rgb"], min=0.0), acc=res["incoming_acc"])
