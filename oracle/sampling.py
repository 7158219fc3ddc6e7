"""Oracle restatement of internal/sampling.py ProposalVolumeSampler (TEST INFRASTRUCTURE ONLY).

Config: configs/ngp_yobo.gin:178-242 and configs/nerf_ngp_yobo.gin:521-562:
sampling_strategy ((0,0,64),(1,1,64),(2,2,32)); anneal_slope 10, anneal_end 1,
anneal_clip 0.4; resample_padding 1e-5; dilation 0 (dead code under these
configs); single_jitter True; stop_level_grad True; ray_shape 'cone';
raydist_fn power_ladder(p=-1.5, premult=2) used for secondary rays only
(NeRFModel.use_raydist_for_secondary_only=True, nerf_ngp_yobo.gin:94) or for all
rays in the transient model.
"""
import numpy as np
import torch

from . import coord, geometry, ref_math, render, stepfun

GRID_PARAMS = (
    dict(hash_map_size=524288, max_grid_size=512, num_features=1),
    dict(hash_map_size=524288, max_grid_size=1024, num_features=1),
    dict(hash_map_size=524288, max_grid_size=2048, num_features=4),
)
MLP_PARAMS = (
    dict(disable_density_normals=False, enable_pred_normals=False, normals_for_filter_only=True),
    dict(disable_density_normals=False, enable_pred_normals=False, normals_for_filter_only=True),
    dict(disable_density_normals=False, enable_pred_normals=True, normals_for_filter_only=False),
)


class ProposalVolumeSampler:
    def __init__(
        self,
        sampling_strategy=((0, 0, 64), (1, 1, 64), (2, 2, 32)),
        grid_params_per_level=GRID_PARAMS,
        mlp_params_per_level=MLP_PARAMS,
        anneal_slope=10.0,
        anneal_end=1.0,
        anneal_clip=0.4,
        resample_padding=1e-5,
        single_jitter=True,
        warp_c=2.0,
        bbox_scaling=1.0,
        raydist=(-1.5, 2.0),
        opaque_background=False,
        shadow_normal_eps_dot_min=None,
    ):
        self.sampling_strategy = sampling_strategy
        self.mlps = [
            geometry.DensityMLP(grid_params=g, warp_c=warp_c, bbox_scaling=bbox_scaling, **m)
            for g, m in zip(grid_params_per_level, mlp_params_per_level)
        ]
        self.anneal_slope, self.anneal_end, self.anneal_clip = anneal_slope, anneal_end, anneal_clip
        self.resample_padding = resample_padding
        self.single_jitter = single_jitter
        self.raydist = raydist
        self.opaque_background = opaque_background

    def init(self, gen, table_init_range=None, bias_range=0.0):
        return {f"MLP_{i}": m.init(gen, table_init_range, bias_range) for i, m in enumerate(self.mlps)}

    def anneal(self, train_frac):
        """internal/sampling.py:326-336."""
        if self.anneal_slope > 0:
            bias = lambda x, s: (s * x) / ((s - 1) * x + 1)
            return float(np.clip(bias(train_frac / self.anneal_end, self.anneal_slope), 0.0, self.anneal_clip))
        return self.anneal_clip

    def __call__(self, params, rays, u01_per_level, train_frac=1.0, use_raydist_fn=False, normals_all_levels=False,
                 create_graph=False, weights_only=False):
        """internal/sampling.py:142-649.

        rays: dict(origins[R,3], directions[R,3], viewdirs[R,3], radii[R,1], near[R,1], far[R,1]).
        u01_per_level: list of [R,1] uniform draws in [0,1) (one jitter per ray per level).
        """
        near, far = rays["near"], rays["far"]
        if use_raydist_fn:
            _, s_to_t = coord.power_ladder_warps(near, far, *self.raydist)
        else:
            _, s_to_t = coord.construct_ray_warps(None, near, far)
        sdist = torch.cat([torch.zeros_like(near), torch.ones_like(far)], dim=-1)
        resample_weights = torch.ones_like(near)
        anneal = self.anneal(train_frac)
        history = []
        for i_level, (i_mlp, _, num_samples) in enumerate(self.sampling_strategy):
            mlp = self.mlps[i_mlp]
            logits = anneal * ref_math.safe_log(resample_weights + self.resample_padding)  # :340
            sdist = stepfun.sample_intervals(
                u01_per_level[i_level], sdist, logits, num_samples, single_jitter=self.single_jitter,
                domain=(0.0, 1.0),
            )  # :343-350
            sdist = sdist.detach()  # :353-354 stop_level_grad
            tdist = s_to_t(sdist)  # :358
            means, covs = render.cast_rays(
                tdist, rays["origins"], rays["directions"], rays["radii"], "cone", diag=False
            )  # :361-368
            want_normals = (normals_all_levels or not mlp.normals_for_filter_only) and not weights_only
            saved = mlp.disable_density_normals
            if not want_normals:
                # levels 0-1 compute and immediately discard the analytic normals
                # (geometry.py:581-584); XLA dead-code-eliminates them.
                mlp.disable_density_normals = True
            res = mlp(params[f"MLP_{i_mlp}"], means, viewdirs=rays["viewdirs"], origins=rays["origins"],
                      create_graph=create_graph)
            mlp.disable_density_normals = saved
            # rectified normals (:519-526)
            for k in list(res.keys()):
                if k.startswith("normals") and res[k] is not None:
                    p = torch.sum(res[k] * rays["viewdirs"][..., None, :], dim=-1, keepdim=True)
                    res[k + "_rectified"] = res[k] * torch.where(p > 0, -1.0, 1.0)
            weights, alphas, trans = render.compute_alpha_weights(
                res["density"], tdist, rays["directions"], opaque_background=self.opaque_background
            )  # :529-534
            resample_weights = weights
            res.update(points=means, means=means, covs=covs, tdist=tdist, sdist=sdist, weights=weights,
                       alphas=alphas, trans=trans)
            history.append(res)
        return history
