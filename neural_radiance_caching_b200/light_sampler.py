"""Host-side mirror of the light sampler: internal/light_sampler.py LightMLP (predict_lighting
:162-214, get_vmfs :135-160) under configs/ngp_yobo.gin:335-352, and train_utils.light_sampling_loss
(internal/train_utils.py:1985-2071) over render_utils.vmf_loss_fn (render_utils.py:1493-1550).

  light_grid (L=8, F=4) -> Dense 64 ReLU -> Dense 64 ReLU -> output layer 128 lobes x 5   (tcgen05 chain, bf16)
  -> nrc_vmf_head_{fwd,bwd}; the loss is nrc_vmf_loss.  `means_random` (a fixed-key jax.random.normal in the
reference) is an input, like every random draw at the C ABI."""
import torch

from . import _lib, coord, grid_utils, mlp_chain, nerf

LIGHT_GRID = dict(hash_map_size=524288, max_grid_size=2048, num_features=4)   # configs/ngp_yobo.gin:346-352
NUM_COMPONENTS = 128                                                         # configs/ngp_yobo.gin:337
VMF_SCALE = 20.0                                                             # configs/ngp_yobo.gin:336


class _VmfHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, means_random, positions, K, scale):
        P = raw.shape[0]
        dev = raw.device
        raw = raw.contiguous()
        mr = means_random.contiguous()
        per_point = 1 if mr.dim() == 3 else 0
        means = torch.empty((P, K, 3), device=dev, dtype=torch.float32)
        kappas = torch.empty((P, K, 1), device=dev, dtype=torch.float32)
        logits = torch.empty((P, K, 1), device=dev, dtype=torch.float32)
        _lib.call("nrc_vmf_head_fwd", _lib.stream_ptr(), _lib.ptr(raw), _lib.ptr(mr), per_point,
                  _lib.ptr(positions.contiguous()), P, K, float(scale), _lib.ptr(means), _lib.ptr(kappas), _lib.ptr(logits))
        ctx.save_for_backward(raw)
        ctx.K, ctx.scale = K, scale
        return means, kappas, logits

    @staticmethod
    def backward(ctx, g_means, g_kappas, g_logits):
        (raw,) = ctx.saved_tensors
        P = raw.shape[0]
        z = lambda g, shape: g.contiguous() if g is not None else torch.zeros(shape, device=raw.device)
        g_raw = torch.empty_like(raw)
        _lib.call("nrc_vmf_head_bwd", _lib.stream_ptr(), _lib.ptr(raw), _lib.ptr(z(g_means, (P, ctx.K, 3))),
                  _lib.ptr(z(g_kappas, (P, ctx.K, 1))), _lib.ptr(z(g_logits, (P, ctx.K, 1))), P, ctx.K, float(ctx.scale),
                  _lib.ptr(g_raw))
        return g_raw, None, None, None, None


class _VmfLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, kappas, logits, normals, dirs, pdf, weight, fvals, lossmult, srgb):
        P, K = means.shape[0], means.shape[1]
        S = dirs.shape[1]
        dev = means.device
        c = lambda t: t.contiguous()
        loss = torch.zeros((), device=dev, dtype=torch.float32)
        g_m, g_k, g_l = torch.empty_like(means), torch.empty_like(kappas), torch.empty_like(logits)
        _lib.call("nrc_vmf_loss", _lib.stream_ptr(), _lib.ptr(c(means)), _lib.ptr(c(kappas)), _lib.ptr(c(logits)),
                  _lib.ptr(c(normals)), _lib.ptr(c(dirs)), _lib.ptr(c(pdf)), _lib.ptr(c(weight)), _lib.ptr(c(fvals)), P, K, S,
                  float(lossmult), int(bool(srgb)), _lib.ptr(loss), _lib.ptr(g_m), _lib.ptr(g_k), _lib.ptr(g_l))
        ctx.save_for_backward(g_m, g_k, g_l)
        return loss

    @staticmethod
    def backward(ctx, g):
        g_m, g_k, g_l = ctx.saved_tensors
        return (g_m * g, g_k * g, g_l * g) + (None,) * 7


def vmf_loss_fn(vmf_vars, sample_normals, sample_dirs, pdf, weight, function_vals, lossmult, srgb=True):
    """render_utils.vmf_loss_fn with a constant per-sample lossmult (1 / S at the call site)."""
    return _VmfLossFn.apply(vmf_vars[0], vmf_vars[1], vmf_vars[2], sample_normals, sample_dirs, pdf, weight,
                            function_vals, lossmult, srgb)


def light_sampling_loss(vmfs, sample_dirs, pdf, weight, radiance_in, srgb=True):
    """train_utils.light_sampling_loss for one suffix present (multiplier 2 and the / 2 inside the loop cancel)."""
    fv = torch.linalg.norm(radiance_in.detach(), dim=-1)
    S = fv.shape[-1]
    K = vmfs["vmf_means"].shape[-2]
    v = (vmfs["vmf_means"].reshape(-1, K, 3), vmfs["vmf_kappas"].reshape(-1, K, 1), vmfs["vmf_logits"].reshape(-1, K, 1))
    return vmf_loss_fn(v, vmfs["vmf_normals"].reshape(-1, 3), sample_dirs.detach().reshape(-1, S, 3),
                       pdf.detach().reshape(-1, S), weight.detach().reshape(-1, S), fv.reshape(-1, S), 1.0 / S, srgb)


class LightMLP:
    """Render path: the whole stack is one chain program, the output layer split into 128-column head groups (the
    kernel's accumulator groups are at most 128 columns wide).  Training: per-layer GEMMs."""

    def __init__(self, warp_c=2.0, bbox_scaling=1.0, num_components=NUM_COMPONENTS, vmf_scale=VMF_SCALE, bf16=True):
        self.grid = grid_utils.HashEncoding(bbox_scaling=bbox_scaling, scale_supersample=1.0, **LIGHT_GRID)
        self.warp_c = warp_c if warp_c is not None else 0.0
        self.num_components, self.vmf_scale, self.bf16 = num_components, vmf_scale, bf16
        self.out_dim = num_components * 5
        self.splits = [(c, min(128, self.out_dim - c)) for c in range(0, self.out_dim, 128)]
        self.chain = mlp_chain.ChainSpec(in_widths=[self.grid.num_outputs],
                                         hidden=[("layers_0", 64, False), ("layers_1", 64, False)],
                                         heads=[[(f"output_layer_{i}", w)] for i, (_, w) in enumerate(self.splits)])
        self._pack_cache = mlp_chain.PackCache()

    def init(self, device, generator=None, table_init_range=0.1):
        """Random-init parameters in the reference's layout (he_uniform kernels, zero biases)."""
        def layer(fi, fo):
            lim = float((6.0 / fi) ** 0.5)
            return {"kernel": torch.empty((fi, fo), device=device).uniform_(-lim, lim, generator=generator),
                    "bias": torch.zeros((fo,), device=device)}
        _, arena = self.grid.init(device, generator=generator, init_range=table_init_range)
        return {"light_grid": dict(self.grid.views(arena), _arena=arena), "layers_0": layer(self.grid.num_outputs, 64),
                "layers_1": layer(64, 64), "output_layer": layer(64, self.out_dim)}

    def from_oracle(self, p, device):
        names = [n for (n, _, _, _) in self.grid.level_layout]
        arena = torch.cat([p["light_grid"][n].detach().reshape(-1) for n in names]).to(device)
        out = {k: {kk: vv.detach().to(device).contiguous() for kk, vv in v.items()} for k, v in p.items() if k != "light_grid"}
        out["light_grid"] = dict(self.grid.views(arena), _arena=arena)
        return out

    def _split_params(self, p):
        q = {"layers_0": p["layers_0"], "layers_1": p["layers_1"]}
        for i, (c, w) in enumerate(self.splits):
            q[f"output_layer_{i}"] = {"kernel": p["output_layer"]["kernel"][:, c:c + w].contiguous(),
                                      "bias": p["output_layer"]["bias"][c:c + w].contiguous()}
        return q

    def predict_lighting(self, p, means, means_random, normals=None, weights=None):
        """-> dict(vmf_means [...,K,3] (relative to the position), vmf_kappas / vmf_logits [...,K,1], vmf_origins,
        vmf_normals, weights) for shaded points `means` [...,3]."""
        lead = means.shape[:-1]
        m2 = means.reshape(-1, 3)
        z = coord._ContractFn.apply(m2, self.warp_c)
        enc = self.grid(p["light_grid"], z).reshape(-1, self.grid.num_outputs)
        if self.bf16 and not torch.is_grad_enabled():
            # render path: one tcgen05 chain program (the 640-wide output layer as five 128-column head groups)
            outs = mlp_chain.forward_cached(self.chain, self._split_params(p), [enc], self._pack_cache)
            raw = torch.cat(list(outs), dim=-1)
        else:
            # training: per-layer GEMMs (bf16 mma.sync or fp32): the data-gradient chain program keeps every head
            # group's dY tile resident and five groups do not fit its shared-memory slots
            x = nerf.dense(p["layers_0"], enc, relu=True, bf16=self.bf16)
            x = nerf.dense(p["layers_1"], x, relu=True, bf16=self.bf16)
            raw = nerf.dense(p["output_layer"], x, bf16=self.bf16)
        K = self.num_components
        vm, vk, vl = _VmfHeadFn.apply(raw, means_random, m2.detach(), K, self.vmf_scale)
        out = dict(vmf_means=vm.reshape(lead + (K, 3)), vmf_kappas=vk.reshape(lead + (K, 1)),
                   vmf_logits=vl.reshape(lead + (K, 1)), vmf_origins=means.detach()[..., None, :])
        if normals is not None:
            out["vmf_normals"] = normals.detach()[..., None, :]
        if weights is not None:
            out["weights"] = weights.detach()[..., None, None]
        return out
