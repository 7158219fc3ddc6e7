"""Host-side mirror of the cache shader: internal/nerf.py NeRFMLP (passive path),
internal/shading.py BaseShader.predict_appearance_feature, the IDE-form
internal/surface_light_field.py SurfaceLightFieldMLP (`SurfaceLightField`, shader-level `EnvMap`)
and internal/ref_utils.py generate_ide_fn / reflect.  Dense layers and the IDE run in the CUDA
library (nrc_dense_*, nrc_ide_*); the appearance grid uses the encode kernels.  The remaining
per-point scalar glue (softplus / sigmoid / clip / products of [P,1..3] tensors) is torch
elementwise code for now.
"""
import ctypes as C
import math as pymath

import numpy as np
import torch

from . import _lib, coord, grid_utils, mlp_chain


# ----------------------------------------------------------------------------- Dense
def _pad4(t):
    """Row stride multiple of 4 floats (16-byte rows) so the GEMM tile loader can use LDG.128."""
    k = t.shape[-1]
    if k % 4 == 0:
        return t.contiguous(), k
    k4 = (k + 3) // 4 * 4
    return torch.nn.functional.pad(t, (0, k4 - k)).contiguous(), k4


class _DenseFn(torch.autograd.Function):
    """custom_vjp analogue of flax.linen.Dense (+ optional ReLU) over nrc_dense_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, x, kernel, bias, relu, bf16):
        K, N = kernel.shape
        x2, ldx = _pad4(x.reshape(-1, K))
        M = x2.shape[0]
        y = torch.empty((M, N), device=x.device, dtype=torch.float32)
        _lib.call("nrc_dense_fwd", _lib.stream_ptr(), _lib.ptr(x2), ldx, _lib.ptr(kernel), _lib.ptr(bias), M, K, N,
                  int(relu), int(bf16), _lib.ptr(y), N)
        ctx.save_for_backward(x2, kernel, y if relu else None)
        ctx.bias_ref = bias
        ctx.meta = (x.shape, relu, bf16, ldx)
        return y.reshape(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, g):
        x2, kernel, y = ctx.saved_tensors
        xshape, relu, bf16, ldx = ctx.meta
        K, N = kernel.shape
        M = x2.shape[0]
        g2 = g.reshape(M, N).contiguous()
        if relu:
            gpre = torch.empty_like(g2)
            _lib.call("nrc_relu_bwd", _lib.stream_ptr(), _lib.ptr(y), N, _lib.ptr(g2), N, M, N, _lib.ptr(gpre), N)
            g2 = gpre
        g2, ldg = _pad4(g2)
        gx = torch.empty((M, K), device=g.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        want_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gk = gb = None
        sunk = False
        if want_w:
            gk, gb = _lib.grad_sink(kernel), _lib.grad_sink(ctx.bias_ref)
            sunk = gk is not None and gb is not None
            if not sunk:
                gk = torch.zeros_like(kernel)
                gb = torch.zeros((N,), device=g.device, dtype=torch.float32)
        _lib.call("nrc_dense_bwd", _lib.stream_ptr(), _lib.ptr(x2), ldx, _lib.ptr(kernel), _lib.ptr(g2), ldg, M, K, N,
                  int(bf16), _lib.ptr(gx), K, 0, _lib.ptr(gk), _lib.ptr(gb))
        if sunk:
            gk = gb = None   # accumulated into the registered sinks
        return (gx.reshape(xshape) if gx is not None else None), gk, gb, None, None


def dense(p, x, relu=False, bf16=False):
    """flax.linen.Dense with params {'kernel': [in,out], 'bias': [out]}."""
    return _DenseFn.apply(x, p["kernel"], p["bias"], relu, bf16)


# ----------------------------------------------------------------------------- IDE
def _generalized_binomial_coeff(a, k):
    return np.prod(a - np.arange(k)) / pymath.factorial(k)


def _assoc_legendre_coeff(l, m, k):
    return ((-1) ** m * 2**l * pymath.factorial(l) / pymath.factorial(k) / pymath.factorial(l - k - m)
            * _generalized_binomial_coeff(0.5 * (l + k + m - 1.0), l))


def _sph_harm_coeff(l, m, k):
    return np.sqrt((2.0 * l + 1.0) * pymath.factorial(l - m) / (4.0 * np.pi * pymath.factorial(l + m))
                   ) * _assoc_legendre_coeff(l, m, k)


class _IdeTables:
    """Host tables of ref_utils.generate_ide_fn (internal/ref_utils.py:117-158,181)."""

    _cache = {}

    def __init__(self, deg_view):
        if deg_view > 5:
            raise ValueError("Only deg_view of at most 5 is numerically stable.")
        ml = []
        for i in range(deg_view):
            l = 2**i
            for m in range(l + 1):
                ml.append((m, l))
        self.n_sh = len(ml)
        l_max = 2 ** (deg_view - 1)
        mat = np.zeros((l_max + 1, self.n_sh))
        for i, (m, l) in enumerate(ml):
            for k in range(l - m + 1):
                mat[k, i] = _sph_harm_coeff(l, m, k)
        self.mat_host = mat.astype(np.float32)
        self.m = (C.c_int32 * self.n_sh)(*[m for m, _ in ml])
        self.l = (C.c_int32 * self.n_sh)(*[l for _, l in ml])
        self.sigma = (C.c_float * self.n_sh)(*[0.5 * l * (l + 1) for _, l in ml])
        self._mat_dev = {}

    def mat(self, device):
        key = str(device)
        if key not in self._mat_dev:
            self._mat_dev[key] = torch.from_numpy(self.mat_host).to(device).contiguous()
        return self._mat_dev[key]

    @classmethod
    def get(cls, deg_view):
        if deg_view not in cls._cache:
            cls._cache[deg_view] = cls(deg_view)
        return cls._cache[deg_view]


class _IdeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, kappa_inv, deg_view):
        t = _IdeTables.get(deg_view)
        x2 = xyz.reshape(-1, 3).contiguous()
        k2 = kappa_inv.reshape(-1).contiguous()
        P = x2.shape[0]
        out = torch.empty((P, 2 * t.n_sh), device=xyz.device, dtype=torch.float32)
        _lib.call("nrc_ide_fwd", _lib.stream_ptr(), t.n_sh, t.m, t.l, t.sigma, _lib.ptr(t.mat(xyz.device)),
                  _lib.ptr(x2), _lib.ptr(k2), P, _lib.ptr(out), 2 * t.n_sh)
        ctx.save_for_backward(x2, k2)
        ctx.meta = (deg_view, xyz.shape, kappa_inv.shape)
        return out.reshape(xyz.shape[:-1] + (2 * t.n_sh,))

    @staticmethod
    def backward(ctx, g):
        x2, k2 = ctx.saved_tensors
        deg_view, xshape, kshape = ctx.meta
        t = _IdeTables.get(deg_view)
        P = x2.shape[0]
        g2 = g.reshape(P, 2 * t.n_sh).contiguous()
        gx = torch.empty_like(x2)
        gk = torch.empty_like(k2)
        _lib.call("nrc_ide_bwd", _lib.stream_ptr(), t.n_sh, t.m, t.l, t.sigma, _lib.ptr(t.mat(g.device)),
                  _lib.ptr(x2), _lib.ptr(k2), _lib.ptr(g2), 2 * t.n_sh, P, _lib.ptr(gx), _lib.ptr(gk))
        return gx.reshape(xshape), gk.reshape(kshape), None


def generate_ide_fn(deg_view):
    """ref_utils.generate_ide_fn (internal/ref_utils.py:131-192): returns f(xyz, kappa_inv)."""
    _IdeTables.get(deg_view)
    return lambda xyz, kappa_inv: _IdeFn.apply(xyz, kappa_inv, deg_view)


def reflect(viewdirs, normals):
    """ref_utils.reflect (internal/ref_utils.py:25-42)."""
    return 2.0 * torch.sum(normals * viewdirs, dim=-1, keepdim=True) * normals - viewdirs


# ----------------------------------------------------------------------------- SLF (IDE form)
class SurfaceLightFieldMLP:
    """IDE-form SurfaceLightFieldMLP (internal/surface_light_field.py:782-1069) as configured by
    configs/nerf_ngp_yobo.gin:232-251 (`SurfaceLightField`) and :299-343 (shader `EnvMap`)."""

    def __init__(self, deg_view, use_shader_bottleneck, bottleneck_width=128, net_width_viewdirs=128,
                 net_depth_viewdirs=4, skip_layer_dir=2, ambient_rgb_bias=-1.0, ambient_rgb_max=float("inf"),
                 bf16=False):
        self.deg_view = deg_view
        self.use_shader_bottleneck = use_shader_bottleneck
        self.width, self.depth, self.skip = net_width_viewdirs, net_depth_viewdirs, skip_layer_dir
        self.ambient_rgb_bias, self.ambient_rgb_max = ambient_rgb_bias, ambient_rgb_max
        self.dir_enc_fn = generate_ide_fn(deg_view)
        self.bf16 = bf16
        # tcgen05 chain (bf16 variant): the whole Dense stack is one fused program per 128-point tile
        ide_dim = 2 * _IdeTables.get(deg_view).n_sh
        names = self.layer_names()
        self.chain = mlp_chain.ChainSpec(
            in_widths=([bottleneck_width] if use_shader_bottleneck else []) + [ide_dim],
            hidden=[(n, net_width_viewdirs, (i % skip_layer_dir == 0 and i > 0)) for i, n in enumerate(names)],
            heads=[[("output_ambient_rgb_layer", 3)]]) if net_width_viewdirs <= 128 else None

    def layer_names(self):
        return [f"layer_{i}" for i in range(self.depth - 1)] + ["layer_bottleneck"]

    def __call__(self, p, refdirs, roughness, shader_bottleneck):
        x = []
        if self.use_shader_bottleneck:
            x.append(shader_bottleneck)
        x.append(self.dir_enc_fn(refdirs, roughness))
        if self.bf16 and self.chain is not None:
            lead = x[0].shape[:-1]
            (raw,) = mlp_chain.apply(self.chain, p, [t.reshape(-1, t.shape[-1]) for t in x])
            ambient = torch.nn.functional.softplus(raw.reshape(lead + (3,)) + self.ambient_rgb_bias)
            return dict(incoming_ambient_rgb=torch.clamp(ambient, 0.0, self.ambient_rgb_max),
                        incoming_acc=torch.ones(lead, device=raw.device, dtype=torch.float32))
        x = torch.cat(x, dim=-1) if len(x) > 1 else x[0]
        inputs = x
        for i, name in enumerate(self.layer_names()):  # run_surface_lightfield_network :480-500
            x = dense(p[name], x, relu=True, bf16=self.bf16)
            if i % self.skip == 0 and i > 0:
                x = torch.cat([x, inputs], dim=-1)
        ambient = torch.nn.functional.softplus(
            dense(p["output_ambient_rgb_layer"], x, bf16=self.bf16) + self.ambient_rgb_bias)
        acc = torch.ones_like(x[..., 0])
        return dict(incoming_ambient_rgb=torch.clamp(ambient, 0.0, self.ambient_rgb_max), incoming_acc=acc)


# ----------------------------------------------------------------------------- cache shader
APPEARANCE_GRID = dict(hash_map_size=524288, max_grid_size=2048, num_features=4)


def shader_pack(shader, names, flat):
    """bf16 operand images of the four stacks (one launch): (packed buffer, per-stack views)."""
    params = _unflatten_shader(names, flat)
    specs = [(shader.trunk_chain, params[""]), (shader.brdf_chain, params[""]),
             (shader.surface_lf.chain, params["SurfaceLightField"]), (shader.env_map.chain, params["EnvMap"])]
    return mlp_chain.pack_weights_many(specs)


def shader_encode(shader, means, arena):
    """contract + appearance-grid encode of the shaded points (shading.py:133-220): (z [P,3], enc [P,32])."""
    P = means.numel() // 3
    m2 = means.reshape(P, 3).contiguous()
    z = torch.empty_like(m2)
    _lib.call("nrc_contract_fwd", _lib.stream_ptr(), _lib.ptr(m2), P, float(shader.warp_c), _lib.ptr(z))
    enc = torch.empty((P, shader.grid.num_outputs), device=means.device, dtype=torch.float32)
    desc = shader.grid._descriptor(shader.grid.tables(shader.grid.views(arena)), None)
    _lib.call("nrc_encode_fwd", _lib.stream_ptr(), C.byref(desc), _lib.ptr(z), P, _lib.ptr(enc))
    return z, enc


def shader_fused_forward(shader, names, flat, viewdirs, means, density_feature, normals, arena, train, packed=None,
                         encoded=None, env_stream=None, want_bottleneck=True, defer_out=False):
    """Forward schedule of the bf16 cache shader (no autograd): 1 weight pack, contract + appearance-grid
    encode, trunk stack, per-point `mid` stage, integrated-BRDF / EnvMap / SurfaceLightField stacks,
    per-point `out` stage.  `packed` / `encoded` may be supplied by a caller that produced them earlier on
    another stream; defer_out=True leaves the per-point `out` stage to the caller (the training step folds it into
    nrc_shade_render_loss: outputs[0:2] are None and saved carries the raw stack outputs + the stage's constants);
    `env_stream` is accepted for compatibility (the EnvMap stack now shares one launch with the integrated-BRDF and SurfaceLightField stacks).  Returns (outputs,
    saved-for-backward, meta)."""
    lead = means.shape[:-1]
    P = means.numel() // 3
    spr = int(lead[-1]) if len(lead) > 1 else 1
    dev = means.device
    params = _unflatten_shader(names, flat)
    packed, views = packed if packed is not None else shader_pack(shader, names, flat)
    z, enc = encoded if encoded is not None else shader_encode(shader, means, arena)
    feat = density_feature.reshape(P, 64).contiguous()
    nrm = normals.reshape(P, 3).contiguous()
    vd = viewdirs.reshape(-1, 3).contiguous()
    # Stack inputs travel as bf16 tile images in the chains' operand layout: the trunk's epilogue writes the bottleneck
    # atoms, the mid stage the IDE / n.v atoms; the downstream chains bulk-copy them (no fp32 rows are re-read and
    # converted, the 128-wide fp32 bottleneck is only written when the caller wants it).
    Img = mlp_chain.ImgRef
    img_in = mlp_chain.new_image(P, 4, dev)     # atoms 0-1 bottleneck, 2-3 IDE_5 (72 -> 80 columns)
    img_aux = mlp_chain.new_image(P, 2, dev)    # atom 0: n.v (16 columns), atom 1: IDE_4 (38 -> 48 columns)
    (bott, heads), _, act_t = mlp_chain.run_forward(shader.trunk_chain, params[""], [feat, enc], views[0], save=train,
                                                    head_images={0: Img(img_in, 0, 2, 4)}, head_fp32={0: want_bottleneck})
    t5, t4 = _IdeTables.get(5), _IdeTables.get(4)
    rough = torch.empty((P,), device=dev, dtype=torch.float32)
    refdirs = torch.empty((P, 3), device=dev, dtype=torch.float32)
    images = _lib.nrc_shader_images_t(img_in.data_ptr(), 4, 2, img_aux.data_ptr(), 2, 1, img_aux.data_ptr(), 2, 0)
    _lib.call("nrc_shader_mid_fwd", _lib.stream_ptr(), t5.n_sh, t5.m, t5.l, t5.sigma, _lib.ptr(t5.mat(dev)), t4.n_sh,
              _lib.ptr(heads), heads.shape[1], _lib.ptr(nrm), _lib.ptr(vd), P, spr, -1.0, _lib.ptr(rough),
              None, _lib.ptr(refdirs), None, None, C.byref(images))
    # the three stacks behind the mid stage are independent: ONE launch, their tiles dealt to the CTA pairs together
    batch = mlp_chain.Batch()
    (fbuf,), _, act_b = mlp_chain.run_forward(shader.brdf_chain, params[""], [Img(img_in, 0, 2, 4), Img(img_aux, 0, 1, 2)],
                                              views[1], save=train, P=P, batch=batch)
    (ebuf,), _, _ = mlp_chain.run_forward(shader.env_map.chain, params["EnvMap"], [Img(img_aux, 1, 1, 2)], views[3], save=False,
                                          P=P, batch=batch)
    (sbuf,), _, act_s = mlp_chain.run_forward(shader.surface_lf.chain, params["SurfaceLightField"],
                                              [Img(img_in, 0, 2, 4), Img(img_in, 2, 2, 4)], views[2], save=train, P=P, batch=batch)
    batch.flush()
    lb = float(shader.surface_lf.ambient_rgb_bias)
    rgb = extras = None
    if not defer_out:
        rgb = torch.empty((P, 3), device=dev, dtype=torch.float32)
        extras = torch.empty((P, 22), device=dev, dtype=torch.float32)
        _lib.call("nrc_shader_out_fwd", _lib.stream_ptr(), _lib.ptr(heads), heads.shape[1], _lib.ptr(fbuf), fbuf.shape[1],
                  _lib.ptr(sbuf), sbuf.shape[1], _lib.ptr(ebuf), ebuf.shape[1], P, float(shader.rgb_max), -2.0, lb,
                  float(np.log(3.0)), _lib.ptr(rgb), _lib.ptr(extras))
    outs = (rgb.reshape(lead + (3,)) if rgb is not None else None,
            extras.reshape(lead + (22,)) if extras is not None else None, rough.reshape(lead + (1,)),
            bott.reshape(lead + (128,)) if bott is not None else None, refdirs.reshape(lead + (3,)),
            enc.reshape(lead + (-1,)))
    saved = (z, nrm, vd, heads, fbuf, sbuf, act_t, act_b, act_s, packed)
    if defer_out:
        saved = saved + ((ebuf, (float(shader.rgb_max), -2.0, lb, float(np.log(3.0)))),)
    return outs, saved, (lead, P, spr)


def shader_fused_backward(shader, names, flat, saved, meta, arena, g_rgb, need_arena_grad=True, on_data_grads=None,
                          on_grid_grads=None, g_out=None):
    """Backward schedule: per-point `out` VJP, SurfaceLightField and integrated-BRDF data-gradient chains
    (the EnvMap's gradient is exactly zero: 1 - incoming_acc == 0), per-point `mid` VJP (IDE), trunk
    data-gradient chain, appearance-grid scatter, ONE weight-gradient launch for the three stacks.
    Returns (d_density_feature [P,64], d_normals [P,3], g_arena | None, sinks, sunk).
    `on_data_grads(d_feat, g_nrm)` is called (in stream order) as soon as the gradients that leave the shader towards
    the density field are final, i.e. BEFORE the weight-gradient launch and the appearance-grid scatter: the caller
    forks the final sampler level's backward there.  g_out = (g_heads, g_f, g_s) [P,16] each: the `out` stage's VJP was
    already taken by the caller (nrc_shade_render_loss) and g_rgb is ignored."""
    lead, P, spr = meta
    z, nrm, vd, heads, fbuf, sbuf, act_t, act_b, act_s, packed = saved[:10]
    dev = z.device
    params = _unflatten_shader(names, flat)
    b0, views = 0, []
    for spec in (shader.trunk_chain, shader.brdf_chain, shader.surface_lf.chain, shader.env_map.chain):
        views.append(packed[b0 * (mlp_chain.ATOM_BYTES // 2):])
        b0 += mlp_chain._built(spec).num_chunks
    new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    lb = float(shader.surface_lf.ambient_rgb_bias)
    if g_out is not None:
        g_heads, g_f, g_s = g_out
    else:
        g2 = g_rgb.reshape(P, 3).contiguous()
        g_heads, g_f, g_s = new(P, 16), new(P, 16), new(P, 16)
        _lib.call("nrc_shader_out_bwd", _lib.stream_ptr(), _lib.ptr(heads), heads.shape[1], _lib.ptr(fbuf), fbuf.shape[1],
                  _lib.ptr(sbuf), sbuf.shape[1], P, float(shader.rgb_max), -2.0, lb, float(np.log(3.0)), _lib.ptr(g2),
                  _lib.ptr(g_heads), 16, _lib.ptr(g_f), 16, _lib.ptr(g_s), 16)
    # d(bottleneck) leaves the SurfaceLightField and integrated-BRDF gradient chains as bf16 atoms (two images; the
    # trunk's gradient chain takes their SUM as its upstream gradient: dX and dW are linear in dY)
    Img = mlp_chain.ImgRef
    img_db = mlp_chain.new_image(P, 4, dev)
    g_ide5, g_dot = new(P, 72), new(P, 1)
    batch = mlp_chain.Batch()
    dy_b = mlp_chain.run_backward_data(shader.brdf_chain, params[""], [g_f], act_b, views[1], P,
                                       [Img(img_db, 2, 2, 4), (g_dot, False)], batch=batch)
    dy_s = mlp_chain.run_backward_data(shader.surface_lf.chain, params["SurfaceLightField"], [g_s], act_s, views[2], P,
                                       [Img(img_db, 0, 2, 4), (g_ide5, False)], batch=batch)
    batch.flush()
    g_nrm = new(P, 3)
    t5, t4 = _IdeTables.get(5), _IdeTables.get(4)
    _lib.call("nrc_shader_mid_bwd", _lib.stream_ptr(), t5.n_sh, t5.m, t5.l, t5.sigma, _lib.ptr(t5.mat(dev)), t4.n_sh,
              _lib.ptr(heads), heads.shape[1], _lib.ptr(nrm), _lib.ptr(vd), P, spr, -1.0, _lib.ptr(g_dot), 1,
              _lib.ptr(g_ide5), 72, None, 0, _lib.ptr(g_heads), 16, _lib.ptr(g_nrm))
    d_feat, d_enc = new(P, 64), new(P, 32)
    dy_t = mlp_chain.run_backward_data(shader.trunk_chain, params[""], [[Img(img_db, 0, 2, 4), Img(img_db, 2, 2, 4)], g_heads],
                                       act_t, views[0], P, [(d_feat, False), (d_enc, False)])
    if on_data_grads is not None:
        on_data_grads(d_feat, g_nrm)
    # appearance-grid scatter first: it is the large half of the shader's gradients (a data-parallel harness starts that
    # bucket's all-reduce from `on_grid_grads`, beside the weight-gradient launch)
    g_arena = None
    if need_arena_grad:
        sink = _lib.grad_sink(arena)
        g_arena = sink if sink is not None else torch.zeros_like(arena)
        desc = shader.grid._descriptor(shader.grid.tables(shader.grid.views(arena)),
                                       shader.grid.tables(shader.grid.views(g_arena)))
        _lib.call("nrc_encode_bwd", _lib.stream_ptr(), C.byref(desc), _lib.ptr(z), _lib.ptr(d_enc), P, None)
        if sink is not None:
            g_arena = None
    if on_grid_grads is not None:
        on_grid_grads()
    named = {(scope, name): (flat[2 * i], flat[2 * i + 1]) for i, (scope, name) in enumerate(names)}
    sinks, sunk = mlp_chain.resolve_sinks({k: v for k, v in named.items() if k[0] != "EnvMap"})
    wptrs = mlp_chain._Ptrs()
    layers = []
    for spec, scope, act, dy in ((shader.surface_lf.chain, "SurfaceLightField", act_s, dy_s),
                                 (shader.brdf_chain, "", act_b, dy_b), (shader.trunk_chain, "", act_t, dy_t)):
        local = {n: sinks[(sc, n)] for (sc, n) in sinks if sc == scope}
        extra = {0: [mlp_chain.ImgRef(img_db, 2, 2, 4)]} if spec is shader.trunk_chain else None
        layers += mlp_chain.wgrad_layers(spec, act, dy, local, wptrs, extra_head_dy=extra)
    mlp_chain.wgrad_launch(layers, wptrs, P)
    return d_feat, g_nrm, g_arena, sinks, sunk


class _ShaderBf16Fn(torch.autograd.Function):
    """The whole cache shader (bf16 tensor-core variant) as ONE custom VJP over
    shader_fused_forward / shader_fused_backward."""

    @staticmethod
    def forward(ctx, shader, names, viewdirs, means, density_feature, normals, arena, *flat):
        train = any(t.requires_grad for t in flat) or density_feature.requires_grad or normals.requires_grad
        outs, saved, meta = shader_fused_forward(shader, names, flat, viewdirs, means, density_feature, normals, arena,
                                                 train)
        ctx.shader, ctx.names, ctx.meta = shader, names, meta
        # the saved tuple mixes tensors with ActImage records (tile images + where each input atom lives): tensors go
        # through save_for_backward (version checks), the records stay on ctx
        ctx.saved_layout = [isinstance(t, torch.Tensor) or t is None for t in saved]
        ctx.saved_objs = [None if is_t else t for t, is_t in zip(saved, ctx.saved_layout)]
        ctx.save_for_backward(*[t for t, is_t in zip(saved, ctx.saved_layout) if is_t], arena, *flat)
        ctx.mark_non_differentiable(*[o for o in outs[1:] if o is not None])
        return outs

    @staticmethod
    def backward(ctx, g_rgb, *_unused):
        shader, names = ctx.shader, ctx.names
        lead = ctx.meta[0]
        n_t = sum(ctx.saved_layout)
        tensors = list(ctx.saved_tensors[:n_t])
        saved = tuple(tensors.pop(0) if is_t else obj for is_t, obj in zip(ctx.saved_layout, ctx.saved_objs))
        arena, flat = ctx.saved_tensors[n_t], ctx.saved_tensors[n_t + 1:]
        d_feat, g_nrm, g_arena, sinks, sunk = shader_fused_backward(
            shader, names, flat, saved, ctx.meta, arena, g_rgb, need_arena_grad=ctx.needs_input_grad[6])
        grads = [None, None, None, None, d_feat.reshape(lead + (64,)), g_nrm.reshape(lead + (3,)), g_arena]
        for key in names:
            if key[0] == "EnvMap" or sunk[key]:
                grads += [None, None]     # exactly zero (EnvMap: 1 - incoming_acc == 0), or sunk
            else:
                grads += list(sinks[key])
        return tuple(grads)


def _unflatten_shader(names, flat):
    params = {"": {}, "SurfaceLightField": {}, "EnvMap": {}}
    for i, (scope, name) in enumerate(names):
        params[scope][name] = {"kernel": flat[2 * i], "bias": flat[2 * i + 1]}
    return params


class NeRFMLP:
    """Cache shader (internal/nerf.py:561-689,940-1090) under configs/ngp_yobo.gin:143-176 and
    configs/nerf_ngp_yobo.gin:491-506."""

    def __init__(self, warp_c=2.0, bbox_scaling=1.0, rgb_max=10000.0, bf16=False):
        self.grid = grid_utils.HashEncoding(bbox_scaling=bbox_scaling, scale_supersample=1.0, **APPEARANCE_GRID)
        self.warp_c = warp_c if warp_c is not None else 0.0
        self.rgb_max = rgb_max
        self.bf16 = bf16
        self.fused = True   # bf16 variant: whole shader as one custom VJP (_ShaderBf16Fn)
        self.surface_lf = SurfaceLightFieldMLP(5, True, bf16=bf16)
        self.env_map = SurfaceLightFieldMLP(4, False, bf16=bf16)
        # bf16 variant: bottleneck + the four heads are ONE GEMM over the shared 96-wide feature
        self.trunk_chain = mlp_chain.ChainSpec(
            in_widths=[64, 32], hidden=[],
            heads=[[("bottleneck_layer", 128)],
                   [("roughness_layer", 1), ("ambient_irradiance_layer", 3), ("irradiance_layer", 3), ("tint_layer", 3)]])
        self.brdf_chain = mlp_chain.ChainSpec(
            in_widths=[128, 1], hidden=[("integrated_brdf_layers_0", 64, False), ("integrated_brdf_layers_1", 64, False)],
            heads=[[("output_integrated_brdf_layer", 1)]])

    def fused_params(self, p):
        """((scope, name) list, [kernel, bias, ...]) of every Dense layer, in stack order."""
        names, flat = [], []
        for scope, spec in (("", self.trunk_chain), ("", self.brdf_chain), ("SurfaceLightField", self.surface_lf.chain),
                            ("EnvMap", self.env_map.chain)):
            src = p[scope] if scope else p
            for name in [h[0] for h in spec.hidden] + [n for grp in spec.heads for n, _ in grp]:
                names.append((scope, name))
                flat += [src[name]["kernel"], src[name]["bias"]]
        return tuple(names), flat

    def _call_fused(self, p, viewdirs, means, density_feature, normals, return_feature):
        names, flat = self.fused_params(p)
        rgb, ex, rough, bott, refdirs, enc = _ShaderBf16Fn.apply(
            self, names, viewdirs, means, density_feature, normals, p["appearance_grid"]["_arena"], *flat)
        out = dict(rgb=rgb, diffuse_rgb=ex[..., 0:3], specular_rgb=ex[..., 3:6], ambient_rgb=ex[..., 6:9],
                   indirect_rgb=ex[..., 9:12], albedo_rgb=ex[..., 12:15], integrated_brdf=ex[..., 15:16],
                   env_rgb=ex[..., 16:19], ref_rgb=ex[..., 19:22], roughness=rough, bottleneck=bott, refdirs=refdirs,
                   feature=torch.cat([density_feature, enc], dim=-1) if return_feature else None)
        return out

    def from_oracle(self, p, device):
        def mv(t):
            if isinstance(t, dict):
                return {k: mv(v) for k, v in t.items()}
            return t.detach().to(device).contiguous()

        out = {k: mv(v) for k, v in p.items() if k != "appearance_grid"}
        names = [n for (n, _, _, _) in self.grid.level_layout]
        arena = torch.cat([p["appearance_grid"][n].detach().reshape(-1) for n in names]).to(device)
        out["appearance_grid"] = dict(self.grid.views(arena), _arena=arena)
        return out

    def predict_appearance_feature(self, p, density_feature, means):
        """shading.py:133-220."""
        z = coord._ContractFn.apply(means, self.warp_c)
        enc = self.grid(p["appearance_grid"], z)
        return torch.cat([density_feature, enc], dim=-1)

    def __call__(self, p, viewdirs, means, density_feature, normals, return_feature=False):
        sp = torch.nn.functional.softplus
        b = self.bf16
        if b and self.fused:
            return self._call_fused(p, viewdirs, means, density_feature, normals, return_feature)
        if b:
            z = coord._ContractFn.apply(means, self.warp_c)
            enc = self.grid(p["appearance_grid"], z)
            lead = enc.shape[:-1]
            bott, r_raw, ai_raw, irr_raw, tint_raw = mlp_chain.apply(
                self.trunk_chain, p, [density_feature.reshape(-1, 64), enc.reshape(-1, 32)])
            # the concatenated feature only exists inside the chain's shared-memory tile
            feature = torch.cat([density_feature, enc], dim=-1) if return_feature else None
            bottleneck = bott.reshape(lead + (128,))
            r_raw, ai_raw, irr_raw, tint_raw = (t.reshape(lead + (t.shape[-1],)) for t in (r_raw, ai_raw, irr_raw, tint_raw))
        else:
            feature = self.predict_appearance_feature(p, density_feature, means)
            bottleneck = dense(p["bottleneck_layer"], feature)
            r_raw = dense(p["roughness_layer"], feature)
            ai_raw = dense(p["ambient_irradiance_layer"], feature)
            irr_raw = dense(p["irradiance_layer"], feature)
            tint_raw = dense(p["tint_layer"], feature)
        roughness = sp(r_raw - 1.0)
        ambient_diffuse = torch.clamp(sp(ai_raw - 2.0), 0.0, self.rgb_max)
        tint = torch.sigmoid(tint_raw)
        dotprod = torch.sum(normals * (-viewdirs[..., None, :]), dim=-1, keepdim=True)
        if b:
            (f_raw,) = mlp_chain.apply(self.brdf_chain, p, [bottleneck.reshape(-1, 128), dotprod.reshape(-1, 1)])
            f_raw = f_raw.reshape(dotprod.shape)
        else:
            x = torch.cat([bottleneck, dotprod], dim=-1)
            x = dense(p["integrated_brdf_layers_0"], x, relu=True)
            x = dense(p["integrated_brdf_layers_1"], x, relu=True)
            f_raw = dense(p["output_integrated_brdf_layer"], x)
        F = torch.sigmoid(f_raw + float(np.log(3.0)))
        refdirs = reflect(-viewdirs[..., None, :], normals)
        env_rgb = self.env_map(p["EnvMap"], refdirs, roughness, None)["incoming_ambient_rgb"]
        indirect_diffuse = torch.clamp(sp(irr_raw - 2.0), 0.0, self.rgb_max)
        inc = self.surface_lf(p["SurfaceLightField"], refdirs, roughness, bottleneck)
        ref_rgb = inc["incoming_ambient_rgb"]
        ref_acc = inc["incoming_acc"][..., None]
        ambient_specular = torch.clamp(tint * F * (env_rgb * (1.0 - ref_acc)), 0.0, self.rgb_max)
        indirect_specular = torch.clamp(tint * F * (ref_rgb * ref_acc), 0.0, self.rgb_max)
        ambient = ambient_diffuse + ambient_specular
        indirect = indirect_diffuse + indirect_specular
        return dict(
            rgb=ambient + indirect, diffuse_rgb=ambient_diffuse + indirect_diffuse,
            specular_rgb=ambient_specular + indirect_specular, ambient_rgb=ambient, indirect_rgb=indirect,
            albedo_rgb=tint, roughness=roughness, integrated_brdf=F, env_rgb=env_rgb, ref_rgb=ref_rgb,
            bottleneck=bottleneck, feature=feature, refdirs=refdirs,
        )


# ----------------------------------------------------------------------------- transient heads (row 22)
class TransientIndirectHead:
    """TransientNeRFMLP.get_indirect (internal/nerf.py:1757-1777): [shader feature 96 || pos_enc(light position,
    0..2) 15] = 111 -> Dense 64 + ReLU -> Dense 64 + ReLU -> transient_indirect_layer 64 -> n_bins * C (raw;
    softplus(. + irradiance_bias), indirect_scale, validity masks and the time shift are fused into
    render.volumetric_transient_rendering).  The 2100-wide output layer runs on the generic GEMM
    (nrc_dense_fwd); the [P, n_bins*C] result is the only large tensor of the transient path."""

    def __init__(self, n_bins=700, channels=3, deg_lights=2, width=64, bf16=False):
        self.n_bins, self.channels, self.deg, self.width, self.bf16 = n_bins, channels, deg_lights, width, bf16
        self.in_dim = 96 + 3 + 6 * deg_lights

    def init(self, device, generator=None):
        def layer(fi, fo):
            lim = float(np.sqrt(6.0 / fi))
            return {"kernel": torch.empty((fi, fo), device=device).uniform_(-lim, lim, generator=generator),
                    "bias": torch.zeros((fo,), device=device)}
        return {"irradiance_layers_0": layer(self.in_dim, self.width), "irradiance_layers_1": layer(self.width, self.width),
                "transient_indirect_layer": layer(self.width, self.n_bins * self.channels)}

    def hidden(self, p, feature, lights):
        """The irradiance stack up to its last ReLU (run_irradiance_network, internal/nerf.py:1757-1772): [P,64].  The
        fused time-resolved kernel (render.volumetric_transient_rendering_fused) applies transient_indirect_layer itself."""
        P = feature.shape[0]
        l2 = lights.reshape(P, 3).contiguous()
        enc = torch.empty((P, 3 + 6 * self.deg), device=feature.device, dtype=torch.float32)
        _lib.call("nrc_pos_enc", _lib.stream_ptr(), _lib.ptr(l2), P, 3, 0, self.deg, 1, _lib.ptr(enc), enc.shape[1])
        x = torch.cat([feature.reshape(P, -1), enc], dim=-1)
        x = dense(p["irradiance_layers_0"], x, relu=True, bf16=self.bf16)
        return dense(p["irradiance_layers_1"], x, relu=True, bf16=self.bf16)

    def __call__(self, p, feature, lights):
        """feature [P,96], lights [P,3] -> raw transient indirect [P, n_bins, C]."""
        P = feature.shape[0]
        l2 = lights.reshape(P, 3).contiguous()
        enc = torch.empty((P, 3 + 6 * self.deg), device=feature.device, dtype=torch.float32)
        _lib.call("nrc_pos_enc", _lib.stream_ptr(), _lib.ptr(l2), P, 3, 0, self.deg, 1, _lib.ptr(enc), enc.shape[1])
        x = torch.cat([feature.reshape(P, -1), enc], dim=-1)
        x = dense(p["irradiance_layers_0"], x, relu=True, bf16=self.bf16)
        x = dense(p["irradiance_layers_1"], x, relu=True, bf16=self.bf16)
        return dense(p["transient_indirect_layer"], x, bf16=self.bf16).reshape(P, self.n_bins, self.channels)
