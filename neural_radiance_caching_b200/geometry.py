"""Host-side mirror of internal/geometry.py DensityMLP (CUDA bodies).

Parameter names follow the Flax auto-names of the reference's setup() attributes
(`density_layers_{i}`, `output_density_layer`, `pred_normals_layer`, `density_grid/...`;
internal/geometry.py:123-151), kernels [in,out], biases [out], fp32.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, grid_utils, mlp_chain


def _mlp_desc(p, in_dim, use_pred_normals):
    d = _lib.nrc_density_mlp_t()
    d.d_w0 = _lib.ptr(p["density_layers_0"]["kernel"]).value
    d.d_b0 = _lib.ptr(p["density_layers_0"]["bias"]).value
    d.d_w1 = _lib.ptr(p["density_layers_1"]["kernel"]).value
    d.d_b1 = _lib.ptr(p["density_layers_1"]["bias"]).value
    d.d_wd = _lib.ptr(p["output_density_layer"]["kernel"]).value
    d.d_bd = _lib.ptr(p["output_density_layer"]["bias"]).value
    if use_pred_normals:
        d.d_wn = _lib.ptr(p["pred_normals_layer"]["kernel"]).value
        d.d_bn = _lib.ptr(p["pred_normals_layer"]["bias"]).value
    d.in_dim = in_dim
    d.width = 64
    return d


_MLP_KEYS = ("density_layers_0", "density_layers_1", "output_density_layer", "pred_normals_layer")


class _RunNetworkFn(torch.autograd.Function):
    """custom_vjp analogue over nrc_density_mlp_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, mlp, x, bf16, *flat):
        p = mlp._unflatten(flat)
        x2 = x.reshape(-1, mlp.in_dim).contiguous()
        P = x2.shape[0]
        raw = torch.empty((P,), device=x.device, dtype=torch.float32)
        feat = torch.empty((P, 64), device=x.device, dtype=torch.float32)
        gp = torch.empty((P, 3), device=x.device, dtype=torch.float32) if mlp.enable_pred_normals else None
        desc = _mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals)
        _lib.call("nrc_density_mlp_fwd", _lib.stream_ptr(), C.byref(desc), _lib.ptr(x2), P, int(bf16),
                  _lib.ptr(raw), _lib.ptr(feat), _lib.ptr(gp))
        ctx.mlp = mlp
        ctx.bf16 = bool(bf16)
        ctx.save_for_backward(x2, *flat)
        ctx.lead = x.shape[:-1]
        lead = x.shape[:-1]
        outs = (raw.reshape(lead), feat.reshape(lead + (64,)))
        if gp is not None:
            outs = outs + (gp.reshape(lead + (3,)),)
        return outs

    @staticmethod
    def backward(ctx, g_raw, g_feat, g_gp=None):
        mlp = ctx.mlp
        x2, *flat = ctx.saved_tensors
        p = mlp._unflatten(flat)
        P = x2.shape[0]
        dev = x2.device
        f = lambda g, shape: g.reshape(shape).contiguous() if g is not None else None
        g_raw2 = f(g_raw, (P,)) if g_raw is not None else torch.zeros((P,), device=dev)
        g_feat2 = f(g_feat, (P, 64))
        g_gp2 = f(g_gp, (P, 3)) if mlp.enable_pred_normals else None
        g_enc = torch.empty_like(x2) if ctx.needs_input_grad[1] else None
        want_w = any(ctx.needs_input_grad[3:])
        gflat = [torch.zeros_like(t) for t in flat] if want_w else None
        gd = None
        if want_w:
            gp_ = mlp._unflatten(gflat)
            gd = _lib.nrc_density_mlp_grad_t()
            gd.d_w0 = _lib.ptr(gp_["density_layers_0"]["kernel"]).value
            gd.d_b0 = _lib.ptr(gp_["density_layers_0"]["bias"]).value
            gd.d_w1 = _lib.ptr(gp_["density_layers_1"]["kernel"]).value
            gd.d_b1 = _lib.ptr(gp_["density_layers_1"]["bias"]).value
            gd.d_wd = _lib.ptr(gp_["output_density_layer"]["kernel"]).value
            gd.d_bd = _lib.ptr(gp_["output_density_layer"]["bias"]).value
            if mlp.enable_pred_normals:
                gd.d_wn = _lib.ptr(gp_["pred_normals_layer"]["kernel"]).value
                gd.d_bn = _lib.ptr(gp_["pred_normals_layer"]["bias"]).value
        desc = _mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals)
        _lib.call("nrc_density_mlp_bwd", _lib.stream_ptr(), C.byref(desc), _lib.ptr(x2), _lib.ptr(g_raw2), None,
                  _lib.ptr(g_feat2), _lib.ptr(g_gp2), P, int(ctx.bf16), _lib.ptr(g_enc),
                  C.byref(gd) if gd is not None else None)
        gx = g_enc.reshape(ctx.lead + (mlp.in_dim,)) if g_enc is not None else None
        return (None, gx, None) + (tuple(gflat) if want_w else (None,) * len(flat))


def _grad_desc(mlp, gp_):
    gd = _lib.nrc_density_mlp_grad_t()
    gd.d_w0 = _lib.ptr(gp_["density_layers_0"]["kernel"]).value
    gd.d_b0 = _lib.ptr(gp_["density_layers_0"]["bias"]).value
    gd.d_w1 = _lib.ptr(gp_["density_layers_1"]["kernel"]).value
    gd.d_b1 = _lib.ptr(gp_["density_layers_1"]["bias"]).value
    gd.d_wd = _lib.ptr(gp_["output_density_layer"]["kernel"]).value
    gd.d_bd = _lib.ptr(gp_["output_density_layer"]["bias"]).value
    if mlp.enable_pred_normals:
        gd.d_wn = _lib.ptr(gp_["pred_normals_layer"]["kernel"]).value
        gd.d_bn = _lib.ptr(gp_["pred_normals_layer"]["bias"]).value
    return gd


class _DensityQueryFn(torch.autograd.Function):
    """custom_vjp analogue of the fused training query.
    forward : nrc_density_query_fwd (contract + encode + MLP + activation), saving the encoded
              features [P, L*F];
    backward: nrc_density_mlp_bwd (data + weight gradients; the safe_exp / bbox-mask VJP is the
              multiplication by the saved density) then nrc_encode_bwd (scatter-add into one
              arena-shaped table gradient).  No gradient flows to the sample positions
              (stop_level_grad, internal/sampling.py:353-354)."""

    @staticmethod
    def forward(ctx, mlp, means, want_feat, arena, *flat):
        p = mlp._unflatten(flat)
        m2 = means.reshape(-1, 3).contiguous()
        P = m2.shape[0]
        dev = m2.device
        density = torch.empty((P,), device=dev, dtype=torch.float32)
        feat = torch.empty((P, 64), device=dev, dtype=torch.float32) if want_feat else None
        gp = torch.empty((P, 3), device=dev, dtype=torch.float32) if mlp.enable_pred_normals else None
        enc_out = torch.empty((P, mlp.in_dim), device=dev, dtype=torch.float32)
        enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(arena)), None)
        desc = _mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals)
        _lib.call("nrc_density_query_fwd", _lib.stream_ptr(), C.byref(enc), C.byref(desc), _lib.ptr(m2), P,
                  float(mlp.warp_c), float(mlp.density_bias), int(mlp.bf16), _lib.ptr(density), None,
                  _lib.ptr(feat), _lib.ptr(gp), None, _lib.ptr(enc_out))
        ctx.mlp = mlp
        ctx.save_for_backward(m2, enc_out, density, arena, *flat)
        lead = means.shape[:-1]
        ctx.lead = lead
        ctx.has = (want_feat, gp is not None)
        outs = [density.reshape(lead)]
        outs.append(feat.reshape(lead + (64,)) if want_feat else None)
        outs.append(gp.reshape(lead + (3,)) if gp is not None else None)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_density, g_feat, g_gp):
        mlp = ctx.mlp
        m2, enc_out, density, arena, *flat = ctx.saved_tensors
        p = mlp._unflatten(flat)
        P = m2.shape[0]
        dev = m2.device
        c = lambda g, shape: g.reshape(shape).contiguous() if g is not None else None
        g_d = c(g_density, (P,)) if g_density is not None else torch.zeros((P,), device=dev)
        g_f = c(g_feat, (P, 64)) if ctx.has[0] else None
        g_g = c(g_gp, (P, 3)) if ctx.has[1] else None
        want_w = any(ctx.needs_input_grad[4:])
        want_t = ctx.needs_input_grad[3]
        gflat, w_sunk = None, False
        if want_w:
            sinks = [_lib.grad_sink(t) for t in flat]
            w_sunk = all(s_ is not None for s_ in sinks)
            gflat = sinks if w_sunk else [torch.zeros_like(t) for t in flat]
        gd = _grad_desc(mlp, mlp._unflatten(gflat)) if want_w else None
        g_enc = torch.empty_like(enc_out) if want_t else None
        desc = _mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals)
        _lib.call("nrc_density_mlp_bwd", _lib.stream_ptr(), C.byref(desc), _lib.ptr(enc_out), _lib.ptr(g_d),
                  _lib.ptr(density), _lib.ptr(g_f), _lib.ptr(g_g), P, int(mlp.bf16), _lib.ptr(g_enc),
                  C.byref(gd) if gd is not None else None)
        g_arena = None
        if want_t:
            t_sink = _lib.grad_sink(arena)
            g_arena = t_sink if t_sink is not None else torch.zeros_like(arena)
            enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(arena)),
                                       mlp.grid.tables(mlp.grid.views(g_arena)))
            _lib.call("nrc_encode_bwd_warped", _lib.stream_ptr(), C.byref(enc), _lib.ptr(m2), float(mlp.warp_c),
                      _lib.ptr(g_enc), P)
            if t_sink is not None:
                g_arena = None
        if w_sunk:
            gflat = None
        return (None, None, None, g_arena) + (tuple(gflat) if gflat is not None else (None,) * len(flat))


class _RawGradFn(torch.autograd.Function):
    """Analytic raw-density gradient d raw / d means (internal/geometry.py:442-460: jax.vjp of predict_density
    w.r.t. the means) WITH its parameter VJP, i.e. the second-order path the predicted-normal loss takes
    (SURVEY 8f-1).  forward: nrc_density_query_fwd (raw-gradient output only); backward:
    nrc_density_normals_bwd.  No gradient to the sample positions (stop_level_grad)."""

    @staticmethod
    def forward(ctx, mlp, means, arena, *flat):
        p = mlp._unflatten(flat)
        m2 = means.reshape(-1, 3).contiguous()
        P = m2.shape[0]
        rg = torch.empty((P, 3), device=m2.device, dtype=torch.float32)
        enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(arena)), None)
        desc = _mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals)
        _lib.call("nrc_density_query_fwd", _lib.stream_ptr(), C.byref(enc), C.byref(desc), _lib.ptr(m2), P,
                  float(mlp.warp_c), float(mlp.density_bias), int(mlp.bf16), None, None, None, None, _lib.ptr(rg), None)
        ctx.mlp = mlp
        ctx.save_for_backward(m2, arena, *flat)
        ctx.lead = means.shape[:-1]
        return rg.reshape(ctx.lead + (3,))

    @staticmethod
    def backward(ctx, g_rg):
        mlp = ctx.mlp
        m2, arena, *flat = ctx.saved_tensors
        g = g_rg.reshape(-1, 3).contiguous()
        sinks = [_lib.grad_sink(t) for t in flat]
        w_sunk = all(s_ is not None for s_ in sinks)
        gflat = sinks if w_sunk else [torch.zeros_like(t) for t in flat]
        t_sink = _lib.grad_sink(arena)
        g_arena = t_sink if t_sink is not None else torch.zeros_like(arena)
        density_normals_bwd(mlp, mlp._unflatten(flat), arena, m2, g, mlp._unflatten(gflat), g_arena)
        return (None, None, None if t_sink is not None else g_arena) + (
            (None,) * len(flat) if w_sunk else tuple(gflat))


def density_normals_bwd(mlp, p, arena, means, g_raw_grad, grad_p, grad_arena, enc_out=None):
    """nrc_density_normals_bwd: accumulate d/d theta <g_raw_grad, d raw / d means> into `grad_p` (a parameter-
    shaped tree of gradient buffers; only the three kernels receive anything) and `grad_arena`.
    With the bf16 variant and `enc_out` (the primal features [P, L*F] saved by the query) the term runs on tensor cores:
    tangent gather -> the MLP's backward pass on the tangent network -> weighted table scatter (three launches)."""
    enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(arena)), mlp.grid.tables(mlp.grid.views(grad_arena)))
    desc = _mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals)
    gd = _grad_desc(mlp, grad_p)
    P = means.shape[0]
    if mlp.bf16 and enc_out is not None and os.environ.get("NRC_NORMALS2_TC", "1") == "1":
        edot = torch.empty((P, mlp.in_dim), device=means.device, dtype=torch.float32)
        ge = torch.empty_like(edot)
        _lib.call("nrc_encode_tangent_fwd", _lib.stream_ptr(), C.byref(enc), _lib.ptr(means), _lib.ptr(g_raw_grad), P,
                  float(mlp.warp_c), _lib.ptr(edot))
        _lib.call("nrc_density_mlp_bwd_tangent", _lib.stream_ptr(), C.byref(desc), _lib.ptr(enc_out), _lib.ptr(edot), P,
                  _lib.ptr(ge), C.byref(gd))
        _lib.call("nrc_encode_tangent_bwd", _lib.stream_ptr(), C.byref(enc), _lib.ptr(means), _lib.ptr(g_raw_grad),
                  _lib.ptr(ge), P, float(mlp.warp_c))
        return
    _lib.call("nrc_density_normals_bwd", _lib.stream_ptr(), C.byref(enc), C.byref(desc), _lib.ptr(means),
              _lib.ptr(g_raw_grad), P, float(mlp.warp_c), C.byref(gd))


class DensityMLP:
    """BaseDensityMLP/DensityMLP (internal/geometry.py:36-593) as configured by
    configs/ngp_yobo.gin:137-140,206-230: depth 2, width 64, ReLU, safe_exp activation,
    unscented basis 'mean', contraction warp, hash-grid input."""

    def __init__(self, grid_params, net_depth=2, net_width=64, density_bias=-1.0, warp_c=2.0, bbox_scaling=1.0,
                 enable_pred_normals=False, disable_density_normals=False, normals_for_filter_only=False,
                 bf16=False):
        if net_depth != 2 or net_width != 64:
            raise NotImplementedError("the fused kernels are compiled for net_depth=2, net_width=64")
        self.grid = grid_utils.HashEncoding(bbox_scaling=bbox_scaling, scale_supersample=1.0, **grid_params)
        self.in_dim = self.grid.num_outputs
        self.density_bias = density_bias
        self.warp_c = warp_c if warp_c is not None else 0.0
        self.enable_pred_normals = enable_pred_normals
        self.disable_density_normals = disable_density_normals
        self.normals_for_filter_only = normals_for_filter_only
        self.bf16 = bf16

    def bbox_tensors(self, device):
        """Device copies of the grid bbox corners (cached: no H2D copy inside CUDA graphs)."""
        key = str(device)
        cache = self.__dict__.setdefault("_bbox_cache", {})
        if key not in cache:
            bbox = self.grid.bbox
            cache[key] = (torch.tensor(bbox[0].astype(np.float32), device=device),
                          torch.tensor(bbox[1].astype(np.float32), device=device))
        return cache[key]

    # -- parameters -------------------------------------------------------------
    def from_oracle(self, p, device):
        """Move a parameter tree (oracle/reference layout) to the device, tables in one arena."""
        out = {}
        names = [n for (n, _, _, _) in self.grid.level_layout]
        arena = torch.cat([p["density_grid"][n].reshape(-1) for n in names]).to(device)
        out["density_grid"] = dict(self.grid.views(arena), _arena=arena)
        for k in _MLP_KEYS:
            if k in p:
                out[k] = {kk: vv.to(device).contiguous() for kk, vv in p[k].items()}
        return out

    def _flatten(self, p):
        keys = _MLP_KEYS if self.enable_pred_normals else _MLP_KEYS[:3]
        return [p[k][kk] for k in keys for kk in ("kernel", "bias")]

    def _unflatten(self, flat):
        keys = _MLP_KEYS if self.enable_pred_normals else _MLP_KEYS[:3]
        it = iter(flat)
        return {k: {"kernel": next(it), "bias": next(it)} for k in keys}

    # -- reference methods ------------------------------------------------------
    def run_network(self, p, x, means=None):
        """internal/geometry.py:155-168 -> (raw_density, x) [+ grad_pred when enabled]."""
        outs = _RunNetworkFn.apply(self, x, self.bf16, *self._flatten(p))
        return outs

    def query_train(self, p, means, want_feat=True):
        """Fused training query through the custom VJP: (density, feature|None, grad_pred|None).
        Needs the tables in one arena (`p["density_grid"]["_arena"]`)."""
        arena = p["density_grid"].get("_arena")
        if arena is None:
            raise ValueError("the fused training path needs the level tables in one arena")
        return _DensityQueryFn.apply(self, means, bool(want_feat), arena, *self._flatten(p))

    def raw_grad_density(self, p, means):
        """d raw / d means with the second-order parameter VJP attached (internal/geometry.py:442-460)."""
        arena = p["density_grid"].get("_arena")
        if arena is None:
            raise ValueError("the fused training path needs the level tables in one arena")
        return _RawGradFn.apply(self, means, arena, *self._flatten(p))

    def chain_spec(self):
        """The density stack as a tensor-core chain: two 64-wide ReLU layers, density (+ predicted-normal) head."""
        if getattr(self, "_chain_spec", None) is None:
            head = [("output_density_layer", 1)] + ([("pred_normals_layer", 3)] if self.enable_pred_normals else [])
            self._chain_spec = mlp_chain.ChainSpec(
                [self.in_dim], [("density_layers_0", 64, False), ("density_layers_1", 64, False)], [head])
        return self._chain_spec

    def supports_query_tc(self):
        return self.in_dim <= 32 and self.grid.num_features in (1, 2, 4)

    def query_tc(self, p, means, density, feat=None, grad_pred=None, enc_out=None, cache=None):
        """predict_density + convert_raw_density (internal/geometry.py:199-341,442-460) through nrc_chain_query:
        the hash-grid gather feeds tcgen05 GEMMs tile by tile.  Inference path; outputs are written in place.
        `cache` (mlp_chain.PackCache) keeps the packed bf16 weights across calls."""
        spec = self.chain_spec()
        names = [n for grp in spec.heads for n, _ in grp]

        def build():
            hb = torch.cat([p[n]["bias"] for n in names]) if len(names) > 1 else p[names[0]]["bias"]
            return mlp_chain.pack_weights(spec, p), hb

        keys = [p[n][k] for n in ("density_layers_0", "density_layers_1", *names) for k in ("kernel", "bias")]
        packed, hb = cache.get(keys, build) if cache is not None else build()
        enc = self.grid._descriptor(self.grid.tables(p["density_grid"]), None)
        mlp_chain.run_density_query(spec, p, enc, means.reshape(-1, 3), packed, hb, self.warp_c, self.density_bias,
                                    density, feat, grad_pred if self.enable_pred_normals else None, enc_out)

    def query(self, p, means, want_feat=True, want_normals=False):
        """Fused predict_density + convert_raw_density (+ analytic raw gradient):
        internal/geometry.py:199-341,442-460.  Inference path (no autograd)."""
        m2 = means.reshape(-1, 3).contiguous()
        P = m2.shape[0]
        dev = m2.device
        density = torch.empty((P,), device=dev, dtype=torch.float32)
        raw = torch.empty((P,), device=dev, dtype=torch.float32)
        feat = torch.empty((P, 64), device=dev, dtype=torch.float32) if want_feat else None
        gp = torch.empty((P, 3), device=dev, dtype=torch.float32) if self.enable_pred_normals else None
        rg = torch.empty((P, 3), device=dev, dtype=torch.float32) if want_normals else None
        enc = self.grid._descriptor(self.grid.tables(p["density_grid"]), None)
        mlp = _mlp_desc(p, self.in_dim, self.enable_pred_normals)
        _lib.call("nrc_density_query_fwd", _lib.stream_ptr(), C.byref(enc), C.byref(mlp), _lib.ptr(m2), P,
                  float(self.warp_c), float(self.density_bias), int(self.bf16), _lib.ptr(density), _lib.ptr(raw),
                  _lib.ptr(feat), _lib.ptr(gp), _lib.ptr(rg), None)
        lead = means.shape[:-1]
        r = lambda t, *s: t.reshape(lead + s) if t is not None else None
        return dict(density=r(density), raw_density=r(raw), feature=r(feat, 64), grad_pred=r(gp, 3),
                    raw_grad_density=r(rg, 3))
