"""Static schedule of the cache-stage training step (BASELINE config 2) over the C ABI.

The autograd mirrors (sampling.py / nerf.py / render.py / models.py) keep the reference's call
structure; this module runs the SAME kernels as one hand-ordered launch sequence -- forward, fused
loss (+ its gradients), backward -- with every gradient accumulated straight into the flat gradient
arena.  No elementwise glue runs between the launches, so a step is ~50 kernels and is captured in one
CUDA graph by bench.py.  tests/test_engine_gpu.py pins it against the autograd path.

Schedule (reference call sites in brackets):
  per level l = 0,1,2   nrc_ray_sample_intervals  [sampling.py:326-354, stepfun.py:207-250]
                        nrc_ray_cast              [coord.py:223-260, render.py:106-131]
                        nrc_density_query_fwd     [geometry.py:199-341,442-479]
                        nrc_ray_alpha_weights_fwd [render.py:134-169]
  nrc_normals_fwd x2 (predicted + analytic)       [geometry.py:442-479]
  shader_fused_forward                            [nerf.py:561-689,940-1090, surface_light_field.py:782-1069]
  nrc_ray_composite_fwd                           [render.py:172-247]
  nrc_charb_srgb_loss, nrc_interlevel_loss x2     [loss + d loss / d rgb, d loss / d proposal weights;
                                                   image.py:192-200, loss_utils.py:74-108]
  nrc_ray_composite_bwd, shader_fused_backward, nrc_normals_bwd
  per level l = 2,1,0   nrc_ray_alpha_weights_bwd, nrc_density_mlp_bwd, nrc_encode_bwd_warped
"""
import ctypes as C
import os

import torch

from . import _lib, geometry, mlp_chain, nerf, stepfun


class FusedCacheStep:
    def __init__(self, model, params, charb_padding=0.001, interlevel_mults=(0.01, 0.01), interlevel_blurs=(0.03, 0.003),
                 geometry_mults=(0.01, 0.001, 0.01), predicted_normal_stopgrad_weight=0.1, mask_weights=(1.0, 1.0),
                 backward_mask_weight=0.1, distortion=(0.01, -0.25, 10000.0), density_grid_regularizer=1.0):
        self.model, self.params = model, params
        self.charb_padding = charb_padding
        self.interlevel_mults, self.interlevel_blurs = interlevel_mults, interlevel_blurs
        # geometry / mask losses (configs/nerf_ngp_yobo_lego.gin:7-11, nerf_ngp_yobo.gin:59-72,367-376)
        self.geometry_mults = geometry_mults            # orientation, predicted normals, reverse; None: off
        self.sg_w = predicted_normal_stopgrad_weight
        self.mask_weights = mask_weights                # opaque, empty; None: off
        self.backward_mask_weight = backward_mask_weight
        self.distortion = distortion                    # mult, power_ladder p, premult; None: off
        self.density_grid_regularizer = density_grid_regularizer   # Config.param_regularizers['density_grid']; None: off
        self._bg = {}
        self._side = None
        self.final_grads_announced = False   # last step(): on_final_grads was invoked (the tail fork ran)
        self.concurrent = True   # independent branches of the schedule on side streams (fork/join events)
        if model.sampler.opaque_background:
            # nrc_ray_alpha_weights_bwd recomputes a finite density * delta for the last sample: with an opaque background
            # its gradients would be wrong (render.compute_alpha_weights raises for the same combination)
            raise NotImplementedError("FusedCacheStep: opaque_background=True has no backward pass")

    def _constants(self, R, dev):
        """Per-batch-size constants, created and filled on the CALLER's stream before any side stream is forked: a side
        stream that first-used them would leave the main stream reading memory it never ordered itself behind."""
        self._bg_ones(R, dev)
        self._initial_step_function(R, dev)
        self._zero_mask(R, dev)

    def _streams(self, n):
        if self._side is None:
            self._side = []
        while len(self._side) < n:          # extend, never replace: streams handed out earlier stay valid
            self._side.append(torch.cuda.Stream())
        return self._side

    def _bg_ones(self, R, dev):
        key = (R, str(dev))
        if key not in self._bg:
            self._bg[key] = torch.ones((R, 3), device=dev, dtype=torch.float32)
        return self._bg[key]

    def _initial_step_function(self, R, dev):
        """sdist = [0, 1], weights = [1] per ray (sampling.py:300-309), built once per batch size."""
        key = ("s0", R, str(dev))
        if key not in self._bg:
            sd = torch.zeros((R, 2), device=dev, dtype=torch.float32)
            sd[:, 1] = 1.0
            self._bg[key] = (sd, torch.ones((R, 1), device=dev, dtype=torch.float32))
        return self._bg[key]

    def step(self, rays, u01, target_rgb, train_frac=1.0, extra=None, zero_grad=None, on_shader_grads=None,
             on_proposal_grads=None, on_final_grads=None, on_grid_grads=None):
        """One forward + loss + backward; gradients land in the registered sinks.  Returns the loss
        (device scalar) and leaves the per-level sampler state in self.last (for tests).
        `extra` = (rays, u01) of the backward-mask pass (train_utils.py:3348-3401) or None.
        `zero_grad` = callable that clears the gradient sinks (the 110 MB arena memset): issued here on a side
        stream beside the sampler's forward instead of in front of the step.
        `on_shader_grads` = callable invoked (in stream order on the main stream) as soon as every gradient of the
        `Shader` parameters is final: a data-parallel harness forks that bucket's all-reduce there.
        `on_proposal_grads` = the same for the proposal levels' parameters (every sampler level but the last), invoked
        on the proposal branch's side stream right after that branch's backward.  `on_grid_grads` = the same for the
        appearance grid alone (its scatter precedes the stacks' weight gradients), `on_final_grads` = for the final
        sampler level's parameters, invoked on the stream that ran that level's backward (only when it was forked off
        beside the shader's weight gradients: self.final_grads_announced tells)."""
        hi = self._hi_stream()
        if hi is None:
            state = self.step_front(rays, u01, target_rgb, train_frac, fork_proposals=True, extra=extra, zero_grad=zero_grad,
                                    on_shader_grads=on_shader_grads, on_proposal_grads=on_proposal_grads, on_final_grads=on_final_grads,
                                    on_grid_grads=on_grid_grads)
            self.step_back(state)
            self.final_grads_announced = state.get("final_grads_announced", False)
            return state["loss"]
        # The step's main chain (sampler forward -> shader forward / backward -> final level's backward) is the critical
        # path; the proposal supervision, the geometry losses and the backward-mask pass run beside it on the side
        # streams.  The main chain is issued on a HIGH-priority stream (the priority is recorded in the captured kernel
        # nodes), so that its CTAs - the chain kernel needs a whole SM each - are placed before the side branches'.
        outer = torch.cuda.current_stream()
        hi.wait_stream(outer)
        with torch.cuda.stream(hi):
            state = self.step_front(rays, u01, target_rgb, train_frac, fork_proposals=True, extra=extra, zero_grad=zero_grad,
                                    on_shader_grads=on_shader_grads, on_proposal_grads=on_proposal_grads, on_final_grads=on_final_grads,
                                    on_grid_grads=on_grid_grads)
            self.step_back(state)
        outer.wait_stream(hi)
        self.final_grads_announced = state.get("final_grads_announced", False)
        return state["loss"]

    def _hi_stream(self):
        if not self.concurrent or os.environ.get("NRC_HI_PRIO", "1") != "1":
            return None
        if getattr(self, "_hi", None) is None:
            self._hi = torch.cuda.Stream(priority=-1)
        return self._hi

    def _weights_only_pass(self, rays, u01, train_frac, loss, grads_ready=None):
        """Backward-mask term: sampler-only forward on the extra rays (weights_only=True), mask loss against a zero
        mask on acc = sum(weights), and the final level's backward (the proposal levels receive nothing from this
        pass: they are supervised by the main rays' interlevel loss only and positions are stop-gradiented)."""
        sampler = self.model.sampler
        sp = self.params["Sampler"]
        R, dev = rays["near"].shape[0], rays["near"].device
        st = _lib.stream_ptr
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        anneal = sampler.anneal(train_frac)
        sdist, weights = self._initial_step_function(R, dev)
        prev, fuse_w = None, os.environ.get("NRC_FUSE_WEIGHTS", "1") == "1"
        nl = len(sampler.sampling_strategy)
        lv = None
        for i_level, (i_mlp, _, n) in enumerate(sampler.sampling_strategy):
            mlp, p = sampler.mlps[i_mlp], sp[f"MLP_{i_mlp}"]
            last = i_level == nl - 1
            sdist, tdist, means = sampler.sample_and_cast(u01[i_level], sdist, weights, n, anneal, rays, False, prev=prev)
            prev = None
            P = R * n
            density = new(P)
            enc_out = new(P, mlp.in_dim) if last else None
            arena = p["density_grid"]["_arena"]
            enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(arena)), None)
            desc = geometry._mlp_desc(p, mlp.in_dim, False)
            _lib.call("nrc_density_query_fwd", st(), C.byref(enc), C.byref(desc), _lib.ptr(means), P, float(mlp.warp_c),
                      float(mlp.density_bias), int(mlp.bf16), _lib.ptr(density), None, None, None, None, _lib.ptr(enc_out))
            weights = new(R, n)
            if last or not fuse_w:
                _lib.call("nrc_ray_alpha_weights_fwd", st(), _lib.ptr(density), _lib.ptr(tdist), _lib.ptr(rays["directions"]),
                          R, n, int(sampler.opaque_background), _lib.ptr(weights), None, None)
            else:
                prev = (density, tdist)
            if last:
                lv = dict(mlp=mlp, p=p, n=n, tdist=tdist, means=means, density=density, enc_out=enc_out, weights=weights,
                          arena=arena, flat=mlp._flatten(p), desc=geometry._mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals))
        g_w = new(R, lv["n"])
        _lib.call("nrc_mask_loss", st(), _lib.ptr(lv["weights"]), lv["n"], self._zero_mask(R, dev), R,
                  float(self.charb_padding), 0.0, float(self.backward_mask_weight), _lib.ptr(loss), _lib.ptr(g_w))
        if grads_ready is not None:     # the gradient arena's memset (side stream) must be done before the first scatter
            torch.cuda.current_stream().wait_event(grads_ready)
        self._level_backward(lv, rays, g_w, None, None, R)
        self.last_extra = lv

    def _zero_mask(self, R, dev):
        key = ("m0", R, str(dev))
        if key not in self._bg:
            self._bg[key] = torch.zeros((R,), device=dev, dtype=torch.float32)
        return _lib.ptr(self._bg[key])

    def step_front(self, rays, u01, target_rgb, train_frac=1.0, fork_proposals=False, extra=None, zero_grad=None,
                   on_shader_grads=None, on_proposal_grads=None, on_final_grads=None, on_grid_grads=None):
        """Forward, loss and the SHADER's backward: when this returns (in stream order) every gradient of the
        `Shader` parameters (appearance grid + all stacks) is final, so a data-parallel harness can start
        all-reducing that half of the gradient arena while step_back() produces the sampler's half."""
        sampler, shader = self.model.sampler, self.model.shader
        sp = self.params["Sampler"]
        near = rays["near"]
        R, dev = near.shape[0], near.device
        st = _lib.stream_ptr
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        anneal = sampler.anneal(train_frac)
        main = torch.cuda.current_stream()
        self._constants(R, dev)
        s_pack, s_enc, s_env, s_prop = self._streams(6)[:4] if self.concurrent else (None,) * 4
        shp = self.params["Shader"]
        names, sflat = shader.fused_params(shp)
        app_arena = shp["appearance_grid"]["_arena"]
        loss = torch.zeros((), device=dev, dtype=torch.float32)
        reg_on = self.density_grid_regularizer is not None and self.geometry_mults is not None

        def regularize(init=False):   # param_regularizer_loss on the three density grids: atomic, order-free;
            if not reg_on:            # init=True: plain stores that double as the zero-fill of these gradient tables
                return
            for i_mlp, mlp in enumerate(sampler.mlps):
                arena = sp[f"MLP_{i_mlp}"]["density_grid"]["_arena"]
                t_sink = _lib.grad_sink(arena)
                enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(arena)), mlp.grid.tables(mlp.grid.views(t_sink)))
                _lib.call("nrc_grid_regularizer_init" if init else "nrc_grid_regularizer", _lib.stream_ptr(), C.byref(enc),
                          float(self.density_grid_regularizer), _lib.ptr(loss))

        # Gradient-arena memset and the parameter regularizer: nothing reads or writes gradients before the first backward
        # launch, so they run on the packing stream beside the sampler's forward; every gradient-writing stream waits on
        # `grads_ready`.  When the caller hands over the memset (zero_grad), the regularizer INITIALISES the density grids'
        # gradient tables (its gradient is dense over them) and zero_grad(skip_density_grids=True) clears only the rest.
        grads_ready = None
        if zero_grad is not None:
            if s_pack is not None:
                s_pack.wait_stream(main)
                with torch.cuda.stream(s_pack):
                    zero_grad(skip_density_grids=reg_on)
                    regularize(init=True)
                    grads_ready = torch.cuda.Event()
                    grads_ready.record()
            else:
                zero_grad(skip_density_grids=reg_on)
                regularize(init=True)
        # backward-mask pass: independent rays, its own stream (joined in step_back); in split mode
        # (fork_proposals False: two graphs) it is issued by step_back instead
        s_x = None
        if extra is not None and fork_proposals:
            if self.concurrent:
                s_x = self._streams(6)[4]
                s_x.wait_stream(main)
                with torch.cuda.stream(s_x):
                    self._weights_only_pass(extra[0], extra[1], train_frac, loss, grads_ready)
            else:
                self._weights_only_pass(extra[0], extra[1], train_frac, loss)
            extra = None
        # weight packing does not depend on the rays: it runs beside the sampler (with the regularizer, when the caller
        # cleared the gradients itself)
        if s_pack is not None:
            s_pack.wait_stream(main)
            with torch.cuda.stream(s_pack):
                packed = nerf.shader_pack(shader, names, sflat)
                if zero_grad is None:
                    regularize()
        else:
            packed = nerf.shader_pack(shader, names, sflat)
            if zero_grad is None:
                regularize()
        # ------------------------------------------------------------------ forward: proposal sampler
        sdist, weights = self._initial_step_function(R, dev)
        levels = []
        prev, fuse_w = None, os.environ.get("NRC_FUSE_WEIGHTS", "1") == "1"
        nl = len(sampler.sampling_strategy)
        for i_level, (i_mlp, _, n) in enumerate(sampler.sampling_strategy):
            mlp, p = sampler.mlps[i_mlp], sp[f"MLP_{i_mlp}"]
            last = i_level == nl - 1
            # the previous level's alpha-compositing weights are computed in the head of this level's resampling launch
            sdist, tdist, means = sampler.sample_and_cast(u01[i_level], sdist, weights, n, anneal, rays, False, prev=prev)
            prev = None
            P = R * n
            if last:   # the appearance-grid gather only needs the final sample positions
                if s_enc is not None:
                    s_enc.wait_stream(main)
                    with torch.cuda.stream(s_enc):
                        encoded = nerf.shader_encode(shader, means, app_arena)
                else:
                    encoded = nerf.shader_encode(shader, means, app_arena)
            density, enc_out = new(P), new(P, mlp.in_dim)
            feat = new(P, 64) if last else None
            gp = new(P, 3) if mlp.enable_pred_normals else None
            want_normals = (not mlp.normals_for_filter_only) and not mlp.disable_density_normals
            # The analytic normals (d raw / d means: the query kernel's in-kernel back-propagation + corner re-gather,
            # half of the final level's launch) are consumed by the geometry branch only: there the query is issued a
            # second time for that output alone, beside the shader, and the launch on the critical path stays
            # forward-only.  Measured (profiles/r02_ab_runs.txt): 0.692 vs 0.681 ms - the step is bound by total kernel work, the extra
            # forward costs more than the shorter critical path gains - so it is OFF by default (NRC_SPLIT_NORMALS=1).
            defer_rg = (want_normals and last and self.geometry_mults is not None and self.concurrent
                        and os.environ.get("NRC_SPLIT_NORMALS", "0") == "1")
            rg = new(P, 3) if (want_normals and not defer_rg) else None
            arena = p["density_grid"]["_arena"]
            enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(arena)), None)
            flat = mlp._flatten(p)
            desc = geometry._mlp_desc(p, mlp.in_dim, mlp.enable_pred_normals)
            _lib.call("nrc_density_query_fwd", st(), C.byref(enc), C.byref(desc), _lib.ptr(means), P, float(mlp.warp_c),
                      float(mlp.density_bias), int(mlp.bf16), _lib.ptr(density), None, _lib.ptr(feat), _lib.ptr(gp),
                      _lib.ptr(rg), _lib.ptr(enc_out))
            weights = new(R, n)
            if last or not fuse_w:
                _lib.call("nrc_ray_alpha_weights_fwd", st(), _lib.ptr(density), _lib.ptr(tdist), _lib.ptr(rays["directions"]),
                          R, n, int(sampler.opaque_background), _lib.ptr(weights), None, None)
            else:
                prev = (density, tdist)      # `weights` is filled by the next level's nrc_ray_weights_sample_cast
            levels.append(dict(mlp=mlp, p=p, n=n, sdist=sdist, tdist=tdist, means=means, density=density, enc_out=enc_out,
                               feat=feat, gp=gp, rg=rg, weights=weights, arena=arena, flat=flat, desc=desc, enc=enc,
                               defer_rg=defer_rg))
        L2 = levels[-1]
        P2 = R * L2["n"]
        normals_pred = new(P2, 3)
        _lib.call("nrc_normals_fwd", st(), _lib.ptr(L2["gp"]), P2, _lib.ptr(normals_pred))
        if L2["rg"] is not None:   # analytic normals: computed like the reference, consumed by the 8f losses
            L2["normals"] = new(P2, 3)
            _lib.call("nrc_normals_fwd", st(), _lib.ptr(L2["rg"]), P2, _lib.ptr(L2["normals"]))
        # ------------------------------------------------------------------ geometry branch (side stream)
        # distortion / orientation / predicted-normal losses need only the final level's step function and normals:
        # they, the l2_normalize VJP and the second-order kernel (nrc_density_normals_bwd, the longest kernel of the
        # step) start here and run beside the shader; their g_w / g_normals_pred terms are added after the shader's.
        k = L2["n"]
        if grads_ready is not None:     # long done by now (memset ~20 us vs the sampler's forward); orders every later fork
            main.wait_event(grads_ready)
        geo, s_geo = None, None
        if self.geometry_mults is not None:
            def geometry_branch():
                if L2["defer_rg"]:
                    mlp2 = L2["mlp"]
                    L2["rg"] = new(P2, 3)
                    _lib.call("nrc_density_query_fwd", _lib.stream_ptr(), C.byref(L2["enc"]), C.byref(L2["desc"]),
                              _lib.ptr(L2["means"]), P2, float(mlp2.warp_c), float(mlp2.density_bias), int(mlp2.bf16),
                              None, None, None, None, _lib.ptr(L2["rg"]), None)
                    L2["normals"] = new(P2, 3)
                    _lib.call("nrc_normals_fwd", _lib.stream_ptr(), _lib.ptr(L2["rg"]), P2, _lib.ptr(L2["normals"]))
                gw_geo = torch.zeros((R, k), device=dev, dtype=torch.float32)
                gnp_geo = torch.zeros((P2, 3), device=dev, dtype=torch.float32)
                if self.distortion is not None:
                    dm, dp, dpre = self.distortion
                    _lib.call("nrc_distortion_loss", _lib.stream_ptr(), _lib.ptr(L2["tdist"]), _lib.ptr(L2["weights"]), k,
                              R, float(dp), float(dpre), float(dm), _lib.ptr(loss), _lib.ptr(gw_geo))
                has_n = L2.get("normals") is not None
                g_na = new(P2, 3) if has_n else None
                mo, mp, mr = self.geometry_mults
                _lib.call("nrc_geometry_losses", _lib.stream_ptr(), _lib.ptr(L2["weights"]),
                          _lib.ptr(L2["normals"]) if has_n else None, _lib.ptr(normals_pred), _lib.ptr(rays["viewdirs"]),
                          R, k, float(mo), float(mp), float(mr), float(self.sg_w), _lib.ptr(loss), _lib.ptr(gw_geo),
                          _lib.ptr(gnp_geo), _lib.ptr(g_na))
                if has_n:   # second-order path of the predicted-normal loss: d/d theta <g_rg, d raw / d means>
                    g_rg = new(P2, 3)
                    _lib.call("nrc_normals_bwd", _lib.stream_ptr(), _lib.ptr(L2["rg"]), _lib.ptr(g_na), P2, _lib.ptr(g_rg))
                    mlp2 = L2["mlp"]
                    sinks = [_lib.grad_sink(t) for t in L2["flat"]]
                    geometry.density_normals_bwd(mlp2, L2["p"], L2["arena"], L2["means"].reshape(P2, 3), g_rg,
                                                 mlp2._unflatten(sinks), _lib.grad_sink(L2["arena"]),
                                                 enc_out=L2["enc_out"])
                    return gw_geo, gnp_geo, g_na, g_rg
                return gw_geo, gnp_geo, g_na, None
            if self.concurrent:
                s_geo = self._streams(6)[5]
                s_geo.wait_stream(main)
                with torch.cuda.stream(s_geo):
                    geo = geometry_branch()
            else:
                geo = geometry_branch()
        # ------------------------------------------------------------------ proposal supervision (side stream)
        # The spline interlevel loss and the proposal levels' backward depend only on the final level's step
        # function (sdist, weights): they start here and run beside the shader's forward / backward.
        g_w = [new(R, lv["n"]) for lv in levels]

        def interlevel():   # loss_utils.spline_interlevel_loss, one launch per proposal level
            for i_level in range(nl - 1):
                lv = levels[i_level]
                _lib.call("nrc_interlevel_loss", _lib.stream_ptr(), _lib.ptr(L2["sdist"]), _lib.ptr(L2["weights"]), k,
                          _lib.ptr(lv["sdist"]), _lib.ptr(lv["weights"]), lv["n"], R, float(self.interlevel_blurs[i_level]),
                          float(self.interlevel_mults[i_level]), 1e-5, _lib.ptr(loss), _lib.ptr(g_w[i_level]), None)

        if s_prop is not None:
            s_prop.wait_stream(main)
            with torch.cuda.stream(s_prop):
                interlevel()
                if fork_proposals:
                    for i_level in range(nl - 2, -1, -1):
                        self._level_backward(levels[i_level], rays, g_w[i_level], None, None, R)
                    if on_proposal_grads is not None:
                        on_proposal_grads()
        else:
            interlevel()
        # ------------------------------------------------------------------ forward: shader + integrator + loss
        if self.concurrent:
            main.wait_stream(s_pack)
            main.wait_stream(s_enc)
        fuse_out = os.environ.get("NRC_FUSE_OUT", "1") == "1"
        outs, saved, meta = nerf.shader_fused_forward(
            shader, names, sflat, rays["viewdirs"], L2["means"], L2["feat"].reshape(R, L2["n"], 64),
            normals_pred.reshape(R, L2["n"], 3), app_arena, True, packed=packed, encoded=encoded, env_stream=s_env,
            want_bottleneck=False, defer_out=fuse_out)
        bg = self._bg_ones(R, dev)
        out_rgb, acc, dist = new(R, 3), new(R), None
        g_rgb = g_acc = gv = g_out = None
        mw = self.mask_weights
        if fuse_out:
            # `out` stage, volumetric rendering (rgb, acc), data term, mask loss, the compositing VJP and the `out`
            # stage's VJP: one launch, one warp per ray (nrc_shade_render_loss)
            heads, fbuf, sbuf = saved[3], saved[4], saved[5]
            ebuf, (rgb_max, d_bias, l_bias, b_bias) = saved[10]
            rgb_s = new(R, k, 3)
            g_out = (new(R * k, 16), new(R * k, 16), new(R * k, 16))
            _lib.call("nrc_shade_render_loss", st(), _lib.ptr(heads), heads.shape[1], _lib.ptr(fbuf), fbuf.shape[1],
                      _lib.ptr(sbuf), sbuf.shape[1], _lib.ptr(ebuf), ebuf.shape[1], rgb_max, d_bias, l_bias, b_bias,
                      _lib.ptr(L2["weights"]), _lib.ptr(bg), _lib.ptr(target_rgb), None, R, k, float(self.charb_padding),
                      0 if mw is None else 1, 0.0 if mw is None else float(mw[0]), 0.0 if mw is None else float(mw[1]),
                      _lib.ptr(loss), _lib.ptr(rgb_s), _lib.ptr(out_rgb), _lib.ptr(acc), _lib.ptr(g_w[-1]),
                      _lib.ptr(g_out[0]), 16, _lib.ptr(g_out[1]), 16, _lib.ptr(g_out[2]), 16)
        else:
            rgb_s = outs[0].reshape(R, L2["n"], 3)
            # volumetric rendering (rgb, acc), data term, mask loss and the compositing VJP: one launch
            gv = new(R, k, 3)
            _lib.call("nrc_render_loss", st(), _lib.ptr(rgb_s), _lib.ptr(L2["weights"]), _lib.ptr(bg), _lib.ptr(target_rgb), None,
                      R, k, float(self.charb_padding), 0 if mw is None else 1, 0.0 if mw is None else float(mw[0]),
                      0.0 if mw is None else float(mw[1]), _lib.ptr(loss), _lib.ptr(out_rgb), _lib.ptr(acc), _lib.ptr(gv),
                      _lib.ptr(g_w[-1]))
        # The final level's backward needs only the data gradients that leave the shader (d_feat, g_nrm): in the
        # single-graph mode it forks off as soon as the trunk's data-gradient chain is done and runs beside the
        # weight-gradient launch and the appearance-grid scatter instead of behind them (NRC_TAIL_FORK=0: old order).
        tail = {}

        def join_geometry(g_nrm):
            if geo is not None:
                if s_geo is not None:
                    torch.cuda.current_stream().wait_stream(s_geo)
                g_w[-1].add_(geo[0])
                g_nrm.add_(geo[1].view_as(g_nrm))

        def fork_tail(d_feat, g_nrm):
            s_tail = self._streams(7)[6]
            s_tail.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_tail):
                join_geometry(g_nrm)
                self._final_level_backward(L2, rays, g_w[-1], d_feat, g_nrm, R)
                if on_final_grads is not None:
                    on_final_grads()
                    tail["final_grads_announced"] = True
            tail["stream"] = s_tail

        use_tail = fork_proposals and self.concurrent and os.environ.get("NRC_TAIL_FORK", "1") == "1"
        d_feat, g_nrm, _, _, _ = nerf.shader_fused_backward(shader, names, sflat, saved, meta, app_arena, gv, True,
                                                            on_data_grads=fork_tail if use_tail else None,
                                                            on_grid_grads=on_grid_grads, g_out=g_out)
        if on_shader_grads is not None:
            on_shader_grads()
        if not use_tail:
            join_geometry(g_nrm)
        if s_prop is not None and not fork_proposals:
            main.wait_stream(s_prop)     # split mode: the side stream only ran the interlevel losses
        self.last = dict(levels=levels, rgb=out_rgb, acc=acc, dist=dist, shader_rgb=rgb_s)
        return dict(loss=loss, levels=levels, rays=rays, g_w=g_w, d_feat=d_feat, g_nrm=g_nrm, R=R,
                    forked=s_prop if fork_proposals else None, keep=(saved, gv, g_rgb, g_acc, geo), extra=extra,
                    extra_stream=s_x, train_frac=train_frac, tail_stream=tail.get("stream"),
                    final_grads_announced=tail.get("final_grads_announced", False))

    def step_back(self, state):
        """Backward of the proposal sampler (three levels) from the state of step_front()."""
        levels, rays, g_w, R = state["levels"], state["rays"], state["g_w"], state["R"]
        nl = len(levels)
        L2 = levels[-1]
        P2 = R * L2["n"]
        dev = L2["density"].device
        main = torch.cuda.current_stream()
        s_prop = state["forked"]
        own_fork = None
        if s_prop is None and self.concurrent:      # split mode: fork the proposal levels here
            own_fork = self._streams(4)[3]
            own_fork.wait_stream(main)
            with torch.cuda.stream(own_fork):
                for i_level in range(nl - 2, -1, -1):
                    self._level_backward(levels[i_level], rays, g_w[i_level], None, None, R)
        x_fork = None
        if state.get("extra") is not None:      # split mode: the backward-mask pass starts here
            if self.concurrent:
                x_fork = self._streams(6)[4]
                x_fork.wait_stream(main)
                with torch.cuda.stream(x_fork):
                    self._weights_only_pass(state["extra"][0], state["extra"][1], state["train_frac"], state["loss"])
            else:
                self._weights_only_pass(state["extra"][0], state["extra"][1], state["train_frac"], state["loss"])
        elif state.get("extra_stream") is not None:
            x_fork = state["extra_stream"]
        if state.get("tail_stream") is not None:    # already issued beside the shader's weight gradients (step_front)
            main.wait_stream(state["tail_stream"])
        else:
            self._final_level_backward(L2, rays, g_w[nl - 1], state["d_feat"], state["g_nrm"], R)
        if x_fork is not None:
            main.wait_stream(x_fork)
        if s_prop is not None:
            main.wait_stream(s_prop)
        elif own_fork is not None:
            main.wait_stream(own_fork)
        else:
            for i_level in range(nl - 2, -1, -1):
                self._level_backward(levels[i_level], rays, g_w[i_level], None, None, R)

    def _final_level_backward(self, L2, rays, g_weights, d_feat, g_nrm, R):
        """Predicted-normal VJP + the final sampler level's backward (density features and normals from the shader)."""
        P2 = R * L2["n"]
        g_gp = torch.empty((P2, 3), device=L2["density"].device, dtype=torch.float32)
        _lib.call("nrc_normals_bwd", _lib.stream_ptr(), _lib.ptr(L2["gp"]), _lib.ptr(g_nrm), P2, _lib.ptr(g_gp))
        self._level_backward(L2, rays, g_weights, d_feat, g_gp if L2["gp"] is not None else None, R)

    def _level_backward(self, lv, rays, g_weights, g_feat, g_gp, R):
        """alpha-weights VJP -> fused density-MLP VJP -> hash-grid scatter of one sampler level."""
        mlp, n = lv["mlp"], lv["n"]
        P = R * n
        dev = lv["density"].device
        st = _lib.stream_ptr
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        g_density = new(P)
        _lib.call("nrc_ray_alpha_weights_bwd", st(), _lib.ptr(lv["density"]), _lib.ptr(lv["tdist"]),
                  _lib.ptr(rays["directions"]), _lib.ptr(g_weights), None, None, R, n, _lib.ptr(g_density))
        sinks = [_lib.grad_sink(t) for t in lv["flat"]]
        t_sink = _lib.grad_sink(lv["arena"])
        if t_sink is None or any(s is None for s in sinks):
            raise _lib.NrcError("FusedCacheStep needs registered gradient sinks for every parameter")
        gd = geometry._grad_desc(mlp, mlp._unflatten(sinks))
        g_enc = new(P, mlp.in_dim)
        _lib.call("nrc_density_mlp_bwd", st(), C.byref(lv["desc"]), _lib.ptr(lv["enc_out"]), _lib.ptr(g_density),
                  _lib.ptr(lv["density"]), _lib.ptr(g_feat), _lib.ptr(g_gp), P, int(mlp.bf16), _lib.ptr(g_enc), C.byref(gd))
        enc = mlp.grid._descriptor(mlp.grid.tables(mlp.grid.views(lv["arena"])), mlp.grid.tables(mlp.grid.views(t_sink)))
        _lib.call("nrc_encode_bwd_warped", st(), C.byref(enc), _lib.ptr(lv["means"].reshape(P, 3)), float(mlp.warp_c),
                  _lib.ptr(g_enc), P)


class FusedCacheQuery:
    """Forward-only static schedule of one radiance-cache query (render path; BASELINE configs 1, 3, 5):
    proposal sampler -> [categorical resample] -> cache shader -> volumetric rendering, as ~25 launches with
    no elementwise glue.  It is the body of the material stage's `radiance_cache_fn`
    (internal/material.py:2174-2231 -> models.py:656-774 with is_secondary=True, resample=True) and of the
    primary-ray cache render.  Same kernels as models.NeRFModel.__call__, pinned against it by
    tests/test_engine_gpu.py.  Quantities nobody downstream reads on this path (analytic normals, rectified
    normals, per-level alpha / transmittance) are not computed -- the reference's XLA program dead-code
    eliminates them too."""

    def __init__(self, model):
        self.model = model
        self._const = {}
        self._pack_cache = mlp_chain.PackCache()
        self._density_packs = {}
        # NRC_QUERY_TC=1: density queries as tcgen05 chains with the hash-grid gather as their front end
        # (nrc_chain_query).  Measured 15-25 % slower than the mma.sync query kernel (the gather warps idle while
        # their tile's three accumulator round trips complete), so the default stays the mma.sync kernel.
        self.tensor_core_query = os.environ.get("NRC_QUERY_TC", "0") == "1"

    def _initial(self, R, dev):
        key = (R, str(dev))
        if key not in self._const:
            sd = torch.zeros((R, 2), device=dev, dtype=torch.float32)
            sd[:, 1] = 1.0
            self._const[key] = (sd, torch.ones((R, 1), device=dev, dtype=torch.float32),
                                torch.ones((R, 3), device=dev, dtype=torch.float32),
                                torch.zeros((R, 3), device=dev, dtype=torch.float32))
        return self._const[key]

    def __call__(self, params, rays, u01, gumbel=None, is_secondary=False, resample=False, train_frac=1.0):
        """rays: dict of contiguous [R,·] tensors; u01: 3 x [R,1]; gumbel [R,n_last,k] when resample.
        Returns dict(rgb [R,3], acc [R], distance [R,4], extras [R,k|n,22], weights, means, normals, inds)."""
        sampler, shader = self.model.sampler, self.model.shader
        sp, shp = params["Sampler"], params["Shader"]
        R, dev = rays["near"].shape[0], rays["near"].device
        st = _lib.stream_ptr
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        anneal = sampler.anneal(train_frac)
        sdist, weights, bg_one, bg_zero = self._initial(R, dev)
        prev, fuse_w = None, os.environ.get("NRC_FUSE_WEIGHTS", "1") == "1"
        nl = len(sampler.sampling_strategy)
        for i_level, (i_mlp, _, n) in enumerate(sampler.sampling_strategy):
            mlp, p = sampler.mlps[i_mlp], sp[f"MLP_{i_mlp}"]
            last = i_level == nl - 1
            sdist, tdist, means = sampler.sample_and_cast(u01[i_level], sdist, weights, n, anneal, rays, is_secondary, prev=prev)
            prev = None
            P = R * n
            density = new(P)
            feat = new(P, 64) if last else None
            gp = new(P, 3) if (last and mlp.enable_pred_normals) else None
            if self.tensor_core_query and mlp.bf16 and mlp.supports_query_tc():
                mlp.query_tc(p, means, density, feat, gp, cache=self._density_packs.setdefault(i_mlp, mlp_chain.PackCache()))
            else:
                enc = mlp.grid._descriptor(mlp.grid.tables(p["density_grid"]), None)
                desc = geometry._mlp_desc(p, mlp.in_dim, gp is not None)
                _lib.call("nrc_density_query_fwd", st(), C.byref(enc), C.byref(desc), _lib.ptr(means), P,
                          float(mlp.warp_c), float(mlp.density_bias), int(mlp.bf16), _lib.ptr(density), None,
                          _lib.ptr(feat), _lib.ptr(gp), None, None)
            weights = new(R, n)
            if last or not fuse_w:
                _lib.call("nrc_ray_alpha_weights_fwd", st(), _lib.ptr(density), _lib.ptr(tdist), _lib.ptr(rays["directions"]),
                          R, n, int(sampler.opaque_background), _lib.ptr(weights), None, None)
            else:
                prev = (density, tdist)      # the next level's resampling launch computes (and stores) these weights
        n_last = n
        inds = None
        if resample:
            k = gumbel.shape[-1]
            inds = torch.empty((R, k), device=dev, dtype=torch.int32)
            w_sh = new(R, k)
            _lib.call("nrc_ray_resample", st(), _lib.ptr(weights), _lib.ptr(gumbel), R, n_last, k,
                      float(self.model.weights_bias), 1.0, _lib.ptr(inds), _lib.ptr(w_sh))

            def take(field, c):
                out = new(R, k, c)
                _lib.call("nrc_ray_resample_gather", st(), _lib.ptr(field), _lib.ptr(inds), R, n_last, k, c, _lib.ptr(out))
                return out

            means_sh, feat_sh, gp_sh = take(means, 3), take(feat, 64), take(gp, 3)
        else:
            k, w_sh, means_sh, feat_sh, gp_sh = n_last, weights, means, feat.reshape(R, n_last, 64), gp
        Psh = R * k
        normals = new(Psh, 3)
        _lib.call("nrc_normals_fwd", st(), _lib.ptr(gp_sh), Psh, _lib.ptr(normals))
        names, sflat = shader.fused_params(shp)
        packed = self._pack_cache.get(sflat[0::2], lambda: nerf.shader_pack(shader, names, sflat))
        outs, _, _ = nerf.shader_fused_forward(shader, names, sflat, rays["viewdirs"], means_sh.reshape(R, k, 3),
                                               feat_sh.reshape(R, k, 64), normals.reshape(R, k, 3),
                                               shp["appearance_grid"]["_arena"], False, packed=packed, want_bottleneck=False)
        rgb_s = outs[0].reshape(R, k, 3)
        out_rgb, acc, dist = new(R, 3), new(R), new(R, 4)
        _lib.call("nrc_ray_composite_fwd", st(), _lib.ptr(rgb_s), _lib.ptr(w_sh), k, _lib.ptr(weights) if resample else None,
                  _lib.ptr(tdist), _lib.ptr(bg_zero if is_secondary else bg_one), R, n_last, 3, 1, _lib.ptr(out_rgb),
                  _lib.ptr(acc), _lib.ptr(dist))
        return dict(rgb=out_rgb, acc=acc, distance=dist, extras=outs[1], weights=w_sh, weights_all=weights,
                    means=means_sh.reshape(R, k, 3), normals=normals.reshape(R, k, 3), inds=inds, tdist=tdist)
