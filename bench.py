#!/usr/bin/env python
"""Benchmark of the radiance-cache query path (contract: see the task statement / DESIGN.md 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--rays R]

b200 arm   : one step = one pass of the cache-stage hot path (forward + backward) over one
             batch of synthetic rays, captured in a CUDA graph.  `value` = cache samples/s
             with inputs resident in HBM; `e2e` = the same through the public host API with a
             pinned host->device copy of the ray batch and a device->host read of the loss
             inside the timed region.
reference  : the reference's own algorithm (restated CPU oracle, PyTorch-CPU fp32 -- JAX is
             not installable in this image) on the host cores, same config and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLES_PER_RAY = 160  # 64 + 64 + 32 (configs/nerf_ngp_yobo.gin:521-545)
METRIC = "cache_samples_per_sec"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=1024, help="rays per GPU per step (config 2: 1024)")
    ap.add_argument("--bf16", type=int, default=1,
                    help="1 (default): bf16 tensor-core MLPs (tcgen05), north-star tolerance 2e-2; 0: fp32 parity variant")
    ap.add_argument("--cpu-rays", type=int, default=1024, help="rays per step of the CPU arm (config 2: 1024, the same batch)")
    ap.add_argument("--no-others", action="store_true", help="skip the `others` table (configs 1, 3, 5, fp32, 16384 rays)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true", help="print per-ABI-call device times")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"],
                    help="config2 (default, BASELINE metric): cache training step; config3: material-stage chunk "
                         "(1024 points x 32 secondary rays); config4: time-resolved cache render chunk (700 bins); config5: "
                         "full-view render in row bands")
    ap.add_argument("--image", type=int, default=800, help="config5: image side in pixels")
    ap.add_argument("--ncu-mode", action="store_true",
                    help="minimal run for an ncu launch list: 1 eager step, graph capture, 2 replays, no JSON")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference_throughput(rays, steps, warmup, threads=None):
    """The reference's algorithm (oracle port) for the same step: full cache training step."""
    from oracle import loss_utils as oloss, models as omodels
    from neural_radiance_caching_b200 import workload

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    g = np.random.Generator(np.random.PCG64(workload.SEED))
    model = omodels.NeRFModel()
    params = model.init(g, table_init_range=0.1)
    leaves = []

    def collect(p):
        for k, v in p.items():
            if isinstance(v, dict):
                collect(v)
            else:
                v.requires_grad_(True)
                leaves.append(v)

    collect(params)
    rn = workload.make_rays_np(g, rays)
    rt = {k: torch.from_numpy(v) for k, v in rn.items()}
    u = [torch.from_numpy(g.uniform(size=(rays, 1)).astype(np.float32)) for _ in range(3)]
    target = torch.from_numpy(g.uniform(size=(rays, 3)).astype(np.float32))
    xt = {k: torch.from_numpy(v) for k, v in workload.backward_mask_rays_np(g, rn).items()}
    ux = [torch.from_numpy(g.uniform(size=(rays, 1)).astype(np.float32)) for _ in range(3)]

    def step():
        # the same objective as the CUDA arm: data + interlevel + geometry losses (analytic normals with
        # create_graph: second-order path) + mask + the backward-mask weights_only pass on the extra rays
        for t in leaves:
            t.grad = None
        res = model(params, rt, u, create_graph=True)
        extra_acc = model.weights_only(params, xt, ux)
        loss = workload.cache_loss(res, target, rays=rt, extra_acc=extra_acc, lib=oloss,
                                   reg_tables=workload.density_grid_tables(params))
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return rays * SAMPLES_PER_RAY / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    val, dt, threads = cpu_reference_throughput(args.cpu_rays, steps, warm)
    sample = f"{args.cpu_rays} rays x 160 samples per step, {steps} steps after {warm} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": args.cpu_rays,
                   "note": "restated reference (oracle), PyTorch-CPU fp32 -- not JAX/XLA (not installable here)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_sec": val / SAMPLES_PER_RAY,
    }
    print(json.dumps(line))


WORKLOAD = ("config2 nerf_ngp_yobo_lego cache training step: proposal sampler (64,64,32) hash-grid + density MLP, "
            "cache shader (appearance grid + bottleneck/heads/int-BRDF/IDE SurfaceLightField/EnvMap MLPs) on the "
            "32 final samples, volumetric rendering; losses: Charbonnier-sRGB data, spline interlevel on both proposal "
            "levels, distortion, orientation + predicted-normal + reverse (analytic normals, second-order path), mask, "
            "density-grid parameter regularizer, and the "
            "backward-mask weights_only pass on one extra ray per training ray; fwd+bwd (grads for 4 grids + all MLPs)")


# ----------------------------------------------------------------------------- b200 arm
def run_b200_main(args):
    import torch.distributed as dist

    from neural_radiance_caching_b200 import _lib, dist as ndist, workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    pk, pk_kind = peaks()

    R = args.rays
    step_obj = workload.CacheTrainStep(dev, bf16=bool(args.bf16))
    g = np.random.Generator(np.random.PCG64(workload.SEED + rank))
    n_batches = 4
    host = []
    for _ in range(n_batches):
        rn = workload.make_rays_np(g, R)
        u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
        tgt = g.uniform(size=(R, 3)).astype(np.float32)
        xr = workload.backward_mask_rays_np(g, rn)     # extra rays of the backward-mask loss + their jitter draws
        ux = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
        host.append(torch.from_numpy(workload.pack_batch(rn, u, tgt, xr, ux)).pin_memory())
    dbuf = torch.empty_like(host[0], device=dev)
    dbuf.copy_(host[0])
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    grads_flat = None

    # NRC_DP_INGRAPH=1 (default with the peer-memory arena): both buckets' all-reduces are part of the step and of its
    # CUDA graph, the Shader bucket beside the sampler's backward (workload.CacheTrainStep.step).
    ingraph = (world > 1 and getattr(step_obj, "peer", None) is not None and step_obj.engine is not None
               and os.environ.get("NRC_DP_INGRAPH", "1") == "1" and os.environ.get("NRC_DP_OVERLAP", "0") != "1")

    def allreduce_grads():
        # gradient all-reduce (mean) over all tables + MLP weights, the reference's lax.pmean
        # (internal/train_utils.py:3132-3136): one collective over the flat gradient arena.
        if not ingraph:
            step_obj.allreduce_grads()

    def compute_step():
        rays, u01, target, extra = workload.unpack_batch(dbuf)
        return step_obj.step(rays, u01, target, extra, fused_allreduce=ingraph)

    def one_step():
        loss = compute_step()
        allreduce_grads()
        return loss

    # eager warm-up on a side stream (also counts kernel launches per step)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(1 if args.ncu_mode else 3):
            before = _lib.launch_count
            loss = one_step()
            launches = _lib.launch_count - before
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()

    if args.ncu_mode:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            compute_step()
        torch.cuda.synchronize()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        print(json.dumps({"ncu_mode": True, "launches_per_step": launches}))
        return False, None

    per_kernel = profile_calls(one_step, _lib)

    graph = torch.cuda.CUDAGraph()
    use_graph = True
    # Data parallel: ONE graph + ONE peer-memory all-reduce of the whole arena after it (default), or
    # NRC_DP_OVERLAP=1: two graphs with the shader bucket's all-reduce overlapping the sampler's backward.  Measured at
    # N = 2 (profiles/r01j_ab_runs.txt): the split loses more inside the step (proposal supervision no longer beside the shader,
    # the collective's CTAs competing with the backward kernels) than the overlap hides.
    overlap = world > 1 and step_obj.engine is not None and os.environ.get("NRC_DP_OVERLAP", "0") == "1"
    if not overlap:
        with torch.cuda.graph(graph):
            static_loss = compute_step()

        def run_step():
            graph.replay()
            allreduce_grads()
            return static_loss
    else:
        # Data parallel: the step is captured as TWO graphs.  When the first one (forward, loss, shader backward)
        # has run, the Shader half of the gradient arena is final: its all-reduce starts on a communication
        # stream and overlaps the second graph (the sampler's backward); the Sampler half follows.
        graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            rays_, u01_, target_, extra_ = workload.unpack_batch(dbuf)
            state = step_obj.step_front(rays_, u01_, target_, extra_)
        with torch.cuda.graph(graph_b, pool=graph.pool()):
            static_loss = step_obj.step_back(state)
        comm = torch.cuda.Stream()
        so = step_obj.shader_offset

        def run_step():
            cur = torch.cuda.current_stream()
            graph.replay()
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                step_obj.allreduce_grads(so, None, channel=1)
            graph_b.replay()
            step_obj.allreduce_grads(0, so, channel=0)
            cur.wait_stream(comm)
            return static_loss

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) -------------------------------------------------
    for _ in range(max(3, args.warmup)):
        run_step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    evs = []
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the timed events)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run_step()
        e.record()
        evs.append((s, e))
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(s.elapsed_time(e) for s, e in evs) / args.steps

    # ---- end-to-end timing (e2e): pinned H2D of the ray batch + step + D2H of the loss ----
    for i in range(2):
        dbuf.copy_(host[i % n_batches], non_blocking=True)
        loss_host.copy_(run_step(), non_blocking=True)
    barrier()
    evs2 = []
    for i in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        dbuf.copy_(host[i % n_batches], non_blocking=True)
        loss_host.copy_(run_step(), non_blocking=True)
        e.record()
        evs2.append((s, e))
    barrier()
    e2e_ms = sum(s.elapsed_time(e) for s, e in evs2) / args.steps
    clk = clocks.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_samples = world * R * SAMPLES_PER_RAY
    value = total_samples / (dev_ms * 1e-3)
    e2e_value = total_samples / (e2e_ms * 1e-3)

    if rank != 0:
        return None, (step_obj, graph)

    roof = roofline(per_kernel, R, pk, pk_kind, bool(args.bf16), l2_gather_probe(dev) if per_kernel else None)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt, threads = cpu_reference_throughput(args.cpu_rays, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_rays} rays x 160 samples per step, 3 steps after 1 warm-up "
                         "(restated reference, PyTorch-CPU fp32; JAX not installable here)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.bf16 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu_per_step": R, "samples_per_ray": [64, 64, 32],
                   "shaded_points_per_ray": 32,
                   "backward_mask_rays_per_step": R,   # 160 more density samples each; NOT counted in `value`
                   "precision": ("MLP operands bf16 on tcgen05 tensor cores with fp32 accumulation (north-star bf16-MLP "
                                 "variant, tolerance 2e-2); hash grids, gathers, scatters and ray kernels fp32")
                   if args.bf16 else "fp32 parity variant (1e-5)",
                   "tables": "MLP_0/1/2 density grids + appearance grid (7.5+9.6+46.7+46.7 MB fp32), "
                             "U(+-0.1) trained-like init",
                   "l2": "flushed between timed steps (256 MiB memset outside the event pairs)",
                   "cuda_graph": bool(use_graph),
                   "parallelism": f"dp{world} rays, params replicated" + (
                       "; gradient all-reduce in 2 buckets, the shader bucket overlapping the sampler's backward"
                       if overlap else ""),
                   "allreduce": (step_obj.allreduce_kind + ("; both buckets inside the step's CUDA graph, the shader bucket "
                                 "beside the sampler's backward" if ingraph else "")) if world > 1 else None},
        "rays_per_sec": value / SAMPLES_PER_RAY,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host[0].numel() * 4) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms},
        "gpu_launches": launches * args.steps,
        "gpu_launches_per_step": launches,
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernel_ms": {k: round(v, 5) for k, v in (per_kernel or {}).items()},
        "kernel_ms_note": "per-launch device time from back-to-back replays of each call (warm caches, no host gap), summed per "
                          "entry point over one step; agrees with the ncu launch list (profiles/), not with eager per-call events",
        "wall_s": wall,
    }
    return line, (step_obj, graph)


def measure_train_variant(dev, R, bf16, steps, warmup, pk, pk_kind, forward_only=False):
    """One single-GPU line of the config-2 step at another batch size / precision, or (forward_only) of config 1:
    the cache stage's forward evaluation of 1024 rays x (64, 64, 32) samples with the shader on the final 32."""
    from neural_radiance_caching_b200 import _lib, workload

    step_obj = workload.CacheTrainStep(dev, bf16=bool(bf16))
    g = np.random.Generator(np.random.PCG64(workload.SEED + 7))
    rn = workload.make_rays_np(g, R)
    u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
    tgt = g.uniform(size=(R, 3)).astype(np.float32)
    xr = workload.backward_mask_rays_np(g, rn)
    ux = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
    dbuf = torch.from_numpy(workload.pack_batch(rn, u, tgt, xr, ux)).to(dev)

    def compute():
        rays, u01, target, extra = workload.unpack_batch(dbuf)
        if forward_only:
            with torch.no_grad():
                return step_obj.render(rays, u01)["rgb"]
        return step_obj.step(rays, u01, target, extra)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            before = _lib.launch_count
            compute()
            launches = _lib.launch_count - before
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    per_kernel = profile_calls(compute, _lib, iters=2)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        compute()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    ms = _timed(graph.replay, steps, warmup, flush, torch.cuda.synchronize)
    roof = roofline(per_kernel, R, pk, pk_kind, bool(bf16)) if not forward_only else None
    if forward_only and "nrc_density_query_fwd" in per_kernel:
        pts = {0: 64 * R, 1: 64 * R, 2: 32 * R}
        fwd = sum(pts[i] * (12 + 8 * F * 4 * L + 4 * L * F) for i, (L, F) in {0: (6, 1), 1: (7, 1), 2: (8, 4)}.items())
        ach = fwd / (per_kernel["nrc_density_query_fwd"] * 1e-3) / 1e9
        roof = {"kernel": "nrc_density_query_fwd", "bound": "hbm", "unit": "GB/s", "peak": pk["hbm_gbs"], "achieved": ach,
                "frac": ach / pk["hbm_gbs"], "peak_source": pk_kind, "traffic": None,
                "note": "3 launches; algorithmic gather bytes / summed per-launch device time; tables are L2-resident"}
    del graph, step_obj
    torch.cuda.empty_cache()
    return {"metric": METRIC, "value": R * SAMPLES_PER_RAY / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": 1,
            "steps": steps, "dtype": "bf16" if bf16 else "f32", "rays_per_step": R, "gpu_launches_per_step": launches,
            "roofline": roof, "kernel_ms": {k: round(v, 5) for k, v in per_kernel.items()}}


def run_b200(args):
    import torch.distributed as dist

    line, keep = run_b200_main(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.ncu_mode or line is False:
        if world > 1:
            dist.destroy_process_group()
        return
    del keep
    torch.cuda.empty_cache()
    others = None
    if not (args.no_others or args.profile_kernels):
        # The rest of BASELINE's configs beside the headline (config 2, bf16, 1024 rays).  config 5 runs on every rank
        # (strong scaling: row bands + one all_gather of the tiles); the single-GPU variants only at N = 1.
        others = {}
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        pk, pk_kind = peaks()

        def guarded(name, fn):
            try:
                res = fn()
                if rank == 0:
                    others[name] = res
            except Exception as ex:        # a failing side measurement must not lose the headline
                if rank == 0:
                    others[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

        guarded("config5_800x800_frame", lambda: _slim(render_line(args, "config5", world, rank, dev, 2, 1, rays=1024, image=800),
                                                        frame=True))
        if world == 1:
            guarded("config3_chunk", lambda: _slim(render_line(args, "config3", 1, 0, dev, 5, 3, rays=1024)))
            guarded("config1_forward", lambda: measure_train_variant(dev, 1024, 1, 10, 3, pk, pk_kind, forward_only=True))
            guarded("config2_fp32", lambda: measure_train_variant(dev, 1024, 0, 10, 3, pk, pk_kind))
            guarded("config2_bf16_16384rays", lambda: measure_train_variant(dev, 16384, 1, 5, 3, pk, pk_kind))
            guarded("config4_chunk", lambda: _slim(render_line(args, "config4", 1, 0, dev, 5, 3, rays=1024)))
    if rank == 0:
        line["others"] = others
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _slim(line, frame=False):
    if line is None:
        return None
    out = {k: line[k] for k in ("metric", "value", "unit", "ms_per_step", "n_gpus", "steps", "dtype", "scaling",
                                "gpu_launches_per_step", "roofline", "kernel_ms", "e2e")}
    out["workload"] = line["config"]["workload"]
    for k in ("train", "slf_variate"):
        if k in line:
            out[k] = line[k]
    if frame:
        out["seconds_per_frame"] = line["ms_per_step"] * 1e-3
    return out


def profile_calls(one_step, _lib, iters=3, repeats=6):
    """Device time of every C-ABI call of one step.  Each call is issued once for real and then `repeats` more times
    back to back inside one CUDA-event pair on the launching stream (same arguments, the tensors still alive): the
    average is the kernel's duration with warm caches and no host launch gap in front of it - what the kernel costs inside
    the step's CUDA graph, and what the ncu launch list shows (profiles/*_launch_shares*).  A single eager call timed by
    its own event pair (the previous method) included the host's launch latency: +20-40 % on 15-60 us kernels.
    Repeats are harmless: outputs are rewritten, gradient sinks only accumulate more.  Collectives are not repeated."""
    acc = {}
    orig = _lib.call

    def timed(name, *a):
        orig(name, *a)
        if "allreduce" in name:
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(repeats):
            orig(name, *a)
        e.record()
        acc.setdefault(name, []).append((s, e))

    _lib.call = timed
    # the host mirrors resolve `_lib.call` at call time (module attribute), so this is seen
    try:
        for _ in range(iters):
            one_step()
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
    out = {}
    for name, evs in acc.items():
        out[name] = sum(s.elapsed_time(e) for s, e in evs) / iters / repeats
    return out


def l2_gather_probe(dev):
    """SURVEY 8d L2 roofline denominator: uniformly random row gathers (4-byte and 16-byte rows, 8 in flight per
    thread) over a 32 MiB table that stays L2-resident (126 MB L2), timed with CUDA events.  Rates in G rows/s."""
    import ctypes as C
    from neural_radiance_caching_b200 import _lib
    table = torch.rand(8 << 20, device=dev)          # 32 MiB
    sink = torch.zeros(1, device=dev)
    out = {"table_mib": 32, "in_flight_per_thread": 8}
    n_thr, per = 148 * 2048 * 4, 64
    for row_bytes in (4, 16):
        rows = table.numel() * 4 // row_bytes
        call = lambda: _lib.call("nrc_probe_gather", _lib.stream_ptr(), _lib.ptr(table), rows, row_bytes, n_thr, per,
                                 _lib.ptr(sink))
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        rate = n_thr * per / (best * 1e-3) / 1e9
        out[f"row{row_bytes}_grows_s"] = rate
        out[f"row{row_bytes}_sector_gbs"] = rate * 32.0     # every random row costs one 32-byte sector
    return out


def _traffic(kernel):
    """Measured DRAM bytes per launch of `kernel` from the committed ncu --set full summary, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except Exception:
        return None


# Dense MACs per shaded point of the cache shader's stacks (SURVEY 8d / DESIGN.md 5)
SHADER_MAC_FWD = 96 * 138 + (129 * 64 + 64 * 64 + 64) + (200 * 128 + 128 * 128 + 128 * 128 + 328 * 128 + 128 * 3) + (
    38 * 128 + 128 * 128 + 128 * 128 + 166 * 128 + 128 * 3)
SHADER_MAC_ENV = 38 * 128 + 128 * 128 + 128 * 128 + 166 * 128 + 128 * 3


def roofline(per_kernel, R, pk, pk_kind, bf16, l2=None):
    """Roofline of the dominant entry point of the step (per-call CUDA events on the launching stream).
    Algorithmic work:
      encode fwd per point = 12 (x) + 8*F*4*L (corner rows) + 4*L*F (features) bytes;
      encode bwd per point = 12 + 4*L*F (upstream grad) + 2 * 8*F*4*L (atomic read-modify-write) bytes;
      shader stacks per shaded point = SHADER_MAC_FWD MACs forward, the same minus the EnvMap (whose
      gradient is exactly zero) for the data-gradient pass, and again for the weight gradients."""
    if not per_kernel:
        return None
    per_kernel = dict(per_kernel)
    if "nrc_chain_run_multi" in per_kernel:   # the shader stacks run through both entry points: one roofline entry
        per_kernel["nrc_chain_run"] = per_kernel.get("nrc_chain_run", 0.0) + per_kernel.pop("nrc_chain_run_multi")
    if "nrc_encode_bwd_warped" in per_kernel:  # the density grids' scatters (means in, contraction in the kernel) + the appearance grid's
        per_kernel["nrc_encode_bwd"] = per_kernel.get("nrc_encode_bwd", 0.0) + per_kernel.pop("nrc_encode_bwd_warped")
    top = max(per_kernel, key=per_kernel.get)
    pts = {0: 64 * R, 1: 64 * R, 2: 32 * R}
    LF = {0: (6, 1), 1: (7, 1), 2: (8, 4)}
    fwd = sum(pts[i] * (12 + 8 * F * 4 * L + 4 * L * F) for i, (L, F) in LF.items())
    bwd = sum(pts[i] * (12 + 4 * L * F + 2 * 8 * F * 4 * L) for i, (L, F) in LF.items())
    app_fwd = 32 * R * (12 + 8 * 4 * 4 * 8 + 4 * 32)          # appearance grid, L=8 F=4
    app_bwd = 32 * R * (12 + 4 * 32 + 2 * 8 * 4 * 4 * 8)
    # nrc_encode_fwd is only the appearance grid (density grids are gathered inside the fused
    # query); nrc_encode_bwd scatters into all four grids.
    # the backward-mask pass repeats the three query levels on the extra rays and back-propagates its final level
    bwd_extra = pts[2] * (12 + 4 * 32 + 2 * 8 * 4 * 4 * 8)
    alg_bytes = {"nrc_encode_fwd": app_fwd, "nrc_encode_bwd": bwd + app_bwd + bwd_extra, "nrc_density_query_fwd": 2 * fwd}
    shaded = 32 * R
    alg_flops = {"nrc_chain_run": 2.0 * shaded * (2 * SHADER_MAC_FWD - SHADER_MAC_ENV),
                 "nrc_chain_wgrad": 2.0 * shaded * (SHADER_MAC_FWD - SHADER_MAC_ENV)}
    if not bf16:
        alg_flops = {"nrc_dense_fwd": 2.0 * shaded * SHADER_MAC_FWD, "nrc_dense_bwd": 4.0 * shaded * SHADER_MAC_FWD}
    res = {"kernel": top, "peak_source": pk_kind, "traffic": _traffic(top)}
    if top in alg_bytes:
        ach = alg_bytes[top] / (per_kernel[top] * 1e-3) / 1e9
        res.update(bound="hbm", unit="GB/s", peak=pk["hbm_gbs"], achieved=ach, frac=ach / pk["hbm_gbs"],
                   note="all launches of this entry point in one step; algorithmic bytes / summed per-launch device time "
                        "(back-to-back replays)")
    elif top in alg_flops:
        peak = pk["bf16_tflops"] if bf16 else None
        ach = alg_flops[top] / (per_kernel[top] * 1e-3) / 1e12
        res.update(bound="tensor", unit="TFLOP/s", peak=peak, achieved=ach, frac=(ach / peak) if peak else None,
                   note="all launches of nrc_chain_run + nrc_chain_run_multi in one step (shader stacks fwd + data-gradient); "
                        "algorithmic FLOPs / summed per-launch device time (back-to-back replays); peak = measured cuBLAS "
                        "bf16 burst")
    else:
        res.update(bound="hbm", unit="GB/s", peak=pk["hbm_gbs"], achieved=None, frac=None,
                   note="dominant entry point has no closed-form work model; see the per-kernel list")
    for k in alg_bytes:
        if k in per_kernel:
            a = alg_bytes[k] / (per_kernel[k] * 1e-3) / 1e9
            res[k] = {"bound": "hbm", "achieved": a, "frac": a / pk["hbm_gbs"], "ms": per_kernel[k], "unit": "GB/s"}
    if l2:
        # corner rows gathered per step (8 per point and level), by row size; time the measured L2 random-gather
        # rate would need for them = the bound the gather-side of these entry points is held against
        g4 = 8 * (pts[0] * 6 + pts[1] * 7)
        g16 = 8 * pts[2] * 8
        rows = {"nrc_density_query_fwd": (2 * g4, 2 * g16), "nrc_encode_fwd": (0, 8 * 32 * R * 8)}
        res["l2_gather"] = dict(l2)
        for k, (a4, a16) in rows.items():
            if k in per_kernel and k in res:
                t_bound = a4 / (l2["row4_grows_s"] * 1e9) + a16 / (l2["row16_grows_s"] * 1e9)
                res[k]["l2_gather_frac"] = t_bound / (per_kernel[k] * 1e-3)
                res[k]["grows_s"] = (a4 + a16) / (per_kernel[k] * 1e-3) / 1e9
    for k in alg_flops:
        if k in per_kernel and bf16:
            a = alg_flops[k] / (per_kernel[k] * 1e-3) / 1e12
            res[k] = {"bound": "tensor", "achieved": a, "frac": a / pk["bf16_tflops"], "ms": per_kernel[k],
                      "unit": "TFLOP/s"}
    return res


# ----------------------------------------------------------------------------- configs 3 and 5
def _timed(fn, steps, warmup, flush, barrier):
    for _ in range(max(3, warmup)):
        fn()
    barrier()
    evs = []
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        evs.append((s, e))
    barrier()
    return sum(s.elapsed_time(e) for s, e in evs) / steps


def _dist_ctx():
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def run_render(args):
    """Render-path workloads (not the BASELINE headline metric; reported for the 8d table)."""
    import torch.distributed as dist

    world, rank, local, dev = _dist_ctx()
    line = render_line(args, args.workload, world, rank, dev, args.steps, args.warmup)
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def render_line(args, workload_name, world, rank, dev, steps, warmup, rays=None, image=None):
    """One bench line (dict, on rank 0; None elsewhere) of a render-path workload: config3 or config5."""
    import torch.distributed as dist

    from neural_radiance_caching_b200 import _lib, dist as ndist, render_image as ri, workload

    _lib.load()
    pk, pk_kind = peaks()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    extra = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    S = 32
    use_graph = False
    if workload_name == "config3":
        R = rays or args.rays
        stage = workload.MaterialRenderStep(dev, bf16=bool(args.bf16))
        g = np.random.Generator(np.random.PCG64(workload.SEED + rank))
        host = torch.from_numpy(np.concatenate([a.reshape(-1) for a in workload.make_surface_np(g, R)])).pin_memory()
        dbuf = host.to(dev)
        draws = stage.draws(R)
        out_host = torch.zeros((R, 3), dtype=torch.float32).pin_memory()

        def step():
            m, v, n = (dbuf[i * R * 3:(i + 1) * R * 3].view(R, 3) for i in range(3))
            lobes = stage.light_lobes(R, m, n)      # LightMLP at the shaded points: part of the chunk
            return stage.render(m, v, n, draws, lobes)["rgb"]

        before = _lib.launch_count
        step()
        launches = _lib.launch_count - before
        if args.ncu_mode:
            step()
            torch.cuda.synchronize()
            print(json.dumps({"ncu_mode": True, "launches_per_step": launches}))
            return None
        per_kernel = profile_calls(step, _lib, iters=3)
        # the chunk is a static launch sequence (inputs and random draws resident): capture it once
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_rgb = step()
        use_graph = True

        def gstep():
            graph.replay()
            return static_rgb

        dev_ms = _timed(gstep, steps, warmup, flush, barrier)

        def e2e_step():
            dbuf.copy_(host, non_blocking=True)
            out_host.copy_(gstep(), non_blocking=True)

        e2e_ms = _timed(e2e_step, steps, warmup, flush, barrier)
        # the same chunk with the surface-light-field control variate (MaterialModel.slf_variate, nerf_ngp_yobo.gin:91):
        # the light-field memory queried on the 32 768 secondary rays of the cache pass, its integral subtracted
        try:
            stage_v = workload.MaterialRenderStep(dev, bf16=bool(args.bf16), slf_variate=True)
            draws_v = stage_v.draws(R)

            def vstep():
                m, v, n = (dbuf[i * R * 3:(i + 1) * R * 3].view(R, 3) for i in range(3))
                return stage_v.render(m, v, n, draws_v, stage_v.light_lobes(R, m, n))["rgb"]

            before_v = _lib.launch_count
            vstep()
            launches_v = _lib.launch_count - before_v
            side_v = torch.cuda.Stream()
            side_v.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side_v):
                vstep()
            torch.cuda.current_stream().wait_stream(side_v)
            torch.cuda.synchronize()
            graph_v = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_v):
                static_v = vstep()
            v_ms = _timed(lambda: graph_v.replay(), steps, warmup, flush, barrier)
            extra["slf_variate"] = {"ms_per_step": v_ms, "gpu_launches_per_step": launches_v,
                                    "note": "one CUDA graph like the plain chunk; adds distance grid + 3 stacks + "
                                            "nrc_slf_points_fwd + reflectance grid at 8 points per ray + nrc_slf_reduce_fwd + 2 "
                                            "GGX integrations on the same 32 768 secondary rays"}
            del graph_v, static_v
            del stage_v
        except Exception as ex:      # a failing side measurement must not lose the line
            extra["slf_variate"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        units = world * R * S * SAMPLES_PER_RAY
        wl = ("config3 material_light_from_scratch_resample chunk: %d shaded points x 32 secondary rays (16 microfacet + 8 "
              "cosine + 8 vMF-mixture/128 lobes, MIS), cache query per ray (64,64,32, power-ladder warp, categorical "
              "resample, cache shader), 256-wide env map, GGX/Lambert integration; forward (render) path" % R)
        h2d, d2h = int(host.numel() * 4) * world, R * 3 * 4 * world
        scaling = "weak"
        top = max(per_kernel, key=per_kernel.get)
        per_ray = 64 * (12 + 192 + 24) + 64 * (12 + 224 + 28) + 32 * (12 + 1024 + 128)
        roof = {"kernel": top, "peak_source": pk_kind, "traffic": _traffic(top)}
        if top == "nrc_density_query_fwd":
            ach = R * S * per_ray / (per_kernel[top] * 1e-3) / 1e9
            roof.update(bound="hbm", unit="GB/s", peak=pk["hbm_gbs"], achieved=ach, frac=ach / pk["hbm_gbs"],
                        note="3 launches (proposal levels) on 32768 secondary rays; algorithmic gather bytes / summed "
                             "CUDA-event time; tables are L2-resident, so values above the HBM peak are possible")
            l2 = l2_gather_probe(dev)
            rays = R * S
            t_bound = (rays * 8 * 64 * (6 + 7)) / (l2["row4_grows_s"] * 1e9) + (rays * 8 * 32 * 8) / (l2["row16_grows_s"] * 1e9)
            roof["l2_gather"] = l2
            roof["l2_gather_frac"] = t_bound / (per_kernel[top] * 1e-3)
            roof["grows_s"] = rays * 8 * (64 * 13 + 32 * 8) / (per_kernel[top] * 1e-3) / 1e9
        else:
            roof.update(bound="hbm", unit="GB/s", peak=pk["hbm_gbs"], achieved=None, frac=None)
    elif workload_name == "config4":
        R = rays or args.rays
        stage = workload.TransientRenderStep(dev, bf16=bool(args.bf16))
        g = np.random.Generator(np.random.PCG64(workload.SEED + rank))
        rn = stage.make_rays(g, R)
        keys = sorted(rn.keys())
        host = {k: torch.from_numpy(np.ascontiguousarray(rn[k], dtype=np.float32)).pin_memory() for k in keys}
        drays = {k: v.to(dev) for k, v in host.items()}
        u01 = [torch.rand((R, 1), device=dev) for _ in range(3)]
        out_host = torch.zeros((R, stage.n_bins, 3), dtype=torch.float32).pin_memory()

        def step():
            return stage.render(drays, u01)["rgb"]

        before = _lib.launch_count
        step()
        launches = _lib.launch_count - before
        if args.ncu_mode:
            step()
            torch.cuda.synchronize()
            print(json.dumps({"ncu_mode": True, "launches_per_step": launches}))
            return None
        per_kernel = profile_calls(step, _lib, iters=2)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_rgb = step()
        use_graph = True

        def gstep():
            graph.replay()
            return static_rgb

        dev_ms = _timed(gstep, steps, warmup, flush, barrier)

        def e2e_step():
            for k in keys:
                drays[k].copy_(host[k], non_blocking=True)
            out_host.copy_(gstep(), non_blocking=True)

        e2e_ms = _timed(e2e_step, steps, warmup, flush, barrier)
        # the training path of the same chunk: heads as GEMMs, [R,32,700,3] histograms in HBM, nrc_transient_render_{fwd,bwd},
        # the Dense / encoding / sampler VJPs under autograd (eager launches, no graph); a synthetic target histogram
        target = (gstep() * 0.5).clone()
        train_ms = _timed(lambda: stage.loss_and_grads(drays, u01, target)[0], 3, 3, flush, barrier)
        extra["train"] = {"ms_per_step": train_ms, "value": world * R * SAMPLES_PER_RAY / (train_ms * 1e-3), "unit": UNIT,
                          "note": "forward + backward of one chunk through workload.TransientRenderStep.loss_and_grads (unfused, "
                                  "eager); mean-squared-error objective on the time-resolved rgb"}
        units = world * R * SAMPLES_PER_RAY
        wl = ("config4 transient_simulation_ngp_yobo_cornell time-resolved cache, one chunk of %d primary rays: proposal sampler "
              "(64,64,32), transient shader on the 32 final samples (appearance grid, bottleneck / roughness / tint / albedo / "
              "integrated BRDF, irradiance stack 111->64->64 and transient SurfaceLightField 200->128x4), both transient heads' "
              "last layers (64->2100, 128->2101) fused with activation, zero_invalid_bins, sub-bin shift and the weighted "
              "reduction into [R,700,3] (the [R,32,700,3] histograms never reach HBM), direct splat, Gaussian temporal filter "
              "(sigma 3 bins); direct term without shadow rays / light BRDF network; forward (render) path" % R)
        h2d, d2h = int(sum(v.numel() for v in host.values()) * 4) * world, R * stage.n_bins * 3 * 4 * world
        scaling = "weak"
        top = max(per_kernel, key=per_kernel.get)
        head_flops = 2.0 * R * 32 * (64 * stage.n_bins * 3 + 128 * (stage.n_bins * 3 + 1))
        roof = {"kernel": top, "peak_source": pk_kind, "traffic": None}
        if top == "nrc_transient_head_render_fwd":
            ach = head_flops / (per_kernel[top] * 1e-3) / 1e12
            roof.update(bound="tensor", unit="TFLOP/s", peak=pk["bf16_tflops"], achieved=ach, frac=ach / pk["bf16_tflops"],
                        note="403 k MAC per shaded sample (SURVEY 8d) in the two head layers / per-launch device time; the "
                             "kernel is mma.sync + shared-memory gather (the reduction over samples, not the GEMM, is its cost)")
        else:
            roof.update(bound="hbm", unit="GB/s", peak=pk["hbm_gbs"], achieved=None, frac=None)
    else:
        H = W = image or args.image
        fr = workload.FrameRenderer(dev, bf16=bool(args.bf16))
        c2w = ri.orbit_camera()
        focal = 1111.0 * H / 800.0

        graphs = {}    # the captured chunk (camera rays -> cache + material stage -> band stores -> counter) per band

        def step():
            return ri.render_image(fr.render_chunk, H, W, focal, c2w, dev, chunk=rays or args.rays, _graphs=graphs)["rgb"]

        before = _lib.launch_count
        img = step()
        launches = _lib.launch_count - before
        # the kernels of ONE chunk (the frame is ceil(H*W / chunk / N) replays of the same graph per rank)
        chunk_rays = {k: v[:(rays or args.rays)].contiguous() for k, v in ri.pinhole_rays(H, W, focal, c2w, dev, rows=(0, 8)).items()}
        per_kernel = profile_calls(lambda: fr.render_chunk(chunk_rays), _lib, iters=2)
        dev_ms = _timed(step, steps, warmup, flush, barrier)
        out_host = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory()

        def e2e_step():
            out_host.copy_(step(), non_blocking=True)

        e2e_ms = _timed(e2e_step, steps, warmup, flush, barrier)
        units = H * W * (SAMPLES_PER_RAY + S * SAMPLES_PER_RAY)
        wl = ("config5 full %dx%d view: per 1024-ray chunk cache stage on the primary rays (config 1, resampled) + "
              "material stage (config 3); image row bands over the ranks, one all_gather of the bands; the chunk loop is one CUDA "
              "graph replayed per chunk with the chunk counter and the camera-ray generation on the device" % (H, W))
        h2d, d2h = 0, H * W * 3 * 4
        scaling = "strong"
        use_graph = True
        top = max(per_kernel, key=per_kernel.get)
        roof = {"kernel": top, "peak_source": pk_kind, "traffic": _traffic(top), "bound": "hbm", "unit": "GB/s",
                "peak": pk["hbm_gbs"], "achieved": None, "frac": None, "per": "one 1024-ray chunk"}
        if top == "nrc_density_query_fwd":
            Rc = rays or args.rays
            nray = Rc * (1 + S)          # primary rays + 32 secondary rays per shaded point
            per_ray = 64 * (12 + 192 + 24) + 64 * (12 + 224 + 28) + 32 * (12 + 1024 + 128)
            ach = nray * per_ray / (per_kernel[top] * 1e-3) / 1e9
            l2 = l2_gather_probe(dev)
            t_bound = (nray * 8 * 64 * (6 + 7)) / (l2["row4_grows_s"] * 1e9) + (nray * 8 * 32 * 8) / (l2["row16_grows_s"] * 1e9)
            roof.update(achieved=ach, frac=ach / pk["hbm_gbs"], l2_gather=l2, l2_gather_frac=t_bound / (per_kernel[top] * 1e-3),
                        note="density queries of one chunk (primary + secondary rays); algorithmic gather bytes / summed "
                             "per-launch device time; tables are L2-resident: the L2 random-gather rate is the bound")
    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return None
    return {"metric": METRIC, "value": units / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(3, warmup), "ms_per_step": dev_ms, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "bf16" if args.bf16 else "f32", "data": "synthetic",
            "config": {"workload": wl, "l2": "flushed between timed steps", "cuda_graph": use_graph,
                       "parallelism": f"dp{world}"},
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": launches * steps, "gpu_launches_per_step": launches, "roofline": roof,
            "kernel_ms": {k: round(v, 5) for k, v in per_kernel.items()}, **extra}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload != "config2":
        run_render(a)
    else:
        run_b200(a)
