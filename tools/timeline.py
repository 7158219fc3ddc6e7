"""Kernel timeline of ONE replay of the config-2 step graph (CUPTI activity records through torch.profiler):
start / end of every kernel per stream, so the critical path and the idle gaps are visible.
usage (GPU box): python tools/timeline.py [--rays 1024] > gpurun_out/timeline.txt"""
import argparse, json, os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_radiance_caching_b200 import _lib, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1024)
ap.add_argument("--bf16", type=int, default=1)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
_lib.load()
R = args.rays
step_obj = workload.CacheTrainStep(dev, bf16=bool(args.bf16))
g = np.random.Generator(np.random.PCG64(workload.SEED))
rn = workload.make_rays_np(g, R)
u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
tgt = g.uniform(size=(R, 3)).astype(np.float32)
xr = workload.backward_mask_rays_np(g, rn)
ux = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
dbuf = torch.from_numpy(workload.pack_batch(rn, u, tgt, xr, ux)).to(dev)


def compute_step():
    rays, u01, target, extra = workload.unpack_batch(dbuf)
    return step_obj.step(rays, u01, target, extra, fused_allreduce=False)


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        compute_step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    compute_step()
torch.cuda.synchronize()
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        graph.replay()
        torch.cuda.synchronize()
path = "/tmp/nrc_timeline_trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
k = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "ts" in e]
k.sort(key=lambda e: e["ts"])
# split into replays by large gaps
groups, cur = [], []
for e in k:
    if cur and e["ts"] - (cur[-1]["ts"] + cur[-1]["dur"]) > 200 and e["ts"] - cur[0]["ts"] > 300:
        groups.append(cur); cur = []
    cur.append(e)
if cur: groups.append(cur)
last = groups[-1]
t0 = last[0]["ts"]
streams = sorted({e["args"].get("stream") for e in last})
sid = {s: i for i, s in enumerate(streams)}
end = max(e["ts"] + e["dur"] for e in last) - t0
print(f"replay: {len(last)} launches, {end:.1f} us from first kernel start to last kernel end, {len(streams)} streams")
busy = {s: 0.0 for s in streams}
print(f"{'start':>8} {'end':>8} {'dur':>7}  st  name")
for e in last:
    s = e["args"].get("stream")
    busy[s] += e["dur"]
    nm = e["name"].replace("void ", "").replace("nrc::", "")
    nm = nm.split("(")[0][:70]
    print(f"{e['ts']-t0:8.1f} {e['ts']+e['dur']-t0:8.1f} {e['dur']:7.1f}  {sid[s]:>2}  {'    ' * sid[s]}{nm}")
for s in streams:
    print(f"stream {sid[s]}: busy {busy[s]:.1f} us")
# union coverage: time with at least one kernel running
iv = sorted((e["ts"] - t0, e["ts"] + e["dur"] - t0) for e in last)
cov, ce = 0.0, 0.0
for a, b in iv:
    if b > ce:
        cov += b - max(a, ce); ce = b
print(f"time with >=1 kernel running: {cov:.1f} us of {end:.1f}")
