"""Data-parallel step check (torchrun, N >= 2): the in-step bucketed all-reduce (peer-memory kernels forked from the
backward pass) must leave the same averaged gradients as the plain step followed by one all-reduce of the whole arena.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/test_dp_step.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_radiance_caching_b200 import _lib, workload  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
_lib.load()
R = 1024
step = workload.CacheTrainStep(dev, bf16=True)
g = np.random.Generator(np.random.PCG64(workload.SEED + rank))
rn = workload.make_rays_np(g, R)
u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
tgt = g.uniform(size=(R, 3)).astype(np.float32)
xr = workload.backward_mask_rays_np(g, rn)
ux = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
dbuf = torch.from_numpy(workload.pack_batch(rn, u, tgt, xr, ux)).to(dev)


def run(fused):
    rays, u01, target, extra = workload.unpack_batch(dbuf)
    loss = step.step(rays, u01, target, extra, fused_allreduce=fused)
    if not fused:
        step.allreduce_grads()
    torch.cuda.synchronize()
    dist.barrier()
    return float(loss), step.flat_grad.clone()


ok = True
if step.peer is None:
    print("no peer arena: nothing to compare")
else:
    run(True)                       # warm-up (stream / arena creation)
    la, ga = run(True)
    lb, gb = run(False)
    so, fo, go = step.shader_offset, step.final_level_offset, step.shader_grid_end
    for name, lo, hi in (("proposal", 0, fo), ("final", fo, so), ("grid", so, go), ("stacks", go, ga.numel())):
        a, b = ga[lo:hi], gb[lo:hi]
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        good = err <= 2e-5 and bool(torch.isfinite(a).all())
        ok = ok and good
        if rank == 0:
            print(f"bucket {name:9s} [{lo}, {hi}) rel err {err:.2e} {'ok' if good else 'MISMATCH'}")
    # all ranks hold the same averaged gradients
    ref = ga.clone()
    dist.broadcast(ref, 0)
    same = bool((ref == ga).all())
    ok = ok and same
    if rank == 0:
        print(f"loss fused {la:.6f} plain {lb:.6f}; ranks identical: {same}; announced final bucket: {step.engine.final_grads_announced}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
