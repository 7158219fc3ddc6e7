"""Peer-memory gradient all-reduce against NCCL (run under torchrun, N >= 2):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/test_allreduce.py
Checks both kernels (multicast if available, peer loads/stores) on the full arena and on an offset bucket, then times
them and NCCL at the config-2 arena size (110.6 MB)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from neural_radiance_caching_b200 import dist as ndist


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 27_650_048            # floats: the config-2 gradient arena (110.6 MB)
    res = {"world": world}
    for mode in ("multicast", "peer"):
        os.environ["NRC_ALLREDUCE"] = mode
        arena = ndist.PeerArena.create(n, dev)
        if arena is None:
            res[mode] = "unavailable"
            continue
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        src = torch.randn(n, device=dev, generator=g)
        ref = src.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        arena.buf[:n].copy_(src)
        arena.allreduce_mean_(0, n)
        torch.cuda.synchronize()
        err = float((arena.buf[:n] - ref).abs().max())
        # offset bucket: only [lo, hi) may change
        lo, hi = 16_908_288, n
        arena.buf[:n].copy_(src)
        arena.allreduce_mean_(lo, hi - lo, channel=1)
        torch.cuda.synchronize()
        err2 = float((arena.buf[lo:n] - ref[lo:]).abs().max())
        untouched = bool(torch.equal(arena.buf[:lo], src[:lo]))
        ms = timeit(lambda: arena.allreduce_mean_(0, n))
        sweep = {}
        if os.environ.get("NRC_AR_SWEEP"):
            for ctas in (16, 32, 64, 148, 296, 592):
                sweep[ctas] = round(timeit(lambda: arena.allreduce_mean_(0, n, num_ctas=ctas), iters=10, warm=3), 4)
        res[mode if arena.mode == mode else mode + "->" + arena.mode] = {"max_abs_err_vs_nccl": err, "bucket_err": err2,
                                                         "outside_bucket_untouched": untouched, "ms_110MB": ms,
                                                         "algbw_GBs": n * 4 / ms / 1e6, "ms_by_ctas": sweep}
        del arena
    x = torch.randn(n, device=dev)
    ms = timeit(lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG))
    res["nccl"] = {"ms_110MB": ms, "algbw_GBs": n * 4 / ms / 1e6}
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
