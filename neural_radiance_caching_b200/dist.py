"""Data-parallel plumbing of the query path: one process per GPU, rays / image row-bands
sharded over ranks, parameters replicated, no collective on the forward path.

  * gradient all-reduce (mean) = the reference's lax.pmean over the gradient pytree
    (internal/train_utils.py:3132-3136): ONE collective over the flat gradient arena;
  * tile gather = the reference's lax.all_gather of rendered chunks (internal/train_utils.py:3795-3815):
    image row-bands, one per rank (do not split a 1024-ray chunk across ranks).
Backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def row_bands(height, world_size):
    """[(row0, row1)] per rank: contiguous bands whose sizes differ by at most one row."""
    base, extra = divmod(height, world_size)
    bands, r0 = [], 0
    for r in range(world_size):
        r1 = r0 + base + (1 if r < extra else 0)
        bands.append((r0, r1))
        r0 = r1
    return bands


def shard_rays(num_rays, rank, world_size):
    """Half-open ray range of `rank` when a batch of `num_rays` is split evenly (utils.shard, utils.py:333-335)."""
    if num_rays % world_size:
        raise ValueError(f"batch of {num_rays} rays does not divide over {world_size} ranks")
    per = num_rays // world_size
    return rank * per, (rank + 1) * per


def allreduce_mean_(flat):
    """In-place mean over ranks of the flat gradient arena (one collective)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat


def gather_tiles(band, height):
    """band [rows_of_this_rank, W, C] -> full image [height, W, C] on every rank (row-band order)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return band
    ws = dist.get_world_size()
    bands = row_bands(height, ws)
    max_rows = max(b - a for a, b in bands)
    pad = torch.zeros((max_rows,) + tuple(band.shape[1:]), device=band.device, dtype=band.dtype)
    pad[:band.shape[0]] = band
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[:b - a] for o, (a, b) in zip(out, bands)], dim=0)
