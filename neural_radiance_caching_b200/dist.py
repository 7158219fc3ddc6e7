"""Data-parallel plumbing of the query path: one process per GPU, rays / image row-bands
sharded over ranks, parameters replicated, no collective on the forward path.

  * gradient all-reduce (mean) = the reference's lax.pmean over the gradient pytree
    (internal/train_utils.py:3132-3136): ONE collective over the flat gradient arena;
  * tile gather = the reference's lax.all_gather of rendered chunks (internal/train_utils.py:3795-3815):
    image row-bands, one per rank (do not split a 1024-ray chunk across ranks).
Backend-agnostic (NCCL on the GPU box, gloo in the CPU tests); on NVLink / NVSwitch boxes the gradient arena lives
in symmetric memory and the all-reduce is the library's own peer-memory kernel (PeerArena)."""
import ctypes
import os
import sys

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


class PeerArena:
    """Flat fp32 gradient arena in symmetric memory (every rank maps every rank's copy over NVLink) with an in-place
    mean all-reduce that is ONE kernel of this library per bucket (csrc/allreduce.cu): multicast / in-switch reduction
    when the box has an NVSwitch multicast object, plain peer loads and stores otherwise.  torch's symmetric-memory
    module is used for the allocation, the handle exchange and the stream-ordered cross-rank barrier only.

    PeerArena.create() returns None when symmetric memory cannot be set up (single process, gloo, no peer access):
    the caller then keeps a plain tensor and allreduce_mean_ (NCCL)."""

    def __init__(self, buf, hdl, mode):
        self.buf, self.hdl, self.mode = buf, hdl, mode
        self.rank, self.world = hdl.rank, hdl.world_size
        self._peers = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        self.multicast_ptr = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
        self._next_channel = 0

    @staticmethod
    def create(numel, device):
        want = os.environ.get("NRC_ALLREDUCE", "peer")
        if want == "nccl" or not dist.is_initialized() or dist.get_world_size() == 1 or dist.get_backend() != "nccl":
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            numel = (int(numel) + 63) // 64 * 64
            buf = symm.empty(numel, dtype=torch.float32, device=device)
            buf.zero_()
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            mc = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
            # Which kernel by default (NRC_ALLREDUCE=peer|multicast forces one):
            #   N = 2: plain peer loads / stores (110.6 MB arena: 0.186 ms vs 0.238 ms NCCL vs 0.314 ms multimem,
            #          profiles/r01j_allreduce_n2.json);
            #   N >= 4: multimem (in-switch reduction).  The two-shot peer scheme moves 2 (N-1)/N of the arena through every
            #          GPU's links and needs >= 148 CTAs to keep them busy; the multimem kernel moves the arena once, is as
            #          fast at N = 8 (0.304 vs 0.320 ms, profiles/r01j_allreduce_n8.json) and is flat from 16 CTAs up
            #          (profiles/r01k_allreduce_cta_sweep_n2.json) - so a bucket that overlaps the backward pass leaves
            #          the SMs to it.
            if want == "multicast":
                mode = "multicast" if mc else "peer"
            elif want == "peer-only":
                mode = "peer"
            else:
                mode = "multicast" if (mc and dist.get_world_size() >= 4) else "peer"
            return PeerArena(buf, hdl, mode)
        except Exception as e:   # no NVLink peer access / symmetric memory unavailable: NCCL path
            if dist.get_rank() == 0:
                print(f"[nrc] symmetric-memory gradient arena unavailable ({type(e).__name__}: {e}); using NCCL",
                      file=sys.stderr)
            return None

    def allreduce_mean_(self, offset=0, count=None, channel=0, num_ctas=0, mode=None, leading_barrier=True,
                        trailing_barrier=True):
        """Mean over ranks of buf[offset:offset+count] in place, on the current stream.  Concurrent calls (different
        streams) must use different `channel`s (each call uses barrier channels 2*channel and 2*channel+1).
        `mode` overrides the arena's default kernel for this bucket ('peer' | 'multicast').
        leading_barrier=False: the caller has already passed a barrier that covers this range's producers (several
        ranges behind one barrier); trailing_barrier=False: the caller issues ONE barrier() after all its buckets
        instead of one per bucket (nothing may read the reduced range before that)."""
        from . import _lib
        count = self.buf.numel() - offset if count is None else count
        if offset % 4 or count % 4:
            raise ValueError("offset and count must be multiples of 4 floats")
        mode = mode or self.mode
        if mode == "multicast" and not self.multicast_ptr:
            mode = "peer"
        if leading_barrier:
            self.hdl.barrier(channel=2 * channel)        # every rank's gradients are complete (stream ordered)
        if mode == "multicast":
            _lib.call("nrc_allreduce_mean_multicast", _lib.stream_ptr(), self.multicast_ptr, int(offset),
                      int(count), self.rank, self.world, int(num_ctas))
        else:
            _lib.call("nrc_allreduce_mean_peer", _lib.stream_ptr(), ctypes.cast(self._peers, ctypes.c_void_p), int(offset),
                      int(count), self.rank, self.world, int(num_ctas))
        if trailing_barrier:
            self.hdl.barrier(channel=2 * channel + 1)    # every slice has been written everywhere
        return self.buf

    def barrier(self, channel=6):
        """Stream-ordered cross-rank barrier (closes a group of allreduce_mean_(trailing_barrier=False) calls)."""
        self.hdl.barrier(channel=channel)


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
