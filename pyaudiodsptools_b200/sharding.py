"""Channel sharding across the GPUs of one box (one process per GPU).

Channels (the reference's "one device object per mono channel",
Example2.py:13-22) are independent, so the path shards with NO data-path
collective: every rank filters its own contiguous channel range.  The only
exchanges are the trivial ones of SURVEY.md §8(e): scatter of input rows from a
root rank and gather of output rows, done by ``Communicator`` (our own NCCL
communicator in libadt_b200.so, NVLink / NVSwitch) on device buffers.  The
``*_host`` helpers do the same over an already-initialised torch.distributed
group (gloo on CPU) and exist for host-side plumbing and for the CPU tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native


def channel_range(n_channels: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) with even boundaries (stereo pairs stay on one GPU)."""
    pairs = (n_channels + 1) // 2
    base, rem = divmod(pairs, world)
    lo_p = rank * base + min(rank, rem)
    hi_p = lo_p + base + (1 if rank < rem else 0)
    return min(2 * lo_p, n_channels), min(2 * hi_p, n_channels)


# ---- host-side plumbing over torch.distributed (gloo / CPU tensors) ---------------------
def _dist():
    import torch
    import torch.distributed as dist
    return torch, dist


def scatter_channels_host(full, n_channels, n_samples, src=0):
    torch, dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = channel_range(n_channels, world, rank)
    out = torch.empty((hi - lo, n_samples), dtype=torch.float32)
    if rank == src:
        parts = []
        for r in range(world):
            a, b = channel_range(n_channels, world, r)
            parts.append(torch.from_numpy(np.ascontiguousarray(full[a:b], dtype=np.float32)))
        out.copy_(parts[src])
        reqs = [dist.isend(parts[r], r) for r in range(world) if r != src and parts[r].numel()]
        for q in reqs:
            q.wait()
    elif out.numel():
        dist.recv(out, src)
    return out.numpy()


def gather_channels_host(shard, n_channels, dst=0):
    torch, dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    shard = np.ascontiguousarray(shard, dtype=np.float32)
    if rank != dst:
        if shard.size:
            dist.send(torch.from_numpy(shard), dst)
        return None
    out = np.empty((n_channels, shard.shape[1]), dtype=np.float32)
    for r in range(world):
        a, b = channel_range(n_channels, world, r)
        if r == dst:
            out[a:b] = shard
        elif b > a:
            buf = torch.empty((b - a, shard.shape[1]), dtype=torch.float32)
            dist.recv(buf, r)
            out[a:b] = buf.numpy()
    return out


def max_over_ranks(value: float) -> float:
    torch, dist = _dist()
    t = torch.tensor([value], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- device-side: our own NCCL communicator (adt_comm_*) --------------------------------
class Communicator:
    """NCCL communicator owned by libadt_b200.so.  The 128-byte unique id is created on rank 0 and
    handed to the other ranks through the existing torch.distributed (or any other) rendezvous."""

    def __init__(self, ctx: _native.Context, rank: int, world: int, unique_id: bytes | None = None):
        self.ctx, self.rank, self.world = ctx, rank, world
        lib = ctx.lib
        if unique_id is None:
            _, dist = _dist()
            box = [None]
            if rank == 0:
                buf = C.create_string_buffer(_native.NCCL_UNIQUE_ID_BYTES)
                rc = lib.adt_comm_unique_id(buf)
                if rc != 0:
                    raise _native.AdtError(rc, "adt_comm_unique_id failed (libnccl.so.2 not loadable?)")
                box[0] = buf.raw
            dist.broadcast_object_list(box, src=0)
            unique_id = box[0]
        h = C.c_void_p()
        ctx.check(lib.adt_comm_create(ctx.h, unique_id, rank, world, C.byref(h)))
        self.h = h

    def scatter_rows(self, full_dev, shard_dev, rows_per_rank, pitch, root=0):
        self.ctx.check(self.ctx.lib.adt_comm_scatter_rows(self.h, full_dev, shard_dev, rows_per_rank, pitch, root))

    def gather_rows(self, shard_dev, full_dev, rows_per_rank, pitch, root=0):
        self.ctx.check(self.ctx.lib.adt_comm_gather_rows(self.h, shard_dev, full_dev, rows_per_rank, pitch, root))

    def channel_counts(self, n_channels):
        """Rows per rank of the documented partition (``channel_range``), as the int64 array the C ABI takes."""
        counts = [b - a for a, b in (channel_range(n_channels, self.world, r) for r in range(self.world))]
        return (C.c_int64 * self.world)(*counts)

    def scatter_channels(self, full_dev, shard_dev, n_channels, pitch, root=0):
        """Scatter the root's [n_channels][pitch] device matrix by ``channel_range`` (uneven shards allowed)."""
        self.ctx.check(self.ctx.lib.adt_comm_scatterv_rows(self.h, full_dev, shard_dev, self.channel_counts(n_channels),
                                                           pitch, root))

    def gather_channels(self, shard_dev, full_dev, n_channels, pitch, root=0):
        self.ctx.check(self.ctx.lib.adt_comm_gatherv_rows(self.h, shard_dev, full_dev, self.channel_counts(n_channels),
                                                          pitch, root))

    def broadcast(self, buf_dev, nbytes, root=0):
        self.ctx.check(self.ctx.lib.adt_comm_broadcast(self.h, buf_dev, nbytes, root))

    def barrier(self):
        self.ctx.check(self.ctx.lib.adt_comm_barrier(self.h))

    def close(self):
        if self.h:
            self.ctx.lib.adt_comm_destroy(self.h)
            self.h = None
