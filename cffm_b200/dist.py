"""Data-parallel plumbing: ``torch.distributed`` carries the rendezvous (and, in tests, gloo
collectives); the training step's own collectives run inside the CUDA library over NCCL.

One process per GPU.  Every rank holds a full replica (tables included) and trains on its shard
of the global batch; per step the library all-reduces the loss sum and the dense gradients and
all-gathers the (id, gradient-row) lists, then every rank applies the identical sorted
segment-sum + Adagrad update, so replicas stay bit-identical."""
from __future__ import annotations

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, near-equal split of n samples; rank r owns [lo, hi)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(ids, labels, rank, world):
    lo, hi = shard_bounds(len(labels), rank, world)
    return np.ascontiguousarray(ids[lo:hi]), np.ascontiguousarray(labels[lo:hi])


def bind_engine(engine, dist=None):
    """Create the NCCL communicator of ``engine``: rank 0 makes the unique id, torch.distributed
    broadcasts it.  Must be called before the first training step."""
    if dist is None:
        import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return
    from .engine import comm_unique_id
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.comm_init(box[0], rank, world)
