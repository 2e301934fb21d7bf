"""Data-parallel plumbing: ``torch.distributed`` carries the rendezvous (and, in tests, gloo
collectives); the training step's own collectives run inside the CUDA library over NCCL.

One process per GPU, each training on its shard of the global batch.  Dense variables are replicated: per
step the library all-reduces the loss sum and the dense gradients.  The three embedding tables are either

* replicated (small tables): the (id, gradient-row) lists of all ranks are all-gathered and every rank applies the
  identical sorted segment-sum + Adagrad update, so replicas stay bit-identical, or
* row-sharded (``Engine(..., shard=(rank, world))``; the 10M-row Criteo-shaped table): rank r owns the rows with
  ``row % world == r``; rows travel to the ranks that need them and gradient sums travel back by all-to-all
  (cffm_b200/csrc/shard.cu)."""
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


def owner_of(rows, world):
    """Owner rank and local row number of global table rows under the row-sharded layout."""
    rows = np.asarray(rows)
    return rows % world, rows // world


def gather_table(engine, name, dist=None, accum=False):
    """Assemble the full [features_M, K] table of a row-sharded engine on every rank (tests / checkpoints of small
    tables; a production checkpoint writes each rank's shard, see ``Engine.state_dict``)."""
    if dist is None:
        import torch.distributed as dist
    local = engine.get_param(name, accum)
    if not engine.shard or name not in engine.TABLES:
        return local            # dense variables are replicated
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, local)
    full = np.empty((engine.features_M,) + local.shape[1:], dtype=local.dtype)
    for r, part in enumerate(parts):
        full[r::world] = part
    return full
