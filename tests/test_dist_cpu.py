"""world_size-2 gloo test (CPU) of the data-parallel scheme the CUDA library implements: shard the
batch, all-reduce the loss sum and the dense gradients, all-gather the (id, gradient-row) lists,
apply the identical de-duplicated update on every rank.  The arithmetic here is the ORACLE's (this
is test infrastructure); the check is that G ranks x B/G samples equal 1 rank x B samples."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cffm_b200.dist import shard_batch, shard_bounds
from oracle.cffm_ref import CFFMRef


def test_shard_bounds_cover_the_batch():
    for n in (7, 8, 256, 8191):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _model():
    m = CFFMRef(80, 4, 8, 8, activation="selu", dtype=torch.float64, seed=5)
    g = torch.Generator().manual_seed(2)
    m.params["feature_bias"] = torch.randn(80, 1, generator=g, dtype=torch.float64) * 0.2
    m.params["outer_embeddings"] = torch.randn(80, 8, generator=g, dtype=torch.float64) * 0.3
    return m


def _batch():
    rng = np.random.default_rng(3)
    ids = rng.integers(0, 80, (12, 4)).astype(np.int32)
    ids[7] = ids[2]  # a row touched on both ranks
    return ids, rng.choice([-1.0, 1.0], 12)


def _dp_step(m, ids, y, world):
    """One data-parallel step of the default (RMSE) loss with gloo collectives."""
    ids_t = torch.as_tensor(ids, dtype=torch.long)
    yt = torch.as_tensor(y, dtype=torch.float64).reshape(-1, 1)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in m.params.items()
              if k not in CFFMRef.SPARSE_TABLES}
    rows = {k: m.params[k][ids_t].detach().clone().requires_grad_(True) for k in CFFMRef.SPARSE_TABLES}
    from oracle.cffm_ref import _GatherProxy
    p2 = dict(leaves)
    for k in rows:
        p2[k] = _GatherProxy(m.params[k], ids_t, rows[k])
    out = m.forward(ids, p2)
    # global loss: the sum of squared residuals crosses ranks BEFORE the backward pass
    ssum = ((yt - out) ** 2).sum().detach().reshape(1)
    n = torch.tensor([float(len(y))], dtype=torch.float64)
    dist.all_reduce(ssum); dist.all_reduce(n)
    L = torch.sqrt(ssum / n + 1e-10)
    gout = ((out - yt) / (n * L)).detach()
    wrt = [v for v in leaves.values()] + list(rows.values())
    grads = torch.autograd.grad(out, wrt, grad_outputs=gout, allow_unused=True)
    for (k, v), g in zip(leaves.items(), grads[: len(leaves)]):
        if g is None:
            continue
        g = g.contiguous()
        dist.all_reduce(g)  # dense gradients: sum over ranks
        m._apply_dense(k, g)
    # touched rows: every rank gathers every rank's (ids, gradient rows) in rank order
    flat = torch.as_tensor(ids.reshape(-1), dtype=torch.long)
    all_ids = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(all_ids, flat)
    gid = torch.cat(all_ids).numpy()
    uniq, inv = np.unique(gid, return_inverse=True)
    for k, g in zip(rows, grads[len(leaves):]):
        vals = g.reshape(flat.shape[0], -1).contiguous()
        parts = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(parts, vals)
        allv = torch.cat(parts).numpy()
        summed = np.zeros((len(uniq), allv.shape[1]))
        np.add.at(summed, inv, allv)
        m._apply_sparse(k, uniq, torch.from_numpy(summed))
    return float(L)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    m = _model()
    ids, y = _batch()
    sid, sy = shard_batch(ids, y, rank, world)
    losses = [_dp_step(m, sid, sy, world) for _ in range(2)]
    q.put((rank, losses, {k: v.numpy().copy() for k, v in m.params.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_two_ranks_equal_one_rank_on_the_global_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=200) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
    ref = _model()
    ids, y = _batch()
    want = [ref.train_step(ids, y) for _ in range(2)]
    for rank, losses, params in res:
        assert np.allclose(losses, want, rtol=1e-10), (rank, losses, want)
        for k, v in ref.params.items():
            assert np.allclose(params[k], v.numpy(), rtol=1e-9, atol=1e-12), (rank, k)
    # replicas are identical to each other bit for bit
    for k in res[0][2]:
        assert np.array_equal(res[0][2][k], res[1][2][k]), k


# ---------------------------------------------------------------------------------------------
# Row-sharded tables (cffm_b200/csrc/shard.cu): the exchange protocol restated with numpy + gloo.
# rank r owns rows with row % world == r (local row = row // world).  Forward: unique ids of the local batch in
# owner-major order -> owners gather -> rows back; backward: gradient rows summed per unique id on the requesting rank
# first, sums to the owners, owner adds the ranks' contributions in rank order and applies the sparse update.
def _owner_major_unique(ids, world, mloc_max):
    flat = np.asarray(ids).reshape(-1).astype(np.int64)
    keys = (flat % world) * mloc_max + flat // world          # k_shard_keys
    uniq_keys, inv = np.unique(keys, return_inverse=True)     # sort + unique (ascending key = owner-major)
    return uniq_keys // mloc_max, uniq_keys % mloc_max, inv   # owner, local row at the owner, position of every id


def _sharded_step(m, ids, y, rank, world, M):
    """m.params tables hold THIS rank's rows only."""
    from oracle.cffm_ref import _GatherProxy
    mloc_max = (M + world - 1) // world
    owner, lrow, inv = _owner_major_unique(ids, world, mloc_max)
    # ---- ids to the owners (all_gather_object stands in for the all-to-all), rows back ----
    requests = [None] * world
    dist.all_gather_object(requests, [lrow[owner == o] for o in range(world)])
    asked = [requests[p][rank] for p in range(world)]                       # what every rank wants from me, rank-major
    replies = {k: [m.params[k][torch.as_tensor(a, dtype=torch.long)].numpy() for a in asked] for k in CFFMRef.SPARSE_TABLES}
    got = [None] * world
    dist.all_gather_object(got, replies)
    mini = {k: torch.from_numpy(np.concatenate([got[o][k][rank] for o in range(world)], axis=0)) for k in CFFMRef.SPARSE_TABLES}
    # ---- the step runs on the received rows with renumbered ids ----
    remap = torch.as_tensor(inv.reshape(np.asarray(ids).shape), dtype=torch.long)
    yt = torch.as_tensor(y, dtype=torch.float64).reshape(-1, 1)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in m.params.items() if k not in CFFMRef.SPARSE_TABLES}
    rows = {k: mini[k][remap].detach().clone().requires_grad_(True) for k in CFFMRef.SPARSE_TABLES}
    p2 = dict(leaves)
    for k in rows:
        p2[k] = _GatherProxy(mini[k], remap, rows[k])
    out = m.forward(remap.numpy(), p2)
    ssum = ((yt - out) ** 2).sum().detach().reshape(1)
    n = torch.tensor([float(len(y))], dtype=torch.float64)
    dist.all_reduce(ssum); dist.all_reduce(n)
    L = torch.sqrt(ssum / n + 1e-10)
    gout = ((out - yt) / (n * L)).detach()
    wrt = list(leaves.values()) + list(rows.values())
    grads = torch.autograd.grad(out, wrt, grad_outputs=gout, allow_unused=True)
    for (k, v), g in zip(leaves.items(), grads[: len(leaves)]):
        if g is None:
            continue
        g = g.contiguous()
        dist.all_reduce(g)
        m._apply_dense(k, g)
    # ---- gradient rows: per-unique-id sums on this rank, then to the owners ----
    U = len(lrow)
    sums = {}
    for k, g in zip(rows, grads[len(leaves):]):
        vals = g.reshape(inv.shape[0], -1).numpy()
        s = np.zeros((U, vals.shape[1]))
        np.add.at(s, inv, vals)                                            # order of appearance
        sums[k] = [s[owner == o] for o in range(world)]
    inbox = [None] * world
    dist.all_gather_object(inbox, sums)
    all_rows = np.concatenate(asked)                                        # rank-major, as received in the forward pass
    if len(all_rows):
        uniq, inv2 = np.unique(all_rows, return_inverse=True)
        for k in CFFMRef.SPARSE_TABLES:
            vals = np.concatenate([inbox[p][k][rank] for p in range(world)], axis=0)
            summed = np.zeros((len(uniq), vals.shape[1]))
            np.add.at(summed, inv2, vals)
            m._apply_sparse(k, uniq, torch.from_numpy(summed))
    return float(L)


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    m = _model()
    M = 80
    for k in CFFMRef.SPARSE_TABLES:
        m.params[k] = m.params[k][rank::world].clone()                      # this rank's rows
    ids, y = _batch()
    sid, sy = shard_batch(ids, y, rank, world)
    losses = [_sharded_step(m, sid, sy, rank, world, M) for _ in range(2)]
    q.put((rank, losses, {k: v.numpy().copy() for k, v in m.params.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_tables_equal_one_rank_on_the_global_batch(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=200) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
    ref = _model()
    ids, y = _batch()
    want = [ref.train_step(ids, y) for _ in range(2)]
    for rank, losses, params in res:
        assert np.allclose(losses, want, rtol=1e-10), (rank, losses, want)
        for k, v in ref.params.items():
            full = v.numpy()
            if k in CFFMRef.SPARSE_TABLES:
                assert np.allclose(params[k], full[rank::world], rtol=1e-9, atol=1e-12), (rank, k)   # the owner's rows
            else:
                assert np.allclose(params[k], full, rtol=1e-9, atol=1e-12), (rank, k)


def test_owner_major_keys_group_unique_rows_by_owner():
    rng = np.random.default_rng(0)
    M, world = 1003, 4
    ids = rng.integers(0, M, (64, 7))
    mmax = (M + world - 1) // world
    owner, lrow, inv = _owner_major_unique(ids, world, mmax)
    assert np.all(np.diff(owner) >= 0)                                    # contiguous per owner: one send per peer
    back = lrow * world + owner                                           # owner-side row number -> global row
    assert np.array_equal(back[inv].reshape(ids.shape), ids)
    assert len(back) == len(np.unique(ids))
    from cffm_b200.dist import owner_of
    o2, l2 = owner_of(back, world)
    assert np.array_equal(o2, owner) and np.array_equal(l2, lrow)
