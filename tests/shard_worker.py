"""Launched by tests/test_gpu_multi.py under torch.distributed.run: training with ROW-SHARDED tables on WORLD_SIZE
GPUs (all-to-all of rows / gradient sums, cffm_b200/csrc/shard.cu) must equal single-GPU training on the global
batch with replicated tables."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cffm_b200 import Engine  # noqa: E402
from cffm_b200.dist import bind_engine, gather_table, shard_batch  # noqa: E402


def main():
    out_path, precision = sys.argv[1], sys.argv[2]
    optimizer = sys.argv[3] if len(sys.argv) > 3 else "AdagradOptimizer"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M, F, B = 701, 10, 64 * world                     # M not a multiple of world: shards of unequal size
    rng = np.random.default_rng(4)
    ids = rng.integers(0, M, (3, B, F)).astype(np.int32)
    ids[:, 5] = ids[:, B - 3]                         # rows touched by several ranks
    ids[:, :, 0] = ids[:, :, 0] % 3                   # a heavily duplicated field
    y = rng.choice([-1.0, 1.0], (3, B)).astype(np.float32)
    fb = rng.normal(0, 0.05, (M, 1)).astype(np.float32)
    lr = 0.05 if optimizer == "AdagradOptimizer" else 0.01
    eng = Engine(M, F, 32, 32, activation="selu", max_batch=B // world, precision=precision, device=local, seed=7,
                 optimizer=optimizer, lr=lr, shard=(rank, world))
    eng.set_table_from_global("feature_bias", fb)
    bind_engine(eng, dist)
    res = {"local_rows": int(eng.get_param("inner_embeddings").shape[0])}
    losses = []
    for s in range(3):
        sid, sy = shard_batch(ids[s], y[s], rank, world)
        losses.append(eng.train_step(sid, sy))
    # scoring is collective too: every rank scores its part of the last batch
    sid, sy = shard_batch(ids[2], y[2], rank, world)
    pred = eng.forward(sid)
    res["losses"] = losses
    # checkpoint of a sharded run: every rank saves / restores its own shard (+ replicated dense variables, slots, step)
    state = eng.state_dict()
    eng2 = Engine(M, F, 32, 32, activation="selu", max_batch=B // world, precision=precision, device=local, seed=99,
                  optimizer=optimizer, lr=lr, shard=(rank, world))
    bind_engine(eng2, dist)
    eng2.load_state_dict(state)
    sid0, sy0 = shard_batch(ids[0], y[0], rank, world)
    la, lb = eng.train_step(sid0, sy0), eng2.train_step(sid0, sy0)
    sa, sb = eng.state_dict(), eng2.state_dict()
    res["resume_identical"] = bool(la == lb and all(np.array_equal(np.asarray(sa[k]), np.asarray(sb[k])) for k in sa))
    res["shard_rows_in_state"] = int(state["w:inner_embeddings"].shape[0])
    eng2.close()
    ok = torch.tensor([1.0 if res["resume_identical"] else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res["resume_identical"] = bool(ok.item() == 1.0)
    w = {k: gather_table(eng, k, dist) for k in eng.param_infos()}
    preds = [None] * world
    dist.all_gather_object(preds, pred)
    if rank == 0:
        ref = Engine(M, F, 32, 32, activation="selu", max_batch=B, precision=precision, device=local, seed=7,
                     optimizer=optimizer, lr=lr)
        # identical initial weights: the sharded initialiser draws the values of the replicated layout
        ref.set_param("feature_bias", fb)
        ref_losses = [ref.train_step(ids[s], y[s]) for s in range(3)]
        ref_pred = ref.forward(ids[2])
        ref.train_step(ids[0], y[0])                       # the extra step of the resume check above
        rw = ref.get_weights()
        res["ref_losses"] = ref_losses
        res["pred_max_abs_diff"] = float(np.max(np.abs(np.concatenate(preds) - ref_pred)))
        res["pred_scale"] = float(np.max(np.abs(ref_pred)))
        res["max_abs_diff"] = {k: float(np.max(np.abs(w[k].astype(np.float64) - rw[k]))) for k in w}
        res["frac_over_2e-3"] = {k: float(np.mean(np.abs(w[k].astype(np.float64) - rw[k]) > 2e-3)) for k in w}
        touched = np.unique(ids)
        mask = np.ones(M, dtype=bool); mask[touched] = False
        fresh = Engine(M, F, 32, 32, activation="selu", max_batch=4, precision=precision, device=local, seed=7)
        w0 = fresh.get_param("inner_embeddings")
        res["untouched_rows_bit_identical"] = bool(np.array_equal(w["inner_embeddings"][mask], w0[mask]))
        res["init_identical"] = bool(np.array_equal(w["outer_embeddings"][mask], fresh.get_param("outer_embeddings")[mask]))
        with open(out_path, "w") as f:
            json.dump(res, f)
        ref.close(); fresh.close()
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
