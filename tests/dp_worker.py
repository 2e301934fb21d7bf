"""Launched by tests/test_gpu_multi.py under torch.distributed.run: data-parallel training on
WORLD_SIZE GPUs must equal single-GPU training on the global batch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cffm_b200 import Engine  # noqa: E402
from cffm_b200.dist import bind_engine, shard_batch  # noqa: E402


def main():
    out_path, precision = sys.argv[1], sys.argv[2]
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M, F, B = 700, 10, 64 * world
    rng = np.random.default_rng(4)
    ids = rng.integers(0, M, (3, B, F)).astype(np.int32)
    ids[:, 5] = ids[:, B - 3]          # rows touched by several ranks
    y = rng.choice([-1.0, 1.0], (3, B)).astype(np.float32)
    fb = rng.normal(0, 0.05, (M, 1)).astype(np.float32)
    eng = Engine(M, F, 32, 32, activation="selu", max_batch=B // world, precision=precision, device=local, seed=7)
    eng.set_param("feature_bias", fb)
    bind_engine(eng, dist)
    losses = []
    for s in range(3):
        sid, sy = shard_batch(ids[s], y[s], rank, world)
        losses.append(eng.train_step(sid, sy))
    w = eng.get_weights()
    res = {"losses": losses}
    # replicas must agree bit for bit: compare a checksum of every tensor across ranks
    sums = torch.tensor([float(np.float64(v.astype(np.float64).sum())) for v in w.values()], device="cuda", dtype=torch.float64)
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["replicas_identical"] = bool(torch.equal(lo, hi))
    if rank == 0:
        ref = Engine(M, F, 32, 32, activation="selu", max_batch=B, precision=precision, device=local, seed=7)
        ref.set_param("feature_bias", fb)
        ref_losses = [ref.train_step(ids[s], y[s]) for s in range(3)]
        rw = ref.get_weights()
        res["ref_losses"] = ref_losses
        res["max_abs_diff"] = {k: float(np.max(np.abs(w[k].astype(np.float64) - rw[k]))) for k in w}
        res["frac_over_2e-3"] = {k: float(np.mean(np.abs(w[k].astype(np.float64) - rw[k]) > 2e-3)) for k in w}
        with open(out_path, "w") as f:
            json.dump(res, f)
        ref.close()
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
