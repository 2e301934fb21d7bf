"""Generates ``trajectory_frappe_mini.npz``: validation-RMSE curves of the CPU oracle trained with the reference's
loop (CFFM.py:181-207: per-epoch sklearn shuffle with random_state=2021, ``int(N/batch)`` contiguous blocks starting
at ``randint(0, N - batch)``) on the Frappe fixture, PR1 configuration (F=10, K=32, B=256, selu, Adagrad lr 0.05,
BASELINE.json configs[0]), 50 epochs (README.md:18-28), in TWO arithmetics (fp64 and fp32) and TWO accumulator
settings:

  ``ref``    initial_accumulator_value = 1e-8, the reference's (CFFM.py:523-524)
  ``stable`` the same run started from accumulators 0.1 (TensorFlow's default), loaded through the checkpoint API

The fp64-vs-fp32 gap of the SAME oracle is the arithmetic noise floor of this training run: the test
(tests/test_gpu_trajectory.py) holds the CUDA paths against these curves and against that floor.
The oracle needs ~4 min per run on 8 cores, so the curves are committed rather than recomputed on the GPU box.

    python tests/golden/make_trajectory.py            # all four runs (two at a time)
"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

F, K, B, EPOCHS, SEED, BLOCK_SEED = 10, 32, 256, 50, 11, 0


def load_split():
    from oracle.libfm_ref import LoadDataRef
    d = LoadDataRef(os.path.join(HERE, "frappe_mini") + "/", "frappe", "square_loss")
    X, Y = np.array(d.Train_data["X"], dtype=np.int32), np.array(d.Train_data["Y"], dtype=np.float64)
    Xv, Yv = np.array(d.Validation_data["X"], dtype=np.int32), np.array(d.Validation_data["Y"], dtype=np.float64)
    return d.features_M, X, Y, Xv, Yv


def batch_plan(n):
    """The block starts of all epochs: one RandomState(BLOCK_SEED) stream (the reference draws from the unseeded
    global numpy RNG, Q12)."""
    rng = np.random.RandomState(BLOCK_SEED)
    return [[int(rng.randint(0, n - B)) for _ in range(n // B)] for _ in range(EPOCHS)]


def run(args):
    dtype_name, acc0 = args
    import torch
    from oracle.cffm_ref import CFFMRef
    torch.set_num_threads(4)
    dt = {"fp64": torch.float64, "fp32": torch.float32}[dtype_name]
    M, X, Y, Xv, Yv = load_split()
    m = CFFMRef(M, F, K, K, activation="selu", dtype=torch.float64, seed=SEED)
    init = {k: v.numpy().copy() for k, v in m.params.items()}
    for k in m.params:
        m.params[k] = m.params[k].to(dt)
    m.dtype = dt
    m.state["accumulator"] = {k: torch.full_like(v, acc0) for k, v in m.params.items()}
    plan = batch_plan(len(Y))
    curve = []
    for ep in range(EPOCHS):
        perm = np.random.RandomState(2021).permutation(len(Y))      # CFFM.py:183, :556-558
        X, Y = X[perm], Y[perm]
        for st in plan[ep]:
            m.train_step(X[st:st + B], Y[st:st + B])
        with torch.no_grad():
            p = m.predict(Xv).reshape(-1).double().numpy()
        pb = np.clip(p, Yv.min(), Yv.max())                          # CFFM.py:609-611
        curve.append(float(np.sqrt(np.mean((Yv - pb) ** 2))))
        print(dtype_name, acc0, ep, curve[-1], flush=True)
    return dtype_name, acc0, np.array(curve), init


if __name__ == "__main__":
    jobs = [("fp64", 1e-8), ("fp32", 1e-8), ("fp64", 0.1), ("fp32", 0.1)]
    out = {}
    with ProcessPoolExecutor(max_workers=2) as ex:
        for dtype_name, acc0, curve, init in ex.map(run, jobs):
            out["curve/%s/%s" % ("ref" if acc0 < 1e-4 else "stable", dtype_name)] = curve
            for k, v in init.items():
                out["w0/" + k] = v.astype(np.float32)
    out["meta"] = np.array([F, K, B, EPOCHS, SEED, BLOCK_SEED], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "trajectory_frappe_mini.npz"), **out)
    for k in sorted(out):
        if k.startswith("curve/"):
            print(k, np.round(out[k][[0, 1, 2, 4, 9, 19, 29, 39, 49]], 4).tolist())
