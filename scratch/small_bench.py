"""Graph-replayed training steps at the reference's dataset shapes: ms per step (pipelined host API, no L2 flush)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from cffm_b200 import Engine, synth
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
out = []
for wl in ("frappe", "ml-tag", "book-crossing"):
    w = synth.WORKLOADS[wl]; B = w["batch"]
    ids, M = synth.make_ids(wl, 16 * B, seed=1); y = synth.make_labels(16 * B, seed=1)
    eng = Engine(M, ids.shape[1], 32, 32, activation=w["activation"], max_batch=B, precision=prec, seed=1)
    pool = [(np.ascontiguousarray(ids[i * B:(i + 1) * B]), np.ascontiguousarray(y[i * B:(i + 1) * B])) for i in range(16)]
    for i in range(50): eng.train_submit(*pool[i % 16])
    eng.train_flush(); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(400): eng.train_submit(*pool[i % 16])
        eng.train_flush(); torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 400 * 1e3)
    out.append("%s %.4f" % (wl, best))
    eng.close()
print(os.environ.get("CFFM_SIDE_MASK", "15"), prec, " | ".join(out))
