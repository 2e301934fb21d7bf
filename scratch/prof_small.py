"""A few eager training steps of one workload shape for ncu (CFFM_GRAPH=0): prof_small.py <workload> <precision> <steps>."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
os.environ["CFFM_GRAPH"] = "0"
from cffm_b200 import Engine, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "frappe"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = synth.WORKLOADS[wl]
B = int(sys.argv[4]) if len(sys.argv) > 4 else w["batch"]
ids, M = synth.make_ids(wl, steps * B, seed=1)
y = synth.make_labels(steps * B, seed=1)
eng = Engine(M, ids.shape[1], 32, 32, activation=w["activation"], max_batch=B, precision=prec, seed=1)
for s in range(steps):
    print("loss", eng.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B]))
print("launches", eng.launch_count())
eng.close()
