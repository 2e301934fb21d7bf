"""Two eager training steps at the Criteo shape for ncu (CFFM_GRAPH=0)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
os.environ["CFFM_GRAPH"] = "0"
from cffm_b200 import Engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
ids, M = synth.make_ids("criteo", 2 * B, seed=1)
y = synth.make_labels(2 * B, seed=1)
eng = Engine(M, 39, 32, 32, activation="relu", max_batch=B, precision=prec, seed=1)
for s in range(2):
    print("loss", eng.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B]))
eng.close()
