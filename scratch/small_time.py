"""Event-bracketed per-kernel times of a training step at one of the reference's dataset shapes."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from cffm_b200 import Engine, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "frappe"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
w = synth.WORKLOADS[wl]; B = w["batch"]
ids, M = synth.make_ids(wl, 8 * B, seed=1); y = synth.make_labels(8 * B, seed=1)
eng = Engine(M, ids.shape[1], 32, 32, activation=w["activation"], max_batch=B, precision=prec, seed=1)
for s in range(3): eng.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B])
eng.profile(True); eng.profile_report(reset=True)
for s in range(3, 8): eng.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B])
rep = eng.profile_report(reset=True)
tot = 0.0
for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print("%-22s %7.1f us x%d" % (k, 1e3 * ms / 5, n // 5)); tot += ms / 5
print("total %.1f us" % (1e3 * tot))
eng.close()
