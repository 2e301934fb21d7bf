"""Time the bf16 forward pass kernels at the Criteo shape (profile report of forward-only calls)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from cffm_b200 import Engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
PREC = sys.argv[2] if len(sys.argv) > 2 else "bf16"
ids, M = synth.make_ids("criteo", B, seed=1)
eng = Engine(M, 39, 32, 32, activation="relu", max_batch=B, precision=PREC, seed=1)
out0 = eng.forward(ids)
eng.profile(True); eng.profile_report(reset=True)
for _ in range(5): out = eng.forward(ids)
rep = eng.profile_report(reset=True)
for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:4]: print(k, n, round(ms / n, 4))
print("checksum", float(np.abs(out).sum()), bool(np.array_equal(out, out0)))
eng.close()
