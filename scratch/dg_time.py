"""Time the layer-0 data gradient of a bf16x3 training step at the Criteo shape (profile report)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from cffm_b200 import Engine, synth
B = 8192
ids, M = synth.make_ids("criteo", B, seed=1)
y = synth.make_labels(B, seed=1)
eng = Engine(M, 39, 32, 32, activation="relu", max_batch=B, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16x3", seed=1)
eng.train_step(ids, y)
eng.profile(True); eng.profile_report(reset=True)
for _ in range(3): eng.train_step(ids, y)
rep = eng.profile_report(reset=True)
print(os.environ.get("CFFM_DFACT_ABLATE", "0"), {k: round(ms / n, 3) for k, (n, ms) in rep.items() if k in ("conv_dgrad_l0", "conv_wgrad_l0", "conv_fwd_l0")})
eng.close()
