"""Diagnostic: first two epochs of the 'stable' trajectory variant, step by step, in every arithmetic."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cffm_b200 import Engine, LoadData
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
z = np.load(os.path.join(G, "trajectory_frappe_mini.npz"))
d = LoadData(os.path.join(G, "frappe_mini") + "/", "frappe", "square_loss")
F, K, B, EPOCHS, SEED, BLOCK_SEED = [int(v) for v in z["meta"]]
for variant in ("stable", "ref"):
    for precision in ("fp32", "bf16x3", "bf16"):
        X, Y = np.array(d.Train_data["X"], dtype=np.int32), np.array(d.Train_data["Y"], dtype=np.float32)
        Xv, Yv = np.array(d.Validation_data["X"], dtype=np.int32), np.array(d.Validation_data["Y"], dtype=np.float32)
        eng = Engine(d.features_M, F, K, K, activation="selu", max_batch=B, precision=precision, seed=1)
        for name in eng.param_infos():
            eng.set_param(name, z["w0/" + name])
        if variant == "stable":
            for name, (shape, numel, _) in eng.param_infos().items():
                eng.set_slot(name, 1, np.full(numel, 0.1, dtype=np.float32))
        acc = eng.get_slot("outer_layer_conv_weight_0", 1)
        print(variant, precision, "acc0", float(acc.min()), float(acc.max()), "init rmse", eng.evaluate(Xv, Yv, B)[0], "graph", eng.uses_graph())
        rng = np.random.RandomState(BLOCK_SEED)
        n = len(Y)
        for ep in range(2):
            starts = [int(rng.randint(0, n - B)) for _ in range(n // B)]
            perm = np.random.RandomState(2021).permutation(n)
            X, Y = X[perm], Y[perm]
            losses = [eng.train_step(X[st:st + B], Y[st:st + B]) for st in starts]
            print("   epoch", ep, "losses", [round(l, 4) for l in losses], "rmse", round(eng.evaluate(Xv, Yv, B)[0], 5))
        eng.close()
