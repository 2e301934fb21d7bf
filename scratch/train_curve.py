import os, sys, numpy as np, time
sys.path.insert(0, os.getcwd())
from cffm_b200 import Engine, LoadData
d = LoadData("tests/golden/frappe_mini/", "frappe", "square_loss")
X, Y = np.array(d.Train_data["X"]), np.array(d.Train_data["Y"])
Xv, Yv = np.array(d.Validation_data["X"]), np.array(d.Validation_data["Y"])
for prec in ("fp32", "bf16"):
    for seed in (11, 12):
        eng = Engine(d.features_M, 10, 32, 32, activation="selu", max_batch=256, precision=prec, seed=seed)
        rng = np.random.RandomState(0)
        curve = []
        t0 = time.time()
        for ep in range(40):
            for s in range(11):
                st = rng.randint(0, 3000 - 256)
                eng.train_step(X[st:st + 256], Y[st:st + 256])
            if ep % 4 == 3:
                curve.append(round(eng.evaluate(Xv, Yv, 256)[0], 4))
        print(prec, seed, curve, round(time.time() - t0, 2), "s")
        eng.close()
