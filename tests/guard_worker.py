"""Run under CFFM_GUARD=1 (tests/test_gpu_guard.py): every device allocation of the library carries a 4 KB pattern band on
either side; after forward / training / scoring in every arithmetic and shape class (partial tiles, odd batches, padded
channels, batches below max_batch, both layer-0 forms, K = 16 / 32 / 64) no band may have been written to."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cffm_b200 import Engine, _lib  # noqa: E402


def check(lib, what, report):
    msg = C.create_string_buffer(256)
    n = lib.cffm_debug_check_guards(msg, 256)
    report.append({"after": what, "damaged": int(n), "msg": msg.value.decode()})
    return n


def main():
    assert os.environ.get("CFFM_GUARD") == "1"
    lib = _lib.load()
    rng = np.random.default_rng(0)
    report = []
    cases = [
        # F, K, M, max_batch, batches, activation, precision, fact
        (10, 32, 400, 64, (64, 37, 1), "selu", "fp32", False),
        (4, 8, 50, 16, (16, 5), "gelu", "fp32", False),
        (6, 64, 300, 9, (9, 4), "relu", "fp32", False),
        (10, 32, 400, 64, (64, 37, 1), "selu", "bf16", False),
        (12, 32, 500, 40, (40, 7), "prelu", "bf16", True),
        (39, 32, 2000, 12, (12, 5), "relu", "bf16", True),
        (30, 32, 900, 9, (9, 3), "relu", "bf16", False),
        (10, 32, 400, 64, (64, 37, 1), "selu", "bf16x3", False),
        (20, 32, 700, 24, (24, 9), "elu", "bf16x3", True),
        (39, 32, 2000, 12, (12, 5), "relu", "bf16x3", False),
    ]
    for (F, K, M, MB, batches, act, prec, fact) in cases:
        if fact:
            os.environ["CFFM_FACT_MIN_BATCH"], os.environ["CFFM_FACT_MIN_FIELDS"] = "1", "1"
        else:
            os.environ["CFFM_FACT_MIN_BATCH"] = "1000000000"; os.environ.pop("CFFM_FACT_MIN_FIELDS", None)
        eng = Engine(M, F, K, K, activation=act, max_batch=MB, precision=prec, seed=1)
        eng.set_param("feature_bias", rng.normal(0, 0.05, (M, 1)).astype(np.float32))
        for B in batches:
            ids = rng.integers(0, M, (B, F)).astype(np.int32)
            ids[:, 0] = ids[0, 0]                                   # one id in every sample: a long segment
            y = rng.choice([-1.0, 1.0], B).astype(np.float32)
            eng.forward(ids)
            if not (prec != "fp32" and (K != 32 or act == "gelu")):
                eng.train_step(ids, y)
                eng.train_step(ids, y)
            eng.evaluate(ids, y, max(1, B // 2))
        tag = "F%d K%d %s %s %s" % (F, K, act, prec, "factorised" if fact else "direct")
        check(lib, tag, report)
        eng.close()
    # scoring on the tensor cores for K = 16 / 64
    os.environ["CFFM_FACT_MIN_BATCH"] = "1000000000"
    for (F, K, B, prec) in ((10, 16, 70, "bf16"), (39, 16, 33, "bf16x3"), (10, 64, 9, "bf16x3"), (6, 64, 20, "bf16")):
        eng = Engine(500, F, K, K, activation="relu", max_batch=B, precision=prec, seed=2)
        ids = rng.integers(0, 500, (B, F)).astype(np.int32)
        eng.forward(ids); eng.forward(ids[: B // 2 + 1])
        check(lib, "scoring F%d K%d %s" % (F, K, prec), report)
        eng.close()
    # resident dataset + pipelined submit
    eng = Engine(400, 10, 32, 32, activation="selu", max_batch=128, precision="bf16", seed=3)
    ids = rng.integers(0, 400, (1000, 10)).astype(np.int32); y = rng.choice([-1.0, 1.0], 1000).astype(np.float32)
    eng.dataset_upload(ids, y); eng.dataset_permute(rng.permutation(1000))
    for st in (0, 500, 872):
        eng.train_block(st, 128)
    eng.dataset_evaluate(128)
    for s in range(3):
        eng.train_submit(ids[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128])
    eng.train_flush()
    check(lib, "resident dataset + pipelined submit", report)
    eng.close()
    json.dump(report, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
