"""Generates the committed golden fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

1. ``libfm_golden.npz`` -- outputs of the REFERENCE's own ``LoadData`` (imported from
   /root/reference, pure Python + numpy) on ``tests/golden/frappe_mini`` (slices of the reference's
   frappe.validation/test libfm files) and on a small ragged file.  This pins the loader.
2. ``cffm_golden.npz`` -- outputs of the fp64 oracle (oracle/cffm_ref.py) on small seeded cases:
   weights, ids, labels -> out, loss, gradients, weights after two Adagrad steps.  The reference
   graph itself cannot be executed (TensorFlow 1.14 is not installable), so these vectors pin the
   CUDA path to the oracle, not to the reference: parity stays "unpinned" for the model.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [
    # name, F, K, M, B, activation, loss, linear_att, inner, outer
    ("f4k8_selu", 4, 8, 50, 6, "selu", "square_loss", 1, 1, 1),
    ("f3k32_elu", 3, 32, 200, 9, "elu", "square_loss", 1, 1, 1),
    ("f10k32_selu", 10, 32, 300, 8, "selu", "square_loss", 1, 1, 1),
    ("f6k16_relu_log", 6, 16, 120, 7, "relu", "log_loss", 1, 1, 1),
    ("f5k8_gelu_mse", 5, 8, 64, 5, "gelu", "mse", 0, 1, 1),
    ("f4k16_prelu_outer", 4, 16, 40, 6, "prelu", "mae", 1, 0, 1),
    ("f7k8_elu_inner", 7, 8, 90, 10, "elu", "hybrid", 1, 1, 0),
    # outer_dims 16 and 64 of the scoring sweep (BASELINE.json configs[4]; conv_depth = int(log2 K), CFFM.py:373;
    # inner flatten generalised from the literal 16*2)
    ("f6k64_selu", 6, 64, 150, 5, "selu", "square_loss", 1, 1, 1),
    ("f10k16_relu", 10, 16, 200, 8, "relu", "square_loss", 1, 1, 1),
    ("f3k64_gelu_log", 3, 64, 80, 6, "gelu", "log_loss", 1, 1, 1),
]


def make_ragged(dirname):
    os.makedirs(os.path.join(dirname, "rag"), exist_ok=True)
    rows = {
        "train": ["1 a:1 b:1 c:1", "-1 a:1 d:1", "1 e:1 b:1 f:2", "0.5 a:1  b:1", "-1 g:1"],
        "validation": ["1 b:1 a:1 h:1", "-1 d:1 a:1"],
        "test": ["-1 c:1 i:1", "1 a:1 b:1 c:1", "1 j:0.5"],
    }
    for k, v in rows.items():
        with open(os.path.join(dirname, "rag", "rag.%s.libfm" % k), "w") as f:
            f.write("\n".join(v) + "\n")


def golden_libfm():
    sys.path.insert(0, "/root/reference")
    import LoadData as REF  # the reference's own loader
    out = {}
    make_ragged(os.path.join(HERE, "ragged"))
    for tag, path, ds in (("frappe", os.path.join(HERE, "frappe_mini") + "/", "frappe"),
                          ("rag", os.path.join(HERE, "ragged") + "/", "rag")):
        for loss in ("square_loss", "log_loss"):
            with contextlib.redirect_stdout(io.StringIO()):
                d = REF.LoadData(path, ds, loss)
            out["%s.%s.features_M" % (tag, loss)] = np.int64(d.features_M)
            toks = sorted(d.features.items(), key=lambda kv: kv[1])
            out["%s.%s.tokens" % (tag, loss)] = np.array([t for t, _ in toks])
            for name, split in (("train", d.Train_data), ("validation", d.Validation_data), ("test", d.Test_data)):
                lens = np.array([len(r) for r in split["X"]], dtype=np.int64)
                out["%s.%s.%s.lens" % (tag, loss, name)] = lens
                out["%s.%s.%s.ids" % (tag, loss, name)] = np.array([i for r in split["X"] for i in r], dtype=np.int32)
                out["%s.%s.%s.y" % (tag, loss, name)] = np.array(split["Y"], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "libfm_golden.npz"), **out)
    print("libfm_golden.npz:", len(out), "arrays")


def golden_model():
    from oracle.cffm_ref import CFFMRef
    out = {}
    for (name, F, K, M, B, act, loss, la, ic, oc) in CASES:
        m = CFFMRef(M, F, K, K, activation=act, loss_type=loss, linear_att=la, inner_conv=ic, outer_conv=oc,
                    dtype=torch.float64, seed=7)
        g = torch.Generator().manual_seed(11)
        # feature_bias is zero-initialised in the reference; give it values so the linear term is exercised
        m.params["feature_bias"] = torch.randn(M, 1, generator=g, dtype=torch.float64) * 0.3
        if oc:  # larger outer rows so the conv path is not numerically negligible
            std = 0.05 if loss == "log_loss" else 0.3
            m.params["outer_embeddings"] = torch.randn(M, K, generator=g, dtype=torch.float64) * std
        if loss == "hybrid":  # log(out) and log(1-out) need the raw output inside (0,1)
            m.params["bias"] = torch.tensor(0.5, dtype=torch.float64)
        rng = np.random.default_rng(5)
        ids = rng.integers(0, M, size=(B, F)).astype(np.int32)
        ids[1] = ids[0]  # duplicated rows inside the batch
        y = rng.choice([-1.0, 1.0], size=B)
        if loss in ("log_loss",):
            y = (y > 0).astype(np.float64)
        if loss == "hybrid":  # keep log(out), log(1-out) finite: the raw output must lie in (0,1)
            y = (y > 0).astype(np.float64)
        for k, v in m.params.items():
            out["%s/w0/%s" % (name, k)] = v.numpy().copy()
        out[name + "/ids"] = ids
        out[name + "/y"] = y
        o, inter = m.forward(ids, return_intermediates=True)
        out[name + "/out"] = o.numpy().reshape(-1)
        out[name + "/pred"] = m.predict(ids).numpy().reshape(-1)
        for k in ("final2", "final", "linear", "t1"):
            if k in inter:
                out["%s/%s" % (name, k)] = inter[k].numpy()
        l, dense, sparse = m.gradients(ids, y)
        out[name + "/loss"] = np.float64(l)
        for k, gk in dense.items():
            if gk is not None:
                out["%s/g/%s" % (name, k)] = gk.numpy()
        for k, (rows, summed, vals) in sparse.items():
            out["%s/grows/%s" % (name, k)] = vals.numpy()
            out["%s/uniq" % name] = rows
        losses = [m.train_step(ids, y), m.train_step(ids, y)]
        out[name + "/losses"] = np.array(losses)
        for k, v in m.params.items():
            out["%s/w2/%s" % (name, k)] = v.numpy().copy()
        for k, v in m.accumulators().items():
            out["%s/a2/%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "cffm_golden.npz"), **out)
    print("cffm_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    golden_libfm()
    golden_model()
