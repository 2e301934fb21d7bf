"""Checks at BASELINE.json's full sizes.  Where the oracle finishes in seconds (ml-tag, Book-Crossing,
Frappe shapes) the full batch is compared directly; at the Criteo shape (F=39, B=8192, 10M rows) a random
subset of the batch is compared with the oracle (samples are independent in the forward pass) and the
update is checked through size-independent properties: untouched rows bit-identical, accumulator growth
equal to the squared segment sums of the gradient rows, run-to-run determinism."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def _oracle_from(eng, M, F, act):
    from oracle.cffm_ref import CFFMRef
    ref = CFFMRef(M, F, 32, 32, activation=act, dtype=torch.float64)
    for k, v in eng.get_weights().items():
        ref.params[k] = torch.from_numpy(v.astype(np.float64)).reshape(ref.params[k].shape)
    return ref


@pytest.mark.parametrize("name", ["ml-tag", "book-crossing", "frappe"])
def test_full_batch_forward_and_loss_vs_oracle(name):
    from cffm_b200 import Engine, synth
    w = synth.make_workload(name, n=synth.WORKLOADS[name]["batch"])
    F, M, B = w["num_field"], w["features_M"], w["batch"]
    eng = Engine(M, F, 32, 32, activation=w["activation"], max_batch=B, seed=2)
    eng.set_param("feature_bias", np.random.default_rng(0).normal(0, 0.05, (M, 1)).astype(np.float32))
    ref = _oracle_from(eng, M, F, w["activation"])
    out = eng.forward(w["ids"])
    assert _rel(out, ref.predict(w["ids"]).numpy()) < 1e-4
    loss = eng.train_step(w["ids"], w["labels"])
    want = float(ref.loss(w["ids"], w["labels"]))
    assert abs(loss - want) < 1e-4 * max(1.0, want)
    assert int(eng.fetch("n_uniq")[0]) == len(np.unique(w["ids"]))
    eng.close()


@pytest.mark.timeout(280)
def test_criteo_shape_properties():
    from cffm_b200 import Engine, synth
    B, F, K = 8192, 39, 32
    ids, M = synth.make_ids("criteo", B, seed=5)
    y = synth.make_labels(B, seed=5)
    assert M == 10_000_000
    eng = Engine(M, F, K, K, activation="relu", max_batch=B, precision="bf16", seed=3)
    rng = np.random.default_rng(1)
    P = F * (F - 1) // 2
    for l in range(5):  # O(1) activations through the stack (the default N(0,1) filters overflow nothing but are extreme)
        eng.set_param("outer_layer_conv_weight_%d" % l, rng.normal(0, 1 / np.sqrt(4 * P), (2, 2, P, P)).astype(np.float32))
    # ---- forward: a random subset of the batch against the oracle (per-sample independence) ----
    out = eng.forward(ids)
    pick = rng.choice(B, 12, replace=False)
    uniq_pick = np.unique(ids[pick])
    remap = {int(r): i for i, r in enumerate(uniq_pick)}
    from oracle.cffm_ref import CFFMRef
    ref = CFFMRef(len(uniq_pick), F, K, K, activation="relu", dtype=torch.float64)
    wts = {k: eng.get_param(k) for k in eng.param_infos() if k not in ("inner_embeddings", "outer_embeddings", "feature_bias")}
    for k, v in wts.items():
        ref.params[k] = torch.from_numpy(v.astype(np.float64)).reshape(ref.params[k].shape)
    tabs = {k: eng.get_param(k) for k in ("inner_embeddings", "outer_embeddings", "feature_bias")}
    for k, v in tabs.items():
        ref.params[k] = torch.from_numpy(v[uniq_pick].astype(np.float64))
    small_ids = np.vectorize(remap.get)(ids[pick])
    want = ref.predict(small_ids).numpy().reshape(-1)
    assert _rel(out[pick], want) < 1e-2
    # batch-split invariance: scoring a slice alone gives the same numbers as inside the big batch
    part = eng.forward(ids[1000:1512])
    assert np.array_equal(part, out[1000:1512])
    # ---- one training step: untouched rows, accumulator identity, determinism ----
    loss = eng.train_step(ids, y)
    assert np.isfinite(loss)
    touched = np.unique(ids)
    assert int(eng.fetch("n_uniq")[0]) == len(touched)
    g_rows = eng.fetch("grad_outer_rows").reshape(B * F, K)
    new_tab, new_acc = eng.get_param("outer_embeddings"), eng.get_param("outer_embeddings", accum=True)
    mask = np.ones(M, dtype=bool); mask[touched] = False
    assert np.array_equal(new_tab[mask], tabs["outer_embeddings"][mask])     # untouched rows: bit identical
    assert np.all(new_acc[mask] == np.float32(1e-8))
    order = np.argsort(ids.reshape(-1), kind="stable")
    sums = np.add.reduceat(g_rows[order].astype(np.float64), np.searchsorted(ids.reshape(-1)[order], touched), axis=0)
    want_acc = 1e-8 + sums ** 2                                              # acc += (segment sum)^2
    assert np.allclose(new_acc[touched], want_acc, rtol=2e-4, atol=1e-12)
    step = tabs["outer_embeddings"][touched] - new_tab[touched]
    assert np.allclose(step, 0.05 * sums / np.sqrt(want_acc), rtol=1e-3, atol=1e-6)  # w -= lr g / sqrt(acc)
    eng2 = Engine(M, F, K, K, activation="relu", max_batch=B, precision="bf16", seed=3)
    for l in range(5):
        eng2.set_param("outer_layer_conv_weight_%d" % l, wts["outer_layer_conv_weight_%d" % l])
    loss2 = eng2.train_step(ids, y)
    assert loss2 == loss                                                     # run-to-run determinism
    eng.close(); eng2.close()


@pytest.mark.timeout(580)
@pytest.mark.parametrize("precision,layer0", [("bf16", "factorised"), ("bf16", "direct"), ("bf16x3", "factorised"), ("fp32", "direct")])
def test_criteo_field_count_gradients_vs_oracle(precision, layer0, monkeypatch):
    """F = 39 (741 pairs, 768 padded channels, three N tiles) with a batch that spans several 8-sample tiles and partial
    ones (B = 44): logits, every conv gradient and the embedding-row gradients against the fp64 oracle -- the gradient
    comparison at the Criteo field count that the B = 6 parity case is too small for."""
    _gradients_vs_oracle(39, 44, precision, layer0, monkeypatch)


@pytest.mark.timeout(580)
@pytest.mark.parametrize("F", [20, 24, 32, 33])
@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_factorised_layer0_operand_widths(F, precision, monkeypatch):
    """The factorised layer-0 kernels pad 2F to a multiple of 16 (KA): F = 20 / 24 / 32 / 33 give KA = 48 / 48 / 64 / 80,
    i.e. one 64-column operand block, a full one, and the second block barely used -- the operand descriptors (K-major
    and MN-major slabs, the E / Z split offsets) at widths the dataset shapes (KA = 16, 32, 80) do not reach."""
    _gradients_vs_oracle(F, 20, precision, "factorised", monkeypatch)


def _gradients_vs_oracle(F, B, precision, layer0, monkeypatch):
    from cffm_b200 import Engine
    from oracle.cffm_ref import CFFMRef
    if layer0 == "factorised":
        monkeypatch.setenv("CFFM_FACT_MIN_BATCH", "1"); monkeypatch.setenv("CFFM_FACT_MIN_FIELDS", "1")
    else:
        monkeypatch.setenv("CFFM_FACT_MIN_BATCH", "1000000000")
    K, M = 32, 3000
    rng = np.random.default_rng(11)
    eng = Engine(M, F, K, K, activation="relu", max_batch=B, precision=precision, seed=3)
    eng.set_param("feature_bias", rng.normal(0, 0.05, (M, 1)).astype(np.float32))
    eng.set_param("outer_embeddings", rng.normal(0, 0.3, (M, K)).astype(np.float32))
    P = F * (F - 1) // 2
    for l in range(5):
        eng.set_param("outer_layer_conv_weight_%d" % l, rng.normal(0, 1 / np.sqrt(4 * P), (2, 2, P, P)).astype(np.float32))
    ref = _oracle_from(eng, M, F, "relu")
    ids = rng.integers(0, M, (B, F)).astype(np.int32)
    ids[:, 0] = ids[:, 0] % 5                     # heavy duplication in one field
    y = rng.choice([-1.0, 1.0], B).astype(np.float32)
    tol = {"fp32": 1e-4, "bf16x3": 1e-4, "bf16": 1e-2}[precision]
    out = eng.forward(ids)
    assert _rel(out, ref.predict(ids).numpy()) < tol
    l_ref, dense, sparse = ref.gradients(ids, y)
    loss = eng.train_step(ids, y)
    assert abs(loss - float(l_ref)) < tol * max(1.0, float(l_ref))

    def rel2(a, b):
        a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
        return float(np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b)))
    gtol = {"fp32": 1e-3, "bf16x3": 5e-3, "bf16": 5e-2}[precision]
    errs = {}
    for l in range(4):
        errs["wgrad%d" % l] = rel2(eng.dense_grad("outer_layer_conv_weight_%d" % l), dense["outer_layer_conv_weight_%d" % l].numpy())
        errs["bgrad%d" % l] = rel2(eng.dense_grad("outer_layer_conv_bias_%d" % l), dense["outer_layer_conv_bias_%d" % l].numpy())
    errs["outer_rows"] = rel2(eng.fetch("grad_outer_rows"), sparse["outer_embeddings"][2].numpy())
    errs["inner_rows"] = rel2(eng.fetch("grad_inner_rows"), sparse["inner_embeddings"][2].numpy())
    errs["bias_rows"] = rel2(eng.fetch("grad_bias_rows"), sparse["feature_bias"][2].numpy())
    print(precision, layer0, {k: "%.2e" % v for k, v in errs.items()})
    # top layer (one output position per sample): a pre-activation within the arithmetic's noise of zero lands on the
    # other side of the relu than in the oracle and moves a whole (sample, channel) term of these two sums -- with a few
    # dozen samples that is up to 1e-1 (bf16) / 1e-2 (bf16x3: measured 8.7e-3 and 1.1e-2 at F = 32, B = 20) relative
    top = {"bf16": 0.12, "bf16x3": 2e-2}.get(precision, gtol)
    bad = {k: v for k, v in errs.items() if v > (top if k.endswith("3") else gtol)}
    assert not bad, (precision, layer0, bad)
    assert int(eng.fetch("n_uniq")[0]) == len(np.unique(ids))
    eng.close()
