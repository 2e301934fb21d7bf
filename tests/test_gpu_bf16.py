"""Parity of the tensor-core (bf16 operands, fp32 accumulation) conv stack against the fp64 oracle.
Tolerance from BASELINE.json's north star: forward logits within 1e-2 relative in bf16 mode;
gradients are compared at 4e-2 relative to each tensor's scale (three bf16 roundings per product)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# name: F, M, B, activation
SHAPES = {
    "frappe_like": (10, 400, 24, "selu"),      # P=45 -> 64 channels, one N tile of 64
    "bx_like": (6, 300, 40, "relu"),           # P=15 -> 64
    "mltag_like": (3, 200, 64, "elu"),         # P=3  -> 64 (mostly padding)
    "criteo_like": (39, 2000, 6, "relu"),      # P=741 -> 768 channels, three N tiles of 256
    "odd_batch": (12, 500, 7, "prelu"),        # P=66 -> 128, BN=128; partial M tiles everywhere
    "overhang": (30, 900, 5, "relu"),          # P=435 -> 448: layer-0 forward covers it with 4 x 128 (64 columns of zero fill)
    "three_by_64": (20, 700, 9, "elu"),        # P=190 -> 192: layer-0 forward pairs two N tiles of 96
}


def _rel(a, b):
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def _rel2(a, b):
    """relative L2 error: a relu mask that flips between bf16 and fp64 moves one term of a short,
    cancellation-heavy batch sum by 100 %, which the max norm reports as a ~1/sqrt(n) error."""
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b)))


def _pair(name, seed=3):
    from cffm_b200 import Engine
    from oracle.cffm_ref import CFFMRef
    F, M, B, act = SHAPES[name]
    eng = Engine(M, F, 32, 32, activation=act, max_batch=B, precision="bf16", seed=seed)
    rng = np.random.default_rng(seed)
    eng.set_param("feature_bias", rng.normal(0, 0.05, (M, 1)).astype(np.float32))
    eng.set_param("outer_embeddings", rng.normal(0, 0.3, (M, 32)).astype(np.float32))
    P = F * (F - 1) // 2
    for l in range(5):  # keep the activations O(1) through four layers: filters ~ N(0, 1/(4P))
        eng.set_param("outer_layer_conv_weight_%d" % l, rng.normal(0, 1.0 / np.sqrt(4 * P), (2, 2, P, P)).astype(np.float32))
    ref = CFFMRef(M, F, 32, 32, activation=act, dtype=torch.float64)
    for k, v in eng.get_weights().items():
        ref.params[k] = torch.from_numpy(v.astype(np.float64)).reshape(ref.params[k].shape)
    ids = rng.integers(0, M, (B, F)).astype(np.int32)
    ids[-1] = ids[0]
    y = rng.choice([-1.0, 1.0], B).astype(np.float32)
    return eng, ref, ids, y


MODES = ["direct", "factorised"]


@pytest.fixture(params=MODES)
def layer0_mode(request, monkeypatch):
    """Layer 0 has two implementations: the direct implicit GEMM over the synthesised cube
    and the factorised form, which the library only picks from num_field >= 16 and batch >= 512 on.  The parity
    cases are small, so the choice is forced here and both are held to the same tolerances."""
    if request.param == "factorised":
        monkeypatch.setenv("CFFM_FACT_MIN_BATCH", "1")
        monkeypatch.setenv("CFFM_FACT_MIN_FIELDS", "1")
    else:
        monkeypatch.setenv("CFFM_FACT_MIN_BATCH", "1000000000")
    return request.param


@pytest.mark.parametrize("name", list(SHAPES))
def test_bf16_forward(name, layer0_mode):
    eng, ref, ids, y = _pair(name)
    out = eng.forward(ids)
    want, inter = ref.forward(ids, return_intermediates=True)
    B = ids.shape[0]
    for l in range(4):  # stored activations X_{l+1} = phi(Y_l)
        got = eng.fetch("conv_%d" % l)
        assert _rel(got, inter["conv_%d" % l].numpy()) < 2e-2, (name, l, _rel(got, inter["conv_%d" % l].numpy()))
    assert _rel(eng.fetch("t1").reshape(B, -1), inter["t1"].numpy()) < 1e-2
    assert _rel(eng.fetch("final"), inter["final"].numpy()) < 1e-2
    assert _rel(out, want.numpy()) < 1e-2, (name, _rel(out, want.numpy()))
    eng.close()


@pytest.mark.parametrize("name", list(SHAPES))
def test_bf16_gradients(name, layer0_mode):
    eng, ref, ids, y = _pair(name)
    l_ref, dense, sparse = ref.gradients(ids, y)
    loss = eng.train_step(ids, y)
    assert abs(loss - float(l_ref)) < 1e-2 * max(1.0, float(l_ref))
    errs = {}
    for l in range(4):
        errs["wgrad%d" % l] = _rel2(eng.dense_grad("outer_layer_conv_weight_%d" % l), dense["outer_layer_conv_weight_%d" % l].numpy())
        errs["bgrad%d" % l] = _rel2(eng.dense_grad("outer_layer_conv_bias_%d" % l), dense["outer_layer_conv_bias_%d" % l].numpy())
    errs["outer_rows"] = _rel2(eng.fetch("grad_outer_rows"), sparse["outer_embeddings"][2].numpy())
    # the parts that stay fp32 only see the bf16 logits through the loss
    errs["inner_rows"] = _rel(eng.fetch("grad_inner_rows"), sparse["inner_embeddings"][2].numpy())
    errs["dense_1"] = _rel(eng.dense_grad("dense_1/kernel"), dense["dense_1/kernel"].numpy())
    print(name, {k: round(v, 4) for k, v in errs.items()})
    # the top layer's sums are the shortest (B*4 terms per channel): one relu mask that flips between
    # the bf16 and the fp64 forward pass moves a whole term, so it gets the widest band
    tol = {"wgrad3": 0.12, "bgrad3": 0.12}
    bad = {k: v for k, v in errs.items() if v > tol.get(k, 5e-2)}
    assert not bad, (name, errs)
    eng.close()


def test_bf16_is_deterministic():
    a, _, ids, y = _pair("frappe_like")
    b, _, _, _ = _pair("frappe_like")
    la = [a.train_step(ids, y) for _ in range(3)]
    lb = [b.train_step(ids, y) for _ in range(3)]
    assert la == lb
    wa, wb = a.get_weights(), b.get_weights()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
    a.close(); b.close()


def test_bf16_training_tracks_fp32():
    """40 epochs of random-block training on the Frappe fixture (CFFM.py:181-200 loop): bf16 and fp32
    end at the same validation RMSE up to the run-to-run spread measured at this size (0.02-0.05 between
    seeds or summation orders: Adagrad with acc0 = 1e-8 makes early steps +-lr*sign(g));
    the 0.002 band of the north star is for the full datasets and epoch counts."""
    import os
    from cffm_b200 import Engine, LoadData
    from conftest import GOLDEN
    d = LoadData(os.path.join(GOLDEN, "frappe_mini") + "/", "frappe", "square_loss")
    X, Y = np.array(d.Train_data["X"]), np.array(d.Train_data["Y"])
    Xv, Yv = np.array(d.Validation_data["X"]), np.array(d.Validation_data["Y"])
    res = {}
    for prec in ("fp32", "bf16"):
        eng = Engine(d.features_M, 10, 32, 32, activation="selu", max_batch=256, precision=prec, seed=11)
        rng = np.random.RandomState(0)
        init = eng.evaluate(Xv, Yv, 256)[0]
        for step in range(40 * 11):
            st = rng.randint(0, 3000 - 256)
            eng.train_step(X[st:st + 256], Y[st:st + 256])
        res[prec] = eng.evaluate(Xv, Yv, 256)[0]
        assert res[prec] < 0.85 < init, (prec, init, res[prec])
        eng.close()
    assert abs(res["fp32"] - res["bf16"]) < 0.07, res


@pytest.mark.parametrize("F,M,B", [(10, 400, 24), (20, 700, 9)])
def test_bf16_gelu_trains(F, M, B):
    """gelu (CFFM.py:149-151): phi'(y) = Phi(r) + r N(r) does not follow from the sign of the stored activation, so the
    forward epilogues store it (bf16) next to X and the data gradient multiplies by it instead of masking."""
    from cffm_b200 import Engine
    from oracle.cffm_ref import CFFMRef
    rng = np.random.default_rng(5)
    eng = Engine(M, F, 32, 32, activation="gelu", max_batch=B, precision="bf16", seed=3)
    eng.set_param("feature_bias", rng.normal(0, 0.05, (M, 1)).astype(np.float32))
    eng.set_param("outer_embeddings", rng.normal(0, 0.3, (M, 32)).astype(np.float32))
    P = F * (F - 1) // 2
    for l in range(5):
        eng.set_param("outer_layer_conv_weight_%d" % l, rng.normal(0, 1.0 / np.sqrt(4 * P), (2, 2, P, P)).astype(np.float32))
    ref = CFFMRef(M, F, 32, 32, activation="gelu", dtype=torch.float64)
    for k, v in eng.get_weights().items():
        ref.params[k] = torch.from_numpy(v.astype(np.float64)).reshape(ref.params[k].shape)
    ids = rng.integers(0, M, (B, F)).astype(np.int32)
    y = rng.choice([-1.0, 1.0], B).astype(np.float32)
    assert _rel(eng.forward(ids), ref.predict(ids).numpy()) < 1e-2
    l_ref, dense, sparse = ref.gradients(ids, y)
    loss = eng.train_step(ids, y)
    assert abs(loss - float(l_ref)) < 1e-2 * max(1.0, float(l_ref))
    errs = {}
    for l in range(4):
        errs["wgrad%d" % l] = _rel2(eng.dense_grad("outer_layer_conv_weight_%d" % l), dense["outer_layer_conv_weight_%d" % l].numpy())
    errs["outer_rows"] = _rel2(eng.fetch("grad_outer_rows"), sparse["outer_embeddings"][2].numpy())
    print("gelu", F, {k: round(v, 4) for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if v > (0.12 if k == "wgrad3" else 5e-2)}
    assert not bad, errs
    eng.close()


def test_bf16_rejects_what_it_cannot_do():
    """Scoring runs on the tensor cores for outer_dims 16 / 32 / 64 and every activation; TRAINING there needs
    outer_dims == 32 and an activation whose derivative follows from the stored post-activation."""
    from cffm_b200 import Engine, CffmError
    ids = np.zeros((4, 4), dtype=np.int32)
    y = np.ones(4, dtype=np.float32)
    for prec, kw in (("bf16", dict(inner_dims=16, outer_dims=16)), ("bf16x3", dict(inner_dims=32, outer_dims=32, activation="gelu"))):
        eng = Engine(100, 4, max_batch=4, precision=prec, **kw)
        assert eng.forward(ids).shape == (4,)
        with pytest.raises(CffmError):
            eng.train_step(ids, y)
        eng.close()
    with pytest.raises(CffmError):
        Engine(100, 4, 8, 8, max_batch=4, precision="bf16")
