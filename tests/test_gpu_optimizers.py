"""the other optimizers of CFFM.py:517-529 and the lamda > 0 regulariser of
CFFM.py:489-491 (Q9), against the oracle's TF-1.14 restatement over three steps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(optimizer, lamda=0.0, act="selu", F=5, K=16, M=90, B=12, lamda_att=1.0, lr=0.05):
    from cffm_b200 import Engine
    from oracle.cffm_ref import CFFMRef
    eng = Engine(M, F, K, K, activation=act, optimizer=optimizer, lamda=lamda, lamda_att=lamda_att, lr=lr, max_batch=B, seed=4)
    rng = np.random.default_rng(2)
    eng.set_param("feature_bias", rng.normal(0, 0.2, (M, 1)).astype(np.float32))
    eng.set_param("outer_embeddings", rng.normal(0, 0.3, (M, K)).astype(np.float32))
    P = F * (F - 1) // 2
    for l in range(int(np.log2(K))):
        eng.set_param("outer_layer_conv_weight_%d" % l, rng.normal(0, 1 / np.sqrt(4 * P), (2, 2, P, P)).astype(np.float32))
    ref = CFFMRef(M, F, K, K, activation=act, optimizer=optimizer, lamda=lamda, lamda_att=lamda_att, lr=lr, dtype=torch.float64)
    for k, v in eng.get_weights().items():
        ref.params[k] = torch.from_numpy(v.astype(np.float64)).reshape(ref.params[k].shape)
    ids = rng.integers(0, M, (3, B, F)).astype(np.int32)
    ids[:, 3] = ids[:, 0]
    y = rng.choice([-1.0, 1.0], (3, B)).astype(np.float32)
    return eng, ref, ids, y


def _check(eng, ref, ids, y, wtol, ltol=1e-4):
    for s in range(3):
        got, want = eng.train_step(ids[s], y[s]), ref.train_step(ids[s], y[s])
        assert abs(got - want) < ltol * max(1.0, abs(want)), (s, got, want)
    for k, v in eng.get_weights().items():
        d = np.abs(v.astype(np.float64) - ref.params[k].numpy().reshape(v.shape))
        assert d.max() < wtol, (k, float(d.max()))


@pytest.mark.parametrize("optimizer", ["GradientDescentOptimizer", "MomentumOptimizer", "AdamOptimizer"])
def test_optimizer_matches_tf_semantics(optimizer):
    eng, ref, ids, y = _pair(optimizer)
    _check(eng, ref, ids, y, wtol=2e-4)
    touched = np.unique(ids)
    w = eng.get_param("inner_embeddings")
    if optimizer != "AdamOptimizer":  # sparse applies leave untouched rows alone; Adam's moves every row [TF-1.14]
        mask = np.ones(w.shape[0], dtype=bool); mask[touched] = False
        fresh, _, _, _ = _pair(optimizer)
        assert np.array_equal(w[mask], fresh.get_param("inner_embeddings")[mask])
        fresh.close()
    eng.close()


def test_adam_slots_are_exposed():
    eng, ref, ids, y = _pair("AdamOptimizer")
    eng.train_step(ids[0], y[0]); ref.train_step(ids[0], y[0])
    m = eng.get_param("bias_W", accum=True)
    shape, numel, _ = eng.param_infos()["bias_W"]
    import ctypes as C
    v = np.empty(numel, dtype=np.float32)
    eng._check(eng.lib.cffm_get_accum(eng.h, b"bias_W:2", v.ctypes.data_as(C.c_void_p), numel), "cffm_get_accum")
    assert np.allclose(m, ref.state["m"]["bias_W"].numpy(), atol=1e-6)
    assert np.allclose(v.reshape(shape), ref.state["v"]["bias_W"].numpy(), atol=1e-8)
    eng.close()


@pytest.mark.parametrize("optimizer", ["GradientDescentOptimizer", "MomentumOptimizer"])
def test_l2_regulariser_makes_table_gradients_dense(optimizer):
    """lamda > 0: loss = l2_loss + lamda/2 |inner|^2 + lamda_att/2 |outer|^2; every table row moves (Q9)."""
    # the lamda > 0 loss is a SUM over the batch: a small step keeps plain SGD / momentum from diverging
    eng, ref, ids, y = _pair(optimizer, lamda=0.01, lamda_att=0.5, lr=1e-4)
    w_before = eng.get_param("outer_embeddings")
    _check(eng, ref, ids, y, wtol=2e-4)
    w_after = eng.get_param("outer_embeddings")
    untouched = np.ones(w_after.shape[0], dtype=bool); untouched[np.unique(ids)] = False
    assert np.all(w_after[untouched] != w_before[untouched])  # the regulariser reaches rows outside the batch
    eng.close()


def test_l2_regulariser_with_adagrad_first_step():
    eng, ref, ids, y = _pair("AdagradOptimizer", lamda=0.01, lamda_att=0.5)
    got, want = eng.train_step(ids[0], y[0]), ref.train_step(ids[0], y[0])
    assert abs(got - want) < 1e-4 * max(1.0, abs(want))
    for k in ("inner_embeddings", "outer_embeddings"):
        a = eng.get_param(k, accum=True)
        assert np.allclose(a, ref.state["accumulator"][k].numpy(), rtol=2e-3, atol=1e-12), k
    eng.close()


def test_log_loss_with_lamda_is_rejected_like_the_reference():
    from cffm_b200 import Engine, CffmError
    with pytest.raises(CffmError):
        Engine(50, 3, 8, 8, loss_type="log_loss", lamda=0.1, max_batch=4)
