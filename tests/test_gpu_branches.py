"""The parallel branches of a step (side streams forked from / joined to the step's stream) change when
a kernel runs, not what it computes: a run with CFFM_SIDE_STREAM=0 (everything on one stream) gives bit-identical
losses, logits and weights -- in eager steps and in graph replays, on the small-step path (weight gradients beside the
data gradients) and on the big-step path (inner path beside the outer one only)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(monkeypatch, side, precision, F, B, steps, pipelined):
    from cffm_b200 import Engine
    monkeypatch.setenv("CFFM_SIDE_STREAM", side)
    M = 500
    rng = np.random.default_rng(5)
    eng = Engine(M, F, 32, 32, activation="relu", max_batch=B, precision=precision, seed=7)
    P = F * (F - 1) // 2
    for l in range(5):
        eng.set_param("outer_layer_conv_weight_%d" % l, rng.normal(0, 1 / np.sqrt(4 * P), (2, 2, P, P)).astype(np.float32))
    losses = []
    for s in range(steps):
        ids = rng.integers(0, M, (B, F)).astype(np.int32)
        y = rng.choice([-1.0, 1.0], B).astype(np.float32)
        if pipelined:
            eng.train_submit(ids, y)
        else:
            losses.append(eng.train_step(ids, y))
    if pipelined:
        losses = [eng.train_flush()]
    ids = rng.integers(0, M, (B, F)).astype(np.int32)
    out = eng.forward(ids)
    w = eng.get_weights()
    eng.close()
    return np.asarray(losses, np.float64), out, w


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16x3"])
@pytest.mark.parametrize("F,B,pipelined", [(10, 64, False), (10, 64, True), (18, 600, True)])
def test_side_streams_do_not_change_results(precision, F, B, pipelined, monkeypatch):
    if F == 18:   # big-step path: layer 0's gradient tensor beyond 48 MB, factorised layer-0 kernels
        monkeypatch.setenv("CFFM_FACT_MIN_BATCH", "1"); monkeypatch.setenv("CFFM_FACT_MIN_FIELDS", "1")
    a = _run(monkeypatch, "1", precision, F, B, 4, pipelined)
    b = _run(monkeypatch, "0", precision, F, B, 4, pipelined)
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1], b[1])
    assert set(a[2]) == set(b[2])
    for k in a[2]:
        assert np.array_equal(a[2][k], b[2][k]), k
