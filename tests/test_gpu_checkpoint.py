"""Checkpoint / resume (CFFM.py:159, :226-228, :241-252): a run that is saved,
loaded into a NEW handle and continued must equal the uninterrupted run bit for bit -- weights, every
optimizer slot (Adam: m AND v) and the optimizer step counter that drives Adam's bias correction."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OPTS = ["AdagradOptimizer", "GradientDescentOptimizer", "MomentumOptimizer", "AdamOptimizer"]


def _data(M, F, n, seed=0):
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, M, (n, F)).astype(np.int32)
    y = rng.choice([-1.0, 1.0], n).astype(np.float32)
    return ids, y


def _engine(opt, seed, precision="fp32", K=16, F=5, M=300, B=32):
    from cffm_b200 import Engine
    return Engine(M, F, K, K, activation="selu", optimizer=opt, lr=0.01 if opt != "AdagradOptimizer" else 0.05,
                  max_batch=B, seed=seed, precision=precision)


@pytest.mark.parametrize("opt", OPTS)
def test_resume_equals_uninterrupted(opt):
    M, F, B = 300, 5, 32
    ids, y = _data(M, F, 5 * B)
    a = _engine(opt, seed=1)
    for s in range(3):
        a.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B])
    state = a.state_dict()
    assert "opt_step" in state
    if opt == "AdamOptimizer":
        assert int(state["opt_step"]) == 3
        assert any(k.startswith("a2:") for k in state)          # second moment is part of the checkpoint
        assert np.any(state["a2:inner_embeddings"] != 0)
    if opt == "GradientDescentOptimizer":
        assert not any(k.startswith("a:") or k.startswith("a2:") for k in state)   # plain SGD keeps no slots
    b = _engine(opt, seed=99)                                   # different initial weights: everything must come from the state
    b.load_state_dict(state)
    la, lb = [], []
    for s in range(3, 5):
        la.append(a.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B]))
        lb.append(b.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B]))
    assert la == lb, (opt, la, lb)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert np.array_equal(np.asarray(sa[k]), np.asarray(sb[k])), (opt, k)
    a.close(); b.close()


def test_adam_without_second_slot_or_step_diverges():
    """The round-1 checkpoint (slot 1 only, no step counter) is NOT enough for Adam: guards the fix."""
    M, F, B = 300, 5, 32
    ids, y = _data(M, F, 4 * B)
    a = _engine("AdamOptimizer", seed=1)
    for s in range(3):
        a.train_step(ids[s * B:(s + 1) * B], y[s * B:(s + 1) * B])
    state = a.state_dict()
    partial = {k: v for k, v in state.items() if not k.startswith("a2:") and k != "opt_step"}
    b = _engine("AdamOptimizer", seed=99)
    b.load_state_dict(partial)
    a.train_step(ids[3 * B:], y[3 * B:]); b.train_step(ids[3 * B:], y[3 * B:])
    assert not np.array_equal(a.get_param("inner_embeddings"), b.get_param("inner_embeddings"))
    a.close(); b.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_model_save_load_npz_round_trip(tmp_path, precision):
    """class CFFM: --pretrain -1 saves after every epoch, --pretrain 1 restores in build_graph (CFFM.py:226-228, :241-252)."""
    from cffm_b200.model import CFFM
    M, F, K, B = 400, 10, 32, 64
    ids, y = _data(M, F, 3 * B, seed=3)
    save = os.path.join(str(tmp_path), "ckpt", "cffm")

    def model(pretrain, seed):
        return CFFM(M, pretrain, save, K, K, "square_loss", 1, B, 0.05, 0, [1.0, 1.0], "AdamOptimizer", 0, 0, 0, F, 1, 0, 1.0,
                    1, 1.0, 1, 1.0, "selu", random_seed=seed, precision=precision)

    a = model(0, 5)
    a.build_graph()
    a.engine.train_step(ids[:B], y[:B]); a.engine.train_step(ids[B:2 * B], y[B:2 * B])
    a.save_state(save + ".npz")
    b = model(1, 77)
    b.build_graph()                      # pretrain_flag > 0: load_state inside build_graph
    assert b.engine.get_opt_step() == 2
    la = a.engine.train_step(ids[2 * B:], y[2 * B:])
    lb = b.engine.train_step(ids[2 * B:], y[2 * B:])
    assert la == lb
    wa, wb = a.engine.get_weights(), b.engine.get_weights()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
    a.engine.close(); b.engine.close()
