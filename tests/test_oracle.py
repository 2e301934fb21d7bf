"""Oracle self-tests (CPU).  The reference ships no tests or golden vectors for the model path
(parity unpinned, see oracle/cffm_ref.py); what pins the oracle is: the shape table of SURVEY §8.3,
closed-form identities, fp64 finite-difference gradient checks and the TF-1.14 Adagrad semantics."""
import math

import numpy as np
import pytest
import torch

from oracle.cffm_ref import CFFMRef, evaluate, eva_termination, get_ordered_block, get_random_block, pair_list


def _model(**kw):
    args = dict(features_M=60, num_field=4, inner_dims=8, outer_dims=8, activation="selu", dtype=torch.float64, seed=3)
    args.update(kw)
    m = CFFMRef(**args)
    g = torch.Generator().manual_seed(1)
    m.params["feature_bias"] = torch.randn(m.M, 1, generator=g, dtype=m.dtype) * 0.2
    if m.outer_conv:
        m.params["outer_embeddings"] = torch.randn(m.M, m.Ko, generator=g, dtype=m.dtype) * 0.3
    return m


def _batch(m, B=5, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, m.M, (B, m.F)), rng.choice([-1.0, 1.0], B)


def test_pair_order():
    assert pair_list(4) == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def test_shape_table_k32():
    """SURVEY §8.3 (K=32, all components on)."""
    m = CFFMRef(500, 10, 32, 32, activation="selu", dtype=torch.float32)
    ids, _ = _batch(m, 3)
    out, inter = m.forward(ids, return_intermediates=True)
    P = 45
    assert out.shape == (3, 1)
    assert inter["inner_max"].shape == (3, 32 * P)
    assert [tuple(inter["conv_%d" % l].shape) for l in range(5)] == [(3, 16, 16, P), (3, 8, 8, P), (3, 4, 4, P), (3, 2, 2, P), (3, 1, 1, P)]
    assert inter["t1"].shape == (3, 62)
    names = list(m.params)
    assert "dense/kernel" in names and m.params["dense/kernel"].shape == (1440, 1)
    assert m.params["dense_1/kernel"].shape == (62, 32) and m.params["dense_2/kernel"].shape == (32, 1)
    assert m.params["dense_3/kernel"].shape == (10, 1)
    assert m.params["outer_layer_conv_weight_4"].shape == (2, 2, P, P)


def test_dense_names_shift_when_components_are_disabled():
    m = CFFMRef(50, 4, 8, 8, inner_conv=0)
    assert m.dense_outer1 == "dense" and m.dense_outer2 == "dense_1" and m.dense_linear == "dense_2"
    m = CFFMRef(50, 4, 8, 8, outer_conv=0, linear_att=0)
    assert m.dense_inner == "dense" and "dense_1/kernel" not in m.params


def test_sum_pooling0_closed_form():
    """sum_pooling[0][b,h] = sum_p o_i[h] * sum_c o_j[c]."""
    m = _model()
    ids, _ = _batch(m)
    _, inter = m.forward(ids, return_intermediates=True)
    o = m.params["outer_embeddings"][torch.as_tensor(ids)]
    S = o.sum(-1)
    want = sum(o[:, i, :] * S[:, j:j + 1] for (i, j) in m.pairs)
    assert torch.allclose(inter["t1"][:, : m.Ko], want, atol=1e-12)


def test_double_activation_is_scaled_relu_for_selu():
    m = _model()
    ids, _ = _batch(m)
    _, inter = m.forward(ids, return_intermediates=True)
    assert (inter["conv_0"] >= 0).all()


@pytest.mark.parametrize("act", ["relu", "elu", "selu", "prelu", "gelu"])
@pytest.mark.parametrize("loss", ["square_loss", "mse", "log_loss"])
def test_finite_difference_gradients(act, loss):
    m = _model(activation=act, loss_type=loss, num_field=3, inner_dims=4, outer_dims=4, features_M=20)
    ids, y = _batch(m, 4, seed=2)
    if loss == "log_loss":
        y = (y > 0).astype(np.float64)
    l0, dense, sparse = m.gradients(ids, y)
    rng = np.random.default_rng(0)
    eps = 1e-6
    for name, g in dense.items():
        if g is None:
            assert name in m.dead_params()
            continue
        flat = m.params[name].reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            old = float(flat[idx])
            flat[idx] = old + eps
            lp = float(m.loss(ids, y))
            flat[idx] = old - eps
            lm = float(m.loss(ids, y))
            flat[idx] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - float(g.reshape(-1)[idx])) < 1e-5 * max(1.0, abs(fd)), (name, fd, float(g.reshape(-1)[idx]))
    for name, (rows, summed, _) in sparse.items():
        tab = m.params[name]
        for r_i in range(min(3, len(rows))):
            r, c = int(rows[r_i]), int(rng.integers(0, tab.shape[1]))
            old = float(tab[r, c])
            tab[r, c] = old + eps
            lp = float(m.loss(ids, y))
            tab[r, c] = old - eps
            lm = float(m.loss(ids, y))
            tab[r, c] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - float(summed[r_i, c])) < 1e-5 * max(1.0, abs(fd)), (name, fd)


def test_dead_parameters_are_never_updated():
    m = _model()
    before = {k: m.params[k].clone() for k in m.dead_params()}
    ids, y = _batch(m)
    m.train_step(ids, y)
    for k, v in before.items():
        assert torch.equal(v, m.params[k])
    assert set(m.dead_params()) == {"outer_W", "outer_b", "outer_layer_conv_weight_2", "outer_layer_conv_bias_2"}


def test_adagrad_tf_semantics():
    """acc0 = 1e-8, no epsilon: the first update of a touched weight is ~ lr*sign(g);
    duplicate ids are summed before the update; untouched rows and slots do not move."""
    m = _model()
    ids, y = _batch(m)
    ids[1] = ids[0]
    w0 = m.params["inner_embeddings"].clone()
    _, _, sparse = m.gradients(ids, y)
    rows, summed, vals = sparse["inner_embeddings"]
    assert len(rows) == len(np.unique(ids))
    m.train_step(ids, y)
    w1 = m.params["inner_embeddings"]
    touched = torch.zeros(m.M, dtype=torch.bool)
    touched[torch.as_tensor(rows)] = True
    assert torch.equal(w0[~touched], w1[~touched])
    acc = m.state["accumulator"]["inner_embeddings"]
    assert torch.all(acc[~touched] == 1e-8)
    g = summed
    big = g.abs() > 1e-2
    step = (w0[torch.as_tensor(rows)] - w1[torch.as_tensor(rows)])
    assert torch.allclose(step[big], 0.05 * torch.sign(g[big]), atol=1e-5)


def test_rmse_loss_definition():
    m = _model()
    ids, y = _batch(m)
    out = m.forward(ids).reshape(-1)
    want = math.sqrt(float(((torch.as_tensor(y) - out) ** 2).mean()) + 1e-10)
    assert abs(float(m.loss(ids, y)) - want) < 1e-12


def test_host_loop_pieces():
    data = {"X": [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]], "Y": [1.0, -1.0, 1.0, -1.0, 1.0]}
    assert get_ordered_block(data, 2, 2) == {"X": [[9, 10]], "Y": [1.0]}
    assert get_ordered_block(data, 2, 3) == {"X": [], "Y": []}
    blk = get_random_block(data, 2, 1)
    assert blk["X"] == [[3, 4], [5, 6]] and blk["Y"] == [[-1.0], [1.0]]
    assert eva_termination([5, 4, 1, 2, 3, 4, 5]) and not eva_termination([1, 2, 3, 4, 5])
    m = _model(num_field=2, inner_dims=4, outer_dims=4, features_M=11)
    rmse, r2 = evaluate(m, data, 2)
    assert rmse >= 0 and r2 <= 1
