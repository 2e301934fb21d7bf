"""Host-side logic that mirrors the reference (CLI surface, batching, shuffling, synthetic shapes)."""
import numpy as np
import pytest

from cffm_b200 import cli, synth
from cffm_b200.model import CFFM, _shuffle_in_unison

REFERENCE_DEFAULTS = dict(  # CFFM.py:24-78
    path='data/', dataset='frappe', epoch=50, pretrain=0, batch_size=1024, inner_dims=32, outer_dims=32, lamda=0,
    keep='[1.0,1.0]', lr=0.05, loss_type='square_loss', optimizer='AdagradOptimizer', verbose=1, batch_norm=0,
    tensorboard=0, num_field=3, linear_att=1, att_dim=0, lamda_att=1.0, inner_conv=1, gamma_inner=1.0, outer_conv=1,
    beta_outer=1.0, activation='relu')


def test_cli_defaults_match_reference():
    a = vars(cli.parse_args([]))
    for k, v in REFERENCE_DEFAULTS.items():
        assert a[k] == v, k


def test_cli_readme_command_lines():
    a = cli.parse_args("--dataset frappe --epoch 50 --batch_size 256 --inner_dims 32 --outer_dims 32 --lamda 0 --lr 0.05 "
                       "--loss_type square_loss --num_field 10 --linear_att 1 --inner_conv 1 --outer_conv 1 "
                       "--activation selu".split())
    assert (a.dataset, a.batch_size, a.num_field, a.activation) == ("frappe", 256, 10, "selu")


def test_shuffle_matches_sklearn():
    from sklearn.utils import shuffle
    x = np.arange(40).reshape(20, 2)
    y = np.arange(20, dtype=np.float32)
    xs, ys = shuffle(x, y, random_state=2021)
    xo, yo = _shuffle_in_unison(x, y, 2021)
    assert np.array_equal(xs, xo) and np.array_equal(ys, yo)
    xl, yl = _shuffle_in_unison([list(r) for r in x], list(y), 2021)
    assert np.array_equal(np.asarray(xl), xs)


def _model(**kw):
    args = [100, 0, "", 8, 8, "square_loss", 1, 4, 0.05, 0, [1.0, 1.0], "AdagradOptimizer", 0, 0, 0, 3, 1, 0, 1.0, 1, 1.0,
            1, 1.0, "relu"]
    return CFFM(*args, **kw)


def test_random_block_is_contiguous_and_seedable():
    m = _model(batch_seed=3)
    data = {"X": np.arange(60, dtype=np.int32).reshape(20, 3), "Y": np.arange(20, dtype=np.float32)}
    blk = m.get_random_block_from_data(data, 4)
    assert blk["X"].shape == (4, 3) and np.all(np.diff(blk["Y"]) == 1)
    start = np.random.RandomState(3).randint(0, 16)
    assert blk["Y"][0] == start


def test_ordered_blocks_cover_the_set_with_a_partial_tail():
    m = _model()
    data = {"X": np.arange(30, dtype=np.int32).reshape(10, 3), "Y": np.arange(10, dtype=np.float32)}
    sizes = []
    i = 0
    while True:
        b = m.get_ordered_block_from_data(data, 4, i)
        if len(b["X"]) == 0:
            break
        sizes.append(len(b["X"]))
        i += 1
    assert sizes == [4, 4, 2]


def test_ragged_block_rules():
    m = _model(batch_seed=0)
    X = [np.array([1, 2, 3]), np.array([4, 5, 6]), np.array([7, 8]), np.array([9, 10])]
    data = {"X": X, "Y": np.array([1, -1, 1, -1], dtype=np.float32)}
    b = m.get_ordered_block_from_data(data, 4, 0)
    assert b["X"].shape == (2, 3)  # stops at the first length change


def test_early_stop_rule():
    m = _model()
    assert m.eva_termination([9, 1, 2, 3, 4, 5]) and not m.eva_termination([1, 2, 3])


@pytest.mark.parametrize("name,F,M", [("frappe", 10, 5382), ("ml-tag", 3, 90445), ("book-crossing", 6, 226336)])
def test_synthetic_shapes(name, F, M):
    w = synth.make_workload(name, n=2048)
    assert w["ids"].shape == (2048, F) and w["ids"].dtype == np.int32 and w["features_M"] == M
    assert w["ids"].min() >= 0 and w["ids"].max() < M
    offs = np.concatenate([[0], np.cumsum(synth.field_cards(name))])
    for f in range(F):  # every field draws from its own id range
        assert offs[f] <= w["ids"][:, f].min() and w["ids"][:, f].max() < offs[f + 1]
    assert set(np.unique(w["labels"])) == {-1.0, 1.0}
    assert 0.25 < (w["labels"] > 0).mean() < 0.42


def test_criteo_cards():
    c = synth.criteo_cards()
    assert len(c) == 39 and sum(c) == 10_000_000 and c[:13] == [100] * 13
    ids, M = synth.make_ids("criteo", 4096)
    assert ids.shape == (4096, 39) and M == 10_000_000
    assert len(np.unique(ids)) < ids.size  # skewed: duplicates inside a batch
