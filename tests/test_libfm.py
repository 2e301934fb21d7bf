"""libfm loader: the native parser (C ABI) and the Python restatement against outputs of the
REFERENCE's own LoadData, committed in tests/golden/libfm_golden.npz (bit-exact)."""
import os

import numpy as np
import pytest

from cffm_b200 import LoadData
from cffm_b200._lib import CffmError
from oracle.libfm_ref import LoadDataRef

from conftest import GOLDEN

CASES = [("frappe", os.path.join(GOLDEN, "frappe_mini") + "/", "frappe"), ("rag", os.path.join(GOLDEN, "ragged") + "/", "rag")]


def _flatten(split):
    X = split["X"]
    lens = np.array([len(r) for r in X], dtype=np.int64)
    ids = np.array([i for r in X for i in r], dtype=np.int32)
    return lens, ids, np.asarray(split["Y"], dtype=np.float64)


def _groups(lens, ids, y):
    out, o = {}, 0
    for n, yy in zip(lens, y):
        out.setdefault(int(n), []).append((tuple(int(i) for i in ids[o:o + n]), float(yy)))
        o += n
    return {k: sorted(v) for k, v in out.items()}


@pytest.mark.parametrize("tag,path,ds", CASES)
@pytest.mark.parametrize("loss", ["square_loss", "log_loss"])
@pytest.mark.parametrize("impl", ["native", "oracle"])
def test_loader_matches_reference(golden_libfm, tag, path, ds, loss, impl):
    d = LoadData(path, ds, loss) if impl == "native" else LoadDataRef(path, ds, loss)
    g = golden_libfm
    assert d.features_M == int(g["%s.%s.features_M" % (tag, loss)])
    for name, split in (("train", d.Train_data), ("validation", d.Validation_data), ("test", d.Test_data)):
        lens, ids, y = _flatten(split)
        glens, gids, gy = (g["%s.%s.%s.%s" % (tag, loss, name, k)] for k in ("lens", "ids", "y"))
        assert np.array_equal(lens, glens)
        if tag == "rag":
            # LoadData.py:109 uses numpy's default (unstable) argsort: rows of EQUAL length may come out
            # in any order; ours is stable.  Compare the groups of equal length as multisets.
            assert _groups(lens, ids, y) == _groups(glens, gids, gy)
        else:
            assert np.array_equal(ids, gids)
            assert np.array_equal(y, gy)
    if impl == "native":
        toks = g["%s.%s.tokens" % (tag, loss)]
        for fid in (0, len(toks) // 2, len(toks) - 1):
            assert d.token(fid) == str(toks[fid])


def test_frappe_fixture_shape():
    d = LoadData(os.path.join(GOLDEN, "frappe_mini") + "/", "frappe", "square_loss")
    assert isinstance(d.Train_data["X"], np.ndarray) and d.Train_data["X"].shape == (3000, 10)
    assert d.Train_data["X"].dtype == np.int32 and d.Train_data["Y"].dtype == np.float32
    assert d.Validation_data["X"].shape == (800, 10) and d.Test_data["X"].shape == (800, 10)
    assert set(np.unique(d.Train_data["Y"])) == {-1.0, 1.0}
    assert d.Train_data["X"].max() < d.features_M


def test_empty_token_and_whitespace(tmp_path):
    """Python's line.strip().split(' ') yields '' for doubled spaces (a vocabulary entry of its own),
    strips \r\n and leading / trailing blanks, and does not need a final newline."""
    p = tmp_path / "w"
    p.mkdir()
    (p / "w.train.libfm").write_text("1 a:1  b:1\r\n  -1 x:1 a:1 \n")
    (p / "w.test.libfm").write_text("1 b:1 y:1")
    (p / "w.validation.libfm").write_text("")
    nat = LoadData(str(tmp_path) + "/", "w", "square_loss")
    ref = LoadDataRef(str(tmp_path) + "/", "w", "square_loss")
    assert nat.features_M == ref.features_M == 5
    assert [nat.token(i) for i in range(5)] == ["a:1", "", "b:1", "x:1", "y:1"]
    for a, b in ((nat.Train_data, ref.Train_data), (nat.Test_data, ref.Test_data)):
        assert [list(r) for r in a["X"]] == b["X"]
        assert list(a["Y"]) == b["Y"]
    assert len(nat.Validation_data["Y"]) == 0


def test_bad_label_is_an_error(tmp_path):
    p = tmp_path / "e"
    p.mkdir()
    (p / "e.train.libfm").write_text("1 a:1\n\n1 b:1\n")  # blank line: float('') raises in the reference
    (p / "e.test.libfm").write_text("1 a:1\n")
    (p / "e.validation.libfm").write_text("1 a:1\n")
    with pytest.raises(CffmError):
        LoadData(str(tmp_path) + "/", "e", "square_loss")
    with pytest.raises(ValueError):
        LoadDataRef(str(tmp_path) + "/", "e", "square_loss")


def test_missing_file_is_an_error(tmp_path):
    with pytest.raises(CffmError):
        LoadData(str(tmp_path) + "/", "nope", "square_loss")
