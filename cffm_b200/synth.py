"""Synthetic workloads with the shapes of the reference's datasets.

ids ``int32 [N, F]``: field f draws from its own contiguous id range with a Zipf-like skew, so
row duplication inside a batch resembles the real files; labels are +/-1 with P(+1) = 1/3
(0/1 for log_loss).  Seeded with ``numpy.random.default_rng(2021)``.
"""
from __future__ import annotations

import numpy as np

WORKLOADS = {
    # name: field cardinalities, batch, activation, K
    "frappe": dict(cards=[957, 4082, 7, 7, 2, 3, 2, 9, 80, 233], batch=256, activation="selu", K=32,
                   n_train=202024),
    "ml-tag": dict(cards=[17045, 23743, 49657], batch=1024, activation="elu", K=32, n_train=1404799),
    "book-crossing": dict(cards=[26847, 134039, 144, 124, 56089, 9093], batch=512, activation="relu", K=32,
                          n_train=849356),
    "criteo": dict(cards=None, batch=8192, activation="relu", K=32, n_train=None),
}


def criteo_cards(total=10_000_000, n_numeric=13, n_cat=26):
    """13 'numeric-bucket' fields of <=100 ids + 26 categorical fields sharing the rest."""
    num = [100] * n_numeric
    rest = total - sum(num)
    w = np.linspace(2.0, 0.2, n_cat)
    w = w / w.sum()
    cat = np.maximum(1000, np.floor(w * rest)).astype(np.int64)
    cat[0] += rest - int(cat.sum())
    return num + [int(c) for c in cat]


def field_cards(name):
    if name == "criteo":
        return criteo_cards()
    return list(WORKLOADS[name]["cards"])


def _zipf_draw(rng, card, n, a=1.05):
    """Bounded Zipf(a) over [0, card) by inverse-CDF on a truncated power law (vectorised)."""
    if card <= 1:
        return np.zeros(n, dtype=np.int64)
    u = rng.random(n)
    if abs(a - 1.0) < 1e-9:
        r = np.exp(u * np.log(card + 1.0)) - 1.0
    else:
        # continuous power law p(x) ~ (x+1)^-a on [0, card)
        lo, hi = 1.0, float(card + 1)
        e = 1.0 - a
        r = (u * (hi ** e - lo ** e) + lo ** e) ** (1.0 / e) - 1.0
    return np.minimum(r.astype(np.int64), card - 1)


def make_ids(name, n, seed=2021, cards=None):
    rng = np.random.default_rng(seed)
    cards = field_cards(name) if cards is None else list(cards)
    offs = np.concatenate([[0], np.cumsum(cards)[:-1]]).astype(np.int64)
    ids = np.empty((n, len(cards)), dtype=np.int32)
    for f, (c, o) in enumerate(zip(cards, offs)):
        perm_key = rng.integers(1, 1 << 30)
        r = _zipf_draw(rng, c, n)
        # spread the popular ranks over the field's range instead of always its first ids
        r = (r * 2654435761 + perm_key) % c if c > 2 else r
        ids[:, f] = (o + r).astype(np.int32)
    return ids, int(sum(cards))


def make_labels(n, seed=2021, loss_type="square_loss"):
    rng = np.random.default_rng(seed + 1)
    pos = rng.random(n) < (1.0 / 3.0)
    if loss_type == "log_loss":
        return pos.astype(np.float32)
    return np.where(pos, 1.0, -1.0).astype(np.float32)


def make_workload(name, n=None, seed=2021, loss_type="square_loss"):
    """Returns dict(ids, labels, features_M, num_field, batch, activation, K)."""
    w = WORKLOADS[name]
    if n is None:
        n = w["batch"] * 8
    ids, M = make_ids(name, n, seed)
    return dict(ids=ids, labels=make_labels(n, seed, loss_type), features_M=M, num_field=ids.shape[1],
                batch=w["batch"], activation=w["activation"], K=w["K"], name=name)
