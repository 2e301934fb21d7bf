"""CPU oracle for the libfm loader -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ``LoadData`` (reference ``LoadData.py:25-112``) in plain Python.  Unlike the model
oracle this one IS pinned: ``tests/golden/make_golden.py`` imports the reference's own
``LoadData`` in the build container and commits its outputs on the fixture files under
``tests/golden/`` (``libfm_golden.npz``); ``tests/test_libfm.py`` checks this restatement and
the native parser against them bit for bit.
"""
from __future__ import annotations

import numpy as np


def map_features(files):
    """LoadData.py:33-55 -- ids are first-appearance order of the whole ``"idx:val"`` token,
    scanned over ``files`` in the order given (the reference passes train, test, validation)."""
    features = {}
    for path in files:
        with open(path) as f:
            for line in f:
                items = line.strip().split(" ")
                for item in items[1:]:
                    if item not in features:
                        features[item] = len(features)
    return features


def read_data(path, features):
    """LoadData.py:81-103."""
    X, Y, Ylog = [], [], []
    with open(path) as f:
        for line in f:
            items = line.strip().split(" ")
            Y.append(1.0 * float(items[0]))
            Ylog.append(1.0 if float(items[0]) > 0 else 0.0)
            X.append([features[item] for item in items[1:]])
    return X, Y, Ylog


def construct_dataset(X, Y):
    """LoadData.py:105-112 -- rows re-ordered by ``np.argsort(row_length)``.  The reference uses
    the default (unstable) quicksort; a stable sort is used here and in the native parser, which
    agrees with it whenever all rows have the same length (every shipped dataset)."""
    lens = [len(r) for r in X]
    order = np.argsort(lens, kind="stable")
    return {"Y": [Y[i] for i in order], "X": [X[i] for i in order]}


class LoadDataRef:
    """Same attributes as the reference's ``LoadData`` object (LoadData.py:25-31)."""

    def __init__(self, path, dataset, loss_type):
        self.path = path + dataset + "/"
        self.trainfile = self.path + dataset + ".train.libfm"
        self.testfile = self.path + dataset + ".test.libfm"
        self.validationfile = self.path + dataset + ".validation.libfm"
        self.features = map_features([self.trainfile, self.testfile, self.validationfile])
        self.features_M = len(self.features)
        out = []
        for fpath in (self.trainfile, self.validationfile, self.testfile):  # LoadData.py:57-79
            X, Y, Ylog = read_data(fpath, self.features)
            out.append(construct_dataset(X, Ylog if loss_type == "log_loss" else Y))
        self.Train_data, self.Validation_data, self.Test_data = out
