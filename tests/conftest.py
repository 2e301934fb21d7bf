import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_model():
    return np.load(os.path.join(GOLDEN, "cffm_golden.npz"))


@pytest.fixture(scope="session")
def golden_libfm():
    return np.load(os.path.join(GOLDEN, "libfm_golden.npz"))


@pytest.fixture(scope="session")
def lib():
    from cffm_b200 import _lib
    return _lib.load()
