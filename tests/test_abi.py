"""The C-ABI library loads and exports every symbol include/cffm.h declares (no compute here)."""
import ctypes as C
import os
import re

import pytest

from cffm_b200 import _lib
from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "cffm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cffm_[A-Za-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(_lib.EXPORTED_SYMBOLS) == set(names), set(_lib.EXPORTED_SYMBOLS) ^ set(names)


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "cffm.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "#include <stdint.h>" in code


def test_create_fails_loudly_without_a_device(lib):
    """No CPU fallback: without CUDA the product path must refuse to run."""
    if lib.cffm_device_available() == 0:
        pytest.skip("a CUDA device is present")
    cfg = _lib.Config(abi_version=_lib.ABI_VERSION, features_M=10, num_field=3, inner_dims=8, outer_dims=8,
                      inner_conv=1, outer_conv=1, linear_att=1, activation=0, loss_type=0, optimizer=0, precision=0,
                      lr=0.05, lamda=0.0, lamda_att=1.0, beta_outer=1.0, max_batch=4, device=0, seed=1)
    h = C.c_void_p()
    rc = lib.cffm_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value
    assert b"CUDA" in lib.cffm_last_error(None)
    from cffm_b200 import Engine, CffmError
    with pytest.raises(CffmError):
        Engine(10, 3, 8, 8, max_batch=4)


def test_config_struct_layout():
    assert C.sizeof(_lib.Config) == 12 * 4 + 4 * 4 + 2 * 4 + 8 + 2 * 4   # ... + shard_world, shard_rank (ABI version 2)
    assert _lib.ABI_VERSION == 2
