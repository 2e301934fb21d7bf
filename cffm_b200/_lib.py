"""ctypes binding of ``libcffm_b200.so`` (the C ABI declared in ``include/cffm.h``) and the nvcc
build recipe.  There is no CPU fallback: if the library is missing or no CUDA device is present,
every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libcffm_b200.so")
BUILD_DIR = os.path.join(ROOT, "build")

CU_SOURCES = ["params.cu", "forward.cu", "backward.cu", "update.cu", "api.cu", "comm.cu", "shard.cu", "gemm_tc.cu", "conv_tc.cu"]
CPP_SOURCES = ["libfm.cpp"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]

ABI_VERSION = 2
ACTIVATIONS = {"relu": 0, "elu": 1, "selu": 2, "prelu": 3, "gelu": 4}
LOSSES = {"square_loss": 0, "log_loss": 1, "mse": 2, "mae": 3, "hybrid": 4}
OPTIMIZERS = {"AdagradOptimizer": 0, "GradientDescentOptimizer": 1, "MomentumOptimizer": 2, "AdamOptimizer": 3}
PRECISIONS = {"fp32": 0, "bf16": 1, "bf16x3": 2}


class CffmError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("features_M", C.c_int32), ("num_field", C.c_int32),
        ("inner_dims", C.c_int32), ("outer_dims", C.c_int32), ("inner_conv", C.c_int32),
        ("outer_conv", C.c_int32), ("linear_att", C.c_int32), ("activation", C.c_int32),
        ("loss_type", C.c_int32), ("optimizer", C.c_int32), ("precision", C.c_int32),
        ("lr", C.c_float), ("lamda", C.c_float), ("lamda_att", C.c_float), ("beta_outer", C.c_float),
        ("max_batch", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64),
        ("shard_world", C.c_int32), ("shard_rank", C.c_int32),
    ]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _source_digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/cffm.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link ``libcffm_b200.so`` in-tree."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    stamp = os.path.join(BUILD_DIR, "cffm.digest")
    digest = _source_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        _record_build(False, digest, [])
        return LIB_PATH
    nvcc = _nvcc()
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets"]

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + common + (ARCH_FLAGS if src.endswith(".cu") else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise CffmError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, CU_SOURCES + CPP_SOURCES))
    tmp = LIB_PATH + ".tmp"
    r = subprocess.run([nvcc, "-shared", "-o", tmp] + objs + ["-ldl", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise CffmError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB_PATH)
    with open(stamp, "w") as fh:
        fh.write(digest)
    _record_build(True, digest, CU_SOURCES + CPP_SOURCES)
    return LIB_PATH


LAST_BUILD = None


def _record_build(nvcc_ran, digest, compiled):
    """Makes the rebuild observable: ``LAST_BUILD`` (and build/build_record.json) say whether this call ran nvcc
    or found the in-tree library up to date with the sources (same digest)."""
    global LAST_BUILD
    import json
    import time
    LAST_BUILD = {"nvcc_ran": bool(nvcc_ran), "compiled": list(compiled), "source_digest": digest, "library": LIB_PATH,
                  "arch": ARCH_FLAGS[1], "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    try:
        with open(os.path.join(BUILD_DIR, "build_record.json"), "w") as fh:
            json.dump(LAST_BUILD, fh, indent=1)
    except OSError:
        pass


_lib = None


def _declare(lib):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    P = C.POINTER
    sigs = {
        "cffm_create": (C.c_int, [P(Config), P(vp)]),
        "cffm_destroy": (C.c_int, [vp]),
        "cffm_last_error": (C.c_char_p, [vp]),
        "cffm_device_available": (C.c_int, []),
        "cffm_param_count": (C.c_int, [vp]),
        "cffm_param_info": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_int, P(i64), P(i32), P(i64), P(i32)]),
        "cffm_get_param": (C.c_int, [vp, C.c_char_p, vp, i64]),
        "cffm_set_param": (C.c_int, [vp, C.c_char_p, vp, i64]),
        "cffm_get_accum": (C.c_int, [vp, C.c_char_p, vp, i64]),
        "cffm_set_accum": (C.c_int, [vp, C.c_char_p, vp, i64]),
        "cffm_init_params": (C.c_int, [vp, C.c_uint64]),
        "cffm_get_opt_step": (C.c_int, [vp, P(i64)]),
        "cffm_set_opt_step": (C.c_int, [vp, i64]),
        "cffm_uses_graph": (C.c_int, [vp]),
        "cffm_debug_check_guards": (C.c_int, [C.c_char_p, i32]),
        "cffm_debug_guard_selftest": (C.c_int, []),
        "cffm_forward_dev": (C.c_int, [vp, vp, i64, vp, vp]),
        "cffm_forward_host": (C.c_int, [vp, vp, i64, vp]),
        "cffm_train_step_dev": (C.c_int, [vp, vp, vp, i64, vp, vp]),
        "cffm_train_step_host": (C.c_int, [vp, vp, vp, i64, P(f32)]),
        "cffm_train_submit_host": (C.c_int, [vp, vp, vp, i64, P(f32), P(i32)]),
        "cffm_train_flush": (C.c_int, [vp, P(f32), P(i32)]),
        "cffm_evaluate_host": (C.c_int, [vp, vp, vp, i64, i64, P(C.c_double), P(C.c_double)]),
        "cffm_dataset_upload": (C.c_int, [vp, vp, vp, i64]),
        "cffm_dataset_permute": (C.c_int, [vp, vp]),
        "cffm_train_block": (C.c_int, [vp, i64, i64]),
        "cffm_last_loss": (C.c_int, [vp, P(f32)]),
        "cffm_dataset_evaluate": (C.c_int, [vp, i64, P(C.c_double), P(C.c_double)]),
        "cffm_synchronize": (C.c_int, [vp]),
        "cffm_launch_count": (i64, [vp]),
        "cffm_profile_enable": (C.c_int, [vp, i32]),
        "cffm_profile_report": (i64, [vp, C.c_char_p, i64, i32]),
        "cffm_op_gather_dev": (C.c_int, [vp, vp, i64, i32, vp, vp]),
        "cffm_op_sparse_adagrad_dev": (C.c_int, [vp, vp, i32, i32, vp, vp, i64, f32, vp, vp, vp]),
        "cffm_op_gemm_bf16_dev": (C.c_int, [vp, vp, vp, i32, i32, i32, vp]),
        "cffm_op_gemm_bf16_tn_dev": (C.c_int, [vp, vp, vp, i32, i32, i32, vp]),
        "cffm_tc_last_error": (C.c_char_p, []),
        "cffm_debug_fetch": (C.c_int, [vp, C.c_char_p, vp, i64, P(i64)]),
        "cffm_debug_dense_grad": (C.c_int, [vp, C.c_char_p, vp, i64]),
        "cffm_comm_unique_id": (C.c_int, [C.c_char_p]),
        "cffm_comm_init": (C.c_int, [vp, C.c_char_p, i32, i32]),
        "cffm_libfm_load": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, P(vp)]),
        "cffm_libfm_features_M": (i64, [vp]),
        "cffm_libfm_split": (C.c_int, [vp, C.c_int, P(i64), P(vp), P(vp), P(vp), P(vp)]),
        "cffm_libfm_token": (C.c_int, [vp, i64, C.c_char_p, C.c_int]),
        "cffm_libfm_free": (C.c_int, [vp]),
        "cffm_libfm_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return sigs


EXPORTED_SYMBOLS = None


def load(build_if_missing=True):
    """Load the shared library (building it first if it is absent and nvcc is available)."""
    global _lib, EXPORTED_SYMBOLS
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise CffmError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        build()
    lib = C.CDLL(LIB_PATH)
    EXPORTED_SYMBOLS = sorted(_declare(lib))
    _lib = lib
    return lib


def check(rc, handle=None, what=""):
    if rc == 0:
        return
    lib = load()
    msg = lib.cffm_last_error(handle)
    msg = msg.decode() if msg else ""
    raise CffmError("%s failed (%d): %s" % (what or "cffm call", rc, msg))


def device_available():
    return load().cffm_device_available() == 0
