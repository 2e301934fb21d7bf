"""``LoadData``-compatible facade over the native libfm parser.

Same constructor and attributes as the reference's ``LoadData`` (LoadData.py:25-31):
``features_M``, ``Train_data`` / ``Validation_data`` / ``Test_data`` = ``{'X': ..., 'Y': ...}``.
'X' is an ``int32 [N, F]`` array when all rows of a split have the same length (every shipped
dataset) and a list of arrays otherwise; 'Y' is a ``float32 [N]`` array.  The arrays are views
of the parser's page-locked CSR buffers, so batches can be handed to ``cudaMemcpyAsync``
without another copy.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import CffmError


class LoadData:
    def __init__(self, path, dataset, loss_type):
        self.path = path + dataset + "/"
        self.trainfile = self.path + dataset + ".train.libfm"
        self.testfile = self.path + dataset + ".test.libfm"
        self.validationfile = self.path + dataset + ".validation.libfm"
        lib = _lib.load()
        self._lib = lib
        self._h = C.c_void_p()
        rc = lib.cffm_libfm_load(self.trainfile.encode(), self.testfile.encode(), self.validationfile.encode(),
                                 C.byref(self._h))
        if rc != 0:
            raise CffmError("cffm_libfm_load failed (%d): %s" % (rc, (lib.cffm_libfm_last_error() or b"").decode()))
        self.features_M = int(lib.cffm_libfm_features_M(self._h))
        use_log = loss_type == "log_loss"  # LoadData.py:59-76
        self.Train_data = self._split(0, use_log)
        self.Validation_data = self._split(1, use_log)
        self.Test_data = self._split(2, use_log)

    def _split(self, which, use_log):
        n = C.c_int64()
        rp, ids, yr, yl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = self._lib.cffm_libfm_split(self._h, which, C.byref(n), C.byref(rp), C.byref(ids), C.byref(yr), C.byref(yl))
        if rc != 0:
            raise CffmError("cffm_libfm_split failed")
        n = n.value
        if n == 0:
            return {"X": np.zeros((0, 0), np.int32), "Y": np.zeros((0,), np.float32)}
        row_ptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int64)), shape=(n + 1,))
        nnz = int(row_ptr[-1])
        flat = np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz]
        y = np.ctypeslib.as_array(C.cast(yl if use_log else yr, C.POINTER(C.c_float)), shape=(n,))
        lens = np.diff(row_ptr)
        if np.all(lens == lens[0]):
            X = flat.reshape(n, int(lens[0]))
        else:
            X = [flat[row_ptr[i]:row_ptr[i + 1]] for i in range(n)]
        return {"X": X, "Y": y, "row_ptr": row_ptr}

    def token(self, feature_id):
        buf = C.create_string_buffer(256)
        n = self._lib.cffm_libfm_token(self._h, int(feature_id), buf, 256)
        if n < 0:
            raise CffmError("no such feature id")
        return buf.value.decode()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            # the arrays handed out are views of these buffers: drop them first
            self.Train_data = self.Validation_data = self.Test_data = None
            self._lib.cffm_libfm_free(self._h)
            self._h = C.c_void_p()
