"""Thin object wrapper over a ``cffm_handle`` (one per device / rank).

``Engine`` plays the role of the reference's ``tf.Session`` + graph (CFFM.py:158-161, :531-541):
``forward`` is ``sess.run(self.out)`` (:596), ``train_step`` is
``sess.run((self.loss, self.optimizer))`` (:200).  All arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib
from ._lib import ACTIVATIONS, LOSSES, OPTIMIZERS, PRECISIONS, CffmError, Config, check


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, features_M, num_field, inner_dims=32, outer_dims=32, activation="relu",
                 loss_type="square_loss", lamda=0.0, lamda_att=1.0, lr=0.05, linear_att=1, att_dim=0,
                 inner_conv=1, outer_conv=1, beta_outer=1.0, optimizer="AdagradOptimizer", max_batch=1024,
                 precision="fp32", device=0, seed=2021, shard=None):
        """``shard=(rank, world)``: row-sharded tables -- this handle owns the rows ``r`` with ``r % world == rank``
        (local row ``r // world``); ``comm_init(id, rank, world)`` must follow before the first forward."""
        self.lib = _lib.load()
        if att_dim not in (0, num_field) and linear_att:
            # tf.matmul([B,F],[att_dim,att_dim]) only type-checks for att_dim == num_field
            raise CffmError("att_dim must be 0 or equal to num_field")
        if activation not in ACTIVATIONS:
            raise CffmError("unknown activation %r" % (activation,))
        if loss_type not in LOSSES:
            raise CffmError("loss_type %r leaves the loss undefined in the reference (SURVEY Q10)" % (loss_type,))
        if optimizer not in OPTIMIZERS:
            raise CffmError("unknown optimizer %r" % (optimizer,))
        cfg = Config(
            abi_version=_lib.ABI_VERSION, features_M=int(features_M), num_field=int(num_field),
            inner_dims=int(inner_dims), outer_dims=int(outer_dims), inner_conv=int(inner_conv),
            outer_conv=int(outer_conv), linear_att=int(linear_att), activation=ACTIVATIONS[activation],
            loss_type=LOSSES[loss_type], optimizer=OPTIMIZERS[optimizer], precision=PRECISIONS[precision],
            lr=float(lr), lamda=float(lamda), lamda_att=float(lamda_att), beta_outer=float(beta_outer),
            max_batch=int(max_batch), device=int(device), seed=int(seed),
            shard_world=int(shard[1]) if shard else 0, shard_rank=int(shard[0]) if shard else 0)
        self.shard = (int(shard[0]), int(shard[1])) if shard and int(shard[1]) > 1 else None
        self.features_M = int(features_M)
        self.cfg = cfg
        self.F = int(num_field)
        self.max_batch = int(max_batch)
        self.h = C.c_void_p()
        check(self.lib.cffm_create(C.byref(cfg), C.byref(self.h)), None, "cffm_create")
        self._params = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.cffm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        check(rc, self.h, what)

    # ------------------------------------------------------------------ variables
    def param_infos(self):
        if self._params is None:
            out = OrderedDict()
            n = self.lib.cffm_param_count(self.h)
            for i in range(n):
                name = C.create_string_buffer(128)
                shape = (C.c_int64 * 4)()
                ndim, numel, tr = C.c_int32(), C.c_int64(), C.c_int32()
                self._check(self.lib.cffm_param_info(self.h, i, name, 128, shape, C.byref(ndim), C.byref(numel),
                                                     C.byref(tr)), "cffm_param_info")
                out[name.value.decode()] = (tuple(shape[k] for k in range(ndim.value)), int(numel.value), bool(tr.value))
            self._params = out
        return self._params

    def _info(self, name):
        infos = self.param_infos()
        if name not in infos:
            raise CffmError("unknown variable: %s" % (name,))
        return infos[name]

    def get_param(self, name, accum=False):
        shape, numel, _ = self._info(name)
        a = np.empty(numel, dtype=np.float32)
        fn = self.lib.cffm_get_accum if accum else self.lib.cffm_get_param
        self._check(fn(self.h, name.encode(), _ptr(a), numel), "cffm_get_param")
        return a.reshape(shape)

    def set_param(self, name, value, accum=False):
        shape, numel, _ = self._info(name)
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float32).reshape(-1))
        if a.size != numel:
            raise CffmError("size mismatch for %s: %d vs %d" % (name, a.size, numel))
        fn = self.lib.cffm_set_accum if accum else self.lib.cffm_set_param
        self._check(fn(self.h, name.encode(), _ptr(a), numel), "cffm_set_param")

    def get_weights(self, accum=False):
        return OrderedDict((k, self.get_param(k, accum)) for k in self.param_infos())

    def set_weights(self, weights, accum=False):
        for k, v in weights.items():
            self.set_param(k, v, accum)

    def opt_slots(self):
        """Optimizer slots this handle keeps per variable: 1 (Adagrad accumulator / momentum / Adam m) and,
        for Adam, 2 (v).  Plain gradient descent has none."""
        opt = int(self.cfg.optimizer)
        return {OPTIMIZERS["GradientDescentOptimizer"]: 0, OPTIMIZERS["AdamOptimizer"]: 2}.get(opt, 1)

    def get_slot(self, name, slot):
        """slot 1: ``<name>``, slot 2: ``<name>:2`` of cffm_get_accum (same shape as the variable)."""
        shape, numel, _ = self._info(name)
        a = np.empty(numel, dtype=np.float32)
        key = name if slot == 1 else "%s:%d" % (name, slot)
        self._check(self.lib.cffm_get_accum(self.h, key.encode(), _ptr(a), numel), "cffm_get_accum")
        return a.reshape(shape)

    def set_slot(self, name, slot, value):
        shape, numel, _ = self._info(name)
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float32).reshape(-1))
        if a.size != numel:
            raise CffmError("size mismatch for %s: %d vs %d" % (name, a.size, numel))
        key = name if slot == 1 else "%s:%d" % (name, slot)
        self._check(self.lib.cffm_set_accum(self.h, key.encode(), _ptr(a), numel), "cffm_set_accum")

    def get_opt_step(self):
        v = C.c_int64()
        self._check(self.lib.cffm_get_opt_step(self.h, C.byref(v)), "cffm_get_opt_step")
        return int(v.value)

    def set_opt_step(self, step):
        self._check(self.lib.cffm_set_opt_step(self.h, int(step)), "cffm_set_opt_step")

    def state_dict(self):
        """Everything a resumed run needs: variables (``w:<name>``), optimizer slots (``a:<name>``,
        ``a2:<name>`` = Adam v) and the optimizer step counter (``opt_step``).  The reference's
        ``tf.train.Saver`` (CFFM.py:159, :226-228) stores all global variables, i.e. the same set."""
        out = OrderedDict()
        for k in self.param_infos():
            out["w:" + k] = self.get_param(k)
        for slot in range(1, self.opt_slots() + 1):
            for k in self.param_infos():
                out[("a:" if slot == 1 else "a2:") + k] = self.get_slot(k, slot)
        out["opt_step"] = np.int64(self.get_opt_step())
        return out

    def load_state_dict(self, state):
        for k in state:
            if k == "opt_step":
                self.set_opt_step(int(state[k]))
                continue
            kind, name = k.split(":", 1)
            if kind == "w":
                self.set_param(name, state[k])
            elif kind in ("a", "a2"):
                slot = 1 if kind == "a" else 2
                if slot <= self.opt_slots():
                    self.set_slot(name, slot, state[k])
            else:
                raise CffmError("unknown checkpoint entry %r" % (k,))

    # ------------------------------------------------------------------ row-sharded tables
    TABLES = ("inner_embeddings", "outer_embeddings", "feature_bias")

    def owned_rows(self):
        """Global row numbers of this handle's table rows, in local order."""
        if not self.shard:
            return np.arange(self.features_M)
        rank, world = self.shard
        return np.arange(rank, self.features_M, world)

    def set_table_from_global(self, name, full, accum=False):
        """Load this handle's rows of a full [features_M, K] table (variable or optimizer slot 1)."""
        self.set_param(name, np.asarray(full)[self.owned_rows()], accum)

    def init_params(self, seed):
        self._check(self.lib.cffm_init_params(self.h, int(seed)), "cffm_init_params")

    # ------------------------------------------------------------------ compute (host buffers)
    def _ids(self, ids):
        a = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
        if a.ndim != 2 or a.shape[1] != self.F:
            raise CffmError("ids must be [B, %d], got %r" % (self.F, a.shape))
        return a

    def forward(self, ids):
        a = self._ids(ids)
        out = np.empty(a.shape[0], dtype=np.float32)
        self._check(self.lib.cffm_forward_host(self.h, _ptr(a), a.shape[0], _ptr(out)), "cffm_forward_host")
        return out

    def train_step(self, ids, labels):
        a = self._ids(ids)
        y = np.ascontiguousarray(np.asarray(labels, dtype=np.float32).reshape(-1))
        if y.shape[0] != a.shape[0]:
            raise CffmError("labels/ids batch mismatch")
        loss = C.c_float()
        self._check(self.lib.cffm_train_step_host(self.h, _ptr(a), _ptr(y), a.shape[0], C.byref(loss)),
                    "cffm_train_step_host")
        return float(loss.value)

    def train_submit(self, ids, labels):
        """Pipelined step: returns the loss of the previously submitted step, or None."""
        a = self._ids(ids)
        y = np.ascontiguousarray(np.asarray(labels, dtype=np.float32).reshape(-1))
        loss, n = C.c_float(), C.c_int32()
        self._check(self.lib.cffm_train_submit_host(self.h, _ptr(a), _ptr(y), a.shape[0], C.byref(loss), C.byref(n)),
                    "cffm_train_submit_host")
        return float(loss.value) if n.value else None

    def train_flush(self):
        loss, n = C.c_float(), C.c_int32()
        self._check(self.lib.cffm_train_flush(self.h, C.byref(loss), C.byref(n)), "cffm_train_flush")
        return float(loss.value) if n.value else None

    def evaluate(self, ids, labels, batch=0):
        a = self._ids(ids)
        y = np.ascontiguousarray(np.asarray(labels, dtype=np.float32).reshape(-1))
        rmse, r2 = C.c_double(), C.c_double()
        self._check(self.lib.cffm_evaluate_host(self.h, _ptr(a), _ptr(y), a.shape[0], int(batch), C.byref(rmse),
                                                C.byref(r2)), "cffm_evaluate_host")
        return float(rmse.value), float(r2.value)

    # ------------------------------------------------------------------ resident training set
    def dataset_upload(self, ids, labels):
        a = self._ids(ids)
        y = np.ascontiguousarray(np.asarray(labels, dtype=np.float32).reshape(-1))
        self._check(self.lib.cffm_dataset_upload(self.h, _ptr(a), _ptr(y), a.shape[0]), "cffm_dataset_upload")

    def dataset_permute(self, perm):
        p = np.ascontiguousarray(np.asarray(perm, dtype=np.int64))
        self._check(self.lib.cffm_dataset_permute(self.h, _ptr(p)), "cffm_dataset_permute")

    def train_block(self, start, B):
        self._check(self.lib.cffm_train_block(self.h, int(start), int(B)), "cffm_train_block")

    def last_loss(self):
        v = C.c_float()
        self._check(self.lib.cffm_last_loss(self.h, C.byref(v)), "cffm_last_loss")
        return float(v.value)

    def dataset_evaluate(self, batch=0):
        rmse, r2 = C.c_double(), C.c_double()
        self._check(self.lib.cffm_dataset_evaluate(self.h, int(batch), C.byref(rmse), C.byref(r2)), "cffm_dataset_evaluate")
        return float(rmse.value), float(r2.value)

    # ------------------------------------------------------------------ compute (device buffers)
    def forward_dev(self, ids_ptr, B, out_ptr, stream=0):
        self._check(self.lib.cffm_forward_dev(self.h, C.c_void_p(ids_ptr), int(B), C.c_void_p(out_ptr),
                                              C.c_void_p(stream)), "cffm_forward_dev")

    def train_step_dev(self, ids_ptr, labels_ptr, B, loss_ptr=0, stream=0):
        self._check(self.lib.cffm_train_step_dev(self.h, C.c_void_p(ids_ptr), C.c_void_p(labels_ptr), int(B),
                                                 C.c_void_p(loss_ptr), C.c_void_p(stream)), "cffm_train_step_dev")

    def synchronize(self):
        self._check(self.lib.cffm_synchronize(self.h), "cffm_synchronize")

    def profile(self, on=True):
        self._check(self.lib.cffm_profile_enable(self.h, 1 if on else 0), "cffm_profile_enable")

    def profile_report(self, reset=True):
        """{tag: (launches, total_ms)} of the event-bracketed launches since the last reset."""
        buf = C.create_string_buffer(1 << 16)
        self.lib.cffm_profile_report(self.h, buf, len(buf), 1 if reset else 0)
        out = {}
        for line in buf.value.decode().splitlines():
            tag, n, ms = line.split()
            out[tag] = (int(n), float(ms))
        return out

    def uses_graph(self):
        """Whether training steps currently replay from a CUDA graph (CFFM_GRAPH=0 or a failed capture turn it off)."""
        return bool(self.lib.cffm_uses_graph(self.h))

    def launch_count(self):
        return int(self.lib.cffm_launch_count(self.h))

    # ------------------------------------------------------------------ introspection / DP
    def fetch(self, what):
        n = C.c_int64()
        self._check(self.lib.cffm_debug_fetch(self.h, what.encode(), None, 0, C.byref(n)), "cffm_debug_fetch")
        a = np.empty(n.value, dtype=np.float32)
        self._check(self.lib.cffm_debug_fetch(self.h, what.encode(), _ptr(a), n.value, C.byref(n)), "cffm_debug_fetch")
        return a

    def dense_grad(self, name):
        shape, numel, _ = self._info(name)
        a = np.empty(numel, dtype=np.float32)
        self._check(self.lib.cffm_debug_dense_grad(self.h, name.encode(), _ptr(a), numel), "cffm_debug_dense_grad")
        return a.reshape(shape)

    def comm_init(self, unique_id, rank, world):
        self._check(self.lib.cffm_comm_init(self.h, unique_id, int(rank), int(world)), "cffm_comm_init")


def comm_unique_id():
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    check(lib.cffm_comm_unique_id(buf), None, "cffm_comm_unique_id")
    return buf.raw
