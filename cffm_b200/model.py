"""Reference-facing model class: same constructor signature, ``train(data)`` / ``evaluate(data)``
and result attributes as class ``CFFM`` in the reference (CFFM.py:97-648), with the TF graph
replaced by the CUDA engine.  Host logic only; there is no arithmetic here beyond batching."""
from __future__ import annotations

import logging
import math
import os
from time import time

import numpy as np

from .engine import Engine
from ._lib import CffmError


def _shuffle_in_unison(x, y, seed):
    """sklearn.utils.shuffle(x, y, random_state=seed) (CFFM.py:556-558): one fixed permutation,
    re-applied to the current order every epoch."""
    n = len(y)
    perm = np.random.RandomState(seed).permutation(n)
    if isinstance(x, np.ndarray):
        return x[perm], np.asarray(y)[perm]
    return [x[i] for i in perm], np.asarray(y)[perm]


class CFFM:
    def __init__(self, features_M, pretrain_flag, save_file, inner_dims, outer_dims, loss_type, epoch, batch_size,
                 learning_rate, lamda_bilinear, keep, optimizer_type, batch_norm, verbose, tensorboard, num_field,
                 linear_att, att_dim, lamda_att, inner_conv, gamma_inner, outer_conv, beta_outer, activation_function,
                 random_seed=2021, precision="fp32", device=0, eval_batch=None, batch_seed=None):
        # CFFM.py:102-130 -- bind params
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.inner_dims = inner_dims
        self.outer_dims = outer_dims
        self.pretrain_flag = pretrain_flag
        self.save_file = save_file
        self.loss_type = loss_type
        self.features_M = features_M
        self.lamda_bilinear = lamda_bilinear
        self.keep = keep                  # accepted, unused
        self.epoch = epoch
        self.random_seed = random_seed
        self.optimizer_type = optimizer_type
        self.batch_norm = batch_norm      # accepted, unused (Q14)
        self.verbose = verbose
        self.tensorboard = tensorboard    # the reference's tensorboard branch crashes; ignored here
        self.num_field = num_field
        self.linear_att = linear_att
        self.att_dim = num_field if att_dim == 0 else att_dim
        self.lamda_att = lamda_att
        self.inner_conv = inner_conv
        self.gamma_inner = gamma_inner    # accepted, unused (Q14)
        self.outer_conv = outer_conv
        self.beta_outer = beta_outer
        self.activation_function = activation_function
        self.num_interactions = int(self.num_field * (self.num_field - 1) / 2)
        self.precision = precision
        self.device = device
        self.eval_batch = eval_batch      # None: evaluate with batch_size like the reference
        self._rng = np.random if batch_seed is None else np.random.RandomState(batch_seed)
        if save_file and not os.path.exists(os.path.dirname(save_file) or "."):
            os.makedirs(os.path.dirname(save_file) or ".", exist_ok=True)  # CFFM.py:637-639
        self.train_rmse, self.valid_rmse, self.test_rmse = [], [], []
        self.train_r2, self.valid_r2, self.test_r2 = [], [], []
        self.engine = None

    # ------------------------------------------------------------------ graph / session
    def build_graph(self):
        """CFFM.py:531-541: variables + inference + loss + optimizer become one engine handle."""
        if self.engine is not None:
            return self.engine
        max_batch = max(int(self.batch_size), int(self.eval_batch or 0))
        self.engine = Engine(
            self.features_M, self.num_field, self.inner_dims, self.outer_dims, activation=self.activation_function,
            loss_type=self.loss_type, lamda=self.lamda_bilinear, lamda_att=self.lamda_att, lr=self.learning_rate,
            linear_att=self.linear_att, att_dim=self.att_dim, inner_conv=self.inner_conv, outer_conv=self.outer_conv,
            beta_outer=float(self.beta_outer), optimizer=self.optimizer_type, max_batch=max_batch,
            precision=self.precision, device=self.device, seed=self.random_seed)
        if self.pretrain_flag > 0:
            self.load_state(self.save_file + ".npz")
        return self.engine

    def calculate_parameters(self):
        """CFFM.py:543-553 counts self.weights only -- not the tf.layers.dense layers."""
        total = 0
        for name, (shape, numel, _) in self.engine.param_infos().items():
            if not name.startswith("dense"):
                total += numel
        if self.verbose > 0:
            logging.info("#params: %d" % total)
        return total

    # ------------------------------------------------------------------ checkpoint (SURVEY next-3)
    def save_state(self, path):
        """Variables, every optimizer slot (Adam: m and v) and the optimizer step counter."""
        np.savez(path, **self.engine.state_dict())

    def load_state(self, path):
        z = np.load(path)
        self.engine.load_state_dict({k: z[k] for k in z.files})

    # ------------------------------------------------------------------ training loop
    def train(self, data):
        self.build_graph()
        eng = self.engine
        self.calculate_parameters()
        if self.verbose > 0:
            t2 = time()
            tr = self.evaluate(data.Train_data)
            va = self.evaluate(data.Validation_data)
            te = self.evaluate(data.Test_data)
            logging.info(("Init_RMSE: train=%.4f,validation=%.4f,test=%.4f | Init_R2: train=%.4f,validation=%.4f,"
                          "test=%.4f [%.1f s] " % (tr[0], va[0], te[0], tr[1], va[1], te[1], time() - t2)))
        # Equal-length rows (every shipped dataset): the training split lives in HBM, an epoch's shuffle is
        # a device gather and a step reads its contiguous block in place -- no per-step host->device copy.
        resident = isinstance(data.Train_data['X'], np.ndarray) and data.Train_data['X'].ndim == 2 \
            and data.Train_data['X'].shape[1] == self.num_field
        if resident:
            eng.dataset_upload(data.Train_data['X'], data.Train_data['Y'])
        for epoch in range(self.epoch):
            t1 = time()
            n_train = len(data.Train_data['Y'])
            if resident:
                perm = np.random.RandomState(self.random_seed).permutation(n_train)  # CFFM.py:183, :556-558
                data.Train_data['X'], data.Train_data['Y'] = data.Train_data['X'][perm], np.asarray(data.Train_data['Y'])[perm]
                eng.dataset_permute(perm)
            else:
                data.Train_data['X'], data.Train_data['Y'] = _shuffle_in_unison(
                    data.Train_data['X'], data.Train_data['Y'], self.random_seed)
            total_batch = int(n_train / self.batch_size)  # :185
            for _ in range(total_batch):
                if resident:
                    eng.train_block(self._rng.randint(0, n_train - self.batch_size), self.batch_size)  # :188, :561
                else:
                    blk = self.get_random_block_from_data(data.Train_data, self.batch_size)
                    eng.train_submit(blk['X'], blk['Y'])  # :200, pipelined: loss values are not consumed by the loop
            if resident:
                eng.synchronize()
            else:
                eng.train_flush()
            t2 = time()
            train_rmse, train_r2 = (eng.dataset_evaluate(int(self.eval_batch or self.batch_size)) if resident
                                    else self.evaluate(data.Train_data))
            valid_rmse, valid_r2 = self.evaluate(data.Validation_data)
            test_rmse, test_r2 = self.evaluate(data.Test_data)
            self.train_rmse.append(train_rmse); self.valid_rmse.append(valid_rmse); self.test_rmse.append(test_rmse)
            self.train_r2.append(train_r2); self.valid_r2.append(valid_r2); self.test_r2.append(test_r2)
            if self.verbose > 0 and epoch % self.verbose == 0:
                logging.info(("Epoch %d [%.1f s] RMSE: train=%.4f,validation=%.4f,Test=%.4f | R2: train=%.4f,"
                              "validation=%.4f,Test=%.4f [%.1f s]" % (epoch + 1, t2 - t1, train_rmse, valid_rmse,
                                                                      test_rmse, train_r2, valid_r2, test_r2,
                                                                      time() - t2)))
            if self.eva_termination(self.valid_rmse):
                break
            if self.pretrain_flag < 0:
                logging.info("Save model to file as pretrain.")
                self.save_state(self.save_file + ".npz")

    def get_random_block_from_data(self, data, batch_size):
        """CFFM.py:560-581: a contiguous block starting at a random index (Q12)."""
        n = len(data['Y'])
        start = self._rng.randint(0, n - batch_size)
        X, Y = data['X'], data['Y']
        if isinstance(X, np.ndarray):  # all rows have the same length: the forward fill always completes
            return {'X': X[start:start + batch_size], 'Y': np.asarray(Y[start:start + batch_size])}
        bx, by = [], []
        i = start
        while len(bx) < batch_size and i < n:  # forward
            if len(X[i]) == len(X[start]):
                bx.append(X[i]); by.append(Y[i]); i += 1
            else:
                break
        i = start
        while len(bx) < batch_size and i >= 0:  # backward (restarts at `start`, as in the reference)
            if len(X[i]) == len(X[start]):
                bx.append(X[i]); by.append(Y[i]); i -= 1
            else:
                break
        return {'X': np.asarray(bx, dtype=np.int32), 'Y': np.asarray(by, dtype=np.float32)}

    def get_ordered_block_from_data(self, data, batch_size, index):
        """CFFM.py:617-629."""
        start = index * batch_size
        X, Y = data['X'], data['Y']
        n = len(Y)
        if isinstance(X, np.ndarray):
            return {'X': X[start:start + batch_size], 'Y': np.asarray(Y[start:start + batch_size])}
        bx, by = [], []
        i = start
        while len(bx) < batch_size and i < n:
            if len(X[i]) == len(X[start]):
                bx.append(X[i]); by.append(Y[i]); i += 1
            else:
                break
        return {'X': np.asarray(bx, dtype=np.int32).reshape(len(bx), -1), 'Y': np.asarray(by, dtype=np.float32)}

    def evaluate(self, data):
        """CFFM.py:583-615 -> (RMSE, R2).  For equal-length rows the whole pass (ordered blocks,
        clipping, both metrics) runs on the device and only two scalars come back."""
        if self.engine is None:
            self.build_graph()
        X, Y = data['X'], data['Y']
        num_example = len(Y)
        if num_example == 0:
            raise CffmError("evaluate() needs at least one example")
        bs = int(self.eval_batch or self.batch_size)
        if isinstance(X, np.ndarray) and X.shape[1] == self.num_field:
            return self.engine.evaluate(X, Y, bs)
        # Ragged rows: the reference's block stops at the first length change while the next block still starts at
        # (k+1)*bs, so rows are skipped and its metric call then fails on the length mismatch (CFFM.py:597-612).
        # Scored here: exactly the rows the blocks covered, each paired with its own label; the clip bounds come
        # from the full label vector as in the reference (:609-611).
        preds, rows, idx = [], [], 0
        blk = self.get_ordered_block_from_data(data, bs, idx)
        while len(blk['X']) > 0:
            preds.append(self.engine.forward(blk['X']))
            rows.append(np.arange(idx * bs, idx * bs + len(blk['X'])))
            idx += 1
            blk = self.get_ordered_block_from_data(data, bs, idx)
        y_pred = np.concatenate(preds).astype(np.float64)
        y_all = np.asarray(Y, dtype=np.float64)
        y_true = y_all[np.concatenate(rows)]
        pb = np.minimum(np.maximum(y_pred, y_all.min()), y_all.max())
        rmse = math.sqrt(float(np.mean((y_true - pb) ** 2)))
        sst = float(np.sum((y_true - y_true.mean()) ** 2))
        r2 = 1.0 - float(np.sum((y_true - pb) ** 2)) / sst if sst > 0 else 0.0
        return rmse, r2

    def eva_termination(self, valid):
        """CFFM.py:631-635."""
        if len(valid) > 5:
            if valid[-1] > valid[-2] > valid[-3] > valid[-4] > valid[-5]:
                return True
        return False
