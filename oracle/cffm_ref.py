"""CPU oracle for the CFFM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``cffm_b200``) never routes through it; it fails loudly when its CUDA library
is missing.

**PARITY UNPINNED.**  The reference's arithmetic lives in ``tensorflow==1.14.0``
(pin: reference ``README.md:9``), which is absent from ``/root/reference`` and not
installable here (Python 3.12, no network).  The reference ships no tests and no
golden vectors for this path.  This file is therefore a restatement of the graph
that ``CFFM.py`` builds, op by op, using TF 1.14's documented semantics
(SURVEY.md App. B); it cannot be checked against outputs of the reference itself.
What pins it: fp64 finite-difference gradient checks, closed-form identities and
the shape table of SURVEY.md §8.3 (``tests/test_oracle.py``).

Every function cites the reference lines (relative to ``/root/reference``) that
it restates.  Quirk numbers (Q1..Q15) refer to SURVEY.md §8.1.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

SELU_SCALE = 1.0507009873554805
SELU_ALPHA = 1.6732632423543772


# --------------------------------------------------------------------------- activations
def act_fn(name):
    """CFFM.py:132-141 (+ :149-155 for gelu / prelu)."""
    if name == "relu":
        return torch.relu
    if name == "elu":
        return torch.nn.functional.elu
    if name == "selu":
        return torch.selu
    if name == "prelu":  # CFFM.py:153-155 -- fixed slope 0.25, not learned
        return lambda x: torch.relu(x) + 0.25 * (-torch.relu(-x))
    if name == "gelu":  # CFFM.py:149-151 -- exact erf form
        return lambda x: x * (0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))))
    raise ValueError("unknown activation %r" % (name,))


def pair_list(num_field):
    """Pair order of the two Python loops at CFFM.py:304-305 / :355-356."""
    return [(i, j) for i in range(num_field) for j in range(i + 1, num_field)]


def truncated_normal(gen, shape, std, dtype):
    """tf.truncated_normal (CFFM.py:460): redraw until within two sigma."""
    n = int(np.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.float64)
    filled = 0
    while filled < n:
        cand = torch.randn(2 * (n - filled) + 16, generator=gen, dtype=torch.float64)
        cand = cand[cand.abs() <= 2.0][: n - filled]
        out[filled:filled + cand.numel()] = cand
        filled += cand.numel()
    return (out * std).reshape(shape).to(dtype)


class CFFMRef:
    """Restatement of class ``CFFM`` (CFFM.py:97-648): parameters, forward graph,
    loss, optimizer.  Tensors are torch CPU tensors of ``dtype``."""

    def __init__(self, features_M, num_field, inner_dims=32, outer_dims=32, activation="relu",
                 loss_type="square_loss", lamda=0.0, lamda_att=1.0, lr=0.05, linear_att=1, att_dim=0,
                 inner_conv=1, outer_conv=1, beta_outer=1.0, optimizer="AdagradOptimizer",
                 dtype=torch.float32, seed=2021):
        self.M, self.F = int(features_M), int(num_field)
        self.Ki, self.Ko = int(inner_dims), int(outer_dims)
        self.activation_name = activation
        self.act = act_fn(activation)
        self.loss_type, self.lamda, self.lamda_att, self.lr = loss_type, float(lamda), float(lamda_att), float(lr)
        self.linear_att, self.inner_conv, self.outer_conv = int(linear_att), int(inner_conv), int(outer_conv)
        self.att_dim = self.F if att_dim == 0 else int(att_dim)  # CFFM.py:121-124
        self.beta_outer = beta_outer
        self.optimizer = optimizer
        self.dtype = dtype
        self.P = int(self.F * (self.F - 1) / 2)  # CFFM.py:130
        self.pairs = pair_list(self.F)
        self.conv_depth = int(math.log(self.Ko, 2)) if self.outer_conv else 0  # CFFM.py:373
        self.params = OrderedDict()
        self.state = {}  # optimizer slots
        self.step_count = 0
        self._init_params(seed)

    # ---------------------------------------------------------------- parameters
    def _init_params(self, seed):
        """CFFM.py:239-293 (initialize_variables), :323, :376-377 (conv weights created inside the
        inference function), :339/:409/:410/:441 (tf.layers.dense, Q6) and Q7 for the initialisers.
        The reference is unseeded; a seed is taken here so tests are reproducible."""
        g = torch.Generator().manual_seed(seed)
        dt, M, F, P = self.dtype, self.M, self.F, self.P
        p = self.params

        def normal(shape, std):
            return (torch.randn(shape, generator=g, dtype=torch.float64) * std).to(dt)

        def glorot(fan_in, fan_out):  # tf.layers.dense default kernel init [TF-1.14]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return ((torch.rand(fan_in, fan_out, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dt)

        if self.inner_conv == 1:
            p["inner_embeddings"] = normal((M, self.Ki), 0.1)  # :257-259
        if self.outer_conv == 1:
            p["outer_embeddings"] = normal((M, self.Ko), 0.01)  # :264-266
            p["outer_W"] = truncated_normal(g, (P, 1), 1.0, dt)  # :271 (unused, Q14)
            p["outer_b"] = truncated_normal(g, (1,), 1.0, dt)  # :272 (unused, Q14)
        p["feature_bias"] = torch.zeros(M, 1, dtype=dt)  # :276-277, stddev 0.0
        if self.linear_att == 1:
            p["bias_W"] = truncated_normal(g, (self.att_dim, self.att_dim), 1.0, dt)  # :281
            p["bias_b"] = truncated_normal(g, (self.att_dim,), 1.0, dt)  # :282
        p["bias"] = torch.zeros((), dtype=dt)  # :284
        dense_idx = [0]

        def dense_name():
            n = "dense" if dense_idx[0] == 0 else "dense_%d" % dense_idx[0]
            dense_idx[0] += 1
            return n

        if self.inner_conv == 1:
            p["inner_layer_conv_weight_0"] = truncated_normal(g, (1, 2, 1, 2), 1.0, dt)  # :323
            p["inner_layer_conv_bias_0"] = torch.full((2,), 0.01, dtype=dt)  # :466
            n = dense_name()  # :339 -- input width P*K (Q5 generalised from the literal 16*2)
            p[n + "/kernel"] = glorot(P * self.Ki, 1)
            p[n + "/bias"] = torch.zeros(1, dtype=dt)
            self.dense_inner = n
        if self.outer_conv == 1:
            for i in range(self.conv_depth):  # :375-377
                p["outer_layer_conv_weight_%d" % i] = truncated_normal(g, (2, 2, P, P), 1.0, dt)
                p["outer_layer_conv_bias_%d" % i] = torch.full((P,), 0.01, dtype=dt)
            n1, n2 = dense_name(), dense_name()  # :409-410
            t1_dim = sum(self.Ko >> l for l in range(self.conv_depth))  # Q1: 2K-2 for K a power of two
            p[n1 + "/kernel"] = glorot(t1_dim, 32)
            p[n1 + "/bias"] = torch.zeros(32, dtype=dt)
            p[n2 + "/kernel"] = glorot(32, 1)
            p[n2 + "/bias"] = torch.zeros(1, dtype=dt)
            self.dense_outer1, self.dense_outer2 = n1, n2
        if self.linear_att == 1:
            n = dense_name()  # :441
            p[n + "/kernel"] = glorot(self.F, 1)
            p[n + "/bias"] = torch.zeros(1, dtype=dt)
            self.dense_linear = n

    def dead_params(self):
        """Variables unreachable from the loss: no gradient, never updated (Q2, Q14)."""
        dead = set()
        if self.outer_conv == 1:
            dead |= {"outer_W", "outer_b"}
            if self.conv_depth >= 1:
                dead |= {"outer_layer_conv_weight_%d" % (self.conv_depth - 1),
                         "outer_layer_conv_bias_%d" % (self.conv_depth - 1)}
        return dead

    # ---------------------------------------------------------------- forward
    def forward(self, ids, params=None, return_intermediates=False):
        """CFFM.py:296-453.  ``ids``: integer [B, F].  Returns ``out`` [B, 1] (pre-sigmoid)."""
        p = self.params if params is None else params
        ids = torch.as_tensor(np.asarray(ids), dtype=torch.long)
        B = ids.shape[0]
        comps = []
        inter = {}
        act = self.act
        if self.inner_conv == 1:
            K = self.Ki
            emb = p["inner_embeddings"][ids]  # :303  [B,F,K]
            inner = [emb[:, i, :] * emb[:, j, :] for (i, j) in self.pairs]  # :304-310
            x = torch.stack(inner)  # :313  [P,B,K]
            x = x.permute(1, 0, 2)  # :315  [B,P,K]
            x = x.unsqueeze(-1)  # :317  [B,P,K,1]
            x = act(x)  # :319 (Q4: full activation on raw products)
            # :323-328 conv2d filter [1,2,1,2], strides [1,1,2,1], VALID, then relu(conv+b) (:475-478)
            w = p["inner_layer_conv_weight_0"]
            xv = x[..., 0].reshape(B, self.P, K // 2, 2)  # taps t=0,1 along the K axis
            conv = torch.einsum("bpwt,to->bpwo", xv, w[0, :, 0, :]) + p["inner_layer_conv_bias_0"]
            c1 = act(torch.relu(conv))  # :478 then :330 (Q3 double activation)
            # :331 max_pool ksize [1,1,2,1] of the ACTIVATED INPUT; ties route to the first tap (Q15)
            a0, a1 = xv[..., 0], xv[..., 1]
            mp = torch.where(a0 >= a1, a0, a1).unsqueeze(-1)  # [B,P,K/2,1]
            r = c1 + mp  # :332 broadcast over the two output channels
            flat = r.reshape(B, self.P * (K // 2) * 2)  # :333 (Q5), order (p, w, o)
            final2 = flat @ p[self.dense_inner + "/kernel"] + p[self.dense_inner + "/bias"]  # :339
            comps.append(final2)
            inter["inner_max"] = flat
            inter["final2"] = final2
        if self.outer_conv == 1:
            K = self.Ko
            oe = p["outer_embeddings"][ids]  # :354
            outer = [oe[:, i, :].unsqueeze(-1) * oe[:, j, :].unsqueeze(1) for (i, j) in self.pairs]  # :355-362
            x = torch.stack(outer).permute(1, 2, 3, 0)  # :365-367  NHWC [B,K,K,P]
            sum_pooling = [x.sum(dim=(2, 3))]  # :381 (Q1: sums W and channels, keeps H)
            for l in range(self.conv_depth):  # :384-391 (layer d-1 is dead, Q2, but built)
                H = x.shape[1]
                if H < 2:
                    break
                Hh = H // 2
                xv = x[:, :2 * Hh, :2 * Hh, :].reshape(B, Hh, 2, Hh, 2, self.P)
                w = p["outer_layer_conv_weight_%d" % l]
                y = torch.einsum("bhiwjc,ijco->bhwo", xv, w) + p["outer_layer_conv_bias_%d" % l]  # :476
                x = act(torch.relu(y))  # :478, :387 (Q3)
                sum_pooling.append(x.sum(dim=(2, 3)))  # :390-391
                inter["conv_%d" % l] = x
            t1 = torch.cat(sum_pooling[: self.conv_depth], dim=1)  # :394-396 (Q1/Q2)
            h = t1 @ p[self.dense_outer1 + "/kernel"] + p[self.dense_outer1 + "/bias"]  # :409
            fin = h @ p[self.dense_outer2 + "/kernel"] + p[self.dense_outer2 + "/bias"]  # :410
            fin = (self.beta_outer * fin).reshape(-1, 1)  # :414
            comps.append(fin)
            inter["t1"] = t1
            inter["final"] = fin
        fb = p["feature_bias"][ids]  # :422 [B,F,1]
        if self.linear_att == 1:
            fb = fb[..., 0]  # :425 (tf.squeeze; B=1 edge case not reproduced, Q8)
            lin = fb @ p["bias_W"] + p["bias_b"]  # :432
            lin = lin / self.lamda_att  # :434
            lin = torch.softmax(lin, dim=-1)  # :436
            lin = (fb * lin).reshape(-1, self.F)  # :438
            lin = lin @ p[self.dense_linear + "/kernel"] + p[self.dense_linear + "/bias"]  # :441
        else:
            lin = fb.sum(dim=1)  # :444
        comps.append(lin)
        inter["linear"] = lin
        comps.append(p["bias"] * torch.ones(B, 1, dtype=self.dtype))  # :449
        out = comps[0]
        for c in comps[1:]:  # :453 tf.add_n
            out = out + c
        if return_intermediates:
            return out, inter
        return out

    def predict(self, ids, params=None):
        """What ``sess.run(self.out)`` returns (CFFM.py:596): sigmoid applied for log_loss (:496)."""
        out = self.forward(ids, params)
        if self.loss_type == "log_loss":
            out = torch.sigmoid(out)
        return out

    # ---------------------------------------------------------------- loss
    @staticmethod
    def _tf_log_loss(pred, y, eps=1e-7):
        """tf.contrib.losses.log_loss [TF-1.14]: mean of -y log(p+eps) - (1-y) log(1-p+eps)."""
        return (-(y * torch.log(pred + eps)) - (1 - y) * torch.log(1 - pred + eps)).mean()

    def loss(self, ids, y, params=None):
        """CFFM.py:486-514."""
        p = self.params if params is None else params
        y = torch.as_tensor(np.asarray(y), dtype=self.dtype).reshape(-1, 1)
        out = self.forward(ids, p)
        lt = self.loss_type
        if lt == "square_loss":
            if self.lamda > 0:  # :489-491 (Q9: outer table regularised by lamda_att)
                l = 0.5 * ((y - out) ** 2).sum()
                l = l + self.lamda * 0.5 * (p["inner_embeddings"] ** 2).sum()
                l = l + self.lamda_att * 0.5 * (p["outer_embeddings"] ** 2).sum()
                return l
            return torch.sqrt(((y - out) ** 2).mean() + 1e-10)  # :493
        if lt == "log_loss":  # :495-504 (lamda>0 branch raises KeyError in the reference, Q10)
            return self._tf_log_loss(torch.sigmoid(out), y)
        if lt == "mse":
            return ((y - out) ** 2).mean()  # :506
        if lt == "mae":
            return (y - out).abs().mean()  # :508
        if lt == "hybrid":  # :511-513 (log_loss of the raw, un-squashed output, as written)
            return 0.5 * (0.5 * ((y - out) ** 2).sum()) + 0.5 * self._tf_log_loss(out, y)
        raise ValueError("loss_type %r leaves self.loss undefined in the reference (Q10)" % (lt,))

    # ---------------------------------------------------------------- gradients
    SPARSE_TABLES = ("inner_embeddings", "outer_embeddings", "feature_bias")

    def gradients(self, ids, y):
        """What ``Optimizer.minimize`` differentiates (CFFM.py:517-529) [TF-1.14].

        Returns (loss, dense_grads{name: tensor}, sparse_grads{name: (unique_rows, summed_rows)}).
        Gathered tables get IndexedSlices gradients which TF de-duplicates by summation
        (unique + unsorted_segment_sum, summation in order of appearance) -- Q11.  When lamda>0 the
        regulariser makes the table gradients dense (Q9); they are then returned in dense_grads."""
        ids_t = torch.as_tensor(np.asarray(ids), dtype=torch.long)
        leaves = OrderedDict()
        gathered = {}
        p2 = OrderedDict()
        dense_tables = self.loss_type == "square_loss" and self.lamda > 0
        for k, v in self.params.items():
            if k in self.SPARSE_TABLES and not (dense_tables and k != "feature_bias"):
                # differentiate w.r.t. the gathered rows: that is the IndexedSlices `values`
                rows = v[ids_t].detach().clone().requires_grad_(True)  # [B,F,*]
                gathered[k] = rows
                p2[k] = _GatherProxy(v, ids_t, rows)
            else:
                leaves[k] = v.detach().clone().requires_grad_(True)
                p2[k] = leaves[k]
        l = self.loss(ids, y, p2)
        wrt = list(leaves.values()) + list(gathered.values())
        grads = torch.autograd.grad(l, wrt, allow_unused=True)
        dense = {}
        for (k, _), g in zip(leaves.items(), grads[: len(leaves)]):
            dense[k] = g  # None for dead params
        sparse = {}
        flat_ids = ids_t.reshape(-1).numpy()
        uniq, inv = np.unique(flat_ids, return_inverse=True)
        for (k, rows), g in zip(gathered.items(), grads[len(leaves):]):
            vals = g.reshape(flat_ids.shape[0], -1).numpy()
            summed = np.zeros((uniq.shape[0], vals.shape[1]), dtype=vals.dtype)
            np.add.at(summed, inv, vals)  # sequential, in order of appearance
            sparse[k] = (uniq.astype(np.int64), torch.from_numpy(summed), torch.from_numpy(vals.copy()))
        return l.detach(), dense, sparse

    # ---------------------------------------------------------------- optimizer
    def train_step(self, ids, y):
        """One ``sess.run((self.loss, self.optimizer))`` (CFFM.py:200).  Returns the loss (pre-update)."""
        l, dense, sparse = self.gradients(ids, y)
        self.step_count += 1
        for k, g in dense.items():
            if g is None:
                continue  # minimize() skips variables with None gradients (Q2)
            self._apply_dense(k, g)
        for k, (rows, g, _) in sparse.items():
            self._apply_sparse(k, rows, g)
        return float(l)

    def _slot(self, name, key, init):
        d = self.state.setdefault(key, {})
        if name not in d:
            d[name] = torch.full_like(self.params[name], init)
        return d[name]

    def _apply_dense(self, name, g):
        """[TF-1.14] dense Apply* kernels for the optimizers at CFFM.py:519-529."""
        w, lr = self.params[name], self.lr
        g = g.reshape(w.shape).to(w.dtype)
        if self.optimizer == "AdagradOptimizer":  # acc0 = 1e-8, no epsilon (Q11)
            acc = self._slot(name, "accumulator", 1e-8)
            acc += g * g
            w -= lr * g * torch.rsqrt(acc)
        elif self.optimizer == "GradientDescentOptimizer":
            w -= lr * g
        elif self.optimizer == "MomentumOptimizer":  # momentum 0.95
            acc = self._slot(name, "momentum", 0.0)
            acc.mul_(0.95).add_(g)
            w -= lr * acc
        elif self.optimizer == "AdamOptimizer":  # beta1 .9, beta2 .999, eps 1e-8
            m, v = self._slot(name, "m", 0.0), self._slot(name, "v", 0.0)
            t = self.step_count
            lr_t = lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
            m.mul_(0.9).add_(0.1 * g)
            v.mul_(0.999).add_(0.001 * g * g)
            w -= lr_t * m / (torch.sqrt(v) + 1e-8)
        else:
            raise ValueError(self.optimizer)

    def _apply_sparse(self, name, rows, g):
        """[TF-1.14] SparseApply* on de-duplicated rows; untouched rows and slots stay as they are
        (Adam is the exception: its sparse apply decays m, v and updates every row)."""
        w, lr = self.params[name], self.lr
        rows_t = torch.as_tensor(rows, dtype=torch.long)
        g = g.reshape((len(rows),) + tuple(w.shape[1:])).to(w.dtype)
        if self.optimizer == "AdagradOptimizer":
            acc = self._slot(name, "accumulator", 1e-8)
            a = acc[rows_t] + g * g
            acc[rows_t] = a
            w[rows_t] = w[rows_t] - lr * g * torch.rsqrt(a)
        elif self.optimizer == "GradientDescentOptimizer":
            w[rows_t] = w[rows_t] - lr * g
        elif self.optimizer == "MomentumOptimizer":
            acc = self._slot(name, "momentum", 0.0)
            a = acc[rows_t] * 0.95 + g
            acc[rows_t] = a
            w[rows_t] = w[rows_t] - lr * a
        elif self.optimizer == "AdamOptimizer":
            full = torch.zeros_like(w)
            full[rows_t] = g
            self._apply_dense(name, full)
        else:
            raise ValueError(self.optimizer)

    # ---------------------------------------------------------------- weight exchange
    def export_params(self):
        return OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in self.params.items())

    def accumulators(self):
        """Adagrad slots keyed like the parameters (1e-8 where a slot was never created)."""
        out = OrderedDict()
        for k, v in self.params.items():
            s = self.state.get("accumulator", {}).get(k)
            out[k] = (torch.full_like(v, 1e-8) if s is None else s).numpy().copy()
        return out


class _GatherProxy:
    """Stands in for a table inside ``forward`` so that ``table[ids]`` returns the leaf holding the
    gathered rows (the IndexedSlices ``values`` TF differentiates)."""

    def __init__(self, table, ids, rows):
        self.table, self.ids, self.rows = table, ids, rows

    def __getitem__(self, ids):
        assert ids is self.ids or torch.equal(ids, self.ids)
        return self.rows

    def __pow__(self, e):  # only reached on the lamda>0 path, which uses dense leaves instead
        raise RuntimeError("regulariser needs the dense table")


# --------------------------------------------------------------------------- host loop pieces
def get_ordered_block(data, batch_size, index):
    """CFFM.py:617-629."""
    start = index * batch_size
    X, Y = [], []
    i = start
    while len(X) < batch_size and i < len(data["X"]):
        if len(data["X"][i]) == len(data["X"][start]):
            Y.append(data["Y"][i])
            X.append(data["X"][i])
            i += 1
        else:
            break
    return {"X": X, "Y": Y}


def get_random_block(data, batch_size, start_index):
    """CFFM.py:560-581 with the (unseeded) ``np.random.randint`` draw passed in (Q12)."""
    X, Y = [], []
    i = start_index
    while len(X) < batch_size and i < len(data["X"]):
        if len(data["X"][i]) == len(data["X"][start_index]):
            Y.append([data["Y"][i]])
            X.append(data["X"][i])
            i += 1
        else:
            break
    i = start_index
    while len(X) < batch_size and i >= 0:
        if len(data["X"][i]) == len(data["X"][start_index]):
            Y.append([data["Y"][i]])
            X.append(data["X"][i])
            i -= 1
        else:
            break
    return {"X": X, "Y": Y}


def evaluate(model, data, batch_size):
    """CFFM.py:583-615: ordered blocks -> predictions -> clip to [min y, max y] -> RMSE, R2
    (sklearn mean_squared_error / r2_score restated as their plain definitions)."""
    n = len(data["Y"])
    preds = []
    idx = 0
    blk = get_ordered_block(data, batch_size, idx)
    while len(blk["X"]) > 0:
        with torch.no_grad():
            preds.append(model.predict(np.asarray(blk["X"])).reshape(-1).double().numpy())
        idx += 1
        blk = get_ordered_block(data, batch_size, idx)
    y_pred = np.concatenate(preds)
    y_true = np.reshape(np.asarray(data["Y"], dtype=np.float64), (n,))
    pb = np.minimum(np.maximum(y_pred, y_true.min()), y_true.max())  # :609-611
    rmse = math.sqrt(float(np.mean((y_true - pb) ** 2)))  # :612
    ss_res = float(np.sum((y_true - pb) ** 2))
    ss_tot = float(np.sum((y_true - y_true.mean()) ** 2))
    r2 = 1.0 - ss_res / ss_tot if ss_tot > 0 else 0.0  # :614
    return rmse, r2


def eva_termination(valid):
    """CFFM.py:631-635."""
    if len(valid) > 5:
        if valid[-1] > valid[-2] > valid[-3] > valid[-4] > valid[-5]:
            return True
    return False
