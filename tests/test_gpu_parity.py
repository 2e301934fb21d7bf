"""Parity of the CUDA path (through the C ABI) against the oracle -- the `-m gpu` gate.

Tolerances: index work bit-exact; forward logits 1e-4 relative in fp32 (north star); gradients
1e-3 relative to the tensor's scale; weights after Adagrad steps: 2e-3 absolute with a small
outlier allowance (acc0 = 1e-8 makes the first update ~lr*sign(g), so a gradient that is ~0 can
legitimately flip a +-0.05 step between fp32 and fp64).
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CASES = {
    # name: F, K, M, B, activation, loss, linear_att, inner, outer
    "f4k8_selu": (4, 8, 50, 6, "selu", "square_loss", 1, 1, 1),
    "f3k32_elu": (3, 32, 200, 9, "elu", "square_loss", 1, 1, 1),
    "f10k32_selu": (10, 32, 300, 8, "selu", "square_loss", 1, 1, 1),
    "f6k16_relu_log": (6, 16, 120, 7, "relu", "log_loss", 1, 1, 1),
    "f5k8_gelu_mse": (5, 8, 64, 5, "gelu", "mse", 0, 1, 1),
    "f4k16_prelu_outer": (4, 16, 40, 6, "prelu", "mae", 1, 0, 1),
    "f7k8_elu_inner": (7, 8, 90, 10, "elu", "hybrid", 1, 1, 0),
    "f6k64_selu": (6, 64, 150, 5, "selu", "square_loss", 1, 1, 1),
    "f10k16_relu": (10, 16, 200, 8, "relu", "square_loss", 1, 1, 1),
    "f3k64_gelu_log": (3, 64, 80, 6, "gelu", "log_loss", 1, 1, 1),
}


def _engine(case, max_batch=None, **kw):
    from cffm_b200 import Engine
    F, K, M, B, act, loss, la, ic, oc = CASES[case]
    return Engine(M, F, K, K, activation=act, loss_type=loss, linear_att=la, inner_conv=ic, outer_conv=oc,
                  max_batch=max_batch or B, **kw)


def _load_golden(eng, g, case, tag="w0"):
    for name in eng.param_infos():
        eng.set_param(name, g["%s/%s/%s" % (case, tag, name)])


def _relerr(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def test_param_registry_matches_oracle(golden_model):
    from oracle.cffm_ref import CFFMRef
    for case, (F, K, M, B, act, loss, la, ic, oc) in CASES.items():
        eng = _engine(case)
        ref = CFFMRef(M, F, K, K, activation=act, loss_type=loss, linear_att=la, inner_conv=ic, outer_conv=oc)
        infos = eng.param_infos()
        assert list(infos) == list(ref.params), case
        for k, (shape, numel, trainable) in infos.items():
            assert tuple(ref.params[k].shape) == tuple(shape), (case, k)
            assert trainable == (k not in ref.dead_params()), (case, k)
        eng.close()


@pytest.mark.parametrize("case", list(CASES))
def test_forward_matches_golden(golden_model, case):
    g = golden_model
    eng = _engine(case)
    _load_golden(eng, g, case)
    ids = g[case + "/ids"]
    pred = eng.forward(ids)
    out = eng.fetch("out")
    F, K, M, B, act, loss, la, ic, oc = CASES[case]
    if ic:
        assert _relerr(eng.fetch("final2"), g[case + "/final2"]) < 1e-4
    if oc:
        assert _relerr(eng.fetch("t1").reshape(B, -1), g[case + "/t1"]) < 1e-4
        assert _relerr(eng.fetch("final"), g[case + "/final"]) < 1e-4
    assert np.max(np.abs(eng.fetch("linear") - g[case + "/linear"].reshape(-1))) < 1e-5
    assert _relerr(out, g[case + "/out"]) < 1e-4
    assert _relerr(pred, g[case + "/pred"]) < 1e-4
    eng.close()


@pytest.mark.parametrize("case", list(CASES))
def test_gradients_match_golden(golden_model, case):
    g = golden_model
    eng = _engine(case)
    _load_golden(eng, g, case)
    ids, y = g[case + "/ids"], g[case + "/y"]
    F, K, M, B, act, loss, la, ic, oc = CASES[case]
    loss_v = eng.train_step(ids, y)
    assert abs(loss_v - float(g[case + "/loss"])) < 1e-4 * max(1.0, abs(float(g[case + "/loss"])))
    # index work is bit exact: sorted ids and the unique-row list
    sorted_ids = eng.fetch("sorted_ids").astype(np.int64)
    assert np.array_equal(sorted_ids, np.sort(ids.reshape(-1), kind="stable"))
    n_uniq = int(eng.fetch("n_uniq")[0])
    seg = eng.fetch("seg_start").astype(np.int64)[:n_uniq]
    assert np.array_equal(sorted_ids[seg], g[case + "/uniq"])
    # per-sample gradient rows (the IndexedSlices values)
    if ic:
        assert _relerr(eng.fetch("grad_inner_rows"), g[case + "/grows/inner_embeddings"]) < 1e-3
    if oc:
        assert _relerr(eng.fetch("grad_outer_rows"), g[case + "/grows/outer_embeddings"]) < 1e-3
    assert _relerr(eng.fetch("grad_bias_rows"), g[case + "/grows/feature_bias"]) < 1e-3
    # dense gradients
    for name, (shape, numel, trainable) in eng.param_infos().items():
        key = "%s/g/%s" % (case, name)
        if key in g.files:
            got = eng.dense_grad(name)
            assert _relerr(got, g[key]) < 1e-3, (case, name, _relerr(got, g[key]))
        elif name not in ("inner_embeddings", "outer_embeddings", "feature_bias"):
            assert not trainable, name
            assert np.all(eng.dense_grad(name) == 0), name  # dead variables never receive a gradient
    eng.close()


@pytest.mark.parametrize("case", list(CASES))
def test_two_steps_match_golden(golden_model, case):
    g = golden_model
    eng = _engine(case)
    _load_golden(eng, g, case)
    ids, y = g[case + "/ids"], g[case + "/y"]
    losses = [eng.train_step(ids, y), eng.train_step(ids, y)]
    want = g[case + "/losses"]
    assert abs(losses[0] - want[0]) < 1e-4 * max(1.0, abs(want[0]))
    assert abs(losses[1] - want[1]) < 2e-2 * max(1.0, abs(want[1]))
    touched = np.unique(ids)
    for name, (shape, numel, trainable) in eng.param_infos().items():
        w2, a2 = eng.get_param(name), eng.get_param(name, accum=True)
        gw, ga = g["%s/w2/%s" % (case, name)], g["%s/a2/%s" % (case, name)]
        if not trainable:
            assert np.array_equal(w2, g["%s/w0/%s" % (case, name)].astype(np.float32).reshape(w2.shape)), name
            assert np.all(a2 == np.float32(1e-8)), name
            continue
        if name in ("inner_embeddings", "outer_embeddings", "feature_bias"):
            mask = np.ones(shape[0], dtype=bool)
            mask[touched] = False
            w0 = g["%s/w0/%s" % (case, name)].astype(np.float32).reshape(w2.shape)
            assert np.array_equal(w2[mask], w0[mask]), name       # untouched rows: bit exact
            assert np.all(a2[mask] == np.float32(1e-8)), name
        diff = np.abs(w2.astype(np.float64) - gw.reshape(w2.shape))
        assert np.mean(diff > 2e-3) < 0.02, (case, name, float(diff.max()), float(np.mean(diff > 2e-3)))
        assert _relerr(a2, ga) < 2e-2, (case, name, _relerr(a2, ga))
    eng.close()


def _frappe_batch(B=256):
    from cffm_b200 import LoadData
    d = LoadData(os.path.join(GOLDEN, "frappe_mini") + "/", "frappe", "square_loss")
    return d.features_M, np.array(d.Train_data["X"][:B]), np.array(d.Train_data["Y"][:B])


def _oracle_like(eng, M, F, K, act, loss="square_loss", dtype=torch.float32, **kw):
    from oracle.cffm_ref import CFFMRef
    ref = CFFMRef(M, F, K, K, activation=act, loss_type=loss, dtype=dtype, **kw)
    for k, v in eng.get_weights().items():
        ref.params[k] = torch.from_numpy(v.astype(np.float64)).to(dtype).reshape(ref.params[k].shape)
    return ref


def test_frappe_fixture_step_vs_live_oracle():
    """PR1 config (BASELINE.json configs[0]) on the reference's own libfm rows: F=10, K=32, B=256,
    selu, default initialisers (device RNG), compared with the fp64 oracle fed the same weights."""
    from cffm_b200 import Engine
    M, ids, y = _frappe_batch(256)
    eng = Engine(M, 10, 32, 32, activation="selu", max_batch=256, seed=5)
    fb = np.random.default_rng(0).normal(0, 0.05, (M, 1)).astype(np.float32)
    eng.set_param("feature_bias", fb)  # zero in the reference: give the linear term something to do
    ref = _oracle_like(eng, M, 10, 32, "selu", dtype=torch.float64)
    out = eng.forward(ids)
    want = ref.predict(ids).reshape(-1).numpy()
    assert _relerr(out, want) < 1e-4
    l_ref, dense, sparse = ref.gradients(ids, y)
    loss = eng.train_step(ids, y)
    assert abs(loss - float(l_ref)) < 1e-4 * max(1.0, float(l_ref))
    assert _relerr(eng.fetch("grad_inner_rows"), sparse["inner_embeddings"][2].numpy()) < 1e-3
    assert _relerr(eng.fetch("grad_outer_rows"), sparse["outer_embeddings"][2].numpy()) < 1e-3
    assert _relerr(eng.fetch("grad_bias_rows"), sparse["feature_bias"][2].numpy()) < 1e-3
    for name, gk in dense.items():
        if gk is not None:
            assert _relerr(eng.dense_grad(name), gk.numpy()) < 1e-3, name
    n_uniq = int(eng.fetch("n_uniq")[0])
    assert n_uniq == len(np.unique(ids))
    eng.close()


def test_initialisers_follow_q7():
    from cffm_b200 import Engine
    eng = Engine(20000, 6, 32, 32, max_batch=8, seed=1)
    w = eng.get_weights()
    assert abs(w["inner_embeddings"].std() - 0.1) < 2e-3 and abs(w["inner_embeddings"].mean()) < 1e-3
    assert abs(w["outer_embeddings"].std() - 0.01) < 2e-4
    assert np.all(w["feature_bias"] == 0) and w["bias"].reshape(-1)[0] == 0
    cw = w["outer_layer_conv_weight_0"]
    assert np.abs(cw).max() <= 2.0 and abs(cw.std() - 0.8796) < 0.05  # truncated normal, stddev 1
    assert np.all(w["outer_layer_conv_bias_0"] == np.float32(0.01))
    lim = np.sqrt(6.0 / (62 + 32))
    assert np.abs(w["dense_1/kernel"]).max() <= lim + 1e-6 and np.abs(w["dense_1/kernel"]).max() > 0.8 * lim
    assert np.all(w["dense_1/bias"] == 0)
    acc = eng.get_weights(accum=True)
    assert all(np.all(a == np.float32(1e-8)) for a in acc.values())
    eng2 = Engine(20000, 6, 32, 32, max_batch=8, seed=1)
    assert all(np.array_equal(w[k], v) for k, v in eng2.get_weights().items())  # seeded: reproducible
    eng.close(); eng2.close()


def test_gather_is_bit_exact():
    from cffm_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for K in (4, 16, 32, 64):
        table = torch.from_numpy(rng.standard_normal((5000, K)).astype(np.float32)).cuda()
        ids = torch.from_numpy(rng.integers(0, 5000, 12345).astype(np.int32)).cuda()
        out = torch.empty(12345, K, dtype=torch.float32, device="cuda")
        rc = lib.cffm_op_gather_dev(C.c_void_p(table.data_ptr()), C.c_void_p(ids.data_ptr()), 12345, K,
                                    C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(out, table[ids.long()])


@pytest.mark.parametrize("n", [20000, 60000])   # short list: one warp per segment; long list: chunked reduction
def test_sparse_adagrad_operator(n):
    """Sort-by-row + segmented sum + SparseApplyAdagrad: unique rows bit exact, untouched rows
    bitwise unchanged, touched rows equal to the fp32 restatement of the kernel's (deterministic) summation order."""
    from cffm_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    rng = np.random.default_rng(1)
    M, K, lr = 3000, 32, 0.05
    tab = rng.standard_normal((M, K)).astype(np.float32)
    acc = np.full((M, K), 1e-8, dtype=np.float32)
    ids = np.minimum((rng.pareto(1.0, n) * 3).astype(np.int64), M - 1).astype(np.int32)  # heavy duplication
    grads = (rng.standard_normal((n, K)) * 0.1).astype(np.float32)
    t_d, a_d = torch.from_numpy(tab).cuda(), torch.from_numpy(acc).cuda()
    i_d, g_d = torch.from_numpy(ids).cuda(), torch.from_numpy(grads).cuda()
    u_d = torch.zeros(n, dtype=torch.int32, device="cuda")
    nu_d = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rc = lib.cffm_op_sparse_adagrad_dev(C.c_void_p(t_d.data_ptr()), C.c_void_p(a_d.data_ptr()), M, K,
                                        C.c_void_p(i_d.data_ptr()), C.c_void_p(g_d.data_ptr()), n, lr,
                                        C.c_void_p(u_d.data_ptr()), C.c_void_p(nu_d.data_ptr()), C.c_void_p(0))
    assert rc == 0
    torch.cuda.synchronize()
    uniq = np.unique(ids)
    assert int(nu_d.item()) == len(uniq)
    assert np.array_equal(u_d[: len(uniq)].cpu().numpy(), uniq)
    # the kernel's summation order, restated.  Up to 32768 entries: a row's gradient rows in order of appearance (rows that
    # appear more than 64 times: eight consecutive pieces, each in order of appearance, then the pieces in order).
    # Longer lists: the stably sorted list is cut into chunks of 32 entries; inside a chunk a row's gradient rows
    # are added in order of appearance, then the chunk pieces in chunk order (fp32)
    order = np.argsort(ids, kind="stable")
    G = np.zeros((M, K), dtype=np.float32)
    if n > 32768:
        pieces = {}
        for t, src in enumerate(order):
            key = (int(ids[src]), t // 32)
            pieces[key] = pieces.get(key, np.zeros(K, dtype=np.float32)) + grads[src]
        for (row, chunk) in sorted(pieces):
            G[row] = G[row] + pieces[(row, chunk)]
    else:
        sorted_ids = ids[order]
        bounds = np.flatnonzero(np.r_[True, sorted_ids[1:] != sorted_ids[:-1], True])
        for a0, a1 in zip(bounds[:-1], bounds[1:]):
            src = order[a0:a1]
            row = int(sorted_ids[a0])
            if len(src) <= 64:
                for p in src:
                    G[row] = G[row] + grads[p]
            else:
                piece = (((len(src) + 7) >> 3) + 7) & ~7
                tot = np.zeros(K, dtype=np.float32)
                for w in range(8):
                    acc_p = np.zeros(K, dtype=np.float32)
                    for p in src[w * piece:min(len(src), (w + 1) * piece)]:
                        acc_p = acc_p + grads[p]
                    tot = tot + acc_p
                G[row] = tot
    a_want = acc + G * G
    w_want = tab - lr * G / np.sqrt(a_want)
    got_w, got_a = t_d.cpu().numpy(), a_d.cpu().numpy()
    mask = np.ones(M, dtype=bool); mask[uniq] = False
    assert np.array_equal(got_w[mask], tab[mask]) and np.array_equal(got_a[mask], acc[mask])
    assert np.allclose(got_a[uniq], a_want[uniq], rtol=1e-6, atol=0)
    assert np.allclose(got_w[uniq], w_want[uniq], rtol=0, atol=1e-6)


@pytest.mark.parametrize("n,M,K", [(1, 10, 4), (31, 7, 32), (1025, 3, 32), (2560, 5382, 32), (4096, 100000, 64), (4097, 100000, 16)])
def test_sparse_update_short_lists(n, M, K):
    """Lists up to 4096 ids (the reference's own datasets) take the one-block sort + unique and the per-segment kernels;
    an id that fills a good part of the batch (Frappe has fields with two or three values) is summed by a whole block
    (eight pieces, in order).  Unique rows bit exact, untouched rows bitwise unchanged, touched rows equal to the fp32
    restatement of that summation order."""
    from cffm_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    rng = np.random.default_rng(n)
    lr = 0.05
    tab = rng.standard_normal((M, K)).astype(np.float32)
    acc = np.full((M, K), 1e-8, dtype=np.float32)
    ids = rng.integers(0, M, n).astype(np.int32)
    grads = (rng.standard_normal((n, K)) * 0.1).astype(np.float32)
    t_d, a_d = torch.from_numpy(tab).cuda(), torch.from_numpy(acc).cuda()
    i_d, g_d = torch.from_numpy(ids).cuda(), torch.from_numpy(grads).cuda()
    u_d = torch.zeros(n, dtype=torch.int32, device="cuda")
    nu_d = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rc = lib.cffm_op_sparse_adagrad_dev(C.c_void_p(t_d.data_ptr()), C.c_void_p(a_d.data_ptr()), M, K,
                                        C.c_void_p(i_d.data_ptr()), C.c_void_p(g_d.data_ptr()), n, lr,
                                        C.c_void_p(u_d.data_ptr()), C.c_void_p(nu_d.data_ptr()), C.c_void_p(0))
    assert rc == 0
    torch.cuda.synchronize()
    uniq = np.unique(ids)
    assert int(nu_d.item()) == len(uniq)
    assert np.array_equal(u_d[: len(uniq)].cpu().numpy(), uniq)
    order = np.argsort(ids, kind="stable")
    G = np.zeros((M, K), dtype=np.float32)
    for row in uniq:
        src = order[ids[order] == row]                      # order of appearance
        if len(src) <= 64 or n > 32768:
            for p in src:
                G[row] = G[row] + grads[p]
        else:                                               # eight pieces of ((len + 7) / 8 rounded up to 8) entries
            piece = (((len(src) + 7) >> 3) + 7) & ~7
            parts = []
            for w in range(8):
                acc_p = np.zeros(K, dtype=np.float32)
                for p in src[w * piece:min(len(src), (w + 1) * piece)]:      # order of appearance inside a piece
                    acc_p = acc_p + grads[p]
                parts.append(acc_p)
            tot = np.zeros(K, dtype=np.float32)
            for acc_p in parts:
                tot = tot + acc_p
            G[row] = tot
    a_want = acc + G * G
    w_want = tab - lr * G / np.sqrt(a_want)
    got_w, got_a = t_d.cpu().numpy(), a_d.cpu().numpy()
    mask = np.ones(M, dtype=bool); mask[uniq] = False
    assert np.array_equal(got_w[mask], tab[mask]) and np.array_equal(got_a[mask], acc[mask])
    assert np.allclose(got_a[uniq], a_want[uniq], rtol=2e-5, atol=1e-12)
    assert np.allclose(got_w[uniq], w_want[uniq], rtol=0, atol=2e-5)


def test_evaluate_matches_oracle():
    from cffm_b200 import Engine
    from oracle.cffm_ref import evaluate
    M, ids, y = _frappe_batch(700)
    eng = Engine(M, 10, 32, 32, activation="selu", max_batch=256, seed=9)
    ref = _oracle_like(eng, M, 10, 32, "selu", dtype=torch.float64)
    rmse, r2 = eng.evaluate(ids, y, 256)  # blocks of 256, 256, 188
    want = evaluate(ref, {"X": [list(r) for r in ids], "Y": list(y)}, 256)
    assert abs(rmse - want[0]) < 1e-4 and abs(r2 - want[1]) < 1e-3
    eng.close()


def test_pipelined_submit_equals_blocking_steps():
    from cffm_b200 import Engine
    M, ids, y = _frappe_batch(512)
    a = Engine(M, 10, 32, 32, activation="selu", max_batch=128, seed=2)
    b = Engine(M, 10, 32, 32, activation="selu", max_batch=128, seed=2)
    la, lb = [], []
    for s in range(4):
        la.append(a.train_step(ids[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128]))
        r = b.train_submit(ids[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128])
        if r is not None:
            lb.append(r)
    lb.append(b.train_flush())
    assert la == lb  # same kernels, same order: bit-identical losses
    wa, wb = a.get_weights(), b.get_weights()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
    a.close(); b.close()


def test_device_pointer_entry_points():
    from cffm_b200 import Engine
    M, ids, y = _frappe_batch(64)
    a = Engine(M, 10, 32, 32, activation="selu", max_batch=64, seed=4)
    b = Engine(M, 10, 32, 32, activation="selu", max_batch=64, seed=4)
    ids_d, y_d = torch.from_numpy(ids).cuda(), torch.from_numpy(y).cuda()
    out_d = torch.empty(64, device="cuda")
    loss_d = torch.empty(1, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    b.forward_dev(ids_d.data_ptr(), 64, out_d.data_ptr(), s)
    torch.cuda.synchronize()
    assert np.array_equal(out_d.cpu().numpy(), a.forward(ids))
    b.train_step_dev(ids_d.data_ptr(), y_d.data_ptr(), 64, loss_d.data_ptr(), s)
    torch.cuda.synchronize()
    assert float(loss_d.item()) == a.train_step(ids, y)
    a.close(); b.close()


def test_errors_are_reported_not_thrown():
    from cffm_b200 import Engine, CffmError
    with pytest.raises(CffmError):
        Engine(100, 3, 24, 32, max_batch=4)  # not a power of two
    with pytest.raises(CffmError):
        Engine(100, 3, 32, 32, max_batch=4, loss_type="crossentropy")  # undefined in the reference (Q10)
    eng = Engine(100, 3, 8, 8, max_batch=4)
    with pytest.raises(CffmError):
        eng.forward(np.zeros((4, 5), dtype=np.int32))
    with pytest.raises(CffmError):
        eng.train_step(np.zeros((8, 3), dtype=np.int32), np.zeros(8))  # B > max_batch
    with pytest.raises(CffmError):
        eng.get_param("no_such_variable")
    eng.close()


def test_resident_dataset_equals_host_batches():
    """cffm_dataset_* (split resident in HBM, device shuffle, in-place blocks) == the host-buffer loop."""
    from cffm_b200 import Engine
    M, ids, y = _frappe_batch(1024)
    a = Engine(M, 10, 32, 32, activation="selu", max_batch=128, seed=6)
    b = Engine(M, 10, 32, 32, activation="selu", max_batch=128, seed=6)
    perm = np.random.RandomState(2021).permutation(1024)
    b.dataset_upload(ids, y)
    b.dataset_permute(perm)
    ids_p, y_p = ids[perm], y[perm]
    starts = [5, 700, 896, 333]
    for st in starts:
        la = a.train_step(ids_p[st:st + 128], y_p[st:st + 128])
        b.train_block(st, 128)
        assert b.last_loss() == la
    wa, wb = a.get_weights(), b.get_weights()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
    ra, rb = a.evaluate(ids_p, y_p, 128), b.dataset_evaluate(128)
    assert ra == rb
    a.close(); b.close()


def test_cli_train_loop_runs_on_the_fixture(tmp_path, monkeypatch):
    """The reference command line (README) end to end on the Frappe fixture: a few epochs, RMSE improves."""
    import os
    from cffm_b200 import cli
    monkeypatch.chdir(tmp_path)
    model = cli.main(["--path", os.path.join(GOLDEN, "frappe_mini") + "/", "--dataset", "frappe", "--epoch", "12",
                      "--batch_size", "256", "--inner_dims", "32", "--outer_dims", "32", "--lamda", "0", "--lr", "0.05",
                      "--loss_type", "square_loss", "--num_field", "10", "--linear_att", "1", "--inner_conv", "1",
                      "--outer_conv", "1", "--activation", "selu", "--verbose", "4"])
    assert len(model.valid_rmse) >= 6 and min(model.valid_rmse) < model.valid_rmse[0]
    assert min(model.valid_rmse) < 1.0
    assert len(model.train_rmse) == len(model.valid_rmse) == len(model.test_rmse) == len(model.valid_r2)


@pytest.mark.parametrize("F,K,N,act", [(3, 64, 70000, "elu"), (10, 16, 66000, "selu")])
def test_evaluate_at_65536_blocks_k16_k64(F, K, N, act):
    """Scoring sweep of BASELINE.json configs[4] (B up to 65 536, K = 16 / 64): evaluate() in blocks of 65 536
    (one full block + a partial one) against the oracle scoring the same rows in blocks of 4 096 -- the metric does
    not depend on the block size (CFFM.py:583-615)."""
    from cffm_b200 import Engine
    from oracle.cffm_ref import evaluate
    rng = np.random.default_rng(7)
    M = 5000
    ids = rng.integers(0, M, (N, F)).astype(np.int32)
    y = rng.choice([-1.0, 1.0], N).astype(np.float32)
    eng = Engine(M, F, K, K, activation=act, max_batch=65536, seed=9)
    eng.set_param("feature_bias", rng.normal(0, 0.3, (M, 1)).astype(np.float32))
    eng.set_param("outer_embeddings", rng.normal(0, 0.2, (M, K)).astype(np.float32))
    ref = _oracle_like(eng, M, F, K, act, dtype=torch.float32)
    rmse, r2 = eng.evaluate(ids, y, 65536)
    small = eng.evaluate(ids, y, 4096)
    assert abs(rmse - small[0]) < 1e-5 and abs(r2 - small[1]) < 1e-4          # block-size independence on the device
    want = evaluate(ref, {"X": ids, "Y": list(y)}, 4096)
    assert abs(rmse - want[0]) < 1e-4 * max(1.0, want[0]) and abs(r2 - want[1]) < 1e-3 * max(1.0, abs(want[1])), (rmse, r2, want)
    out = eng.forward(ids[:65536])
    ref_out = ref.predict(ids[:2048]).reshape(-1).numpy()
    assert _relerr(out[:2048], ref_out) < 1e-4
    eng.close()
