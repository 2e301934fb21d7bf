"""A second, independent pin for the model oracle (TensorFlow 1.14 cannot run here and the
reference ships no golden vectors, so ``oracle/cffm_ref.py`` cannot be checked against the reference itself).

``oracle/cffm_ref.py`` restates the graph in einsum / reshape form.  This file restates the SAME lines of
``CFFM.py`` a second time, op for op, through *library* operators whose semantics match the TensorFlow ops the
reference calls -- so a shared misreading of ``CFFM.py:323-333`` / ``:373-396`` (tap order, HWIO layout, pooling
axes, stride placement) in the einsum form would show up as a disagreement:

  tf.nn.conv2d(NHWC, HWIO, strides, 'VALID')  ->  torch.nn.functional.conv2d on NCHW / OIHW (cross-correlation, no flip)
  tf.nn.max_pool(NHWC, ksize, strides)        ->  torch.nn.functional.max_pool2d
  tf.stack / tf.transpose / tf.reduce_sum     ->  torch.stack / permute / sum with the SAME axis numbers as the reference
  tf.layers.dense                              ->  x @ kernel + bias

Both restatements are compared in fp64: forward values, and every gradient through autograd.  CPU only.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

from oracle.cffm_ref import CFFMRef, act_fn


def tf_conv2d(x_nhwc, w_hwio, strides, padding="VALID"):
    """tf.nn.conv2d [TF-1.14]: NHWC input, HWIO filter, strides = [1, sh, sw, 1], cross-correlation."""
    assert padding == "VALID" and strides[0] == 1 and strides[3] == 1
    y = Fn.conv2d(x_nhwc.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), stride=(strides[1], strides[2]))
    return y.permute(0, 2, 3, 1)


def tf_max_pool(x_nhwc, ksize, strides):
    """tf.nn.max_pool [TF-1.14], VALID."""
    y = Fn.max_pool2d(x_nhwc.permute(0, 3, 1, 2), kernel_size=(ksize[1], ksize[2]), stride=(strides[1], strides[2]))
    return y.permute(0, 2, 3, 1)


def conv_layer(x, W, b, strides):  # CFFM.py:475-478
    return torch.relu(tf_conv2d(x, W, strides) + b)


def library_forward(ref: CFFMRef, p, ids):
    """CFFM.py:296-453 once more, with the reference's own op sequence and axis numbers."""
    act = act_fn(ref.activation_name)
    ids_t = torch.as_tensor(np.asarray(ids), dtype=torch.long)
    F, P = ref.F, ref.P
    comps = []
    if ref.inner_conv == 1:
        emb = p["inner_embeddings"][ids_t]                                   # :303
        inner = []
        for i in range(0, F):                                                # :304-310
            for j in range(i + 1, F):
                inner.append(emb[:, i, :] * emb[:, j, :])
        x = torch.stack(inner)                                               # :313  interaction * None * K
        x = x.permute(1, 0, 2)                                               # :315  perm=[1,0,2]
        x = x.unsqueeze(-1)                                                  # :317
        x = act(x)                                                           # :319
        c1 = conv_layer(x, p["inner_layer_conv_weight_0"], p["inner_layer_conv_bias_0"], [1, 1, 2, 1])  # :327
        c1 = act(c1)                                                         # :330
        mp = tf_max_pool(x, [1, 1, 2, 1], [1, 1, 2, 1])                      # :331
        c1 = c1 + mp                                                         # :332
        flat = c1.reshape(-1, P * (ref.Ki // 2) * 2)                         # :333 (literal 16*2 generalised, Q5)
        comps.append(flat @ p[ref.dense_inner + "/kernel"] + p[ref.dense_inner + "/bias"])   # :339
    if ref.outer_conv == 1:
        oe = p["outer_embeddings"][ids_t]                                    # :354
        outer = []
        for i in range(0, F):                                                # :355-362
            for j in range(i + 1, F):
                fi = oe[:, i, :].unsqueeze(-1)
                fj = oe[:, j, :].unsqueeze(-1).permute(0, 2, 1)
                outer.append(fi * fj)
        x = torch.stack(outer).permute(1, 2, 3, 0)                           # :365-367 perm=[1,2,3,0]
        depth = int(np.log2(ref.Ko))                                         # :373
        pools = [x.sum(dim=(2, 3))]                                          # :381 axis=[2,3]
        for l in range(depth):                                               # :384-391
            x = conv_layer(x, p["outer_layer_conv_weight_%d" % l], p["outer_layer_conv_bias_%d" % l], [1, 2, 2, 1])
            x = act(x)
            pools.append(x.sum(dim=(2, 3)))
        t1 = pools[0]
        for i in range(1, depth):                                            # :394-396
            t1 = torch.cat([t1, pools[i]], dim=1)
        h = t1 @ p[ref.dense_outer1 + "/kernel"] + p[ref.dense_outer1 + "/bias"]      # :409
        fin = h @ p[ref.dense_outer2 + "/kernel"] + p[ref.dense_outer2 + "/bias"]     # :410
        comps.append((ref.beta_outer * fin).reshape(-1, 1))                  # :414
    fb = p["feature_bias"][ids_t]                                            # :422
    if ref.linear_att == 1:
        fb = fb.squeeze(-1)                                                  # :425
        lin = fb @ p["bias_W"] + p["bias_b"]                                 # :432
        lin = torch.softmax(lin / ref.lamda_att, dim=-1)                     # :434-436
        lin = (fb * lin).reshape(-1, F)                                      # :438
        lin = lin @ p[ref.dense_linear + "/kernel"] + p[ref.dense_linear + "/bias"]   # :441
    else:
        lin = fb.sum(dim=1)                                                  # :444
    comps.append(lin)
    comps.append(p["bias"] * torch.ones(ids_t.shape[0], 1, dtype=torch.float64))     # :449
    return sum(comps[1:], comps[0])                                          # :453 add_n


CASES = [
    # F, K, activation, linear_att, inner, outer
    (4, 8, "selu", 1, 1, 1),
    (5, 16, "relu", 1, 1, 1),
    (3, 32, "elu", 0, 1, 1),
    (6, 32, "gelu", 1, 1, 1),
    (4, 64, "prelu", 1, 1, 1),
    (10, 32, "selu", 1, 1, 1),
]


@pytest.mark.parametrize("F,K,act,la,ic,oc", CASES)
def test_einsum_oracle_equals_library_op_restatement(F, K, act, la, ic, oc):
    M, B = 60, 5
    ref = CFFMRef(M, F, K, K, activation=act, linear_att=la, inner_conv=ic, outer_conv=oc, dtype=torch.float64, seed=3)
    g = torch.Generator().manual_seed(4)
    ref.params["feature_bias"] = torch.randn(M, 1, generator=g, dtype=torch.float64) * 0.3
    ref.params["outer_embeddings"] = torch.randn(M, K, generator=g, dtype=torch.float64) * 0.3
    ids = np.random.default_rng(5).integers(0, M, (B, F))
    # ---- forward ----
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in ref.params.items()}
    out_lib = library_forward(ref, leaves, ids)
    leaves2 = {k: v.detach().clone().requires_grad_(True) for k, v in ref.params.items()}
    out_ein = ref.forward(ids, leaves2)
    assert out_lib.shape == out_ein.shape == (B, 1)
    assert torch.allclose(out_lib, out_ein, rtol=1e-11, atol=1e-11), float((out_lib - out_ein).abs().max())
    # ---- gradients of a generic scalar of the output, through both graphs ----
    wvec = torch.randn(B, 1, generator=g, dtype=torch.float64)
    gl = torch.autograd.grad((out_lib * wvec).sum(), list(leaves.values()), allow_unused=True)
    ge = torch.autograd.grad((out_ein * wvec).sum(), list(leaves2.values()), allow_unused=True)
    for k, a, b in zip(leaves, gl, ge):
        dead = k in ref.dead_params()
        if a is None or b is None:
            # the dead last conv layer and outer_W / outer_b are unreachable in BOTH graphs (Q2, Q14) ...
            # except that the library form, like TF, builds the dead conv layer (no gradient flows from `out`)
            assert dead or k in ("outer_W", "outer_b"), k
            continue
        if dead:
            assert float(a.abs().max()) == 0.0 or float(b.abs().max()) == 0.0 or torch.allclose(a, b)
            continue
        scale = max(1e-30, float(b.abs().max()))
        assert float((a - b).abs().max()) / scale < 1e-9, (k, float((a - b).abs().max()), scale)


def test_pooling_axes_and_dead_layer_shapes():
    """SURVEY Q1 / Q2 through the library ops: reduce_sum(axis=[2,3]) of NHWC keeps H; t1 is [B, 2K-2]."""
    F, K, M, B = 4, 16, 30, 3
    ref = CFFMRef(M, F, K, K, activation="relu", dtype=torch.float64, seed=1)
    x = torch.randn(B, K, K, ref.P, dtype=torch.float64)
    assert x.sum(dim=(2, 3)).shape == (B, K)
    y = conv_layer(x, ref.params["outer_layer_conv_weight_0"], ref.params["outer_layer_conv_bias_0"], [1, 2, 2, 1])
    assert y.shape == (B, K // 2, K // 2, ref.P)
    _, inter = ref.forward(np.zeros((B, F), dtype=np.int64), return_intermediates=True)
    assert inter["t1"].shape == (B, 2 * K - 2)


def test_max_pool_tie_goes_to_first_tap():
    """Q15: with relu both taps are often 0; TF routes the gradient to the first maximum, and so do
    torch's max_pool2d and the oracle's torch.where(a0 >= a1, ...)."""
    x = torch.zeros(1, 1, 4, 1, dtype=torch.float64, requires_grad=True)
    tf_max_pool(x, [1, 1, 2, 1], [1, 1, 2, 1]).sum().backward()
    assert x.grad.reshape(-1).tolist() == [1.0, 0.0, 1.0, 0.0]
