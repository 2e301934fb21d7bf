"""tcgen05 pipeline checks (bf16 operands, fp32 accumulation in TMEM) against torch on the same
bf16-rounded inputs: the only differences are fp32 summation order."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(lib, a, b):
    M, K = a.shape
    N = b.shape[0]
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = lib.cffm_op_gemm_bf16_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), M, N, K,
                                   C.c_void_p(0))
    assert rc == 0, lib.cffm_tc_last_error()
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 256), (256, 48, 192), (1000, 768, 3072), (4096, 512, 320),
                                   (77, 16, 64), (20000, 256, 128)])
def test_bf16_gemm_matches_torch(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
    c = _gemm(lib, a, b)
    want = a.float() @ b.float().t()
    err = (c - want).abs().max().item()
    assert not torch.isnan(c).any(), "unwritten outputs"
    assert err < 1e-3 * max(1.0, want.abs().max().item()), err


def test_bf16_gemm_identity_layout(lib):
    """A = one-hot rows exposes any swizzle / descriptor mix-up exactly."""
    K, N, M = 256, 128, 256
    a = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    idx = torch.arange(M, device="cuda") % K
    a[torch.arange(M, device="cuda"), idx] = 1
    b = (torch.arange(N * K, device="cuda", dtype=torch.float32).reshape(N, K) % 251 - 125).to(torch.bfloat16)
    c = _gemm(lib, a, b)
    want = b.float().t()[idx]
    assert torch.equal(c, want)


@pytest.mark.parametrize("M,N,R", [(128, 64, 64), (256, 256, 512), (3072, 768, 4096), (192, 64, 1000), (384, 128, 77)])
def test_bf16_gemm_tn_matches_torch(lib, M, N, R):
    """MN-major operands (the weight-gradient form): C = A^T B with A [R,M], B [R,N]."""
    g = torch.Generator(device="cuda").manual_seed(M + N + R)
    a = torch.randn(R, M, generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randn(R, N, generator=g, device="cuda").to(torch.bfloat16)
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = lib.cffm_op_gemm_bf16_tn_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), M, N, R,
                                      C.c_void_p(0))
    assert rc == 0, lib.cffm_tc_last_error()
    torch.cuda.synchronize()
    want = a.float().t() @ b.float()
    assert not torch.isnan(c).any()
    assert (c - want).abs().max().item() < 1e-3 * max(1.0, want.abs().max().item())
