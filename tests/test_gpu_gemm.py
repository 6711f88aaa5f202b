"""The two contraction cores behind the C ABI against torch on the same inputs (GPU only).
fp32 SIMT core: fp32 operands/accumulate -> tight tolerance.  tcgen05 core: bf16 operands, fp32 TMEM
accumulation -> compared with an fp32 matmul of the same bf16-rounded operands."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    import show_and_tell_b200 as snt
    return snt._lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


SHAPES = [(1, 8, 8), (5, 23, 16), (37, 301, 40), (128, 256, 64), (129, 257, 72), (300, 1000, 512),
          (1024, 2048, 512), (77, 64, 1000)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tA", [0, 1])
@pytest.mark.parametrize("tB", [0, 1])
def test_gemm_f32(M, N, K, tA, tB):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    C0 = torch.randn(M, N, device="cuda", generator=g)
    out = C0.clone()
    L.call("snt_gemm_f32", tA, tB, M, N, K, 0.5, _p(A), A.shape[1], _p(B), B.shape[1], 2.0, _p(out), N, _p(bias),
           L.stream_ptr())
    opA = A.t() if tA else A
    opB = B.t() if tB else B
    ref = 0.5 * (opA.double() @ opB.double()) + 2.0 * C0.double() + bias.double()
    err = (out.double() - ref).norm() / ref.norm()
    assert err < 2e-6, err


def _bf16_case(M, N, K, tA, tB, c_bf16, seed):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = lambda x: (x + 7) // 8 * 8
    # leading dimensions must be multiples of 8: allocate padded, use a view
    if tA:
        Af = torch.randn(K, pad(M), device="cuda", generator=g)
        A = Af.bfloat16(); lda = pad(M); opA = A[:, :M].float().t()
    else:
        Af = torch.randn(M, pad(K), device="cuda", generator=g)
        A = Af.bfloat16(); lda = pad(K); opA = A[:, :K].float()
    if tB:
        Bf = torch.randn(N, pad(K), device="cuda", generator=g)
        B = Bf.bfloat16(); ldb = pad(K); opB = B[:, :K].float().t()
    else:
        Bf = torch.randn(K, pad(N), device="cuda", generator=g)
        B = Bf.bfloat16(); ldb = pad(N); opB = B[:, :N].float()
    bias = torch.randn(N, device="cuda", generator=g)
    ldc = pad(N) if c_bf16 else N
    if c_bf16:
        out = torch.zeros(M, ldc, device="cuda", dtype=torch.bfloat16)
        beta = 0.0
        C0 = None
    else:
        C0 = torch.randn(M, N, device="cuda", generator=g)
        out = C0.clone()
        beta = 1.0
    L.call("snt_gemm_bf16", tA, tB, M, N, K, 1.0, _p(A), lda, _p(B), ldb, beta, _p(out), ldc, 1 if c_bf16 else 0,
           _p(bias), L.stream_ptr())
    torch.cuda.synchronize()
    ref = opA.double() @ opB.double() + bias.double()
    if C0 is not None:
        ref = ref + C0.double()
    got = out[:, :N].double()
    err = ((got - ref).norm() / ref.norm()).item()
    return err


@pytest.mark.parametrize("M,N,K", SHAPES + [(12851, 2048, 256), (2048, 512, 12851), (1024, 10000, 512)])
@pytest.mark.parametrize("tA", [0, 1])
@pytest.mark.parametrize("tB", [0, 1])
def test_gemm_bf16_tcgen05(M, N, K, tA, tB):
    err = _bf16_case(M, N, K, tA, tB, False, M + N + K + tA * 2 + tB)
    assert err < 5e-5, f"tcgen05 GEMM tA={tA} tB={tB} {M}x{N}x{K}: rel err {err}"


@pytest.mark.parametrize("M,N,K", [(129, 257, 72), (1024, 2048, 512)])
def test_gemm_bf16_out_bf16(M, N, K):
    err = _bf16_case(M, N, K, 0, 1, True, 5)
    assert err < 4e-3, err


def test_gemm_bf16_with_sm_reserve():
    """snt_set_sm_reserve shrinks the persistent grids (tiles are redistributed over fewer CTAs); results are unchanged
    bit for bit, and the previous value is returned."""
    L = _lib()
    base = _bf16_case(12851, 2048, 256, 0, 1, False, 11)
    assert L.lib().snt_set_sm_reserve(16) == 0
    try:
        assert _bf16_case(12851, 2048, 256, 0, 1, False, 11) == base
        assert _bf16_case(1024, 10000, 512, 1, 0, False, 12) < 5e-5
    finally:
        assert L.lib().snt_set_sm_reserve(0) == 16
