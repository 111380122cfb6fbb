"""The point-wise convolution engines against an fp64 matmul, shape by shape (every (K,N) class of
the network plus ragged M tails), through the C-ABI validation hook cf_debug_pw_gemm."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 32, 32), (300, 16, 96), (1000, 96, 24), (777, 24, 144), (640, 144, 24), (512, 144, 32), (512, 32, 192),
          (512, 192, 64), (512, 64, 384), (512, 384, 96), (400, 96, 576), (400, 576, 160), (400, 160, 960),
          (400, 960, 320), (400, 320, 24), (128 * 300 + 5, 16, 96), (32, 32, 16), (5, 24, 24)]
# max |err| / (|A|.|W|) allowed per engine: fp32 FFMA, 3xTF32 (hi/lo split), 1xTF32
BOUND = {0: 1e-6, 1: 5e-6, 2: 2e-3}


@pytest.mark.parametrize("engine", [0, 1, 2])
def test_pw_gemm_engines(pkg, engine):
    lib = pkg._lib.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, K, N) in SHAPES:
        A = torch.randn(M, K, device="cuda", generator=g) * 3
        W = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
        res = torch.randn(M, N, device="cuda", generator=g)
        ref = A.double() @ W.double()
        scale = (A.double().abs() @ W.double().abs()).clamp_min(1e-30)
        Wh = np.ascontiguousarray(W.cpu().numpy())
        for epi in (0, 1, 2):  # linear, Swish, + residual
            want = ref if epi == 0 else (ref * torch.sigmoid(ref) if epi == 1 else ref + res.double())
            out = torch.full((M, N), float("nan"), device="cuda")
            rc = lib.cf_debug_pw_gemm(engine, epi, C.c_void_p(A.data_ptr()), C.c_void_p(Wh.ctypes.data),
                                      C.c_void_p(out.data_ptr()), M, K, N, C.c_void_p(res.data_ptr()) if epi == 2 else None,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, lib.cf_last_error()
            assert torch.isfinite(out).all(), (engine, M, K, N, epi)
            err = ((out.double() - want).abs() / scale).max().item()
            assert err < BOUND[engine], (engine, M, K, N, epi, err)


def test_pw_gemm_huge_activations(pkg):
    """The BN-free back-bone reaches |x| ~ 1e13: tf32 keeps fp32's exponent range."""
    lib = pkg._lib.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    M, K, N = 400, 960, 320
    A = torch.randn(M, K, device="cuda", generator=g) * 1e13
    W = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
    ref = A.double() @ W.double()
    scale = A.double().abs() @ W.double().abs()
    Wh = np.ascontiguousarray(W.cpu().numpy())
    out = torch.empty((M, N), device="cuda")
    assert lib.cf_debug_pw_gemm(1, 0, C.c_void_p(A.data_ptr()), C.c_void_p(Wh.ctypes.data), C.c_void_p(out.data_ptr()), M, K, N,
                                None, C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    assert torch.isfinite(out).all()
    assert ((out.double() - ref).abs() / scale).max().item() < 5e-6
