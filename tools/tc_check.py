#!/usr/bin/env python
"""Point-wise GEMM engines against an fp64 matmul on the GPU (development check; the pytest
version lives in tests/test_gpu_pw_gemm.py).  usage: tc_check.py [engine ...]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
pkg.build()
lib = pkg._lib.load()

SHAPES = [  # (M, K, N) : one tile, tails, every (K,N) class of the network
    (128, 32, 32), (128, 32, 64), (256, 64, 96), (300, 16, 96), (1000, 96, 24), (777, 24, 144), (640, 144, 24),
    (512, 144, 32), (512, 32, 192), (512, 192, 64), (512, 64, 384), (512, 384, 96), (400, 96, 576), (400, 576, 160),
    (400, 160, 960), (400, 960, 320), (400, 320, 24), (128 * 300 + 5, 16, 96), (32, 32, 16),
]


def run(engine, epi, A, W, res):
    M, K = A.shape
    N = W.shape[1]
    out = torch.full((M, N), float("nan"), device="cuda")
    Wh = np.ascontiguousarray(W.cpu().numpy())
    rc = lib.cf_debug_pw_gemm(engine, epi, C.c_void_p(A.data_ptr()), C.c_void_p(Wh.ctypes.data), C.c_void_p(out.data_ptr()),
                              M, K, N, C.c_void_p(res.data_ptr()) if res is not None else None,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(lib.cf_last_error().decode())
    torch.cuda.synchronize()
    return out


def main():
    engines = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
    g = torch.Generator(device="cuda").manual_seed(0)
    worst = {e: 0.0 for e in engines}
    for (M, K, N) in SHAPES:
        A = torch.randn(M, K, device="cuda", generator=g) * 3
        W = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
        res = torch.randn(M, N, device="cuda", generator=g)
        ref = A.double() @ W.double()
        scale = (A.double().abs() @ W.double().abs()).clamp_min(1e-30)  # condition-aware error bound
        for epi in (0, 1, 2):
            want = ref if epi == 0 else (ref * torch.sigmoid(ref) if epi == 1 else ref + res.double())
            for e in engines:
                got = run(e, epi, A, W, res if epi == 2 else None)
                bad = ~torch.isfinite(got)
                err = ((got.double() - want).abs() / scale).max().item() if not bad.any() else float("inf")
                worst[e] = max(worst[e], err)
                print(f"M={M:6d} K={K:4d} N={N:4d} epi={epi} engine={e}: max |err|/(|A||W|) = {err:.3e}"
                      + (f"  NON-FINITE x{int(bad.sum())}" if bad.any() else ""), flush=True)
    print("worst:", worst)


if __name__ == "__main__":
    main()
