#!/usr/bin/env python
"""Run single point-wise GEMM shapes through cf_debug_pw_gemm (for `ncu --metrics gpu__time_duration.sum`)."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
lib = pkg._lib.load()
SHAPES = [(3276800, 32, 16), (819200, 144, 24), (819200, 24, 144), (204800, 192, 32), (51200, 384, 96), (12800, 960, 160)]
engine = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for (M, K, N) in SHAPES:
    A = torch.randn(M, K, device="cuda"); W = np.random.randn(K, N).astype(np.float32)
    out = torch.empty(M, N, device="cuda")
    for _ in range(2):
        rc = lib.cf_debug_pw_gemm(engine, 0, C.c_void_p(A.data_ptr()), C.c_void_p(W.ctypes.data), C.c_void_p(out.data_ptr()), M, K, N, None,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.cf_last_error()
    print(M, K, N, "ok", flush=True)
