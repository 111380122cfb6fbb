#!/usr/bin/env python
"""How does tcgen05 kind::tf32 accumulate?  Inputs are pre-rounded to tf32 so every product is exact in
fp32; whatever error remains is the accumulator's.  Positive data makes a rounding bias visible."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
lib = pkg._lib.load()

def tf32(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

def gemm(engine, A, W):
    M, K = A.shape; N = W.shape[1]
    out = torch.empty((M, N), device="cuda")
    Wh = np.ascontiguousarray(W.cpu().numpy())
    rc = lib.cf_debug_pw_gemm(engine, 0, C.c_void_p(A.data_ptr()), C.c_void_p(Wh.ctypes.data), C.c_void_p(out.data_ptr()), M, K, N, None,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.cf_last_error()
    return out

g = torch.Generator(device="cuda").manual_seed(0)
for positive in (True, False):
    for K in (32, 96, 384, 960):
        A = torch.randn(512, K, device="cuda", generator=g); W = torch.randn(K, 64, device="cuda", generator=g)
        if positive: A, W = A.abs(), W.abs()
        A, W = tf32(A.contiguous()), tf32(W.contiguous())
        ref = A.double() @ W.double()
        for e in (0, 2, 1):
            o = gemm(e, A, W).double()
            rel = (o - ref) / ref.abs().clamp_min(1e-30)
            print(f"positive={positive} K={K:4d} engine={e}: mean signed rel err {rel.mean().item():+.3e}  max |rel| {rel.abs().max().item():.3e}"
                  f"  (2^-24 = 5.96e-08, K/8 = {K // 8})", flush=True)
