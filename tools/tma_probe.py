#!/usr/bin/env python
"""Raw TMA streaming rate of a [M][K] fp32 matrix vs ring depth, box height and CTAs per SM (cf_debug_tma_stream)."""
import ctypes as C, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
pkg.build()
lib = pkg._lib.load()
for (M, K) in ((3276800, 32), (819200, 144)):
    A = torch.randn(M, K, device="cuda")
    gb = M * K * 4 / 1e9
    for ctas in (1, 2):
        for rows in (128, 256):
            for stages in (2, 4, 6, 8, 12):
                ms = C.c_float()
                rc = lib.cf_debug_tma_stream(C.c_void_p(A.data_ptr()), M, K, stages, rows, ctas, C.byref(ms))
                if rc != 0:
                    print(M, K, ctas, rows, stages, "skip:", lib.cf_last_error().decode()); continue
                print(f"M={M} K={K} ctas/SM={ctas} box_rows={rows} stages={stages}: {ms.value*1e3:7.1f} us  {gb/ms.value*1e3/1e3:6.2f} TB/s"
                      f"  in flight/SM {ctas*stages*rows*128/1024:.0f} KB", flush=True)
