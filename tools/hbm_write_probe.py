#!/usr/bin/env python
"""HBM bandwidth by direction: fill (write only), sum (read only), copy (read + write) of buffers far larger than the 126 MB L2.
The write-only figure is the roofline of the layers that mostly WRITE (stem: 39 MB in, 419 MB out; the expand convs of the wide maps)."""
import torch

def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3

for mb in (419, 1024, 4096):
    n = mb * (1 << 20) // 4
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    y = torch.empty(n, dtype=torch.float32, device="cuda")
    x.normal_()
    w = t(lambda: y.fill_(1.0))
    m = t(lambda: torch.cuda.memset(y.data_ptr(), 0, n * 4) if hasattr(torch.cuda, "memset") else y.zero_())
    r = t(lambda: x.sum())
    c = t(lambda: y.copy_(x))
    gb = n * 4 / 1e9
    print(f"{mb:5d} MB: fill {gb / w:7.0f} GB/s | zero {gb / m:7.0f} GB/s | read(sum) {gb / r:7.0f} GB/s | copy {2 * gb / c:7.0f} GB/s (read+write bytes)")
