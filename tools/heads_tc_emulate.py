#!/usr/bin/env python
"""CPU emulation of the index logic of tools/heads_tc_probe.cu (no GPU needed): TMA halo tile written with the address-keyed
128-byte swizzle, nine tap-shifted A windows read back as the tensor core reads them (tools/desc_shift_probe.cu measured that rule),
weight image layout, anchor-row -> output-pixel mapping with the garbage rows (window wrap, uninitialised slack lines = NaN here)
dropped.  Prints the number of uncovered outputs and the max error against a direct 3x3 convolution: expected `0` and `0.0`."""
import numpy as np

rng = np.random.RandomState(0)
B, H, W = 2, 20, 35
X = rng.randn(B, H, W, 24)
Wt = rng.randn(9, 24, 15)
bias = rng.randn(15)
TW, TH, WT, HT = 16, 7, 18, 9
tiles_x, tiles_y = -(-W // TW), -(-H // TH)
out = np.full((B, H, W, 15), np.nan)


def line_store(tile, line, vec32):  # TMA: logical 16-byte chunk c of line L lands at chunk c ^ (L & 7)
    for c in range(8):
        pc = c ^ (line & 7)
        tile[line, pc * 4:pc * 4 + 4] = vec32[c * 4:c * 4 + 4]


def line_read(tile, line):  # tcgen05: the same address-keyed rule, whatever line the descriptor starts at
    v = np.empty(32)
    for c in range(8):
        pc = c ^ (line & 7)
        v[c * 4:c * 4 + 4] = tile[line, pc * 4:pc * 4 + 4]
    return v


wimg = np.zeros((9, 16, 32))  # per tap: rows n x 32 k, swizzled by n & 7 (the hi block of the probe's image)
for t in range(9):
    for n in range(15):
        for k in range(24):
            wimg[t, n, (((k >> 2) ^ (n & 7)) << 2) + (k & 3)] = Wt[t, k, n]

for tile_id in range(B * tiles_x * tiles_y):
    tx, r = tile_id % tiles_x, tile_id // tiles_x
    y0, b, x0 = (r % tiles_y) * TH, r // tiles_y, tx * TW
    tile = np.full((168, 32), np.nan)  # lines 162..167 are never written
    for hy in range(HT):
        for hx in range(WT):
            y, x = y0 - 1 + hy, x0 - 1 + hx
            v = np.zeros(32)
            if 0 <= y < H and 0 <= x < W:
                v[:24] = X[b, y, x]
            line_store(tile, hy * WT + hx, v)
    D = np.zeros((128, 16))
    for t in range(9):
        shift = (t // 3) * WT + (t % 3)
        A = np.stack([line_read(tile, shift + m) for m in range(128)])
        Bm = np.stack([line_read(wimg[t], n) for n in range(16)])
        with np.errstate(invalid="ignore"):
            D += A @ Bm.T
    for m in range(128):
        hy, hx = divmod(m, WT)
        y, x = y0 + hy, x0 + hx
        if hx < TW and hy < TH and y < H and x < W:
            out[b, y, x] = D[m, :15] + bias

ref = np.zeros((B, H, W, 15))
Xp = np.pad(X, ((0, 0), (1, 1), (1, 1), (0, 0)))
for ky in range(3):
    for kx in range(3):
        ref += Xp[:, ky:ky + H, kx:kx + W, :] @ Wt[ky * 3 + kx]
ref += bias
print("uncovered outputs:", int(np.isnan(out).sum()), " max |emulation - direct conv|:", float(np.nanmax(np.abs(out - ref))))
