// k_heads_tc: the four collapsed heads (dense 3x3, 24 -> 15, + sigmoid/clamp; model/centernet.py:240-261, centerface.py:43) as ONE
// tcgen05 GEMM per TMA halo tile, the nine taps as nine K blocks whose A operands are shifted windows into the same tile.
//
// Basis: a K-major SWIZZLE_128B shared-memory descriptor may start at ANY 128-byte line of a swizzled tile (base_offset 0;
// tools/desc_shift_probe.cu, profiles/r1d_desc_shift_probe.md).
//
// Geometry.  Halo tile = 9 rows x 18 columns of pixels, 32 channels (24 real, TMA zero-fills 24..31) = 162 lines of 128 B, line
// L = hy * 18 + hx, origin (y0 - 1, x0 - 1): out-of-image pixels are zero-filled by TMA = the convolution's zero padding.  MMA row m
// is the anchor line m: the window of output pixel (y0 + hy, x0 + hx) starts there, tap (ky, kx) reads line m + ky * 18 + kx.  Rows
// with hx >= 16 (window wraps into the next halo row) and rows 126, 127 are garbage and are dropped in the drain: 112 valid outputs
// (7 rows x 16 columns) per 128-row MMA block.
//
// Per tile: TMA -> all threads split the tile into tf32 hi (in place) and lo (second buffer) -> one elected lane issues
// 9 taps x 3 K steps x { A_hi.[W_hi|W_lo] (N = 32: main and correction accumulator), A_lo.W_hi (N = 16: correction) } -> the drain of
// the PREVIOUS tile runs under these MMAs (two accumulator slots) -> bias, sigmoid/clamp, planar stores.  Two CTAs per SM.
// Bound: every MMA streams a 4 KB A slice out of shared memory whatever N is -- 54 MMAs per tile at ~55 cycles each on the tensor
// pipe that the SM's two CTAs share: 7 360 tiles x 54 x 55 cycles / 148 SMs = 77 us of the measured 87.  (A second issuing warp
// with its own accumulator pair changed nothing, round 2: the pipe, not the issuing thread, is the limit.)  The FFMA kernel it
// replaces (k_heads, 3456 FFMA per pixel) stays as the fp32 validation engine's head.
#pragma once
#include "k_dwt.cuh"

namespace cf {

constexpr int HT_TW = 16, HT_TH = 7;                 // outputs per tile
constexpr int HT_WT = HT_TW + 2, HT_HT = HT_TH + 2;  // halo tile
constexpr int HT_LINES = HT_WT * HT_HT;              // 162 lines written by TMA
constexpr int HT_TILE_BYTES = 21504;                 // 168 lines: rows up to 127 + 2 * 18 + 2 = 165 are read (garbage anchors only)
constexpr int HT_NST = 2;
constexpr int HT_W_BYTES = 9 * 4096;                 // per tap: [hi 16 x 128 B | lo 16 x 128 B]
constexpr int HT_THREADS = 256;
constexpr int HT_OFF_LO = HT_NST * HT_TILE_BYTES;
constexpr int HT_OFF_W = HT_OFF_LO + HT_TILE_BYTES;
constexpr int HT_OFF_BARS = HT_OFF_W + HT_W_BYTES;
constexpr int HT_SMEM = HT_OFF_BARS + 64 + 1024;
static_assert(HT_TH * HT_WT <= 128, "one MMA block per tile");
static_assert(HT_TILE_BYTES % 1024 == 0 && HT_TILE_BYTES >= (127 + 2 * HT_WT + 2 + 1) * 128, "tile buffer");

struct HeadsTcParams {
    const float* wimg;  // device, HT_W_BYTES
    float bias[16];
    float *hm, *hm_sig, *wh, *lm, *reg;  // planar outputs
    int B, H, W, tiles_x, tiles_y, n_tiles;
};

__global__ void __launch_bounds__(HT_THREADS, 2) k_heads_tc(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ HeadsTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t lo_sm = base + HT_OFF_LO, w_sm = base + HT_OFF_W, bars = base + HT_OFF_BARS;
    const uint32_t bar_w = bars + 8 * HT_NST, bar_mma = bar_w + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + HT_OFF_BARS + 8 * HT_NST + 24);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        for (int i = 0; i < HT_NST + 2; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // the FPN map is the previous kernel's output; the head planes may still be read by the previous step's decode

    const int my_tiles = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_of = [&](int j, int* b, int* y0, int* x0) {
        const int t = (int)blockIdx.x + j * (int)gridDim.x;
        const int tx = t % p.tiles_x, r = t / p.tiles_x;
        *x0 = tx * HT_TW;
        *y0 = (r % p.tiles_y) * HT_TH;
        *b = r / p.tiles_y;
    };
    auto issue_tile = [&](int j) {  // thread 0 only
        int b, y0, x0;
        tile_of(j, &b, &y0, &x0);
        const int stage = j % HT_NST;
        mbar_expect_tx(bars + 8 * stage, HT_LINES * 128);
        tma_load_4d(base + stage * HT_TILE_BYTES, &tmX, 0, x0 - 1, y0 - 1, b, bars + 8 * stage);
    };
    if (tid == 0 && my_tiles > 0) {
        mbar_expect_tx(bar_w, HT_W_BYTES);
        for (uint32_t off = 0; off < (uint32_t)HT_W_BYTES; off += 16384u)
            bulk_load(w_sm + off, reinterpret_cast<const uint8_t*>(p.wimg) + off, HT_W_BYTES - off < 16384u ? HT_W_BYTES - off : 16384u, bar_w);
        for (int j = 0; j < HT_NST && j < my_tiles; ++j) issue_tile(j);
    }
    const uint32_t idesc16 = umma_idesc_tf32(16), idesc32 = umma_idesc_tf32(32);

    auto drain = [&](int j) {  // tile j's accumulators -> bias, sigmoid/clamp, planar stores; thread = one anchor row x 8 channels
        int b, y0, x0;
        tile_of(j, &b, &y0, &x0);
        const int q = warp & 3, half = warp >> 2;
        const int m = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((j & 1) * 32 + half * 8);
        float v[8], c[8];
        tmem_ld8(taddr, v);
        tmem_ld8(taddr + 16u, c);
        tmem_ld_wait();
        const int hy = m / HT_WT, hx = m - hy * HT_WT;
        const int y = y0 + hy, x = x0 + hx;
        if (hx < HT_TW && hy < HT_TH && y < p.H && x < p.W) {
            const size_t plane = (size_t)p.H * p.W, pix = (size_t)y * p.W + x;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ch = half * 8 + i;
                const float o = v[i] + c[i] + p.bias[ch];
                if (ch == 0) {
                    p.hm[(size_t)b * plane + pix] = o;
                    const float s = 1.f / (1.f + expf(-o));
                    p.hm_sig[(size_t)b * plane + pix] = fminf(fmaxf(s, 1e-4f), 1.f - 1e-4f);
                } else if (ch <= 2) {
                    p.wh[((size_t)b * 2 + (ch - 1)) * plane + pix] = o;
                } else if (ch <= 12) {
                    p.lm[((size_t)b * 10 + (ch - 3)) * plane + pix] = o;
                } else if (ch <= 14) {
                    p.reg[((size_t)b * 2 + (ch - 13)) * plane + pix] = o;
                }
            }
        }
        tc_fence_before();  // the loads are done before the barrier that precedes the MMAs overwriting this slot
    };

    for (int j = 0; j < my_tiles; ++j) {
        const int stage = j % HT_NST;
        mbar_wait(bars + 8 * stage, (uint32_t)(j / HT_NST) & 1u);
        if (j >= 1) {
            // MMAs of tile j-1 are complete: its stage and the lo buffer are free, its accumulator slot is ready
            mbar_wait(bar_mma, (uint32_t)(j - 1) & 1u);
            tc_fence_after();
            if (tid == 0 && j + 1 < my_tiles) issue_tile(j + 1);  // into the stage tile j-1 used
        }
        // ---- split the tile: hi in place, lo beside it (element-wise, so the swizzle is irrelevant here) ----
        uint8_t* raw = sm + stage * HT_TILE_BYTES;
        uint8_t* lob = sm + HT_OFF_LO;
        for (int i = tid; i < HT_LINES * 8; i += HT_THREADS) {
            float4 v = *reinterpret_cast<const float4*>(raw + i * 16);
            float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            *reinterpret_cast<float4*>(raw + i * 16) = h;
            *reinterpret_cast<float4*>(lob + i * 16) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            if (j == 0) mbar_wait(bar_w, 0);
            tc_fence_after();
            const uint32_t d_main = tmem_base + (uint32_t)((j & 1) * 32), d_corr = d_main + 16u;
            const uint32_t a_sm = base + stage * HT_TILE_BYTES;
            if (elect_one()) {
#pragma unroll 1
                for (int t = 0; t < 9; ++t) {
                    const uint32_t shift = (uint32_t)((t / 3) * HT_WT + (t % 3)) * 128u;
                    const uint64_t a_hi = umma_desc(a_sm + shift), a_lo = umma_desc(lo_sm + shift);
                    const uint64_t b_hi = umma_desc(w_sm + (uint32_t)t * 4096u);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {  // K = 24 channels in steps of 8 (the tile's channels 24..31 are TMA zero fill, never read)
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_tf32(d_main, a_hi + ko, b_hi + ko, idesc32, (t > 0 || k > 0) ? 1u : 0u);  // main += hi.hi ; corr += hi.lo
                        umma_tf32(d_corr, a_lo + ko, b_hi + ko, idesc16, 1u);                           // corr += lo.hi
                    }
                }
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        if (j >= 1) drain(j - 1);  // runs under tile j's MMAs
    }
    if (my_tiles > 0) {
        mbar_wait(bar_mma, (uint32_t)(my_tiles - 1) & 1u);
        tc_fence_after();
        drain(my_tiles - 1);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------
struct HeadsTcLaunch {
    CUtensorMap tmX;
    HeadsTcParams p;
    int grid = 0;
};

// weight image of the collapsed head conv (host copy hw[(tap * 24 + k) * 16 + n], n = 15 is padding): per tap
// [hi: 16 rows (n) x 32 k | lo], K-major, SWIZZLE_128B
inline int heads_tc_prepare(PwTcState& st, const float* hw) {
    if (st.heads_img) return CF_OK;
    std::vector<float> img(HT_W_BYTES / 4, 0.f);
    for (int t = 0; t < 9; ++t)
        for (int n = 0; n < 15; ++n)
            for (int k = 0; k < 24; ++k) {
                const float w = hw[(t * 24 + k) * 16 + n], h = tf32_hi(w);
                const int pos = n * 32 + (((k >> 2) ^ (n & 7)) << 2) + (k & 3);
                img[t * 1024 + pos] = h;
                img[t * 1024 + 512 + pos] = tf32_hi(w - h);
            }
    if (cudaMalloc((void**)&st.heads_img, HT_W_BYTES) != cudaSuccess) return fail(CF_ECUDA, "heads_tc_prepare: cudaMalloc failed");
    if (cudaMemcpy(st.heads_img, img.data(), HT_W_BYTES, cudaMemcpyHostToDevice) != cudaSuccess) return fail(CF_ECUDA, "heads_tc_prepare: cudaMemcpy failed");
    return CF_OK;
}

inline int heads_tc_plan(PwTcState& st, const float* X, const float* bias16, float* hm, float* wh, float* lm, float* reg, float* hm_sig, int B,
                         int H, int W, HeadsTcLaunch* hl) {
    if (!st.heads_img) return fail(CF_EINVAL, "heads_tc_plan: weight image was not prepared");
    int rc = xd_make_map(st, &hl->tmX, X, B, H, W, 24, HT_WT, HT_HT);
    if (rc) return rc;
    HeadsTcParams& p = hl->p;
    p.wimg = st.heads_img;
    for (int i = 0; i < 16; ++i) p.bias[i] = bias16[i];
    p.hm = hm, p.hm_sig = hm_sig, p.wh = wh, p.lm = lm, p.reg = reg;
    p.B = B, p.H = H, p.W = W;
    p.tiles_x = (W + HT_TW - 1) / HT_TW, p.tiles_y = (H + HT_TH - 1) / HT_TH;
    const long long nt = (long long)B * p.tiles_x * p.tiles_y;
    if (nt > 0x7fffffffLL) return fail(CF_EINVAL, "heads_tc_plan: too many tiles");
    p.n_tiles = (int)nt;
    hl->grid = p.n_tiles < 2 * st.sms ? p.n_tiles : 2 * st.sms;
    return CF_OK;
}

inline cudaError_t heads_tc_launch(const HeadsTcLaunch& hl, cudaStream_t s) {
    cudaError_t e = smem_optin((const void*)k_heads_tc, HT_SMEM);
    if (e != cudaSuccess) return e;
    return launch_pdl(k_heads_tc, dim3(hl.grid), dim3(HT_THREADS), (size_t)HT_SMEM, s, hl.tmX, hl.p);
}

}  // namespace cf
