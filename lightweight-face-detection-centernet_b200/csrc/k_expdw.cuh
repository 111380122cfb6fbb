// Fused  expand 1x1 + Swish  ->  depth-wise KSxKS stride S + Swish  for the shallow MBConv blocks
// (model/centernet.py:109-112 with Cin <= 32: layer1.0, 1.1, 2.0, 2.1, 3.0).
//
// Why: the expanded tensor (6 x Cin channels at the block's INPUT resolution) is the largest tensor of
// the network (layer1.0: 96 x 320 x 320 = 39 MB per image in fp32) and in the layer-wise engine it is
// written once and read once.  Here it never leaves the SM: a CTA owns a TH x TW output tile, TMA-loads
// the (TH-1)S+KS x (TW-1)S+KS input halo tile of X (zero fill outside the image = the reference's
// ZeroPad2d, and swish(0 . W) = 0 because the expand conv has no bias), and for every chunk of 32
// hidden channels (i) recomputes the expand conv on the halo tile into shared memory, (ii) runs the
// depth-wise conv from shared memory, (iii) writes the 32 depth-wise output channels (128 B per pixel).
//
// The expand conv runs on the fp32 CUDA cores on purpose: K = Cin is 16..32, the phase is bounded by
// the 2 MUFU ops of each Swish (16/clk/SM) as much as by its FFMAs, and exact fp32 keeps the
// precision-critical shallow layers at FFMA-engine accuracy.
//
//   E-producer mapping: lane = (pg = lane>>3, c4 = lane&7); a thread owns 8 pixels x 4 channels,
//     a warp 32 consecutive halo pixels x 32 channels; X rows are read as LDS.128 from the TMA's
//     SWIZZLE_128B image (conflict-free), W rows as LDS.128 (8 distinct addresses = one wavefront).
//   dw mapping: a thread owns a 2x2 block of outputs for one float4 of channels; the (S+KS)^2 input
//     window is streamed row by row from the swizzled E tile.
#pragma once
#include "k_pw_tc.cuh"

namespace cf {

constexpr int XD_THREADS = 512;

template <int KS, int S>
struct XdGeom {
    static constexpr int TH = S == 1 ? 16 : 8, TW = S == 1 ? 16 : 8;
    static constexpr int IH = (TH - 1) * S + KS, IW = (TW - 1) * S + KS;
    static constexpr int NPX = IH * IW;                        // halo pixels (<= 512 = one pass)
    static constexpr int LO = (KS - S) / 2;                    // model/centernet.py:68-70
    static constexpr int XBYTES = ((NPX * 128 + 1023) / 1024) * 1024;
};

struct XdParams {
    const float* We;   // [CIN][hid]
    const float* Wd;   // [KS*KS][hid]
    float* D;          // [B][Ho][Wo][hid]
    int B, Hi, Wi, Ho, Wo, hid;
    int tiles_x, tiles_y, n_items;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// depth-wise KSxKS stride S + Swish from the swizzled E tile -> global D.  A thread owns an XT x YT block of
// outputs for one float4 of channels; the ((YT-1)S+KS) x ((XT-1)S+KS) input window is streamed row by row.
// Block shapes are chosen so that the phase spreads over as many of the 16 warps as the tile allows:
// 16x16 tiles (S=1): 2x2 blocks on 16 warps; 8x8 tiles (S=2): 1x1 on 16 warps (3x3) or 2x1 on 8 warps (5x5).
template <int KS, int S>
struct XdDwShape {
    static constexpr int XT = (S == 1) ? 2 : (KS == 3 ? 1 : 2);
    static constexpr int YT = (S == 1) ? 2 : 1;
};

// SWZ: the halo tile is in the TMA's SWIZZLE_128B layout (needed where the tile doubles as a tcgen05 operand).  A quarter
// warp (8 lanes = the 8 channel vectors of one pixel) reads one whole 128-byte pixel row per LDS.128 wavefront, which is
// bank-conflict free with or without the swizzle; without it every window address is base + an immediate.
template <typename G, int KS, int S, int XT, int YT, bool WD_GLOBAL, int NWARPS, bool SWZ = true>
__device__ __forceinline__ void xd_dw_phase_g(const uint8_t* Es, const float* Wd_s, const XdParams& p, int warp, int pg, int c4,
                                              int cbase, bool cvalid, int b, int ty, int tx) {
    constexpr int NBX = G::TW / XT, NBLK = (G::TH / YT) * NBX;
    static_assert(G::TW % XT == 0 && G::TH % YT == 0, "output blocks must tile the output tile");
    constexpr int NROW = (YT - 1) * S + KS, NCOL = (XT - 1) * S + KS;
    if (!cvalid) return;
    for (int blk = warp * 4 + pg; blk < NBLK; blk += NWARPS * 4) {
        const int by = blk / NBX, bx = blk - by * NBX;
        float4 acc[YT][XT];
#pragma unroll
        for (int a = 0; a < YT; ++a)
#pragma unroll
            for (int c = 0; c < XT; ++c) acc[a][c] = make_float4(0, 0, 0, 0);
        const int r0 = YT * by * S, q0 = XT * bx * S;  // window origin inside the halo tile
        float4 wt[KS][KS];  // tap rows are loaded once, when the first input row needs them, and stay live for YT rows
#pragma unroll
        for (int rr = 0; rr < NROW; ++rr) {
            float4 win[NCOL];
#pragma unroll
            for (int cc = 0; cc < NCOL; ++cc) {
                const int px = (r0 + rr) * G::IW + q0 + cc;
                win[cc] = SWZ ? *reinterpret_cast<const float4*>(Es + px * 128 + ((c4 ^ (px & 7)) << 4))
                              : *reinterpret_cast<const float4*>(Es + px * 128 + (c4 << 4));
            }
            if (rr < KS) {
#pragma unroll
                for (int kx = 0; kx < KS; ++kx)
                    wt[rr][kx] = WD_GLOBAL ? ldg4(Wd_s + (rr * KS + kx) * p.hid + cbase)
                                           : *reinterpret_cast<const float4*>(Wd_s + (rr * KS + kx) * p.hid + cbase);
            }
#pragma unroll
            for (int dy = 0; dy < YT; ++dy) {
                const int ky = rr - dy * S;
                if (ky < 0 || ky >= KS) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
                    for (int dx = 0; dx < XT; ++dx) fma44(acc[dy][dx], win[dx * S + kx], wt[ky][kx]);
                }
            }
        }
        const int yo0 = ty * G::TH + YT * by, xo0 = tx * G::TW + XT * bx;
        float* o0 = p.D + ((size_t)(b * p.Ho + yo0) * p.Wo + xo0) * p.hid + cbase;  // one 64-bit address per block
        const int rstride = p.Wo * p.hid;
#pragma unroll
        for (int dy = 0; dy < YT; ++dy)
#pragma unroll
            for (int dx = 0; dx < XT; ++dx) {
                if (yo0 + dy < p.Ho && xo0 + dx < p.Wo) st4(o0 + dy * rstride + dx * p.hid, swish4(acc[dy][dx]));
            }
    }
}

template <int KS, int S>
__device__ __forceinline__ void xd_dw_phase(const uint8_t* Es, const float* Wd_s, const XdParams& p, int warp, int pg, int c4,
                                            int cbase, bool cvalid, int b, int ty, int tx) {
    xd_dw_phase_g<XdGeom<KS, S>, KS, S, XdDwShape<KS, S>::XT, XdDwShape<KS, S>::YT, false, XD_THREADS / 32>(
        Es, Wd_s, p, warp, pg, c4, cbase, cvalid, b, ty, tx);
}

template <int KS, int S, int CIN>
__global__ void __launch_bounds__(XD_THREADS, 1) k_expdw(const __grid_constant__ CUtensorMap tmX, const XdParams p) {
    using G = XdGeom<KS, S>;
    static_assert(G::NPX <= 512, "halo tile must fit one E-producer pass");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    // smem map: X[2] | E | We | Wd | barriers
    uint8_t* Xs = sm;
    uint8_t* Es = sm + 2 * G::XBYTES;
    float* We_s = reinterpret_cast<float*>(Es + G::XBYTES);
    float* Wd_s = We_s + CIN * p.hid;
    const uint32_t bars = base + 3 * G::XBYTES + (uint32_t)(CIN + KS * KS) * p.hid * 4;  // 16-byte aligned: hid % 4 == 0

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c4 = lane & 7, pg = lane >> 3;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < CIN * p.hid / 4; i += XD_THREADS) reinterpret_cast<float4*>(We_s)[i] = ldg4(p.We + 4 * i);
    for (int i = tid; i < KS * KS * p.hid / 4; i += XD_THREADS) reinterpret_cast<float4*>(Wd_s)[i] = ldg4(p.Wd + 4 * i);
    __syncthreads();

    auto issue = [&](int item, int buf) {  // thread 0 only
        const int tx = item % p.tiles_x;
        const int t2 = item / p.tiles_x;
        const int ty = t2 % p.tiles_y, b = t2 / p.tiles_y;
        mbar_expect_tx(bars + 8 * buf, (uint32_t)G::NPX * 128u);
        tma_load_4d(base + buf * G::XBYTES, &tmX, 0, tx * G::TW * S - G::LO, ty * G::TH * S - G::LO, b, bars + 8 * buf);
    };
    if (tid == 0 && (int)blockIdx.x < p.n_items) issue(blockIdx.x, 0);

    const int nch = (p.hid + 31) >> 5;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        // the other X buffer was last read by the previous item's E-producer, which every thread left
        // through the __syncthreads() below, so it can be refilled while this item computes
        if (tid == 0 && item + (int)gridDim.x < p.n_items) issue(item + gridDim.x, buf ^ 1);
        mbar_wait(bars + 8 * buf, (it >> 1) & 1u);
        const uint8_t* X = Xs + buf * G::XBYTES;

        const int tx = item % p.tiles_x;
        const int t2 = item / p.tiles_x;
        const int ty = t2 % p.tiles_y, b = t2 / p.tiles_y;

        for (int ch = 0; ch < nch; ++ch) {
            const int cbase = ch * 32 + c4 * 4;       // first hidden channel of this thread's float4
            const bool cvalid = cbase < p.hid;
            // ---------------- expand + Swish on the halo tile -> E (swizzled rows of 128 B) ----------------
            {
                const int p0 = warp * 32 + pg;         // thread's pixels: p0 + 4*i, i = 0..7
                if (cvalid && warp * 32 < G::NPX) {
                    float4 acc[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = make_float4(0, 0, 0, 0);
#pragma unroll
                    for (int kq = 0; kq < CIN / 4; ++kq) {
                        const float4 w0 = *reinterpret_cast<const float4*>(We_s + (kq * 4 + 0) * p.hid + cbase);
                        const float4 w1 = *reinterpret_cast<const float4*>(We_s + (kq * 4 + 1) * p.hid + cbase);
                        const float4 w2 = *reinterpret_cast<const float4*>(We_s + (kq * 4 + 2) * p.hid + cbase);
                        const float4 w3 = *reinterpret_cast<const float4*>(We_s + (kq * 4 + 3) * p.hid + cbase);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int px = p0 + 4 * i;
                            const int pxc = px < G::NPX ? px : G::NPX - 1;  // clamp: rows past the tile are never stored
                            const float4 x = *reinterpret_cast<const float4*>(X + pxc * 128 + ((kq ^ (pxc & 7)) << 4));
                            fma4(acc[i], x.x, w0);
                            fma4(acc[i], x.y, w1);
                            fma4(acc[i], x.z, w2);
                            fma4(acc[i], x.w, w3);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int px = p0 + 4 * i;
                        if (px < G::NPX) *reinterpret_cast<float4*>(Es + px * 128 + ((c4 ^ (px & 7)) << 4)) = swish4(acc[i]);
                    }
                }
            }
            __syncthreads();
            xd_dw_phase<KS, S>(Es, Wd_s, p, warp, pg, c4, cbase, cvalid, b, ty, tx);
            __syncthreads();  // E is rewritten by the next chunk; X[buf] by the TMA issued two items later
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// Tensor-core variant: the expand conv of a 32-channel chunk is  E[NPX x 32] = X[NPX x Cin] . We[Cin x 32]
// on tcgen05 (kind::tf32, 3-pass hi/lo split like k_pw_tc).  The TMA's SWIZZLE_128B halo tile IS the
// K-major A operand (row = halo pixel, 128 B = 32 K-floats), so M-blocks of 128 halo pixels are issued
// straight from it; accumulators (main + correction per M-block) live in TMEM; warp w drains TMEM lane
// quarter (w & 3) of M-block (w >> 2), applies Swish and writes the swizzled E rows the depth-wise phase
// reads.  The MMAs of chunk c+1 are issued before the depth-wise phase of chunk c and overlap it.
// This removes the ~512..1024 FFMAs per thread and chunk of the CUDA-core variant; what remains is bounded
// by the two MUFU ops of each Swish.
// ---------------------------------------------------------------------------------------------------
struct XdTcParams {
    XdParams x;
    const float* we_img;   // [chunk][hi 32 x 128 B | lo 32 x 128 B], K-major SWIZZLE_128B (tc_prepare_layer, NC = 32)
    uint32_t off_xlo, off_e, off_we, off_wd, off_bars;  // smem offsets from the 1024-aligned base
    int prefetch;          // 1: two X buffers, the next tile is loaded while this one computes
};

template <int KS, int S, int CIN>
__global__ void __launch_bounds__(XD_THREADS, 1) k_expdw_tc(const __grid_constant__ CUtensorMap tmX, const XdTcParams P) {
    using G = XdGeom<KS, S>;
    constexpr int NMB = (G::NPX + 127) / 128;  // M-blocks of 128 halo pixels (<= 4)
    static_assert(NMB <= 4, "TMEM: 64 columns per M-block, 256 allocated");
    const XdParams& p = P.x;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* Xlo = sm + P.off_xlo;
    uint8_t* Es = sm + P.off_e;
    float* Wd_s = reinterpret_cast<float*>(sm + P.off_wd);
    const uint32_t we_s = base + P.off_we;
    const uint32_t bars = base + P.off_bars;  // [0],[1]: X full; [2]: MMA done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + P.off_bars + 32);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c4 = lane & 7, pg = lane >> 3;
    const int nch = (p.hid + 31) >> 5;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        mbar_init(bars + 16, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < KS * KS * p.hid / 4; i += XD_THREADS) reinterpret_cast<float4*>(Wd_s)[i] = ldg4(p.Wd + 4 * i);
    for (int i = tid; i < nch * 512; i += XD_THREADS)  // 8 KB of hi|lo image per chunk
        reinterpret_cast<float4*>(sm + P.off_we)[i] = ldg4(P.we_img + 4 * i);
    fence_proxy_async();  // the weight image is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto issue_tma = [&](int item, int buf) {  // thread 0 only
        fence_proxy_async();  // the buffer was last written in place by generic-proxy stores (tf32 rounding)
        const int tx = item % p.tiles_x;
        const int t2 = item / p.tiles_x;
        const int ty = t2 % p.tiles_y, b = t2 / p.tiles_y;
        mbar_expect_tx(bars + 8 * buf, (uint32_t)G::NPX * 128u);
        tma_load_4d(base + buf * G::XBYTES, &tmX, 0, tx * G::TW * S - G::LO, ty * G::TH * S - G::LO, b, bars + 8 * buf);
    };
    const uint32_t idesc = umma_idesc_tf32(32);
    auto issue_mma = [&](int ch, uint32_t xa) {  // thread 0 only: all M-blocks of one 32-channel chunk
        const uint64_t b_hi = umma_desc(we_s + ch * 8192), b_lo = umma_desc(we_s + ch * 8192 + 4096);
#pragma unroll
        for (int mb = 0; mb < NMB; ++mb) {
            const uint64_t a_hi = umma_desc(xa + mb * 16384), a_lo = umma_desc(base + P.off_xlo + mb * 16384);
            const uint32_t d_main = tmem_base + mb * 64, d_corr = d_main + 32;
#pragma unroll
            for (int k = 0; k < (CIN + 7) / 8; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);
                umma_tf32(d_corr, a_lo + ko, b_hi + ko, idesc, k > 0);
                umma_tf32(d_corr, a_hi + ko, b_lo + ko, idesc, 1u);
                umma_tf32(d_main, a_hi + ko, b_hi + ko, idesc, k > 0);
            }
        }
        umma_commit(bars + 16);
    };

    if (tid == 0 && (int)blockIdx.x < p.n_items) issue_tma(blockIdx.x, 0);
    uint32_t it = 0, mma_phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int buf = P.prefetch ? (int)(it & 1) : 0;
        if (P.prefetch) {
            if (tid == 0 && item + (int)gridDim.x < p.n_items) issue_tma(item + gridDim.x, buf ^ 1);
            mbar_wait(bars + 8 * buf, (it >> 1) & 1u);
        } else {
            mbar_wait(bars, it & 1u);
        }
        uint8_t* X = sm + buf * G::XBYTES;
        const int tx = item % p.tiles_x;
        const int t2 = item / p.tiles_x;
        const int ty = t2 % p.tiles_y, b = t2 / p.tiles_y;

        // split the halo tile once: X <- rn_tf32(X) in place, Xlo <- the exact remainder
        for (int i = tid; i < G::NPX * 8; i += XD_THREADS) {
            float4* a = reinterpret_cast<float4*>(X) + i;
            const float4 v = *a;
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            *a = h;
            reinterpret_cast<float4*>(Xlo)[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_mma(0, base + buf * G::XBYTES);
        }
        for (int ch = 0; ch < nch; ++ch) {
            const int cbase = ch * 32 + c4 * 4;
            const bool cvalid = cbase < p.hid;
            // ---- drain the accumulators: + correction, Swish, swizzled E rows ----
            mbar_wait(bars + 16, mma_phase);
            mma_phase ^= 1u;
            tc_fence_after();
            {
                // warp w drains columns [8*(w>>2), +8) of TMEM lane quarter (w&3) for every M-block, so all 16
                // warps share the Swish (MUFU) work of the chunk
                const int q = warp & 3, cg = warp >> 2;
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb) {
                    if (mb * 128 + q * 32 >= G::NPX) break;  // warp-uniform: no valid halo pixel in these 32 rows
                    const int px = mb * 128 + q * 32 + lane;
                    float v[8], c[8];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * 64 + cg * 8);
                    tmem_ld8(taddr, v);
                    tmem_ld8(taddr + 32u, c);
                    tmem_ld_wait();
                    if (px < G::NPX) {
                        uint8_t* er = Es + px * 128;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float4 o = swish4(make_float4(v[4 * j] + c[4 * j], v[4 * j + 1] + c[4 * j + 1],
                                                                v[4 * j + 2] + c[4 * j + 2], v[4 * j + 3] + c[4 * j + 3]));
                            *reinterpret_cast<float4*>(er + (((2 * cg + j) ^ (px & 7)) << 4)) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();  // E complete; every TMEM read of this chunk is done
            if (tid == 0 && ch + 1 < nch) {
                tc_fence_after();
                issue_mma(ch + 1, base + buf * G::XBYTES);  // overlaps the depth-wise phase below
            }
            xd_dw_phase<KS, S>(Es, Wd_s, p, warp, pg, c4, cbase, cvalid, b, ty, tx);
            __syncthreads();  // E is rewritten by the next chunk
        }
        if (!P.prefetch && tid == 0 && item + (int)gridDim.x < p.n_items) issue_tma(item + gridDim.x, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------
// fp32 NHWC tensor [B][H][W][C]: box {32 channels (zero filled past C), IW, IH, 1}, SWIZZLE_128B
inline int xd_make_map(PwTcState& st, CUtensorMap* map, const float* ptr, int B, int H, int W, int C, int IW, int IH, bool swizzle = true) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)IW, (cuuint32_t)IH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((PFN_encodeTiled)st.encode)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CF_ECUDA, "cuTensorMapEncodeTiled(4D %dx%dx%dx%d) failed with CUresult %d", B, H, W, C, (int)r);
    return CF_OK;
}

struct XdLaunch {
    CUtensorMap tmX;
    XdParams p;
    XdTcParams tp;
    bool tc = false;  // tensor-core expand
    int ks = 3, s = 1, cin = 16, grid = 0;
    size_t smem = 0;
};

template <int KS, int S, int CIN>
inline cudaError_t xd_launch_t(const XdLaunch& xl, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_expdw<KS, S, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if (xl.tc) {
        static bool attr_tc = false;
        if (!attr_tc) {
            cudaError_t e = cudaFuncSetAttribute(k_expdw_tc<KS, S, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
            if (e != cudaSuccess) return e;
            attr_tc = true;
        }
        k_expdw_tc<KS, S, CIN><<<xl.grid, XD_THREADS, xl.smem, st>>>(xl.tmX, xl.tp);
    } else {
        k_expdw<KS, S, CIN><<<xl.grid, XD_THREADS, xl.smem, st>>>(xl.tmX, xl.p);
    }
    return cudaGetLastError();
}

inline bool xd_supported(int ks, int s, int cin) {
    return (ks == 3 && s == 2 && (cin == 16 || cin == 32)) || (ks == 3 && s == 1 && cin == 24) ||
           (ks == 5 && s == 2 && cin == 24) || (ks == 5 && s == 1 && cin == 32);
}

inline cudaError_t xd_launch(const XdLaunch& xl, cudaStream_t st) {
    if (xl.ks == 3 && xl.s == 2 && xl.cin == 16) return xd_launch_t<3, 2, 16>(xl, st);
    if (xl.ks == 3 && xl.s == 2 && xl.cin == 32) return xd_launch_t<3, 2, 32>(xl, st);
    if (xl.ks == 3 && xl.s == 1 && xl.cin == 24) return xd_launch_t<3, 1, 24>(xl, st);
    if (xl.ks == 5 && xl.s == 2 && xl.cin == 24) return xd_launch_t<5, 2, 24>(xl, st);
    if (xl.ks == 5 && xl.s == 1 && xl.cin == 32) return xd_launch_t<5, 1, 32>(xl, st);
    return cudaErrorInvalidValue;
}

template <int KS, int S>
inline void xd_geom(int* th, int* tw, int* ih, int* iw, int* xbytes) {
    using G = XdGeom<KS, S>;
    *th = G::TH, *tw = G::TW, *ih = G::IH, *iw = G::IW, *xbytes = G::XBYTES;
}

// shared-memory map of k_expdw_tc: X[1 or 2] | Xlo | E | We images | Wd | barriers.  Returns the dynamic smem size
// (> TC_SMEM_MAX if even the single-buffered variant does not fit).
inline size_t xd_tc_layout(int ks, int hid, int npx, int xb, XdTcParams* tp) {
    const uint32_t nch = (uint32_t)(hid + 31) / 32;
    const uint32_t nmb = (uint32_t)(npx + 127) / 128;
    const uint32_t we_bytes = nch * 8192u, wd_bytes = (((uint32_t)(ks * ks * hid * 4) + 1023u) / 1024u) * 1024u;
    // the A operand of the last M-block reads up to nmb*16 KB past a tile's start: keep every tile region that large
    const uint32_t xreg = nmb * 16384u > (uint32_t)xb ? nmb * 16384u : (uint32_t)xb;
    size_t smem = 0;
    for (int pf = 1; pf >= 0; --pf) {
        const uint32_t nx = pf ? 2u : 1u;
        // X buffers are XBYTES apart (the kernel's indexing); the region after the last one absorbs the over-read
        const uint32_t off_xlo = (nx - 1) * (uint32_t)xb + xreg;
        const uint32_t off_e = off_xlo + xreg;
        const uint32_t off_we = off_e + (uint32_t)xb;
        const uint32_t off_wd = off_we + we_bytes;
        const uint32_t off_bars = off_wd + wd_bytes;
        smem = (size_t)off_bars + 64 + 1024;
        if (tp) tp->prefetch = pf, tp->off_xlo = off_xlo, tp->off_e = off_e, tp->off_we = off_we, tp->off_wd = off_wd, tp->off_bars = off_bars;
        if (smem <= (size_t)TC_SMEM_MAX) break;
    }
    return smem;
}

inline bool xd_tc_fits(int ks, int s, int hid) {
    int th, tw, ih, iw, xb;
    if (ks == 3 && s == 1) xd_geom<3, 1>(&th, &tw, &ih, &iw, &xb);
    else if (ks == 3) xd_geom<3, 2>(&th, &tw, &ih, &iw, &xb);
    else if (s == 1) xd_geom<5, 1>(&th, &tw, &ih, &iw, &xb);
    else xd_geom<5, 2>(&th, &tw, &ih, &iw, &xb);
    return xd_tc_layout(ks, hid, ih * iw, xb, nullptr) <= (size_t)TC_SMEM_MAX;
}

inline int xd_plan(PwTcState& st, int ks, int s, const float* X, const float* We, const float* Wd, float* D, int B, int Hi, int Wi,
                   int cin, int hid, XdLaunch* xl, bool tc = false) {
    if (!xd_supported(ks, s, cin)) return fail(CF_EINVAL, "xd_plan: no fused kernel for k=%d s=%d cin=%d", ks, s, cin);
    int th, tw, ih, iw, xb;
    if (ks == 3 && s == 1) xd_geom<3, 1>(&th, &tw, &ih, &iw, &xb);
    else if (ks == 3) xd_geom<3, 2>(&th, &tw, &ih, &iw, &xb);
    else if (s == 1) xd_geom<5, 1>(&th, &tw, &ih, &iw, &xb);
    else xd_geom<5, 2>(&th, &tw, &ih, &iw, &xb);
    int rc = xd_make_map(st, &xl->tmX, X, B, Hi, Wi, cin, iw, ih);
    if (rc) return rc;
    XdParams& p = xl->p;
    p.We = We;
    p.Wd = Wd;
    p.D = D;
    p.B = B;
    p.Hi = Hi;
    p.Wi = Wi;
    p.Ho = Hi / s;
    p.Wo = Wi / s;
    p.hid = hid;
    p.tiles_x = (p.Wo + tw - 1) / tw;
    p.tiles_y = (p.Ho + th - 1) / th;
    p.n_items = B * p.tiles_x * p.tiles_y;
    xl->ks = ks;
    xl->s = s;
    xl->cin = cin;
    xl->grid = p.n_items < st.sms ? p.n_items : st.sms;
    xl->smem = (size_t)3 * xb + (size_t)(cin + ks * ks) * hid * 4 + 64 + 1024;
    xl->tc = tc;
    if (tc) {
        auto it = st.layers.find(We);
        if (it == st.layers.end() || it->second.NC != 32 || it->second.nkb != 1)
            return fail(CF_EINVAL, "xd_plan: expand weights were not prepared as 32-column tensor-core images");
        XdTcParams& tp = xl->tp;
        tp.x = p;
        tp.we_img = it->second.img;
        xl->smem = xd_tc_layout(ks, hid, ih * iw, xb, &tp);
    }
    if (xl->smem > (size_t)TC_SMEM_MAX) return fail(CF_EINVAL, "xd_plan: tile does not fit shared memory (%zu B)", xl->smem);
    return CF_OK;
}

}  // namespace cf
