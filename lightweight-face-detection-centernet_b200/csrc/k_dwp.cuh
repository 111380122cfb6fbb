// k_dwp: fused  depth-wise KSxKS stride S + Swish  ->  project 1x1 (tcgen05, 3xTF32) [+ residual]  for the shallow
// MBConv blocks (model/centernet.py:112-118, :135-137).
//
// In the layer-wise engine the depth-wise output D (hidden channels at the block's OUTPUT resolution) is written to
// HBM by one kernel and read back by the projection GEMM.  Here D never leaves the SM: a CTA owns a 10x10 output
// tile (100 rows of one 128-row MMA block), streams the tile's halo through a TMA ring one 32-channel chunk at a time
// (same producer as k_dwt), computes the depth-wise conv from shared memory, writes the chunk as the K-major,
// 128B-swizzled, tf32 hi/lo-split A operand of  Y[128 x Cout] += D_chunk[128 x 32] . Wp[32 x Cout]  and accumulates over
// the chunks in TMEM (main + correction accumulator as in k_pw_tc).  After the last chunk the accumulator is drained,
// the residual added and Y stored.  HBM traffic per block: the expanded tensor in, the block output out.
#pragma once
#include "k_dwt.cuh"

namespace cf {

struct DwpParams {
    XdParams x;           // Wd, hid = C, Hi/Wi/Ho/Wo, tiles_x/tiles_y; n_items = number of tiles
    const float* wp_img;  // [chunk][hi NCP x 128 B | lo NCP x 128 B], K-major SWIZZLE_128B (tc_prepare_layer, NC = NCP)
    float* Y;             // [B][Ho][Wo][cout]
    const float* res;     // residual [B][Ho][Wo][cout] or NULL
    int cout, ncp;        // real / padded-to-32 output channels
    int nchunk, nst;      // 32-channel chunks, E pipeline stages
    uint32_t off_dhi, off_dlo, off_wp, off_bars;
    uint32_t tmem_cols;
};

template <int KS, int S>
__global__ void __launch_bounds__(DWT_THREADS, 2) k_dwp(const __grid_constant__ CUtensorMap tmX, const DwpParams P) {
    using G = DwtGeom<KS, S, 0>;
    constexpr int XT = 2, YT = 2, NBX = G::TW / XT, NBLK = (G::TH / YT) * NBX;  // 25 blocks of 2x2 outputs
    constexpr int NROW = (YT - 1) * S + KS, NCOL = (XT - 1) * S + KS;
    static_assert(NBLK <= (DWT_THREADS / 32) * 4, "one output block per (warp, lane group)");
    const XdParams& p = P.x;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* Dhi = sm + P.off_dhi;
    uint8_t* Dlo = sm + P.off_dlo;
    const int nst = P.nst, nchunk = P.nchunk;
    const uint32_t bars = base + P.off_bars;  // [0..nst) E full, then MMA done, Wp slot 0/1 full
    const uint32_t bar_mma = bars + 8 * nst, bar_wp = bar_mma + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + P.off_bars + 8 * nst + 32);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c4 = lane & 7, pg = lane >> 3;
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        for (int i = 0; i < nst + 3; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // rows 100..127 of the A operand are never written by the depth-wise phase: zero them once so that the unused
    // accumulator lanes stay finite
    for (int i = tid; i < (128 - G::TH * G::TW) * 8; i += DWT_THREADS) {
        reinterpret_cast<float4*>(Dhi + G::TH * G::TW * 128)[i] = make_float4(0, 0, 0, 0);
        reinterpret_cast<float4*>(Dlo + G::TH * G::TW * 128)[i] = make_float4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_tiles = ((int)blockIdx.x < p.n_items) ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * nchunk;  // job j = (my tile j / nchunk, chunk j % nchunk)
    const uint32_t wp_bytes = (uint32_t)P.ncp * 256u;

    auto tile_of = [&](int job, int* tx, int* ty, int* b) {
        int t = (int)blockIdx.x + (job / nchunk) * (int)gridDim.x;
        *tx = t % p.tiles_x;
        t /= p.tiles_x;
        *ty = t % p.tiles_y;
        *b = t / p.tiles_y;
    };
    auto issue_e = [&](int job) {  // thread 0 only
        int tx, ty, b;
        tile_of(job, &tx, &ty, &b);
        const int stage = job % nst;
        mbar_expect_tx(bars + 8 * stage, (uint32_t)G::NPX * 128u);
        tma_load_4d(base + stage * G::XBYTES, &tmX, (job % nchunk) * 32, tx * G::TW * S - G::LO, ty * G::TH * S - G::LO, b,
                    bars + 8 * stage);
    };
    auto issue_wp = [&](int job) {  // thread 0 only
        mbar_expect_tx(bar_wp + 8 * (job & 1), wp_bytes);
        bulk_load(base + P.off_wp + (job & 1) * wp_bytes, reinterpret_cast<const uint8_t*>(P.wp_img) + (size_t)(job % nchunk) * wp_bytes,
                  wp_bytes, bar_wp + 8 * (job & 1));
    };
    const uint32_t idesc = umma_idesc_tf32(P.ncp);

    if (tid == 0 && total > 0) {
        for (int j = 0; j < nst && j < total; ++j) issue_e(j);
        issue_wp(0);
    }

    for (int job = 0; job < total; ++job) {
        const int ch = job % nchunk;
        int tx, ty, b;
        tile_of(job, &tx, &ty, &b);
        const int stage = job % nst;
        mbar_wait(bars + 8 * stage, (uint32_t)(job / nst) & 1u);
        const uint8_t* Es = sm + stage * G::XBYTES;

        // ---- depth-wise conv of this 32-channel chunk into registers (2x2 outputs x 4 channels per thread) ----
        const int cbase = ch * 32 + c4 * 4;
        const bool cvalid = cbase < p.hid;
        const int blk = warp * 4 + pg;
        const bool active = blk < NBLK;
        const int by = blk / NBX, bx = blk - by * NBX;
        float4 acc[YT][XT];
#pragma unroll
        for (int a = 0; a < YT; ++a)
#pragma unroll
            for (int c = 0; c < XT; ++c) acc[a][c] = make_float4(0, 0, 0, 0);
        if (active && cvalid) {
            const int r0 = YT * by * S, q0 = XT * bx * S;
#pragma unroll
            for (int rr = 0; rr < NROW; ++rr) {
                float4 win[NCOL];
#pragma unroll
                for (int cc = 0; cc < NCOL; ++cc) {
                    const int px = (r0 + rr) * G::IW + q0 + cc;
                    win[cc] = *reinterpret_cast<const float4*>(Es + px * 128 + ((c4 ^ (px & 7)) << 4));
                }
#pragma unroll
                for (int dy = 0; dy < YT; ++dy) {
                    const int ky = rr - dy * S;
                    if (ky < 0 || ky >= KS) continue;
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx) {
                        const float4 wv = ldg4(p.Wd + (ky * KS + kx) * p.hid + cbase);
#pragma unroll
                        for (int dx = 0; dx < XT; ++dx) fma44(acc[dy][dx], win[dx * S + kx], wv);
                    }
                }
            }
#pragma unroll
            for (int dy = 0; dy < YT; ++dy)
#pragma unroll
                for (int dx = 0; dx < XT; ++dx) acc[dy][dx] = swish4(acc[dy][dx]);  // swish(0) = 0 keeps padded channels zero
        }
        // ---- the previous chunk's MMAs must have retired before D is overwritten ----
        if (job > 0) mbar_wait(bar_mma, (uint32_t)(job - 1) & 1u);
        if (active) {
#pragma unroll
            for (int dy = 0; dy < YT; ++dy)
#pragma unroll
                for (int dx = 0; dx < XT; ++dx) {
                    const int r = (YT * by + dy) * G::TW + XT * bx + dx;  // A-operand row = output pixel of the tile
                    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c4 ^ (r & 7)) << 4);
                    const float4 v = acc[dy][dx];
                    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    *reinterpret_cast<float4*>(Dhi + off) = h;
                    *reinterpret_cast<float4*>(Dlo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();  // D complete, E stage drained
        if (tid == 0) {
            if (job + nst < total) issue_e(job + nst);
            if (job + 1 < total) issue_wp(job + 1);  // its slot was read by job-1, which has retired
            mbar_wait(bar_wp + 8 * (job & 1), (uint32_t)(job >> 1) & 1u);
            tc_fence_after();
            const uint32_t wb = base + P.off_wp + (job & 1) * wp_bytes;
            const uint64_t a_hi = umma_desc(base + P.off_dhi), a_lo = umma_desc(base + P.off_dlo);
            const uint64_t b_hi = umma_desc(wb), b_lo = umma_desc(wb + (uint32_t)P.ncp * 128u);
            const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)P.ncp;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);
                const uint32_t accf = (ch > 0 || k > 0) ? 1u : 0u;
                umma_tf32(d_corr, a_lo + ko, b_hi + ko, idesc, accf);
                umma_tf32(d_corr, a_hi + ko, b_lo + ko, idesc, 1u);
                umma_tf32(d_main, a_hi + ko, b_hi + ko, idesc, accf);
            }
            umma_commit(bar_mma);
        }
        if (ch == nchunk - 1) {
            // ---- tile done: drain the accumulator, add the residual, store Y ----
            mbar_wait(bar_mma, (uint32_t)job & 1u);
            tc_fence_after();
            const int q = warp & 3, half = warp >> 2;
            const int r = q * 32 + lane;  // accumulator row = output pixel of the tile
            const int yo = ty * G::TH + r / G::TW, xo = tx * G::TW + r % G::TW;
            const bool rvalid = r < G::TH * G::TW && yo < p.Ho && xo < p.Wo;
            const size_t pix = ((size_t)(b * p.Ho + yo) * p.Wo + xo) * P.cout;
            const int ncols = P.ncp >> 1;  // this warp's half of the padded columns
            for (int c0 = half * ncols; c0 < (half + 1) * ncols; c0 += 8) {
                float v[8], c[8];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                tmem_ld8(taddr, v);
                tmem_ld8(taddr + (uint32_t)P.ncp, c);
                tmem_ld_wait();
                if (rvalid && c0 < P.cout) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float4 o = make_float4(v[4 * j] + c[4 * j], v[4 * j + 1] + c[4 * j + 1], v[4 * j + 2] + c[4 * j + 2],
                                               v[4 * j + 3] + c[4 * j + 3]);
                        if (P.res) {
                            const float4 rr = ldg4(P.res + pix + c0 + 4 * j);
                            o.x += rr.x, o.y += rr.y, o.z += rr.z, o.w += rr.w;
                        }
                        st4(P.Y + pix + c0 + 4 * j, o);
                    }
                }
            }
            tc_fence_before();
            __syncthreads();  // every TMEM read is done before the next tile's first MMA overwrites the accumulator
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------
struct DwpLaunch {
    CUtensorMap tmX;
    DwpParams p;
    int ks = 3, s = 1, grid = 0;
    size_t smem = 0;
};

template <int KS, int S>
inline cudaError_t dwp_launch_t(const DwpLaunch& dl, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_dwp<KS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    k_dwp<KS, S><<<dl.grid, DWT_THREADS, dl.smem, st>>>(dl.tmX, dl.p);
    return cudaGetLastError();
}

inline cudaError_t dwp_launch(const DwpLaunch& dl, cudaStream_t st) {
    if (dl.ks == 3 && dl.s == 1) return dwp_launch_t<3, 1>(dl, st);
    if (dl.ks == 3 && dl.s == 2) return dwp_launch_t<3, 2>(dl, st);
    if (dl.ks == 5 && dl.s == 1) return dwp_launch_t<5, 1>(dl, st);
    return dwp_launch_t<5, 2>(dl, st);
}

inline int dwp_ncp(int cout) { return (cout + 31) / 32 * 32; }
// the accumulator pair must fit 256 TMEM columns so that two CTAs can share an SM
inline bool dwp_supported(int C, int cout) { return C % 4 == 0 && C >= 16 && cout % 8 == 0 && dwp_ncp(cout) <= 128; }

inline int dwp_plan(PwTcState& st, int ks, int s, const float* E, const float* Wd, const float* Wp, float* Y, const float* res, int B,
                    int Hi, int Wi, int C, int cout, DwpLaunch* dl) {
    if (!dwp_supported(C, cout)) return fail(CF_EINVAL, "dwp_plan: C=%d cout=%d unsupported", C, cout);
    const int ncp = dwp_ncp(cout);
    auto it = st.layers.find(Wp);
    if (it == st.layers.end() || it->second.NC != ncp || it->second.nchunks != 1)
        return fail(CF_EINVAL, "dwp_plan: projection weights were not prepared as one %d-column image per K block", ncp);
    int ih, iw, xb;
    int th_, tw_;
    if (ks == 3 && s == 1) dwt_geom<3, 1, 0>(&th_, &tw_, &ih, &iw, &xb);
    else if (ks == 3) dwt_geom<3, 2, 0>(&th_, &tw_, &ih, &iw, &xb);
    else if (s == 1) dwt_geom<5, 1, 0>(&th_, &tw_, &ih, &iw, &xb);
    else dwt_geom<5, 2, 0>(&th_, &tw_, &ih, &iw, &xb);
    int rc = xd_make_map(st, &dl->tmX, E, B, Hi, Wi, C, iw, ih);
    if (rc) return rc;
    DwpParams& P = dl->p;
    XdParams& p = P.x;
    p.We = nullptr;
    p.Wd = Wd;
    p.D = nullptr;
    p.B = B;
    p.Hi = Hi;
    p.Wi = Wi;
    p.Ho = Hi / s;
    p.Wo = Wi / s;
    p.hid = C;
    p.tiles_x = (p.Wo + 9) / 10;
    p.tiles_y = (p.Ho + 9) / 10;
    p.n_items = B * p.tiles_x * p.tiles_y;
    P.wp_img = it->second.img;
    P.Y = Y;
    P.res = res;
    P.cout = cout;
    P.ncp = ncp;
    P.nchunk = (C + 31) / 32;
    P.tmem_cols = 2 * ncp <= 64 ? 64u : (2 * ncp <= 128 ? 128u : 256u);
    const uint32_t fixed = 2 * 16384u + 2 * (uint32_t)ncp * 256u + 1024u /*barriers*/ + 1024u /*alignment*/;
    const int per_cta2 = ((int)(TC_SMEM_MAX / 2) - 1024 - (int)fixed) / xb;
    int ctas_per_sm = 2, nst = per_cta2;
    if (nst < 2) ctas_per_sm = 1, nst = ((int)TC_SMEM_MAX - (int)fixed) / xb;
    if (nst > 6) nst = 6;
    if (nst < 1) return fail(CF_EINVAL, "dwp_plan: tile does not fit shared memory");
    P.nst = nst;
    P.off_dhi = (uint32_t)nst * (uint32_t)xb;
    P.off_dlo = P.off_dhi + 16384u;
    P.off_wp = P.off_dlo + 16384u;
    P.off_bars = P.off_wp + 2 * (uint32_t)ncp * 256u;
    dl->smem = (size_t)P.off_bars + 1024 + 1024;
    dl->ks = ks;
    dl->s = s;
    const int want = ctas_per_sm * st.sms;
    dl->grid = p.n_items < want ? p.n_items : want;
    return CF_OK;
}

}  // namespace cf
