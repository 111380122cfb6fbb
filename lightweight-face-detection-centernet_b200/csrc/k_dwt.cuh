// k_dwt: depth-wise KSxKS stride S + Swish fed by TMA (third-generation depth-wise kernel).
//
// The register/L1 kernels (k_dw, k_dw3 in k_conv.cuh) are latency-bound: ncu shows 48 % of their warp stalls on
// the long scoreboard at 36 % of DRAM peak -- a thread can keep only a handful of 16-byte loads in flight.  Here the
// loads are bulk tensor copies: a CTA walks over (image, 10x10 output tile, 32-channel chunk) items, one elected
// thread keeps NST halo tiles in flight with cp.async.bulk.tensor.4d (box = 32 channels x IW x IH, no swizzle,
// zero fill outside the image = the reference's ZeroPad2d, model/centernet.py:63-70), and all 256 threads compute
// from shared memory with 2x2 (or 2x4) output register blocking (xd_dw_phase_g).  Bytes in flight
// per SM = NST x tile (up to ~190 KB) instead of a few KB of registers.
#pragma once
#include "k_pw_tc.cuh"

namespace cf {

// ---- halo-tile helpers shared with the fused MBConv kernel (k_mbf) ----------------------------------------------------

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

// depth-wise KSxKS stride S + Swish from a halo tile in shared memory -> global D.  A thread owns an XT x YT block of
// outputs for one float4 of channels; the ((YT-1)S+KS) x ((XT-1)S+KS) input window is streamed row by row.
// SWZ: the halo tile is in the TMA's SWIZZLE_128B layout (needed where the tile doubles as a tcgen05 operand).  A quarter
// warp (8 lanes = the 8 channel vectors of one pixel) reads one whole 128-byte pixel row per LDS.128 wavefront, which is
// bank-conflict free with or without the swizzle; without it every window address is base + an immediate.
template <typename G, int KS, int S, int XT, int YT, bool WD_GLOBAL, int NWARPS, bool SWZ = true, bool WCHUNK = false>
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
                    wt[rr][kx] = WCHUNK      ? *reinterpret_cast<const float4*>(Wd_s + (rr * KS + kx) * 32 + c4 * 4)  // the chunk's tap image in shared memory
                                 : WD_GLOBAL ? ldg4(Wd_s + (rr * KS + kx) * p.hid + cbase)
                                             : *reinterpret_cast<const float4*>(Wd_s + (rr * KS + kx) * p.hid + cbase);
            }
#pragma unroll
            for (int dy = 0; dy < YT; ++dy) {
                const int ky = rr - dy * S;
                if (ky < 0 || ky >= KS) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
                    for (int dx = 0; dx < XT; ++dx) fma44p(acc[dy][dx], win[dx * S + kx], wt[ky][kx]);
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
                if (yo0 + dy < p.Ho && xo0 + dx < p.Wo) st4(o0 + dy * rstride + dx * p.hid, swish4p(acc[dy][dx]));
            }
    }
}


// ---- host side --------------------------------------------------------------------------
// fp32 NHWC tensor [B][H][W][C]: box {32 channels (zero filled past C), IW, IH, 1}, SWIZZLE_128B
inline int xd_make_map(PwTcState& st, CUtensorMap* map, const float* ptr, int B, int H, int W, int C, int IW, int IH, int swizzle = 1, int box_c = 32) {  // swizzle: 0 none, 1 SWIZZLE_128B, 2 SWIZZLE_64B (box_c <= 16)
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)IW, (cuuint32_t)IH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((PFN_encodeTiled)st.encode)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CF_ECUDA, "cuTensorMapEncodeTiled(4D %dx%dx%dx%d) failed with CUresult %d", B, H, W, C, (int)r);
    return CF_OK;
}



// Build (once per block) the per-chunk tap image of a depth-wise weight tensor Wd[k*k][hid]: [chunk][tap][32 channels], zero past hid.
inline int mbf_prepare_dw(PwTcState& st, const float* key, const float* hw, int kk, int hid) {
    if (st.dw_imgs.count(key)) return CF_OK;
    const int nch = (hid + 31) / 32;
    std::vector<float> img((size_t)nch * kk * 32, 0.f);
    for (int c = 0; c < nch; ++c)
        for (int t = 0; t < kk; ++t)
            for (int i = 0; i < 32 && c * 32 + i < hid; ++i) img[((size_t)c * kk + t) * 32 + i] = hw[(size_t)t * hid + c * 32 + i];
    float* d = nullptr;
    if (cudaMalloc((void**)&d, img.size() * 4) != cudaSuccess) return fail(CF_ECUDA, "mbf_prepare_dw: cudaMalloc failed");
    if (cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d);
        return fail(CF_ECUDA, "mbf_prepare_dw: cudaMemcpy failed");
    }
    st.dw_imgs[key] = d;
    return CF_OK;
}


constexpr int DWT_THREADS = 256;

// Output tile per item.  GEOM 0: 10 x 10, divides every map of the network (320, 160, 80, 40, 20) but fills only 25 of the
// CTA's 32 block slots (8 warps x 4 pixel groups, one 2x2 output block each); GEOM 1: 8 rows x 16 columns = 32 blocks, every
// thread busy, for the maps it divides (stride-2/4/8 stages).  Tiles that overhang the map are legal either way: TMA zero-fills
// the reads (= the reference's zero padding) and the stores are clipped.
template <int KS, int S, int GEOM>
struct DwtGeom {
    // GEOM 5: 10 rows x 20 columns for the 5x5 stride-1 layers on the 40x40 / 20x20 maps (25 blocks of 2 x 4 outputs)
    // GEOM 6: GEOM 1's 8 x 16 tile with 2 x 4 blocks (16 of the 32 block slots busy, a third fewer shared-memory loads per output): probe
    static constexpr int TH = (GEOM == 0 || GEOM == 5) ? 10 : (GEOM == 1 || GEOM == 3 || GEOM == 6) ? 8 : 16;
    static constexpr int TW = GEOM == 0 ? 10 : GEOM == 5 ? 20 : (GEOM == 1 || GEOM == 2 || GEOM == 6) ? 16 : 32;
    static constexpr int IH = (TH - 1) * S + KS, IW = (TW - 1) * S + KS;
    static constexpr int NPX = IH * IW;
    static constexpr int LO = (KS - S) / 2;
    static constexpr int XBYTES = ((NPX * 128 + 1023) / 1024) * 1024;
    // outputs per thread: 2 x 2, or 2 rows x 4 columns where a 16x16 tile then gives exactly one block per thread slot
    // (3x3 stride 1: 24 window loads + 9 tap loads per 8 outputs instead of 32 + 18)
    // 5x5 stride 1 with 2 x 4 blocks: 48 window loads + 25 tap loads per 8 outputs instead of 36 + 25 per 4 -- the compute phase of
    // these layers is bound by the shared-memory pipe (one LDS.128 wavefront per quarter warp), not by the FMA pipe
    static constexpr int XT = ((S == 1 && (GEOM == 2 || GEOM == 5)) || GEOM == 6) ? 4 : 2, YT = 2;
};

struct DwtParams {
    XdParams x;  // We unused; hid = C
    int nchunk;  // ceil(C / 32); TMA zero-fills the channels past C in the last chunk
    int nst;     // pipeline stages
    const float* wd_img;  // [chunk][tap][32] tap image (mbf_prepare_dw) or NULL: taps through L1 from Wd
    uint32_t sb;          // bytes per stage: halo tile (+ the chunk's taps when wd_img)
    int dbg;     // development only (env CF_DWT_DEBUG): 1 = skip the compute phase (TMA streaming rate of the tiling)
};

template <int KS, int S, int GEOM>
__global__ void __launch_bounds__(DWT_THREADS, 2) k_dwt(const __grid_constant__ CUtensorMap tmX, const DwtParams P) {
    using G = DwtGeom<KS, S, GEOM>;
    const XdParams& p = P.x;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int nst = P.nst;
    const uint32_t SB = P.sb;
    const uint32_t bars = base + (uint32_t)nst * SB;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c4 = lane & 7, pg = lane >> 3;
    pdl_trigger();
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        for (int i = 0; i < nst; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();

    // item = ((b * tiles_y + ty) * tiles_x + tx) * nchunk + chunk : chunks of one tile are adjacent.
    // Thread 0 decodes an item once, when it issues the item's TMA, and leaves (chunk, tx, ty, b) beside the barriers; the
    // other 255 threads read it after the barrier wait instead of repeating three integer divisions each (the SASS of the
    // previous version spent ~35 % of its instructions per item on that decode).
    int4* info = reinterpret_cast<int4*>(sm + (size_t)nst * SB + 64);  // [nst], after the 8 barriers
    auto issue = [&](long long item, int stage) {  // one elected lane of warp 0 only
        const int ch = (int)(item % P.nchunk);
        int t = (int)(item / P.nchunk);
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y, b = t / p.tiles_y;
        info[stage] = make_int4(ch, tx, ty, b);
        // (release: orders the info store before the phase flip)
        mbar_expect_tx(bars + 8 * stage, (uint32_t)G::NPX * 128u + (P.wd_img ? (uint32_t)(KS * KS * 128) : 0u));
        tma_load_4d(base + stage * SB, &tmX, ch * 32, tx * G::TW * S - G::LO, ty * G::TH * S - G::LO, b, bars + 8 * stage);
        // the chunk's KS*KS x 32 taps ride along: read with LDS in the compute phase (through L1 they were long-scoreboard stalls
        // in front of the FFMAs: the same change took 12 % off the fused layer1.1 kernel)
        if (P.wd_img) bulk_load(base + stage * SB + G::XBYTES, P.wd_img + (size_t)ch * (KS * KS * 32), (uint32_t)(KS * KS * 128), bars + 8 * stage);
    };
    if (warp == 0) {  // convergent: one elected lane issues (see elect_one in k_pw_tc.cuh)
        if (elect_one())
            for (int i = 0; i < nst; ++i) {
                const long long item = (long long)blockIdx.x + (long long)i * gridDim.x;
                if (item < p.n_items) issue(item, i);
            }
        __syncwarp();
    }
    int stage = 0;
    uint32_t phase = 0;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        mbar_wait(bars + 8 * stage, phase);
        const int4 inf = info[stage];
        const int ch = inf.x, tx = inf.y, ty = inf.z, b = inf.w;
        const int cbase = ch * 32 + c4 * 4;
        if (!(P.dbg & 1)) {
            if (P.wd_img)
                xd_dw_phase_g<G, KS, S, G::XT, G::YT, false, DWT_THREADS / 32, false, true>(
                    sm + stage * SB, reinterpret_cast<const float*>(sm + stage * SB + G::XBYTES), p, warp, pg, c4, cbase, cbase < p.hid, b, ty, tx);
            else
                xd_dw_phase_g<G, KS, S, G::XT, G::YT, true, DWT_THREADS / 32, false>(sm + stage * SB, p.Wd, p, warp, pg, c4, cbase, cbase < p.hid, b, ty, tx);
        }
        __syncthreads();  // every thread is done with this stage (and has read its info): refill it
        if (warp == 0) {
            const long long nxt = item + (long long)nst * gridDim.x;
            if (nxt < p.n_items && elect_one()) issue(nxt, stage);
            __syncwarp();
        }
        if (++stage == nst) stage = 0, phase ^= 1u;
    }
}

struct DwtLaunch {
    CUtensorMap tmX;
    DwtParams p;
    int ks = 3, s = 1, geom = 0, grid = 0;
    size_t smem = 0;
};

template <int KS, int S, int GEOM>
inline cudaError_t dwt_launch_t(const DwtLaunch& dl, cudaStream_t st) {
    cudaError_t e = smem_optin((const void*)k_dwt<KS, S, GEOM>, (int)(TC_SMEM_MAX));
    if (e != cudaSuccess) return e;
    return launch_pdl(k_dwt<KS, S, GEOM>, dim3(dl.grid), dim3(DWT_THREADS), dl.smem, st, dl.tmX, dl.p);
}

inline cudaError_t dwt_launch(const DwtLaunch& dl, cudaStream_t st) {
#define CF_DWT_CASE(K_, S_)                                                   \
    if (dl.ks == K_ && dl.s == S_) {                                          \
        switch (dl.geom) {                                                    \
            case 1: return dwt_launch_t<K_, S_, 1>(dl, st);                   \
            case 2: return dwt_launch_t<K_, S_, 2>(dl, st);                   \
            case 3: return dwt_launch_t<K_, S_, 3>(dl, st);                   \
            case 4: return dwt_launch_t<K_, S_, 4>(dl, st);                   \
            case 5: return dwt_launch_t<K_, S_, 5>(dl, st);                   \
            case 6: return dwt_launch_t<K_, S_, 6>(dl, st);                   \
            default: return dwt_launch_t<K_, S_, 0>(dl, st);                  \
        }                                                                     \
    }
    CF_DWT_CASE(3, 1) CF_DWT_CASE(3, 2) CF_DWT_CASE(5, 1) CF_DWT_CASE(5, 2)
#undef CF_DWT_CASE
    return cudaErrorInvalidValue;
}

template <int KS, int S, int GEOM>
inline void dwt_geom(int* th, int* tw, int* ih, int* iw, int* xb) {
    using G = DwtGeom<KS, S, GEOM>;
    *th = G::TH, *tw = G::TW, *ih = G::IH, *iw = G::IW, *xb = G::XBYTES;
}

inline bool dwt_supported(int C) { return C % 4 == 0 && C >= 16; }

inline int dwt_plan(PwTcState& st, int ks, int s, const float* X, const float* Wd, float* D, int B, int Hi, int Wi, int C, DwtLaunch* dl) {
    if (!dwt_supported(C)) return fail(CF_EINVAL, "dwt_plan: C=%d is not a multiple of 4", C);
    const int Ho = Hi / s, Wo = Wi / s;
    // Tile geometry per layer, from sweeps on the device at batch 32 @ 640x640 (profiles/r2_tuning.md; us per launch,
    // 10x10 -> chosen): the 3x3 stride-1 layers take 16x16 tiles with two CTAs x two stages per SM (32ch 174 -> 141,
    // 144ch 215 -> 167), the 5x5 layers 8x16 (stride 1, 192ch: 118 -> 100; stride 2, 144ch: 138 -> 122); the 3x3 stride-2
    // layer is TMA/HBM-bound with any tile (209 of its 252 us are the bare tile stream) and the stride-16/32 maps
    // (40x40, 20x20) are not divisible: both keep 10x10.
    const int gth[7] = {10, 8, 16, 8, 16, 10, 8}, gtw[7] = {10, 16, 16, 32, 32, 20, 16};
    auto divides = [&](int g) {  // ... and one halo tile fits shared memory
        const int hh = (gth[g] - 1) * s + ks, hw = (gtw[g] - 1) * s + ks;
        return Ho % gth[g] == 0 && Wo % gtw[g] == 0 && (size_t)((hh * hw * 128 + 1023) / 1024 * 1024) + 2048 <= (size_t)TC_SMEM_MAX;
    };
    // Round 2, late: 2 x 4 output blocks for the 5x5 stride-1 layers as well (their compute phase is bound by the shared-memory
    // pipe: 48 + 25 LDS.128 per 8 outputs instead of 36 + 25 per 4) -- 16x16 tiles on the 80x80 map (192ch: 92 -> 73 us), 10x20 on the
    // 40x40 / 20x20 maps (384ch 58 -> 54, 576ch 82 -> 78, 960ch 39 -> 37).  CF_DWT_W4: bit 0 / bit 1 enable those two (default 7: all three),
    // bit 2 = 10x20 tiles of 2 x 4 blocks for the 3x3 stride-1 layers on the 40x40 / 20x20 maps too (384ch 39 -> 35, 960ch 27 -> 25).
    int w4 = 7;
    if (const char* ev = getenv("CF_DWT_W4")) w4 = atoi(ev);
    int geom = (s == 1 && ks == 3 && divides(2)) ? 2 : (ks == 5 && divides(1)) ? 1 : 0;
    if (ks == 5 && s == 1 && (w4 & 1) && divides(2)) geom = 2;
    else if (ks == 5 && s == 1 && (w4 & 2) && divides(5)) geom = 5;
    else if (ks == 3 && s == 1 && (w4 & 4) && !divides(2) && divides(5)) geom = 5;
    else if (ks == 5 && s == 2 && (w4 & 8) && divides(6)) geom = 6;  // probe
    if (const char* ev = getenv("CF_DWT_GEOM")) {  // development probe: geometry index wherever it divides the map
        const int g = atoi(ev);
        if (g == 0) geom = 0;
        else if (g >= 1 && g <= 6) geom = divides(g) ? g : geom;
    }
    int th = 0, tw = 0, ih = 0, iw = 0, xb = 0;
#define CF_DWT_GEOM_CASE(K_, S_)                                              \
    if (ks == K_ && s == S_) {                                                \
        switch (geom) {                                                       \
            case 1: dwt_geom<K_, S_, 1>(&th, &tw, &ih, &iw, &xb); break;      \
            case 2: dwt_geom<K_, S_, 2>(&th, &tw, &ih, &iw, &xb); break;      \
            case 3: dwt_geom<K_, S_, 3>(&th, &tw, &ih, &iw, &xb); break;      \
            case 4: dwt_geom<K_, S_, 4>(&th, &tw, &ih, &iw, &xb); break;      \
            case 5: dwt_geom<K_, S_, 5>(&th, &tw, &ih, &iw, &xb); break;      \
            case 6: dwt_geom<K_, S_, 6>(&th, &tw, &ih, &iw, &xb); break;      \
            default: dwt_geom<K_, S_, 0>(&th, &tw, &ih, &iw, &xb); break;     \
        }                                                                     \
    }
    CF_DWT_GEOM_CASE(3, 1) CF_DWT_GEOM_CASE(3, 2) CF_DWT_GEOM_CASE(5, 1) CF_DWT_GEOM_CASE(5, 2)
#undef CF_DWT_GEOM_CASE
    if (th == 0) return fail(CF_EINVAL, "dwt_plan: unsupported kernel %dx%d stride %d", ks, ks, s);
    if (iw > 256 || ih > 256 || xb + 2048 > TC_SMEM_MAX) {  // TMA box limit / one stage must fit
        return fail(CF_EINVAL, "dwt_plan: tile geometry %d does not fit (halo %dx%d)", geom, ih, iw);
    }
    int rc = xd_make_map(st, &dl->tmX, X, B, Hi, Wi, C, iw, ih, /*swizzle=*/false);
    if (rc) return rc;
    XdParams& p = dl->p.x;
    p.We = nullptr;
    p.Wd = Wd;
    p.D = D;
    p.B = B;
    p.Hi = Hi;
    p.Wi = Wi;
    p.Ho = Hi / s;
    p.Wo = Wi / s;
    p.hid = C;
    p.tiles_x = (p.Wo + tw - 1) / tw;
    p.tiles_y = (p.Ho + th - 1) / th;
    dl->p.nchunk = (C + 31) / 32;
    const long long items = (long long)B * p.tiles_x * p.tiles_y * dl->p.nchunk;
    if (items > 0x7fffffffLL) return fail(CF_EINVAL, "dwt_plan: too many tiles");
    p.n_items = (int)items;
    {   // taps in shared memory when the block's tap image was prepared (cf_create does it for every block)
        auto it = st.dw_imgs.find(Wd);
        dl->p.wd_img = it == st.dw_imgs.end() ? nullptr : it->second;
        if (const char* ev = getenv("CF_DWT_WDS")) {
            if (atoi(ev) == 0) dl->p.wd_img = nullptr;
        }
        if (dl->p.wd_img) xb += (ks * ks * 128 + 1023) / 1024 * 1024;
    }
    dl->p.sb = (uint32_t)xb;
    // two CTAs per SM when at least three stages fit in half of the shared memory, else one CTA with all of it
    const int per_cta2 = (TC_SMEM_MAX / 2 - 2048) / xb;
    int ctas_per_sm = 2, nst = per_cta2;
    int min2 = (geom == 2 || geom == 5) ? 2 : 3;  // 10x20 tiles of a 5x5 layer: two CTAs x two stages beat one CTA x four (576ch: 78 -> 70 us)
    if (const char* ev = getenv("CF_DWT_MIN2")) min2 = atoi(ev);  // development probe: fewest stages worth two CTAs per SM
    if (nst < min2) ctas_per_sm = 1, nst = (TC_SMEM_MAX - 2048) / xb;
    if (const char* ev = getenv("CF_DWT_CTAS")) {  // development probe (tools/step_times.py)
        if (atoi(ev) == 1) ctas_per_sm = 1, nst = (TC_SMEM_MAX - 2048) / xb;
    }
    if (nst > 8) nst = 8;
    if (const char* ev = getenv("CF_DWT_NST")) {
        if (atoi(ev) >= 1 && atoi(ev) < nst) nst = atoi(ev);
    }
    if (nst < 1) return fail(CF_EINVAL, "dwt_plan: tile does not fit shared memory");
    dl->p.nst = nst;
    dl->p.dbg = 0;
    if (const char* ev = getenv("CF_DWT_DEBUG")) dl->p.dbg = atoi(ev);
    dl->smem = (size_t)nst * xb + 64 + 8 * 16 + 1024;  // stages | 8 barriers | 8 item records | alignment slack
    dl->ks = ks;
    dl->s = s;
    dl->geom = geom;
    const int want = ctas_per_sm * st.sms;
    dl->grid = p.n_items < want ? p.n_items : want;
    return CF_OK;
}

}  // namespace cf
