// k_stem_tc: the stem convolution on the tensor cores.
//
//     ZeroPad2d(0,1,0,1) + conv3x3 s2 3->32 + Swish   (model/centernet.py:224, :58-70), fused u8 normalisation
//     (centerface.py:32-34) through the same 3 x 256 table as k_stem.
//
// The FFMA stem (k_stem, k_conv.cuh) is instruction-issue bound: 864 FFMAs + ~900 other instructions per output
// pixel (SASS), 222 us per 32-image batch against an HBM floor of 70 us.  But the stem IS a dense contraction,
//     out[pixel][32] = sum_{k = (ky*3+kx)*3+c < 27}  x[pixel, k] . w[k][32],
// i.e. a GEMM with K = 27 (padded to one 32-wide K block) and N = 32, so it takes the role-free tcgen05 kernel of the
// narrow point-wise layers (k_pwn) with a different A producer: instead of splitting a TMA box, every thread gathers 16 of
// its pixel's 27 inputs straight from the image (u8 -> table, or fp32 NCHW), splits them into tf32 hi + lo and stores
// them into TMEM (lane = pixel row of the 128-pixel tile, column = k); one thread issues
//     A_hi.[W_hi|W_lo] -> main | correction,   A_lo.W_hi -> correction
// with the weight image (8 KB, built by tc_prepare_layer) resident in shared memory; all threads drain the
// accumulator pair, apply Swish and store their half row.  The next tile's bytes are fetched before the current
// tile's MMA/drain phase so that the global-load latency overlaps it; three CTAs (128 TMEM columns each) share an SM.
//
// Arithmetic: 3xTF32 (error ~2^-21 per product, fp32 accumulation in TMEM), the same as every point-wise convolution
// of this engine; the fp32-FFMA validation engine (CF_PW_SIMT) keeps k_stem.
#pragma once
#include "k_pwn.cuh"

namespace cf {

constexpr int STC_THREADS = 256;
constexpr int STC_NC = 32;                      // output channels = one column chunk
constexpr uint32_t STC_B_BYTES = STC_NC * 256;  // [hi 32 x 128 B | lo 32 x 128 B]
constexpr uint32_t STC_OFF_LUT = STC_B_BYTES;   // 768 floats
constexpr uint32_t STC_OFF_BARS = STC_OFF_LUT + 768 * 4;
constexpr size_t STC_SMEM = STC_OFF_BARS + 64 + 1024;  // + alignment slack

struct StcParams {
    const void* in;     // FMT 1: u8 [B,H,W,3]; FMT 0: fp32 [B,3,H,W]
    const float* bimg;  // tf32 hi|lo image of stem.w [27][32] (K padded to 32)
    const float* lut;   // [3][256]
    float* out;         // [B,H/2,W/2,32]
    int B, H, W, n_tiles;
    long long n_pix;
};

// raw[i] for k = K0 + i: the byte (FMT 1) or the fp32 bit pattern (FMT 0) of input element k of pixel (b, yo, xo);
// bit i of the returned mask is set where the element exists (k < 27 and inside the image; the pad is bottom/right only).
template <int FMT, int K0>
__device__ __forceinline__ uint32_t stc_gather(const StcParams& p, bool valid, int b, int yo, int xo, uint32_t (&raw)[16]) {
    uint32_t mask = 0;
    const bool last_y = (2 * yo + 2 >= p.H), last_x = (2 * xo + 2 >= p.W);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = K0 + i;
        raw[i] = 0;
        if (k >= 27) continue;
        const int ky = k / 9, r = k - ky * 9, kx = r / 3, c = r - kx * 3;
        const bool ok = valid && !(ky == 2 && last_y) && !(kx == 2 && last_x);
        if (ok) {
            if (FMT == 1) {
                const uint8_t* q = (const uint8_t*)p.in + ((size_t)(b * p.H + 2 * yo + ky) * p.W + 2 * xo + kx) * 3 + c;
                raw[i] = __ldg(q);
            } else {
                const float* q = (const float*)p.in + ((size_t)(b * 3 + c) * p.H + 2 * yo + ky) * p.W + 2 * xo + kx;
                raw[i] = __float_as_uint(__ldg(q));
            }
            mask |= 1u << i;
        }
    }
    return mask;
}

// raw -> normalised fp32 -> tf32 hi + exact remainder
template <int FMT, int K0>
__device__ __forceinline__ void stc_split(const uint32_t (&raw)[16], uint32_t mask, const float* lut_s, float (&hi)[16], float (&lo)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        constexpr int c0 = K0 % 3;
        const int c = (c0 + i) % 3;  // k = (ky*3+kx)*3 + c
        float v = 0.f;
        if (mask & (1u << i)) v = FMT == 1 ? lut_s[c * 256 + raw[i]] : __uint_as_float(raw[i]);
        hi[i] = tf32_hi(v);
        lo[i] = v - hi[i];
    }
}

template <int FMT>
__global__ void __launch_bounds__(STC_THREADS, 3) k_stem_tc(const StcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bsm = base;
    float* lut_s = reinterpret_cast<float*>(sm + STC_OFF_LUT);
    const uint32_t bars = base + STC_OFF_BARS;  // B full | accumulator ready
    const uint32_t bar_b = bars, bar_acc = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + STC_OFF_BARS + 40);
    constexpr uint32_t kACol = 64;  // TMEM (128 columns, three CTAs per SM): accumulator pair in [0, 64), A hi | lo in [64, 128)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter; K elements [16 half, 16 half + 16) / output columns likewise
    pdl_trigger();

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (FMT == 1)
        for (int i = tid; i < 768; i += STC_THREADS) lut_s[i] = __ldg(p.lut + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_tiles = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid == 0 && my_tiles > 0) {
        mbar_expect_tx(bar_b, STC_B_BYTES);
        bulk_load(bsm, p.bimg, STC_B_BYTES, bar_b);
    }
    pdl_wait();  // the image may come from the resize kernel; `out` may still be read by the previous forward

    const uint32_t idesc = umma_idesc_tf32(STC_NC), idesc2 = umma_idesc_tf32(2 * STC_NC);
    const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)STC_NC;
    const int row = q * 32 + lane;
    const int Ho = p.H >> 1, Wo = p.W >> 1;

    uint32_t raw[16];
    uint32_t mask = 0;
    auto gather = [&](int tile) {
        const long long pix = (long long)tile * TC_BM + row;
        const bool valid = pix < p.n_pix;
        const int xo = (int)(pix % Wo);
        const int yo = (int)((pix / Wo) % Ho);
        const int b = (int)(pix / ((long long)Wo * Ho));
        mask = half == 0 ? stc_gather<FMT, 0>(p, valid, b, yo, xo, raw) : stc_gather<FMT, 16>(p, valid, b, yo, xo, raw);
    };
    if (my_tiles > 0) gather((int)blockIdx.x);

    for (int j = 0; j < my_tiles; ++j) {
        const int tile = (int)blockIdx.x + j * (int)gridDim.x;
        // ---- this thread's 16 inputs -> normalised fp32 -> tf32 hi / lo -> TMEM ----
        float hi[16], lo[16];
        if (half == 0) stc_split<FMT, 0>(raw, mask, lut_s, hi, lo);
        else stc_split<FMT, 16>(raw, mask, lut_s, hi, lo);
        if (j + 1 < my_tiles) gather(tile + (int)gridDim.x);  // in flight during this tile's MMA + drain
        // the previous tile's MMAs have read the A block: every thread saw its accumulator barrier before the tile-end sync
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kACol + (uint32_t)half * 16u;
        tc_fence_after();
        tmem_st16(ta, hi);
        tmem_st16(ta + 32u, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();  // the A block is in TMEM
        if (tid == 0) {
            if (j == 0) mbar_wait(bar_b, 0);
            tc_fence_after();
            const uint64_t b_hi = umma_desc(bsm);
            const uint32_t a_hi = tmem_base + kACol, a_lo = a_hi + 32u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // K = 27 -> four 8-wide steps (elements 27..31 are zero on both sides)
                const uint64_t ko = (uint64_t)(k * 2);
                umma_tf32_ts(d_main, a_hi + 8u * k, b_hi + ko, idesc2, k > 0 ? 1u : 0u);  // main += hi.hi ; corr += hi.lo
                umma_tf32_ts(d_corr, a_lo + 8u * k, b_hi + ko, idesc, 1u);                 // corr += lo.hi
            }
            umma_commit(bar_acc);
        }
        // ---- drain main + correction, Swish, store this thread's 16 channels of its pixel ----
        mbar_wait(bar_acc, (uint32_t)j & 1u);
        tc_fence_after();
        const long long pix = (long long)tile * TC_BM + row;
        {
            float v[16], c[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16);
            tmem_ld16(taddr, v);
            tmem_ld16(taddr + (uint32_t)STC_NC, c);
            tmem_ld_wait();
            if (pix < p.n_pix) {
                float* o = p.out + (size_t)pix * 32 + half * 16;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    st4(o + 4 * g, swish4(make_float4(v[4 * g] + c[4 * g], v[4 * g + 1] + c[4 * g + 1], v[4 * g + 2] + c[4 * g + 2],
                                                      v[4 * g + 3] + c[4 * g + 3])));
            }
        }
        tc_fence_before();
        __syncthreads();  // every accumulator read is done before the next tile's first MMA overwrites it
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

// ---- second generation: warp-specialised, pipelined ---------------------------------------------------------------
// The role-free kernel above runs gather -> split -> TMEM store -> MMA -> drain -> store as ONE serial chain per
// 128-pixel tile (255 us per batch, slower than the FFMA stem).  Here the chain is a pipeline, as in k_pw_tc:
//   warps 4-11   two producer sets (tiles alternate between them): thread = one pixel, gathers its 27 inputs (prefetched
//                one own-tile ahead), normalises, splits into tf32 hi/lo and tcgen05.st's them into a four-slot TMEM ring
//   warp 1       one lane issues  A_hi.[W_hi|W_lo]  and  A_lo.W_hi  per 8-wide K step into a four-stage accumulator ring
//   warps 12-19  two drain groups (accumulator stages alternate): tcgen05.ld main + correction, Swish, 128-byte row store
//   warp 0       loads the 8 KB weight image once; warp 2 owns the TMEM allocation (512 columns, one CTA per SM)
constexpr int STC2_THREADS = 640;
// table of 768 + entry 768 = 0.0f (taps that fall in the zero padding); the barriers start on their own 128-byte line (the table is
// filled with generic stores after mbarrier.init: compute-sanitizer synccheck tracks barrier validity per line and reported the
// first wait on a barrier sharing the table's last line as "missing init", profiles/r2_sanitizer.md)
constexpr uint32_t STC2_OFF_BARS = (STC_OFF_LUT + 776 * 4 + 127u) & ~127u;
constexpr uint32_t STC2_OFF_STG = STC2_OFF_BARS + 256;      // 8 drain warps x 4 KB store-transpose tiles
constexpr size_t STC2_SMEM = STC2_OFF_STG + 8 * 4096 + 1024;

// raw[k]: FMT 1: the tap's byte, or kPad + k%3 ... see below; FMT 0: the fp32 bit pattern (0 = 0.0f for padding).
// Every load is unconditional and addressed as (one of three row pointers) + immediate: a tap that falls into the zero padding
// (bottom row / right column of ZeroPad2d(0,1,0,1)) or belongs to a row past the last pixel reads a neighbouring in-image byte
// instead and is then replaced by the index of the table's 0.0f entry -- the previous version spent ~10 instructions per tap
// on per-tap predicates and 64-bit address arithmetic.
// FMT 1 encoding: raw[k] = byte (table index = c * 256 + byte, c = k % 3 known at compile time) or 768 - c * 256 (-> entry 768).
template <int FMT>
__device__ __forceinline__ void stc2_gather(const StcParams& p, bool valid, int b, int yo, int xo, uint32_t (&raw)[27]) {
    const bool last_y = (2 * yo + 2 >= p.H), last_x = (2 * xo + 2 >= p.W);
    if (FMT == 1) {
        const int rs = p.W * 3;
        const uint8_t* r0 = (const uint8_t*)p.in + ((size_t)(b * p.H + 2 * yo) * p.W + 2 * xo) * 3;
        const uint8_t* r1 = r0 + rs;
        const uint8_t* r2 = last_y ? r1 : r1 + rs;  // a safe address; the value is discarded below
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const int ky = k / 9, r = k - ky * 9, kx = r / 3, c = r - kx * 3;
            const uint8_t* row = ky == 0 ? r0 : ky == 1 ? r1 : r2;
            const int off = kx == 2 ? (last_x ? 3 + c : 6 + c) : kx * 3 + c;  // kx = 2 at the right edge: re-read column 1
            uint32_t v = __ldg(row + off);
            const uint32_t pad = (uint32_t)(768 - c * 256);
            if (ky == 2) v = last_y ? pad : v;
            if (kx == 2) v = last_x ? pad : v;
            raw[k] = valid ? v : pad;
        }
    } else {
        const size_t plane = (size_t)p.H * p.W;
        const float* q0 = (const float*)p.in + (size_t)b * 3 * plane + (size_t)(2 * yo) * p.W + 2 * xo;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const int ky = k / 9, r = k - ky * 9, kx = r / 3, c = r - kx * 3;
            const int kyy = (ky == 2 && last_y) ? 1 : ky, kxx = (kx == 2 && last_x) ? 1 : kx;
            uint32_t v = __float_as_uint(__ldg(q0 + c * plane + (size_t)kyy * p.W + kxx));
            if (ky == 2) v = last_y ? 0u : v;
            if (kx == 2) v = last_x ? 0u : v;
            raw[k] = valid ? v : 0u;
        }
    }
}

template <int FMT>
__global__ void __launch_bounds__(STC2_THREADS, 1) k_stem_tc2(const StcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bsm = base;
    float* lut_s = reinterpret_cast<float*>(sm + STC_OFF_LUT);
    const uint32_t bars = base + STC2_OFF_BARS;
    const uint32_t bar_aready = bars;            // [4] producers -> MMA   (4 warps arrive)
    const uint32_t bar_aempty = bars + 32;       // [4] MMA -> producers   (tcgen05.commit)
    const uint32_t bar_tfull = bars + 64;        // [4] MMA -> drain       (tcgen05.commit)
    const uint32_t bar_tempty = bars + 96;       // [4] drain -> MMA       (4 warps arrive)
    const uint32_t bar_b = bars + 128;           // weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + STC2_OFF_BARS + 144);
    constexpr uint32_t kACol = 256;              // accumulator pairs: 4 x 64 columns in [0,256); A ring: 4 x (32 hi + 32 lo) above

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();
    if (tid == 0) {
        mbar_init(bar_b, 1);
        for (int a = 0; a < 4; ++a) {
            mbar_init(bar_aready + 8 * a, 4);
            mbar_init(bar_aempty + 8 * a, 1);
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (FMT == 1)
        for (int i = tid; i < 769; i += STC2_THREADS) lut_s[i] = i < 768 ? __ldg(p.lut + i) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = p.n_tiles;
    const int Ho = p.H >> 1, Wo = p.W >> 1;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bar_b, STC_B_BYTES);
            bulk_load(bsm, p.bimg, STC_B_BYTES, bar_b);
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp walks the loop, one elected lane issues =================
        {
            const uint32_t idesc = umma_idesc_tf32(STC_NC), idesc2 = umma_idesc_tf32(2 * STC_NC);
            mbar_wait(bar_b, 0);
            const uint64_t b_hi = umma_desc(bsm);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t s4 = it & 3u, ph = (it >> 2) & 1u;
                mbar_wait(bar_tempty + 8 * s4, ph ^ 1u);
                mbar_wait(bar_aready + 8 * s4, ph);
                tc_fence_after();
                const uint32_t d_main = tmem_base + s4 * 64u, d_corr = d_main + (uint32_t)STC_NC;
                const uint32_t a_hi = tmem_base + kACol + s4 * 64u, a_lo = a_hi + 32u;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // K = 27 -> four 8-wide steps (elements 27..31 are zero on both sides)
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_tf32_ts(d_main, a_hi + 8u * k, b_hi + ko, idesc2, k > 0 ? 1u : 0u);  // main += hi.hi ; corr += hi.lo
                        umma_tf32_ts(d_corr, a_lo + 8u * k, b_hi + ko, idesc, 1u);                 // corr += lo.hi
                    }
                    umma_commit(bar_aempty + 8 * s4);
                    umma_commit(bar_tfull + 8 * s4);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ================= producers: set 0 = warps 4-7, set 1 = warps 8-11 =================
        pdl_wait();  // the image may come from the resize kernel
        const int set = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t raw[27];
        auto gather = [&](int tile) {  // 32-bit index math (stc_plan checks n_pix < 2^31): 64-bit divisions cost ~100 instructions each
            const unsigned pix_raw = (unsigned)tile * TC_BM + (unsigned)row;
            const bool valid = pix_raw < (unsigned)p.n_pix;
            const unsigned pix = valid ? pix_raw : (unsigned)p.n_pix - 1u;  // rows past the end re-read the last pixel (discarded)
            const unsigned t = pix / (unsigned)Wo;
            const int xo = (int)(pix - t * (unsigned)Wo);
            const unsigned b = t / (unsigned)Ho;
            const int yo = (int)(t - b * (unsigned)Ho);
            stc2_gather<FMT>(p, valid, (int)b, yo, xo, raw);
        };
        const int first = (int)blockIdx.x + set * (int)gridDim.x;
        if (first < n_tiles) gather(first);
        uint32_t it = (uint32_t)set;
        for (int tile = first; tile < n_tiles; tile += 2 * (int)gridDim.x, it += 2) {
            float hi[32], lo[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                float v = 0.f;
                if (k < 27) v = FMT == 1 ? lut_s[(k % 3) * 256 + raw[k]] : __uint_as_float(raw[k]);
                hi[k] = tf32_hi(v);
                lo[k] = v - hi[k];
            }
            if (tile + 2 * (int)gridDim.x < n_tiles) gather(tile + 2 * (int)gridDim.x);  // in flight while this tile is handed over
            const uint32_t s4 = it & 3u, ph = (it >> 2) & 1u;
            mbar_wait(bar_aempty + 8 * s4, ph ^ 1u);
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kACol + s4 * 64u;
            tmem_st32(ta, hi);
            tmem_st32(ta + 32u, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready + 8 * s4);
        }
    } else if (warp >= 12) {
        // ================= drain: group g = accumulator stages g, g+2 =================
        pdl_wait();  // `out` may still be read by the previous forward
        const int g = (warp - 12) >> 2, q = warp & 3;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            if ((int)(it & 1u) != g) continue;
            const uint32_t s4 = it & 3u, ph = (it >> 2) & 1u;
            mbar_wait(bar_tfull + 8 * s4, ph);
            tc_fence_after();
            float v[32], c[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s4 * 64u;
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + (uint32_t)STC_NC, c);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * s4);  // the accumulator pair is in registers: hand it back
            // transpose through a per-warp swizzled tile so that one STG.128 writes 4 whole pixels (512 contiguous bytes)
            uint8_t* stg = sm + STC2_OFF_STG + (warp - 12) * 4096;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                    swish4p(make_float4(v[4 * j] + c[4 * j], v[4 * j + 1] + c[4 * j + 1], v[4 * j + 2] + c[4 * j + 2], v[4 * j + 3] + c[4 * j + 3]));
            __syncwarp();
            const long long pix0 = (long long)tile * TC_BM + q * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + (lane >> 3), cc = lane & 7;
                const float4 x = *reinterpret_cast<const float4*>(stg + r * 128 + ((cc ^ (r & 7)) << 4));
                if (pix0 + r < p.n_pix) st4(p.out + (size_t)(pix0 + r) * 32 + cc * 4, x);
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- third generation (u8 input, W % 4 == 0): more warps, fewer instructions per pixel ------------------------------------
// k_stem_tc2 runs at IPC 1.3 of 4 with MUFU, HBM and the tensor pipe all under 45 %: its 20 warps are serial chains (a lone warp
// retires an instruction every ~10 cycles), two producer sets x ~250 instructions and two drain groups x ~250 instructions per
// 128-pixel tile.  Here: 32 warps per SM --
//   warps 4-19   FOUR producer sets, one per TMEM A slot.  A pixel's 3 x 9 input bytes come as three aligned 12-byte windows
//                (9 LDG.32 instead of 27 LDG.U8); the table holds (tf32 hi, remainder) PAIRS, one LDS.64 per tap instead of a load
//                and two ALU operations; taps in the zero padding select the table's zero page through 9 precomputed bases
//                instead of a select per tap.  ~150 instructions per pixel.
//   warps 20-31  THREE drain groups over a three-stage accumulator ring (stage = group), two 16-column rounds each: 48 live
//                registers, which is what the 64-register cap of a 1 024-thread CTA allows.
//   warp 0       weight image + MMA issue (8 MMAs per tile as before, so that u8 and fp32 inputs stay bit-identical);
//                warp 1 owns the TMEM allocation.  (Two issuers over a six-stage single-accumulator ring were tried: same time.)
constexpr int STC3_THREADS = 1024;
// 1 024 entries x 16 copies x (hi, lo): [c][256] for c < 3, page 3 = zeros.  One copy per lane of a half warp: an LDS.64 of 16
// lanes then touches every bank exactly once whatever the bytes are (uniformly random bytes -- the benchmark's synthetic images --
// cost a single table ~3.5 wavefronts per load: 155 us against 129 us on image-like input)
constexpr uint32_t STC3_OFF_LUT = STC_B_BYTES;
constexpr uint32_t STC3_OFF_BARS = STC3_OFF_LUT + 1024 * 128;
constexpr uint32_t STC3_OFF_STG = STC3_OFF_BARS + 256;         // 12 drain warps x 4 KB store-transpose tiles
constexpr size_t STC3_SMEM = STC3_OFF_STG + 12 * 4096 + 1024;

__global__ void __launch_bounds__(STC3_THREADS, 1) k_stem_tc3(const StcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bsm = base;
    float2* lut2 = reinterpret_cast<float2*>(sm + STC3_OFF_LUT);
    const uint32_t bars = base + STC3_OFF_BARS;
    const uint32_t bar_aready = bars;            // [4] producer set -> MMA   (4 warps arrive)
    const uint32_t bar_aempty = bars + 32;       // [4] MMA -> producer set   (tcgen05.commit)
    const uint32_t bar_tfull = bars + 64;        // [3] MMA -> drain group    (tcgen05.commit)
    const uint32_t bar_tempty = bars + 96;       // [3] drain group -> MMA    (4 warps arrive)
    const uint32_t bar_b = bars + 128;           // weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + STC3_OFF_BARS + 144);
    constexpr uint32_t kACol = 192;              // accumulator pairs: 3 x 64 columns in [0,192); A ring: 4 x (32 hi + 32 lo) above

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();
    if (tid == 0) {
        mbar_init(bar_b, 1);
        for (int a = 0; a < 4; ++a) mbar_init(bar_aready + 8 * a, 4), mbar_init(bar_aempty + 8 * a, 1);
        for (int a = 0; a < 3; ++a) mbar_init(bar_tfull + 8 * a, 1), mbar_init(bar_tempty + 8 * a, 4);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 1024 * 16; i += STC3_THREADS) {  // consecutive threads fill consecutive slots (entry-major stores were 32-way conflicts: 8 us)
        const int e = i >> 4;
        const float v = e < 768 ? __ldg(p.lut + e) : 0.f;
        const float h = tf32_hi(v);
        lut2[i] = make_float2(h, v - h);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = p.n_tiles;
    const int Ho = p.H >> 1, Wo = p.W >> 1;

    if (warp == 0) {
        // ================= weights + MMA issuer =================
        if (lane == 0) {
            mbar_expect_tx(bar_b, STC_B_BYTES);
            bulk_load(bsm, p.bimg, STC_B_BYTES, bar_b);
        }
        __syncwarp();
        const uint32_t idesc = umma_idesc_tf32(STC_NC), idesc2 = umma_idesc_tf32(2 * STC_NC);
        mbar_wait(bar_b, 0);
        const uint64_t b_hi = umma_desc(bsm);
        uint32_t sa = 0, pa = 0, st = 0, pt = 0;  // A slot / accumulator stage cursors with their phase parities
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(bar_tempty + 8 * st, pt ^ 1u);
            mbar_wait(bar_aready + 8 * sa, pa);
            tc_fence_after();
            const uint32_t d_main = tmem_base + st * 64u, d_corr = d_main + (uint32_t)STC_NC;
            const uint32_t a_hi = tmem_base + kACol + sa * 64u, a_lo = a_hi + 32u;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // K = 27 -> four 8-wide steps (elements 27..31 are zero on both sides)
                    const uint64_t ko = (uint64_t)(k * 2);
                    umma_tf32_ts(d_main, a_hi + 8u * k, b_hi + ko, idesc2, k > 0 ? 1u : 0u);  // main += hi.hi ; corr += hi.lo
                    umma_tf32_ts(d_corr, a_lo + 8u * k, b_hi + ko, idesc, 1u);                 // corr += lo.hi
                }
                umma_commit(bar_aempty + 8 * sa);
                umma_commit(bar_tfull + 8 * st);
            }
            __syncwarp();
            if (++sa == 4) sa = 0, pa ^= 1u;
            if (++st == 3) st = 0, pt ^= 1u;
        }
    } else if (warp >= 4 && warp < 20) {
        // ================= producers: set = A slot =================
        pdl_wait();  // the image may come from the resize kernel
        const int set = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kACol + (uint32_t)set * 64u;
        uint32_t w[3][3];   // per image row: the aligned 12 bytes that hold the pixel's 9
        uint32_t sh = 0;    // bit offset of the pixel's first byte inside the window (0..24, + 32 at the right edge + ...)
        uint32_t edge = 0;  // bit 0: the kx == 2 taps are padding, bit 1: the ky == 2 taps are
        auto gather = [&](int tile) {
            const unsigned pix_raw = (unsigned)tile * TC_BM + (unsigned)row;
            const unsigned pix = pix_raw < (unsigned)p.n_pix ? pix_raw : (unsigned)p.n_pix - 1u;  // rows past the end re-read the last pixel (their outputs are not stored)
            const unsigned t = pix / (unsigned)Wo;
            const int xo = (int)(pix - t * (unsigned)Wo);
            const unsigned b = t / (unsigned)Ho;
            const int yo = (int)(t - b * (unsigned)Ho);
            const bool last_y = (2 * yo + 2 >= p.H), last_x = (2 * xo + 2 >= p.W);
            edge = (last_x ? 1u : 0u) | (last_y ? 2u : 0u);
            // right edge: the window starts one pixel to the left (its 12 aligned bytes then end with the row: W % 4 == 0); the
            // padded column's bytes are whatever follows in the registers and select the table's zero page
            const unsigned off = ((b * (unsigned)p.H + 2u * (unsigned)yo) * (unsigned)p.W + 2u * (unsigned)xo) * 3u - (last_x ? 3u : 0u);
            const unsigned a0 = off & ~3u, rs = (unsigned)p.W * 3u;
            sh = (off & 3u) * 8u + (last_x ? 24u : 0u);
            const uint8_t* in = (const uint8_t*)p.in;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const unsigned a = a0 + (ky == 2 && last_y ? 1u : (unsigned)ky) * rs;  // bottom edge: re-read row 1 (discarded)
#pragma unroll
                for (int i = 0; i < 3; ++i) w[ky][i] = __ldg(reinterpret_cast<const uint32_t*>(in + a) + i);
            }
        };
        const uint32_t lutb = base + STC3_OFF_LUT + (uint32_t)(lane & 15) * 8u;  // this lane's copy (32-bit shared address: no 64-bit pointer arithmetic per tap)
        const int stride = 4 * (int)gridDim.x;
        const int first = (int)blockIdx.x + set * (int)gridDim.x;
        if (first < n_tiles) gather(first);
        uint32_t ph = 0;
        for (int tile = first; tile < n_tiles; tile += stride, ph ^= 1u) {
            // the pixel's 9 bytes of each row, shifted down to bit 0: v[ky][0] = bytes 0-3, [1] = bytes 4-7; v8 = byte 8 of rows 0, 1, 2
            uint32_t v[3][2], v8 = 0;
            {
                const bool far = sh >= 32u;
                const uint32_t s5 = sh & 31u;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t x0 = far ? w[ky][1] : w[ky][0], x1 = far ? w[ky][2] : w[ky][1], x2 = far ? 0u : w[ky][2];
                    v[ky][0] = __funnelshift_r(x0, x1, s5);
                    v[ky][1] = __funnelshift_r(x1, x2, s5);
                    v8 |= ((x2 >> s5) & 0xffu) << (8 * ky);
                }
            }
            const uint32_t edge_now = edge;
            // the window registers are dead: the next tile's loads go out now and land during this tile's table look-ups
            // (issued at the end of the iteration they were the head of the next one's dependency chain: 2 producer sets -> 167 us)
            if (tile + stride < n_tiles) gather(tile + stride);
            // padded taps read the table's zero page: a byte-offset mask per tap class (page offset c * 32 KB | 96 KB)
            const bool px = edge_now & 1u, py = edge_now & 2u;
            mbar_wait(bar_aempty + 8 * set, ph ^ 1u);
            tc_fence_after();
#pragma unroll
            for (int qt = 0; qt < 4; ++qt) {  // 8 taps per TMEM store: 16 live values (the 64-register cap spilled the prefetched window otherwise)
                float hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = qt * 8 + i;
                    if (k >= 27) {
                        hi[i] = 0.f, lo[i] = 0.f;
                        continue;
                    }
                    const int ky = k / 9, r = k - ky * 9, kx = r / 3, c = r - kx * 3;
                    const uint32_t word = r == 8 ? v8 : v[ky][r >> 2];
                    const int bs = 8 * (r == 8 ? ky : (r & 3)) - 7;  // byte * 128 = the table entry's byte offset
                    const uint32_t boff = (bs < 0 ? word << 7 : word >> bs) & 0x7f80u;
                    // page c, or page 3 (zeros) for a padded tap: c * 32 KB | 96 KB = 96 KB for c = 0..2
                    const bool pad = (ky == 2 && py) || (kx == 2 && px);
                    const uint32_t pg = pad ? 3u * 32768u : (uint32_t)c * 32768u;
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(hi[i]), "=f"(lo[i]) : "r"(lutb + pg + boff));
                }
                tmem_st8(ta + 8u * qt, hi);
                tmem_st8(ta + 32u + 8u * qt, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready + 8 * set);
        }
    } else if (warp >= 20) {
        // ================= drain: group = accumulator stage =================
        pdl_wait();  // `out` may still be read by the previous forward
        const int g = (warp - 20) >> 2, q = warp & 3;
        uint8_t* stg = sm + STC3_OFF_STG + (warp - 20) * 4096;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 64u;
        uint32_t ph = 0;
        for (int tile = (int)blockIdx.x + g * (int)gridDim.x; tile < n_tiles; tile += 3 * (int)gridDim.x, ph ^= 1u) {
            mbar_wait(bar_tfull + 8 * g, ph);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[16], c[16];
                tmem_ld16(taddr + 16u * h, v);
                tmem_ld16(taddr + (uint32_t)STC_NC + 16u * h, c);
                tmem_ld_wait();
                // transpose through a per-warp swizzled tile so that one STG.128 writes 4 whole pixels (512 contiguous bytes)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(stg + lane * 128 + (((4 * h + j) ^ (lane & 7)) << 4)) =
                        swish4p(make_float4(v[4 * j] + c[4 * j], v[4 * j + 1] + c[4 * j + 1], v[4 * j + 2] + c[4 * j + 2], v[4 * j + 3] + c[4 * j + 3]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * g);  // the accumulator pair has been read: hand it back
            const long long pix0 = (long long)tile * TC_BM + q * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + (lane >> 3), cc = lane & 7;
                const float4 x = *reinterpret_cast<const float4*>(stg + r * 128 + ((cc ^ (r & 7)) << 4));
                if (pix0 + r < p.n_pix) st4(p.out + (size_t)(pix0 + r) * 32 + cc * 4, x);
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// k_stem_tc3 reads aligned 32-bit words of the u8 image: W % 4 == 0 (every row then starts and ends on a word) and a 4-byte aligned base
inline bool stc3_supported(const StcParams& p) { return p.W % 4 == 0 && ((uintptr_t)p.in & 3u) == 0 && p.n_pix * 27ll < (1ll << 31); }
inline cudaError_t stc3_launch(const StcParams& p, int sms, cudaStream_t s) {
    cudaError_t e = smem_optin((const void*)k_stem_tc3, (int)STC3_SMEM);
    if (e != cudaSuccess) return e;
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    return launch_pdl(k_stem_tc3, dim3(grid), dim3(STC3_THREADS), STC3_SMEM, s, p);
}

template <int FMT>
inline cudaError_t stc2_launch_t(const StcParams& p, int sms, cudaStream_t s) {
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    return launch_pdl(k_stem_tc2<FMT>, dim3(grid), dim3(STC2_THREADS), STC2_SMEM, s, p);
}

template <int FMT>
inline cudaError_t stc_launch_t(const StcParams& p, int grid, cudaStream_t s) {
    return launch_pdl(k_stem_tc<FMT>, dim3(grid), dim3(STC_THREADS), STC_SMEM, s, p);
}

inline int stc_plan(PwTcState& st, const float* key_w, const void* in, const float* lut, float* out, int B, int H, int W, StcParams* p,
                    int* grid) {
    auto it = st.layers.find(key_w);
    if (it == st.layers.end()) return fail(CF_EINVAL, "stc_plan: the stem weight image was not prepared");
    const TcLayer& L = it->second;
    if (L.NC != STC_NC || L.nchunks != 1 || L.nkb != 1) return fail(CF_EINVAL, "stc_plan: unexpected stem image NC=%d chunks=%d nkb=%d", L.NC, L.nchunks, L.nkb);
    p->in = in;
    p->bimg = L.img;
    p->lut = lut;
    p->out = out;
    p->B = B;
    p->H = H;
    p->W = W;
    p->n_pix = (long long)B * (H / 2) * (W / 2);
    if (p->n_pix + TC_BM >= (1ll << 31)) return fail(CF_EINVAL, "stc_plan: %lld output pixels exceed the 32-bit index math", p->n_pix);
    p->n_tiles = (int)((p->n_pix + TC_BM - 1) / TC_BM);
    const int want = 3 * st.sms;
    *grid = p->n_tiles < want ? p->n_tiles : want;
    return CF_OK;
}

}  // namespace cf
