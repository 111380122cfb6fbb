// Direct-convolution kernels: stem (dense 3x3 s2), depth-wise 3x3/5x5, collapsed heads 3x3.
// Activations are NHWC fp32.  HBM-bound work: coalesced float4 traffic along the channel axis,
// weights staged in shared memory, zero padding folded into index math (ZeroPad2d,
// model/centernet.py:63, is never materialised).
#pragma once
#include "common.cuh"

namespace cf {

// ----------------------------------------------------------------------------------------
// K0 pre-processing: cv2.resize(img, (W', H')) of centerface.py:30, default INTER_LINEAR on 8UC3, on the device
// and bit-exact (OpenCV modules/imgproc/src/resize.cpp; restated and pinned in oracle/centerface_oracle.py::
// resize_linear_u8).  The per-column / per-row source index and the two 11-bit fixed-point weights are built on the
// host with OpenCV's own float arithmetic (ResizeTables) and read here; one thread = one output pixel x 3 channels.
//   tab layout (int32): xofs[dw] | xa0[dw] | xa1[dw] | y0[dh] | y1[dh] | yb0[dh] | yb1[dh]
// area2 = 1: the exact 2x decimation that OpenCV silently switches to INTER_AREA.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize_u8(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                   const int32_t* __restrict__ tab, int B, int sh, int sw, int dh, int dw,
                                                   int area2) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)B * dh * dw) return;
    const int dx = (int)(i % dw);
    const int dy = (int)((i / dw) % dh);
    const int b = (int)(i / ((long long)dw * dh));
    const uint8_t* s = src + (size_t)b * sh * sw * 3;
    uint8_t* o = dst + (size_t)i * 3;
    if (area2) {
        const uint8_t* p0 = s + ((size_t)(2 * dy) * sw + 2 * dx) * 3;
        const uint8_t* p1 = p0 + (size_t)sw * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = (uint8_t)((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
        return;
    }
    const int sx = tab[dx], a0 = tab[dw + dx], a1 = tab[2 * dw + dx];
    const int sx1 = min(sx + 1, sw - 1);
    const int32_t* ty = tab + 3 * dw;
    const int y0 = ty[dy], y1 = ty[dh + dy], b0 = ty[2 * dh + dy], b1 = ty[3 * dh + dy];
    const uint8_t* r0 = s + (size_t)y0 * sw * 3;
    const uint8_t* r1 = s + (size_t)y1 * sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;  // HResizeLinear, int32
        const int h1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
        o[c] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);  // VResizeLinear<uchar>
    }
}

// ----------------------------------------------------------------------------------------
// K0b pre-processing: cv2.warpAffine(img, M, (W', H'), flags=INTER_LINEAR) on 8UC3 with the default constant-0 border --
// the letter-box step of the reference's loader (dataset/dataset.py:130-134) -- on the device and bit-exact (OpenCV
// imgwarp.cpp: fixed-point source positions + remap's 15-bit bilinear table; restated and pinned against cv2 in
// oracle/centerface_oracle.py::warp_affine_linear_u8).  The inverse map's per-column / per-row integer tables are built on
// the host in fp64 exactly as OpenCV does (cf_warp_affine_tables); one thread = one output pixel x 3 channels.
//   tab layout (int32): adelta[dw] | bdelta[dw] | X0[dh] | Y0[dh]   (positions in 1/1024 pixel, round_delta included)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_warp_affine_u8(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                        const int32_t* __restrict__ tab, int B, int sh, int sw, int dh, int dw) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)B * dh * dw) return;
    const int dx = (int)(i % dw);
    const int dy = (int)((i / dw) % dh);
    const int b = (int)(i / ((long long)dw * dh));
    const int X = (tab[2 * dw + dy] + tab[dx]) >> 5;            // AB_BITS - INTER_BITS
    const int Y = (tab[2 * dw + dh + dy] + tab[dw + dx]) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));  // saturate_cast<short>
    const int fx = X & 31, fy = Y & 31;
    int w00 = 32 * (32 - fx) * (32 - fy), w01 = 32 * fx * (32 - fy), w10 = 32 * (32 - fx) * fy, w11 = 32 * fx * fy;
    if (fx == 0 && fy == 0) w00 = 32767, w11 = 1;  // saturate_cast<short>(32768) + the table's sum fix-up
    const uint8_t* s = src + (size_t)b * sh * sw * 3;
    const bool x0 = sx >= 0 && sx < sw, x1 = sx + 1 >= 0 && sx + 1 < sw, y0 = sy >= 0 && sy < sh, y1 = sy + 1 >= 0 && sy + 1 < sh;
    uint8_t* o = dst + (size_t)i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int p00 = (y0 && x0) ? s[((size_t)sy * sw + sx) * 3 + c] : 0;
        const int p01 = (y0 && x1) ? s[((size_t)sy * sw + sx + 1) * 3 + c] : 0;
        const int p10 = (y1 && x0) ? s[((size_t)(sy + 1) * sw + sx) * 3 + c] : 0;
        const int p11 = (y1 && x1) ? s[((size_t)(sy + 1) * sw + sx + 1) * 3 + c] : 0;
        const int v = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;
        o[c] = (uint8_t)min(255, max(0, v));
    }
}

// ----------------------------------------------------------------------------------------
// K1 stem: ZeroPad2d(0,1,0,1) + conv3x3 s2 3->32 (no bias) + Swish   (model/centernet.py:224)
//   FMT 0: fp32 NCHW normalised input (what EfficientNet.forward receives)
//   FMT 1: u8 HWC BGR input; /255, -mean, /std (centerface.py:32-34) applied through a
//          768-entry table built on the host with the reference's own fp32 ops (bit-exact).
// One thread = one output pixel x all 32 output channels.  The 864 weights travel as a by-value
// kernel parameter, i.e. in the constant bank, so every FFMA takes its weight as a c[][] operand:
// no shared-memory or register traffic for weights, the kernel is FFMA-issue bound (864 per pixel).
// ----------------------------------------------------------------------------------------
struct StemW {
    float w[27 * 32];  // [(ky*3+kx)*3+ci][co]; passed by value => lives in the constant bank
};

template <int FMT>
__global__ void __launch_bounds__(128) k_stem(const void* __restrict__ in, const __grid_constant__ StemW sw,
                                              const float* __restrict__ lut, float* __restrict__ out,
                                              int B, int H, int W) {
    __shared__ float lut_s[FMT == 1 ? 768 : 1];
    __shared__ __align__(16) uint8_t stg_s[4][4096];  // per warp: 32 pixels x 128 B, 16-byte chunks XOR-swizzled by row
    pdl_trigger();
    if (FMT == 1) {
        for (int i = threadIdx.x; i < 768; i += 128) lut_s[i] = lut[i];
        __syncthreads();
    }
    pdl_wait();  // `out` may still be read by the previous forward's kernels
    const int Ho = H >> 1, Wo = W >> 1;
    const long long n_pix = (long long)B * Ho * Wo;
    const long long pix_raw = (long long)blockIdx.x * 128 + threadIdx.x;
    const bool live = pix_raw < n_pix;
    const long long pix = live ? pix_raw : n_pix - 1;  // tail threads recompute the last pixel and store nothing
    const int xo = (int)(pix % Wo);
    const int yo = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((long long)Wo * Ho));

    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * yo + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = 2 * xo + kx;
            float v[3] = {0.f, 0.f, 0.f};
            if (iy < H && ix < W) {  // ZeroPad2d(0,1,0,1): only the bottom/right edge is padded
                if (FMT == 1) {
                    const uint8_t* p = (const uint8_t*)in + ((size_t)(b * H + iy) * W + ix) * 3;
                    v[0] = lut_s[__ldg(p)];
                    v[1] = lut_s[256 + __ldg(p + 1)];
                    v[2] = lut_s[512 + __ldg(p + 2)];
                } else {
                    const float* p = (const float*)in + ((size_t)(b * 3) * H + iy) * W + ix;
                    v[0] = __ldg(p);
                    v[1] = __ldg(p + (size_t)H * W);
                    v[2] = __ldg(p + 2 * (size_t)H * W);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int co = 0; co < 32; ++co) acc[co] = fmaf(v[c], sw.w[((ky * 3 + kx) * 3 + c) * 32 + co], acc[co]);
        }
    }
    // A thread owns one pixel = 128 contiguous output bytes; storing them directly makes every STG.128 of a warp touch 32
    // different lines (16 B each).  Transposed through shared memory, a store instruction writes 4 whole pixels = 512
    // contiguous bytes (8 lanes per pixel).
    const int lane = threadIdx.x & 31;
    uint8_t* stg = stg_s[threadIdx.x >> 5];
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
            swish4(make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]));
    __syncwarp();
    const long long pix0 = (long long)blockIdx.x * 128 + (threadIdx.x & ~31);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3), c = lane & 7;
        const float4 v = *reinterpret_cast<const float4*>(stg + r * 128 + ((c ^ (r & 7)) << 4));
        if (pix0 + r < n_pix) st4(out + (size_t)(pix0 + r) * 32 + c * 4, v);
    }
    (void)live;
}

// ----------------------------------------------------------------------------------------
// K3 depth-wise KSxKS stride S + Swish   (ConvReLU with groups=hidden, model/centernet.py:112)
// Padding (model/centernet.py:68-70): total p = KS-S, lo = p/2 on left/top, rest right/bottom.
// Thread = XT consecutive output pixels along x for one float4 of channels; the input row
// window is loaded once per kernel row and reused across the XT outputs in registers, the
// vertical reuse comes from L1 (a CTA covers a TYx(TX*XT) pixel tile).
// w layout: [KS*KS][C].
// ----------------------------------------------------------------------------------------
template <int KS, int S, int XT>
__global__ void __launch_bounds__(256) k_dw(const float* __restrict__ in, const float* __restrict__ w,
                                            float* __restrict__ out, int B, int Hi, int Wi, int C,
                                            int Ho, int Wo, int tiles_x, int tiles_y, int csplit) {
    constexpr int LO = (KS - S) / 2;
    constexpr int WIN = (XT - 1) * S + KS;  // input columns feeding XT outputs
    constexpr int TXS = 4;                  // strips per tile row  -> tile width  = 4*XT
    constexpr int TY = 8;                   // tile height
    const int C4 = C >> 2;
    int bid = blockIdx.x;
    // channel groups [cg0, cg1) of this CTA: small maps (20x20 .. 40x40 with 384-960 channels) have too
    // few spatial tiles to fill 148 SMs, so the channel axis is split across CTAs as well
    const int cs = bid % csplit;
    bid /= csplit;
    const int cg0 = (int)(((long long)C4 * cs) / csplit), cg1 = (int)(((long long)C4 * (cs + 1)) / csplit);
    const int CG = cg1 - cg0;
    const int tx0 = (bid % tiles_x) * (TXS * XT);
    bid /= tiles_x;
    const int ty0 = (bid % tiles_y) * TY;
    const int b = bid / tiles_y;

    const int items = TXS * TY * CG;
    for (int it = threadIdx.x; it < items; it += 256) {
        const int c4 = cg0 + it % CG;
        const int p = it / CG;
        const int xo0 = tx0 + (p % TXS) * XT;
        const int yo = ty0 + p / TXS;
        if (yo >= Ho || xo0 >= Wo) continue;
        float4 acc[XT];
#pragma unroll
        for (int j = 0; j < XT; ++j) acc[j] = make_float4(0, 0, 0, 0);
        const int ix0 = xo0 * S - LO;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
            const int iy = yo * S - LO + ky;
            if (iy < 0 || iy >= Hi) continue;
            const float* row = in + ((size_t)(b * Hi + iy) * Wi) * C + c4 * 4;
            float4 win[WIN];
#pragma unroll
            for (int i = 0; i < WIN; ++i) {
                const int ix = ix0 + i;
                win[i] = (ix >= 0 && ix < Wi) ? ldg4(row + (size_t)ix * C) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                const float4 wv = ldg4(w + (size_t)(ky * KS + kx) * C + c4 * 4);
#pragma unroll
                for (int j = 0; j < XT; ++j) fma44(acc[j], win[j * S + kx], wv);
            }
        }
#pragma unroll
        for (int j = 0; j < XT; ++j) {
            const int xo = xo0 + j;
            if (xo < Wo) st4(out + ((size_t)(b * Ho + yo) * Wo + xo) * C + c4 * 4, swish4(acc[j]));
        }
    }
}

// ----------------------------------------------------------------------------------------
// K3b depth-wise, register-tiled (second generation; profiles/r1_summary.md: the first kernel issues 4.5 (3x3)
// to 16 (5x5) L1 loads per output vector and is LSU-bound, 1.6-2.8 TB/s).  A thread owns XT x TY outputs of
// VEC channels; all KS*KS taps of its channels sit in registers (VEC = 2 for 5x5 so that 25 taps fit), input rows
// are streamed once through a WIN-wide register window and every loaded value feeds up to KS x KS FMAs:
// 3 (3x3) to 4 (5x5 s1) loads per output vector.  Lanes run along the channel axis (CVL consecutive channel
// vectors, >= 128 contiguous bytes per pixel), then along x strips.
// ----------------------------------------------------------------------------------------
template <int VEC>
struct VecT;
template <>
struct VecT<2> {
    typedef float2 T;
};
template <>
struct VecT<4> {
    typedef float4 T;
};
__device__ __forceinline__ float2 vzero(float2*) { return make_float2(0, 0); }
__device__ __forceinline__ float4 vzero(float4*) { return make_float4(0, 0, 0, 0); }
__device__ __forceinline__ void vfma(float2& a, float2 x, float2 w) {
    a.x = fmaf(x.x, w.x, a.x);
    a.y = fmaf(x.y, w.y, a.y);
}
__device__ __forceinline__ void vfma(float4& a, float4 x, float4 w) { fma44(a, x, w); }
__device__ __forceinline__ float2 vswish(float2 v) { return make_float2(swishf(v.x), swishf(v.y)); }
__device__ __forceinline__ float4 vswish(float4 v) { return swish4(v); }

template <int KS, int S, int VEC, int XT, int TY>
__global__ void __launch_bounds__(256) k_dw3(const float* __restrict__ in, const float* __restrict__ w,
                                             float* __restrict__ out, int B, int Hi, int Wi, int C, int Ho, int Wo,
                                             int cvl, int nstrip, int ncvh, int ntile_y, long long ntasks) {
    typedef typename VecT<VEC>::T V;
    constexpr int LO = (KS - S) / 2;
    constexpr int WIN = (XT - 1) * S + KS, ROWS = (TY - 1) * S + KS;
    long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= ntasks) return;
    const int cv_lo = (int)(t % cvl);
    t /= cvl;
    const int strip = (int)(t % nstrip);
    t /= nstrip;
    const int cv_hi = (int)(t % ncvh);
    t /= ncvh;
    const int ty = (int)(t % ntile_y);
    const int b = (int)(t / ntile_y);
    const int c0 = (cv_hi * cvl + cv_lo) * VEC;
    const int xo0 = strip * XT, yo0 = ty * TY;

    V wt[KS * KS];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) wt[i] = __ldg(reinterpret_cast<const V*>(w + (size_t)i * C + c0));
    V acc[TY][XT];
#pragma unroll
    for (int y = 0; y < TY; ++y)
#pragma unroll
        for (int x = 0; x < XT; ++x) acc[y][x] = vzero((V*)nullptr);

    const int ix0 = xo0 * S - LO, iy0 = yo0 * S - LO;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int iy = iy0 + r;
        if (iy < 0 || iy >= Hi) continue;  // zero padding rows contribute nothing
        const float* row = in + ((size_t)(b * Hi + iy) * Wi) * C + c0;
        V win[WIN];
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
            const int ix = ix0 + i;
            win[i] = (ix >= 0 && ix < Wi) ? __ldg(reinterpret_cast<const V*>(row + (size_t)ix * C)) : vzero((V*)nullptr);
        }
#pragma unroll
        for (int y = 0; y < TY; ++y) {
            const int ky = r - y * S;
            if (ky < 0 || ky >= KS) continue;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                for (int x = 0; x < XT; ++x) vfma(acc[y][x], win[x * S + kx], wt[ky * KS + kx]);
        }
    }
#pragma unroll
    for (int y = 0; y < TY; ++y) {
        const int yo = yo0 + y;
        if (yo >= Ho) continue;
#pragma unroll
        for (int x = 0; x < XT; ++x) {
            const int xo = xo0 + x;
            if (xo < Wo) *reinterpret_cast<V*>(out + ((size_t)(b * Ho + yo) * Wo + xo) * C + c0) = vswish(acc[y][x]);
        }
    }
}

// ----------------------------------------------------------------------------------------
// K5 heads.  The reference runs, per head, conv3x3(24->24)+b0 then conv1x1(24->c)+b1 with no
// non-linearity between (model/centernet.py:249-256), so all four heads collapse exactly
// (up to fp rounding) into ONE 3x3 conv 24->15: W' = W1.W0, b' = W1.b0 + b1,
// computed in fp64 by the packer.  Output channel order: hm, wh0, wh1, lm0..9, reg0, reg1, pad.
// Outputs are written planar (NCHW) like the reference's dict, plus
// hm_sig = clamp(sigmoid(hm), 1e-4, 1-1e-4) (centerface.py:43) for the decoders.
// CTA = 128 threads, 32x16 output pixels; thread = 4 pixels (x = sx + 8p) x 16 outputs.
// smem: input tile [18][34][28] (pixel stride 28 floats -> conflict-free LDS.128) + weights.
// ----------------------------------------------------------------------------------------
struct HeadsW {
    float w[216 * 16];  // [(ky*3+kx)*24+ci][16], passed by value => constant bank (13.8 KB of the 32 KB parameter space)
    float b[16];
};
constexpr int HEADS_PS = 28;  // padded pixel stride in floats
constexpr int HEADS_SMEM = (18 * 34 * HEADS_PS) * 4;

__global__ void __launch_bounds__(128) k_heads(const float* __restrict__ in, const __grid_constant__ HeadsW hw,
                                               float* __restrict__ hm,
                                               float* __restrict__ wh, float* __restrict__ lm,
                                               float* __restrict__ reg, float* __restrict__ hm_sig,
                                               int B, int H, int W) {
    extern __shared__ __align__(16) float sm[];
    float* tile = sm;                        // [18][34][28]
    const int tid = threadIdx.x;
    const int x00 = blockIdx.x * 32, y00 = blockIdx.y * 16, b = blockIdx.z;
    pdl_trigger();
    pdl_wait();

    for (int i = tid; i < 18 * 34 * 6; i += 128) {
        const int c4 = i % 6;
        const int px = (i / 6) % 34;
        const int py = i / (6 * 34);
        const int gy = y00 + py - 1, gx = x00 + px - 1;
        float4 v = make_float4(0, 0, 0, 0);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = ldcg4(in + ((size_t)(b * H + gy) * W + gx) * 24 + c4 * 4);
        st4(tile + (py * 34 + px) * HEADS_PS + c4 * 4, v);
    }
    __syncthreads();

    const int sx = tid & 7, sy = tid >> 3;
    float4 acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = make_float4(0, 0, 0, 0);

    // weights come from the constant bank: the tap loop is rolled (code size), its 24 x 16 weights are addressed with a
    // uniform offset, so the inner loop is 4 LDS.128 of inputs per 256 FFMAs instead of 20
#pragma unroll 1
    for (int t = 0; t < 9; ++t) {
        const int ky = t / 3, kx = t - ky * 3;
        const float* trow = tile + ((sy + ky) * 34 + sx + kx) * HEADS_PS;
        const float* wt = hw.w + t * 24 * 16;
#pragma unroll
        for (int c4 = 0; c4 < 6; ++c4) {
            float4 a[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) a[p] = *reinterpret_cast<const float4*>(trow + p * 8 * HEADS_PS + c4 * 4);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float* wq = wt + (c4 * 4 + ci) * 16 + q * 4;
                    const float4 wv = make_float4(wq[0], wq[1], wq[2], wq[3]);
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float av = ci == 0 ? a[p].x : ci == 1 ? a[p].y : ci == 2 ? a[p].z : a[p].w;
                        fma4(acc[p][q], av, wv);
                    }
                }
            }
        }
    }

    const int y = y00 + sy;
    if (y >= H) return;
    const size_t plane = (size_t)H * W;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x00 + sx + 8 * p;
        if (x >= W) continue;
        float o[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[q * 4 + 0] = acc[p][q].x + hw.b[q * 4 + 0];
            o[q * 4 + 1] = acc[p][q].y + hw.b[q * 4 + 1];
            o[q * 4 + 2] = acc[p][q].z + hw.b[q * 4 + 2];
            o[q * 4 + 3] = acc[p][q].w + hw.b[q * 4 + 3];
        }
        const size_t pix = (size_t)y * W + x;
        hm[(size_t)b * plane + pix] = o[0];
        const float s = 1.f / (1.f + expf(-o[0]));
        hm_sig[(size_t)b * plane + pix] = fminf(fmaxf(s, 1e-4f), 1.f - 1e-4f);
        wh[((size_t)b * 2 + 0) * plane + pix] = o[1];
        wh[((size_t)b * 2 + 1) * plane + pix] = o[2];
#pragma unroll
        for (int j = 0; j < 10; ++j) lm[((size_t)b * 10 + j) * plane + pix] = o[3 + j];
        reg[((size_t)b * 2 + 0) * plane + pix] = o[13];
        reg[((size_t)b * 2 + 1) * plane + pix] = o[14];
    }
}

}  // namespace cf
