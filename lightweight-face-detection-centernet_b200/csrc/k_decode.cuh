// Heat-map decode kernels.
//   path C: _nms + _topk + _gather_feat + ctdet_decode      (centerface_ext.py:11-82)
//   paths A/B: threshold decode + greedy IoU NMS + //scale  (centerface.py:73-151, :55-58;
//                                                            eval_widerface.py:92-152)
// HBM/latency-bound integer and compare work: one CTA per image, shared-memory radix select
// and bitonic sort, warp-ballot ordered compaction.  Orders are total (no unspecified ties):
//   path C   : (score desc, flat index asc)
//   paths A/B: (score desc, flat index desc)  == stable ascending argsort reversed.
#pragma once
#include "common.cuh"

namespace cf {

// monotone float -> uint key (works for negative inputs too)
__device__ __forceinline__ uint32_t fkey(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---- K6: 3x3 peak keep (max_pool2d pads with -inf: border pixels only see in-bounds
//      neighbours; equal plateau neighbours are all kept), centerface_ext.py:44-50 -----
__global__ void __launch_bounds__(256) k_peak_mask(const float* __restrict__ heat, float* __restrict__ pk,
                                                   int B, int H, int W) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)B * H * W) return;
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const float* img = heat + (i - (long long)y * W - x);
    const float v = __ldcg(img + y * W + x);
    bool keep = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            keep = keep && !(__ldcg(img + yy * W + xx) > v);
        }
    }
    pk[i] = keep ? v : v * 0.f;  // heat * keep.float()
}

// in-place bitonic sort, descending, n = power of two, keys in shared memory
__device__ __forceinline__ void bitonic_desc(unsigned long long* keys, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---- K7+K8: per-image top-K of the peak map + gather + box assembly ------------------
// grid = B, block = 1024.  4-pass MSB radix select of the K-th largest score, ordered
// (lowest index first) admission of ties at the threshold, bitonic sort of the K winners.
// SM: the image's H*W keys are read ONCE (all loads in flight together) into dynamic shared memory (4 B each, H*W <= 51 200),
// the 3x3 peak keep is applied to that copy (`pk` = the raw heat map), and the four radix passes and the collection pass run on
// it; otherwise `pk` is the map k_peak_mask wrote and every pass re-reads it from L2.
// The histogram increments are warp-aggregated (__match_any_sync): after the peak mask ~95 % of a map is +0.0, i.e. ONE
// radix bin, and 25 600 shared-memory atomics on one address were most of this kernel's 68 us.
constexpr int TOPK_SMEM_MAX_HW = 51200;
template <bool SM>
__global__ void __launch_bounds__(1024) k_topk(const float* __restrict__ pk, const float* __restrict__ wh,
                                               const float* __restrict__ reg, int H, int W, int K,
                                               float* __restrict__ dets, int32_t* __restrict__ inds, int cand_cap, int C, int wh_planes) {
    extern __shared__ uint32_t skeys[];
    __shared__ unsigned hist[256];
    __shared__ unsigned long long buf[1024];
    __shared__ unsigned wcnt[32];
    __shared__ unsigned s_prefix, s_krem, s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // C classes (centerface_ext.py:11-27): top-K per class, then top-K of those C*K candidates = the global top-K of the
    // image's C*H*W scores; ties go to the lower class, then the lower pixel index (the candidate order of :20) = the lower
    // index of the flattened [C][H*W] map, which is the order this kernel admits ties in.  The face model has C = 1.
    const int HW1 = H * W;
    const int HW = C * HW1;
    const int b = blockIdx.x;
    const float* p = pk + (size_t)b * HW;
    pdl_trigger();
    pdl_wait();

    if (SM) {
#pragma unroll 8
        for (int i = tid; i < HW; i += 1024) skeys[i] = fkey(__ldcg(p + i));
        __syncthreads();
        // The 3x3 peak keep (_nms, centerface_ext.py:44-50; k_peak_mask above) on the shared copy: `pk` is the RAW heat map here.
        // A pixel survives iff no in-bounds neighbour of its own plane is larger (max_pool2d pads with -inf; plateau neighbours
        // all survive); the others become heat * 0.  The key order is the float order (-0 < +0 in keys only, and a suppressed
        // +-0 stays +-0), so comparing keys decides exactly what comparing floats does.  One launch and one pass over the map
        // less than k_peak_mask + k_topk (16 + 28 us -> 40 us at 32 x 160 x 160; the keep pass is ~45 instructions per pixel on ONE CTA per image).
        unsigned long long kb = 0ull;  // keep bits of this thread's <= 50 pixels
        int k = 0;
        // (y, x, plane) of pixel i advance by 1024 pixels per iteration without a division; out-of-map neighbours are replaced by
        // clamped ones, i.e. by other members of the same 3x3 window (or the pixel itself): branch-free, nine independent loads
        const int stepy = 1024 / W, stepx = 1024 - stepy * W;
        int pix = tid % HW1, plane0 = tid - pix;
        int y = pix / W, x = pix - y * W;
        for (int i = tid; i < HW; i += 1024, ++k) {
            const uint32_t* pl = skeys + plane0;
            const int ym = max(y - 1, 0) * W, y0 = y * W, yp = min(y + 1, H - 1) * W;
            const int xm = max(x - 1, 0), xp = min(x + 1, W - 1);
            const uint32_t v = pl[y0 + x];
            uint32_t m = max(max(pl[ym + xm], pl[ym + x]), max(pl[ym + xp], pl[y0 + xm]));
            m = max(max(m, pl[y0 + xp]), max(max(pl[yp + xm], pl[yp + x]), pl[yp + xp]));
            if (!(m > v)) kb |= 1ull << k;
            x += stepx, y += stepy;
            if (x >= W) x -= W, ++y;
            while (y >= H) y -= H, plane0 += HW1;  // next class plane (C > 1)
        }
        __syncthreads();
        k = 0;
        for (int i = tid; i < HW; i += 1024, ++k)
            if (!((kb >> k) & 1ull)) skeys[i] = fkey(fkey_inv(skeys[i]) * 0.f);  // heat * keep.float()
        __syncthreads();
    }
    auto key_at = [&](int i) -> uint32_t { return SM ? skeys[i] : fkey(__ldcg(p + i)); };
    const int HWp = (HW + 1023) & ~1023;  // whole warps walk the tail together (ballots below)

    // Candidate list (SM only).  After the peak mask ~90-95 % of a map is +0.0; when at least K scores are positive the K winners
    // are among them, so the positive keys are compacted ONCE, in index order, and the radix passes and the tie admission walk
    // that list (1-3 rows of 1024 instead of 25 per pass).  Fewer than K positives, or more than the list holds: the full walk.
    unsigned* ecnt = SM ? skeys + HWp : nullptr;        // [rows * 32] per (row of 1024, warp) counts, rows = HWp / 1024 <= 50
    unsigned* cand = SM ? ecnt + (HWp >> 5) : nullptr;  // [cand_cap] indices of the positive keys, ascending
    const int rows = HWp >> 10;
    // exclusive scan of ecnt[0 .. rows*32) in place (thread t owns entries 2t, 2t+1); returns the total through s_cnt
    auto scan_ecnt = [&]() {
        const int n = rows * 32;
        const unsigned a = 2 * tid < n ? ecnt[2 * tid] : 0u, b2 = 2 * tid + 1 < n ? ecnt[2 * tid + 1] : 0u;
        unsigned inc = a + b2;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) wcnt[warp] = inc;
        __syncthreads();
        unsigned woff = 0;
        for (int w2 = 0; w2 < warp; ++w2) woff += wcnt[w2];
        const unsigned excl = woff + inc - (a + b2);
        if (2 * tid < n) ecnt[2 * tid] = excl;
        if (2 * tid + 1 < n) ecnt[2 * tid + 1] = excl + a;
        if (tid == 1023) s_cnt = woff + inc;
        __syncthreads();
    };
    int ncand = -1;  // >= 0: the list is in use
    if (SM) {
        constexpr uint32_t KEY0 = 0x80000000u;  // fkey(+0.0f): the masked-out pixels (sigmoid scores are > 0)
        for (int v = 0; v < rows; ++v) {
            const int i = v * 1024 + tid;
            const unsigned m = __ballot_sync(0xffffffffu, i < HW && skeys[i] > KEY0);
            if (lane == 0) ecnt[v * 32 + warp] = __popc(m);
        }
        __syncthreads();
        scan_ecnt();
        const unsigned npos = s_cnt;
        __syncthreads();
        if (npos >= (unsigned)K && npos <= (unsigned)cand_cap) {
            ncand = (int)npos;
            for (int v = 0; v < rows; ++v) {
                const int i = v * 1024 + tid;
                const bool pos = i < HW && skeys[i] > KEY0;
                const unsigned m = __ballot_sync(0xffffffffu, pos);
                if (pos) cand[ecnt[v * 32 + warp] + __popc(m & ((1u << lane) - 1u))] = (unsigned)i;
            }
            __syncthreads();
        }
    }
    const int NW = ncand >= 0 ? ((ncand + 1023) & ~1023) : HWp;  // what the passes walk: list positions or pixels
    auto item = [&](int j, bool* in) -> uint32_t {                // key of walk position j
        if (ncand >= 0) {
            *in = j < ncand;
            return *in ? skeys[cand[j]] : 0u;
        }
        *in = j < HW;
        return *in ? key_at(j) : 0u;
    };

    unsigned prefix = 0, mask = 0, krem = K;
    for (int pass = 3; pass >= 0; --pass) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < NW; i += 1024) {
            bool in;
            const uint32_t key = item(i, &in);
            const bool hit = in && (key & mask) == prefix;
            const unsigned bin = (key >> (8 * pass)) & 255u;
            const unsigned act = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const unsigned same = __match_any_sync(act, bin);
                if (lane == __ffs(same) - 1) atomicAdd(&hist[bin], (unsigned)__popc(same));
            }
        }
        __syncthreads();
        // the bin holding the krem-th largest key: S(bin) = #keys in bins >= bin; pick the one with S >= krem > S - hist[bin]
        unsigned hbin = 0, suf = 0;
        if (tid < 256) {
            hbin = hist[tid];
            suf = hbin;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned t = __shfl_down_sync(0xffffffffu, suf, off);
                if (lane + off < 32) suf += t;
            }
            if (lane == 0) wcnt[warp] = suf;  // this warp's 32 bins
        }
        __syncthreads();
        if (tid < 256) {
            for (int w2 = warp + 1; w2 < 8; ++w2) suf += wcnt[w2];
            if (suf >= krem && suf - hbin < krem) {  // exactly one bin (at least krem keys match the prefix)
                s_krem = krem - (suf - hbin);
                s_prefix = prefix | ((unsigned)tid << (8 * pass));
            }
        }
        __syncthreads();
        krem = s_krem;
        prefix = s_prefix;
        mask |= 0xFFu << (8 * pass);
    }
    const uint32_t T = prefix;         // key of the K-th largest score
    const unsigned G = K - krem;       // winners strictly above T; krem ties admitted by index

    if (tid == 0) s_cnt = 0;
    for (int i = tid; i < 1024; i += 1024) buf[i] = 0ull;
    __syncthreads();
    if (SM && ncand >= 0) {
        // the list is in index order: winners above the threshold in any order, ties AT the threshold lowest list position first
        unsigned eq_run = 0;
        for (int base = 0; base < NW; base += 1024) {
            const int j = base + tid;
            const bool valid = j < ncand;
            const unsigned i = valid ? cand[j] : 0u;
            const uint32_t key = valid ? skeys[i] : 0u;
            const bool gt = valid && key > T, eq = valid && key == T;
            const unsigned mg = __ballot_sync(0xffffffffu, gt);
            if (mg) {
                unsigned basee = 0;
                if (lane == 0) basee = atomicAdd(&s_cnt, (unsigned)__popc(mg));
                basee = __shfl_sync(0xffffffffu, basee, 0);
                if (gt) buf[basee + __popc(mg & ((1u << lane) - 1u))] = ((unsigned long long)key << 32) | (0xFFFFFFFFu - i);
            }
            const unsigned me = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) wcnt[warp] = __popc(me);
            __syncthreads();
            unsigned woff = 0, tot = 0;
#pragma unroll
            for (int w2 = 0; w2 < 32; ++w2) {
                const unsigned c = wcnt[w2];
                woff += (w2 < warp) ? c : 0u;
                tot += c;
            }
            const unsigned rank = eq_run + woff + __popc(me & ((1u << lane) - 1u));
            if (eq && rank < krem) buf[G + rank] = ((unsigned long long)T << 32) | (0xFFFFFFFFu - i);
            eq_run += tot;
            __syncthreads();
        }
    } else if (SM) {
        // winners above the threshold take slots in any order (they are sorted below); ties AT the threshold are admitted
        // lowest index first: per (row of 1024, warp) tie counts -> block-wide exclusive scan -> rank = base + lane rank
        for (int v = 0; v < rows; ++v) {
            const int i = v * 1024 + tid;
            const bool valid = i < HW;
            const uint32_t key = valid ? skeys[i] : 0u;
            const bool gt = valid && key > T;
            const unsigned mg = __ballot_sync(0xffffffffu, gt);
            if (mg) {
                unsigned basee = 0;
                if (lane == 0) basee = atomicAdd(&s_cnt, (unsigned)__popc(mg));
                basee = __shfl_sync(0xffffffffu, basee, 0);
                if (gt) buf[basee + __popc(mg & ((1u << lane) - 1u))] = ((unsigned long long)key << 32) | (0xFFFFFFFFu - (unsigned)i);
            }
            const unsigned me = __ballot_sync(0xffffffffu, valid && key == T);
            if (lane == 0) ecnt[v * 32 + warp] = __popc(me);
        }
        __syncthreads();
        {
            const unsigned gt_total = s_cnt;  // the scan reports its total through s_cnt: keep the winners' slot counter
            __syncthreads();
            scan_ecnt();
            if (tid == 0) s_cnt = gt_total;
            __syncthreads();
        }
        for (int v = 0; v < rows; ++v) {
            const unsigned base_rank = ecnt[v * 32 + warp];
            if (base_rank >= krem) break;  // ranks only grow with the index (warp-uniform)
            const int i = v * 1024 + tid;
            const bool eq = i < HW && skeys[i] == T;
            const unsigned me = __ballot_sync(0xffffffffu, eq);
            const unsigned rank = base_rank + __popc(me & ((1u << lane) - 1u));
            if (eq && rank < krem) buf[G + rank] = ((unsigned long long)T << 32) | (0xFFFFFFFFu - (unsigned)i);
        }
        __syncthreads();
    } else {
        unsigned eq_run = 0;
        for (int base = 0; base < HW; base += 1024) {
            const int i = base + tid;
            const bool valid = i < HW;
            const uint32_t key = valid ? key_at(i) : 0u;
            const bool gt = valid && key > T;
            const bool eq = valid && key == T;
            if (gt) {
                const unsigned slot = atomicAdd(&s_cnt, 1u);
                buf[slot] = ((unsigned long long)key << 32) | (0xFFFFFFFFu - (unsigned)i);
            }
            const unsigned m = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) wcnt[warp] = __popc(m);
            __syncthreads();
            unsigned woff = 0, tot = 0;
    #pragma unroll
            for (int w2 = 0; w2 < 32; ++w2) {
                const unsigned c = wcnt[w2];
                woff += (w2 < warp) ? c : 0u;
                tot += c;
            }
            const unsigned rank = eq_run + woff + __popc(m & ((1u << lane) - 1u));
            if (eq && rank < krem) buf[G + rank] = ((unsigned long long)key << 32) | (0xFFFFFFFFu - (unsigned)i);
            eq_run += tot;
            __syncthreads();
        }
    }
    int P = 1;
    while (P < K) P <<= 1;
    bitonic_desc(buf, P);

    for (int r = tid; r < K; r += 1024) {
        const unsigned long long c = buf[r];
        const int full = (int)(0xFFFFFFFFu - (unsigned)(c & 0xFFFFFFFFull));
        const int cls = C == 1 ? 0 : full / HW1;           // :21 topk_clses
        const int idx = C == 1 ? full : full - cls * HW1;  // :16 topk_inds % (height * width)
        const float score = fkey_inv((uint32_t)(c >> 32));
        float xs = (float)(idx % W), ys = (float)(idx / W);  // centerface_ext.py:18-19
        if (reg) {
            xs = __fadd_rn(xs, __ldcg(reg + ((size_t)b * 2 + 0) * HW1 + idx));  // :62
            ys = __fadd_rn(ys, __ldcg(reg + ((size_t)b * 2 + 1) * HW1 + idx));  // :63
        } else {
            xs += 0.5f;
            ys += 0.5f;
        }
        const int wp = wh_planes == 2 ? 0 : 2 * cls;  // cat_spec_wh (:72-75): the class's own (w, h) planes of wh [B,2C,H,W]
        const float hw = __ldcg(wh + ((size_t)b * wh_planes + wp + 0) * HW1 + idx) / 2.f;
        const float hh = __ldcg(wh + ((size_t)b * wh_planes + wp + 1) * HW1 + idx) / 2.f;
        float* d = dets + ((size_t)b * K + r) * 6;
        d[0] = __fsub_rn(xs, hw);
        d[1] = __fsub_rn(ys, hh);
        d[2] = __fadd_rn(xs, hw);
        d[3] = __fadd_rn(ys, hh);
        d[4] = score;
        d[5] = (float)cls;
        if (inds) inds[(size_t)b * K + r] = idx;
    }
}

// Maps that fit the shared-memory key cache take ONE launch: k_topk<true> reads the raw heat map and applies the peak keep on
// its shared copy.  Larger maps: k_peak_mask into `scratch`, then k_topk<false> re-reads the masked map from L2 in every pass.
inline bool topk_fused(int C, int H, int W) { return (long long)C * H * W <= TOPK_SMEM_MAX_HW; }
inline cudaError_t launch_topk(const float* heat, float* scratch, const float* wh, const float* reg, int B, int H, int W, int K, float* dets,
                               int32_t* inds, cudaStream_t s, int C = 1, int wh_planes = 2) {
    const int HW = C * H * W;
    const float* pk = heat;
    if (!topk_fused(C, H, W)) {
        const long long n = (long long)B * HW;
        cudaError_t e = launch_pdl(k_peak_mask, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, heat, scratch, B * C, H, W);
        if (e != cudaSuccess) return e;
        pk = scratch;
    }
    if (HW <= TOPK_SMEM_MAX_HW) {
        cudaError_t e = smem_optin((const void*)k_topk<true>, (TOPK_SMEM_MAX_HW + TOPK_SMEM_MAX_HW / 32) * 4);
        if (e != cudaSuccess) return e;
        const size_t HWp = ((size_t)HW + 1023) & ~(size_t)1023;
        // candidate list: what the opted-in shared memory leaves beside the keys, at most 8 192 entries
        const size_t room = ((size_t)TOPK_SMEM_MAX_HW + TOPK_SMEM_MAX_HW / 32 - HWp - HWp / 32);
        const int cand_cap = (int)(room < 8192 ? room : 8192);
        return launch_pdl(k_topk<true>, dim3(B), dim3(1024), (HWp + HWp / 32 + (size_t)cand_cap) * 4, s, pk, wh, reg, H, W, K, dets, inds, cand_cap, C, wh_planes);
    }
    return launch_pdl(k_topk<false>, dim3(B), dim3(1024), 0, s, pk, wh, reg, H, W, K, dets, inds, 0, C, wh_planes);
}

// numpy float32 floor_divide (npy_floor_dividef -> npy_divmodf), used by centerface.py:56-58
__device__ __forceinline__ float np_floordiv(float a, float b) {
    if (b == 0.f) return a / b;
    float mod = fmodf(a, b);
    float div = __fdiv_rn(__fsub_rn(a, mod), b);
    if (mod != 0.f) {
        if ((b < 0.f) != (mod < 0.f)) div = __fsub_rn(div, 1.0f);
    }
    if (div != 0.f) {
        float fl = floorf(div);
        if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.0f);
        return fl;
    }
    return copysignf(0.f, __fdiv_rn(a, b));
}

// ---- K9: threshold decode + greedy NMS (paths A and B), one CTA per image ------------
// dynamic smem: cand u64[capP] | box float4[capP] | area float[capP] | keep int[capP] | supp u8[capP]
constexpr int THRESH_MAX_CAP = 4096;
__host__ __device__ inline size_t thresh_smem_bytes(int capP) { return (size_t)capP * (8 + 16 + 4 + 4 + 1) + 16; }

__global__ void __launch_bounds__(1024) k_thresh_nms(const float* __restrict__ hm, const float* __restrict__ wh,
                                                      const float* __restrict__ reg, const float* __restrict__ lm,
                                                      int H, int W, int variant, float thr, float nms_thr,
                                                      int size_h, int size_w, float scale_w, float scale_h,
                                                      int cap, int capP, float* __restrict__ out_dets,
                                                      float* __restrict__ out_lms, int32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) unsigned char smraw[];
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(smraw);
    float4* box = reinterpret_cast<float4*>(cand + capP);
    float* area = reinterpret_cast<float*>(box + capP);
    int* keep = reinterpret_cast<int*>(area + capP);
    unsigned char* supp = reinterpret_cast<unsigned char*>(keep + capP);
    __shared__ unsigned s_cnt;
    __shared__ int s_nk;
    const int tid = threadIdx.x;
    const int HW = H * W;
    const int b = blockIdx.x;
    const float* hmb = hm + (size_t)b * HW;

    if (tid == 0) s_cnt = 0;
    for (int i = tid; i < capP; i += 1024) cand[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < HW; i += 1024) {
        const float s = __ldg(hmb + i);
        if (s > thr) {  // np.where(heatmap > thr): float32 compare
            const unsigned slot = atomicAdd(&s_cnt, 1u);
            if (slot < (unsigned)cap) cand[slot] = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)i;
        }
    }
    __syncthreads();
    const int n = (int)s_cnt;
    if (n > cap) {
        if (tid == 0) out_counts[b] = -n;
        return;
    }
    int P = 1;
    while (P < n) P <<= 1;
    bitonic_desc(cand, P);

    // boxes in the reference's mixed float32/float64 arithmetic, rounded once to float32
    for (int j = tid; j < n; j += 1024) {
        const unsigned long long c = cand[j];
        const int idx = (int)(c & 0xFFFFFFFFull);
        const int row = idx / W, col = idx % W;
        const float s0 = __fmul_rn(__ldg(wh + ((size_t)b * 2 + 0) * HW + idx), 4.f);
        const float s1 = __fmul_rn(__ldg(wh + ((size_t)b * 2 + 1) * HW + idx), 4.f);
        double cx = (double)col, cy = (double)row;
        if (variant == CF_DECODE_B) {  // eval_widerface.py:102: offsets swapped
            cx += (double)__ldg(reg + ((size_t)b * 2 + 1) * HW + idx);
            cy += (double)__ldg(reg + ((size_t)b * 2 + 0) * HW + idx);
        }
        double x1 = (cx + 0.5) * 4.0 - (double)(s0 / 2.f);
        double y1 = (cy + 0.5) * 4.0 - (double)(s1 / 2.f);
        x1 = x1 > 0.0 ? x1 : 0.0;  // max(0, .)
        y1 = y1 > 0.0 ? y1 : 0.0;
        x1 = (double)size_w < x1 ? (double)size_w : x1;  // min(x1, size[1])
        y1 = (double)size_h < y1 ? (double)size_h : y1;
        double x2 = x1 + (double)s0, y2 = y1 + (double)s1;
        x2 = (double)size_w < x2 ? (double)size_w : x2;
        y2 = (double)size_h < y2 ? (double)size_h : y2;
        const float fx1 = __double2float_rn(x1), fy1 = __double2float_rn(y1);
        const float fx2 = __double2float_rn(x2), fy2 = __double2float_rn(y2);
        box[j] = make_float4(fx1, fy1, fx2, fy2);
        area[j] = __fmul_rn(__fadd_rn(__fsub_rn(fx2, fx1), 1.f), __fadd_rn(__fsub_rn(fy2, fy1), 1.f));
        supp[j] = 0;
    }
    if (tid == 0) s_nk = 0;
    __syncthreads();

    // greedy suppression in score order (centerface.py:123-149)
    for (int i = 0; i < n; ++i) {
        if (supp[i]) continue;  // uniform: last write to supp[i] is behind a barrier
        if (tid == 0) keep[s_nk++] = i;
        const float4 bi = box[i];
        const float ai = area[i];
        for (int j = i + 1 + tid; j < n; j += 1024) {
            if (supp[j]) continue;
            const float4 bj = box[j];
            const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
            const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
            float w = __fadd_rn(__fsub_rn(xx2, xx1), 1.f);
            float h = __fadd_rn(__fsub_rn(yy2, yy1), 1.f);
            w = w > 0.f ? w : 0.f;
            h = h > 0.f ? h : 0.f;
            const float inter = __fmul_rn(w, h);
            const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, area[j]), inter));
            if (ovr >= nms_thr) supp[j] = 1;
        }
        __syncthreads();
    }
    __syncthreads();
    const int nk = s_nk;
    const bool resc = (scale_w != 0.f) && (scale_h != 0.f);
    for (int r = tid; r < nk; r += 1024) {
        const int i = keep[r];
        const unsigned long long c = cand[i];
        const int idx = (int)(c & 0xFFFFFFFFull);
        float4 bx = box[i];
        if (resc) {
            bx.x = np_floordiv(bx.x, scale_w);
            bx.z = np_floordiv(bx.z, scale_w);
            bx.y = np_floordiv(bx.y, scale_h);
            bx.w = np_floordiv(bx.w, scale_h);
        }
        float* d = out_dets + ((size_t)b * cap + r) * 5;
        d[0] = bx.x;
        d[1] = bx.y;
        d[2] = bx.z;
        d[3] = bx.w;
        d[4] = __uint_as_float((uint32_t)(c >> 32));
        if (variant == CF_DECODE_A && out_lms && lm) {
            const int row = idx / W, col = idx % W;
            float* l = out_lms + ((size_t)b * cap + r) * 10;
#pragma unroll
            for (int j = 0; j < 5; ++j) {  // centerface.py:97-98
                const double lx = ((double)__ldg(lm + ((size_t)b * 10 + 2 * j) * HW + idx) + (double)col + 0.5) * 4.0;
                const double ly = ((double)__ldg(lm + ((size_t)b * 10 + 2 * j + 1) * HW + idx) + (double)row + 0.5) * 4.0;
                float fx = __double2float_rn(lx), fy = __double2float_rn(ly);
                if (resc) {
                    fx = np_floordiv(fx, scale_w);
                    fy = np_floordiv(fy, scale_h);
                }
                l[2 * j] = fx;
                l[2 * j + 1] = fy;
            }
        }
    }
    if (tid == 0) out_counts[b] = nk;
}

// ---- CenterFace.nms (centerface.py:111-151 == eval_widerface.py:112-152) as a stand-alone call: greedy IoU suppression of
//      arbitrary float32 boxes [n,4] / scores [n], "+1" areas, order = stable ascending argsort reversed (score desc, index desc),
//      suppress when ovr >= thr.  One CTA; keys / areas / flags live in `scratch` (global, n <= any) -- the n of a face detector
//      (tens to a few thousand) makes this a latency kernel either way.  keep[] = ORIGINAL indices in keep order.
__host__ __device__ inline size_t nms_scratch_bytes(int n) {
    size_t P = 1;
    while ((long long)P < n) P <<= 1;
    return P * 8 + (size_t)n * (4 + 4) + 64;  // u64 keys[P] | float area[n] | int supp[n]
}
__global__ void __launch_bounds__(1024) k_nms_keep(const float* __restrict__ boxes, const float* __restrict__ scores, int n, float thr,
                                                   unsigned char* __restrict__ scratch, int32_t* __restrict__ keep, int32_t* __restrict__ count) {
    const int tid = threadIdx.x;
    int P = 1;
    while (P < n) P <<= 1;
    unsigned long long* key = reinterpret_cast<unsigned long long*>(scratch);
    float* area = reinterpret_cast<float*>(key + P);
    int* supp = reinterpret_cast<int*>(area + n);
    __shared__ int s_nk;
    for (int i = tid; i < P; i += 1024) key[i] = i < n ? (((unsigned long long)fkey(scores[i]) << 32) | (unsigned)i) : 0ull;
    for (int i = tid; i < n; i += 1024) {
        const float4 b = reinterpret_cast<const float4*>(boxes)[i];
        area[i] = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
        supp[i] = 0;
    }
    if (tid == 0) s_nk = 0;
    __syncthreads();
    bitonic_desc(key, P);  // n real keys all carry the fkey sign bit or its complement: they sort ahead of the zero padding
    for (int r = 0; r < n; ++r) {
        const int i = (int)(key[r] & 0xFFFFFFFFull);
        if (supp[i]) continue;  // uniform: the last write to supp[] is behind a barrier
        if (tid == 0) keep[s_nk++] = i;
        const float4 bi = reinterpret_cast<const float4*>(boxes)[i];
        const float ai = area[i];
        for (int q = r + 1 + tid; q < n; q += 1024) {
            const int j = (int)(key[q] & 0xFFFFFFFFull);
            if (supp[j]) continue;
            const float4 bj = reinterpret_cast<const float4*>(boxes)[j];
            const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
            const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
            float w = __fadd_rn(__fsub_rn(xx2, xx1), 1.f);
            float h = __fadd_rn(__fsub_rn(yy2, yy1), 1.f);
            w = w > 0.f ? w : 0.f;
            h = h > 0.f ? h : 0.f;
            const float inter = __fmul_rn(w, h);
            const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, area[j]), inter));
            if (ovr >= thr) supp[j] = 1;
        }
        __syncthreads();
    }
    if (tid == 0) *count = s_nk;
}

// ---- ctdet_post_process (utils/post_process.py:83-100): map the two corners of every [x1,y1,x2,y2,score,cls] row
//      through the per-image inverse affine `trans` (2x3, fp64, built on the host exactly like
//      utils/image.py:27-61 does) in fp64 like np.dot, round once to fp32, keep the score -------------------------
__global__ void __launch_bounds__(256) k_affine_boxes(const float* __restrict__ dets, const double* __restrict__ trans,
                                                      float* __restrict__ out, int B, int K) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * K) return;
    const double* t = trans + (size_t)(i / K) * 6;
    const float* d = dets + (size_t)i * 6;
    float* o = out + (size_t)i * 5;
#pragma unroll
    for (int c = 0; c < 2; ++c) {  // affine_transform(pt, t): t . [x, y, 1] with pt as float32 (utils/image.py:64-67)
        const double x = (double)d[2 * c], y = (double)d[2 * c + 1];
        o[2 * c] = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(t[0], x), __dmul_rn(t[1], y)), t[2]));
        o[2 * c + 1] = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(t[3], x), __dmul_rn(t[4], y)), t[5]));
    }
    o[4] = d[4];
}

}  // namespace cf
