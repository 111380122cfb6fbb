// k_pwn: point-wise convolution for the NARROW layers (column chunk NC <= 64, weight image <= 48 KB: the projection
// convs of the shallow blocks and the FPN laterals).
//
// k_pw_tc's warp-specialised pipeline hands a 16 KB tile through five mbarrier hops (producer thread -> 4 splitter
// warps -> MMA thread -> epilogue group -> ...); on these layers every stage has too little work per tile and the
// per-tile hand-off chain, not HBM or the tensor pipe, sets the pace (profiles/r1_fused_kernels.md, section 4:
// the same matrix streams at 6.2 TB/s through a bare TMA ring, the full GEMM runs at 3.8 TB/s).  Here there are no
// roles: a 256-thread CTA walks over its 128-row tiles and ALL threads take part in every phase --
//   wait for the TMA box (K block of 32)  ->  every thread splits 16 elements of its row into tf32 hi + lo and
//   tcgen05.st's them into a two-slot TMEM ring (lane = row)  ->  one thread issues  A_hi.[B_hi|B_lo]  and  A_lo.B_hi
//   with A from TMEM and the resident weight image from shared memory  ->  after the last K block all threads drain
//   the accumulator pair, apply the epilogue and store their rows --
// with CTA-wide barriers in between, and two such CTAs share an SM so one CTA's TMA / MMA latency is the other's
// compute phase.
#pragma once
#include "k_pw_tc.cuh"

namespace cf {

constexpr int PWN_THREADS = 256;

struct PwnParams {
    const float* bimg;  // [kb][hi NC x 128 B | lo NC x 128 B]
    float* out;
    int M, K, N, NC, nkb, n_tiles, nst;
    int three;  // 1: NC <= 32 -- 128 TMEM columns (accumulator pair + ONE A slot) and a shorter A ring, three CTAs per SM
    uint32_t off_b, off_bars, b_bytes;
    EpiArgs ea;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(PWN_THREADS, 2) k_pwn(const __grid_constant__ CUtensorMap tmA, const PwnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int nst = p.nst, nkb = p.nkb;
    const uint32_t bsm = base + p.off_b;
    const uint32_t bars = base + p.off_bars;  // [0..nst) A full | B full | A-slot free x2 | accumulator ready
    const uint32_t bar_b = bars + 8 * nst, bar_afree = bar_b + 8, bar_acc = bar_afree + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + p.off_bars + 8 * nst + 40);
    // TMEM: accumulator pair in columns [0, 2NC), A ring of (32 hi + 32 lo)-column slots above: two slots from column 128
    // (256 columns, two CTAs per SM) or one slot from column 64 (128 columns, three CTAs per SM)
    const uint32_t kACol = p.three ? 64u : 128u, kTmemCols = p.three ? 128u : 256u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter; which 16 of a K block's 32 elements / which column half
    pdl_trigger();

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        for (int i = 0; i < nst + 4; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // IDAUp: the epilogue's constant vectors (weights, not activations: no dependency on the previous kernel) in shared memory,
    // behind the barriers (N <= 32: 6 N floats fit the barrier block's kilobyte)
    float* cst = (EPI == EPI_IDAUP && p.N <= 32 && p.N % 4 == 0) ? reinterpret_cast<float*>(sm + p.off_bars + 256) : nullptr;
    if (cst) {
        for (int i = tid; i < 6 * p.N; i += PWN_THREADS)
            cst[i] = i < p.N ? __ldg(p.ea.bias + i) : i < 2 * p.N ? __ldg(p.ea.tu + i - p.N) : __ldg(p.ea.su + i - 2 * p.N);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_tiles = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const long long total = (long long)my_tiles * nkb;  // job j = (my tile j / nkb, K block j % nkb)
    auto issue_a = [&](long long j) {  // thread 0 only
        const int stage = (int)(j % nst);
        const int tile = (int)blockIdx.x + (int)(j / nkb) * (int)gridDim.x;
        mbar_expect_tx(bars + 8 * stage, TC_A_BYTES);
        tma_load_2d(base + stage * TC_A_BYTES, &tmA, (int)(j % nkb) * TC_BK, tile * TC_BM, bars + 8 * stage);
    };
    if (tid == 0 && total > 0) {
        mbar_expect_tx(bar_b, p.b_bytes);
        for (uint32_t off = 0; off < p.b_bytes; off += 32768u) {
            const uint32_t n = p.b_bytes - off < 32768u ? p.b_bytes - off : 32768u;
            bulk_load(bsm + off, reinterpret_cast<const uint8_t*>(p.bimg) + off, n, bar_b);
        }
    }
    pdl_wait();  // everything above (barriers, TMEM, weights) depends on nothing; A is the previous kernel's output
    if (tid == 0 && total > 0)
        for (long long j = 0; j < nst && j < total; ++j) issue_a(j);
    const uint32_t idesc = umma_idesc_tf32(p.NC), idesc2 = umma_idesc_tf32(2 * p.NC);
    const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)p.NC;
    const uint32_t blk_bytes = (uint32_t)p.NC * 256u;  // one K block of the weight image (hi | lo)

    float4 lowv[4];  // IDAUp: low-resolution operands of the current tile (fetched at its first K block, used after its last)
    int lowq = 0;
    for (long long j = 0; j < total; ++j) {
        const int kb = (int)(j % nkb);
        const int tile = (int)blockIdx.x + (int)(j / nkb) * (int)gridDim.x;
        const int stage = (int)(j % nst);
        const uint32_t aslot = p.three ? 0u : (uint32_t)(j & 1);
        // ---- split: this thread's 16 elements of row (32q + lane) -> tf32 hi / lo -> TMEM ----
        mbar_wait(bars + 8 * stage, (uint32_t)(j / nst) & 1u);
        const int row = q * 32 + lane;
        // IDAUp: this thread's low-resolution operands are fetched now, so the L2 latency runs under the split -> MMA chain
        // instead of after the accumulator wait (up3: 90 us per launch against 48 us for the same GEMM with a linear epilogue)
        if (EPI == EPI_IDAUP && kb == 0) {
            const int grow0 = tile * TC_BM + row;
            if (grow0 < p.M) {
                const int c00 = half * (p.NC >> 1);
                const float* lp = idaup_low_ptr(grow0, c00, p.N, p.ea, &lowq);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    if (g * 4 < (p.NC >> 1) && c00 + 4 * g < p.N) lowv[g] = ldcg4(lp + 4 * g);
            }
        }
        const uint8_t* ar = sm + stage * TC_A_BYTES + row * 128;
        float hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(ar + (((half * 4 + c) ^ (row & 7)) << 4));
            hi[4 * c] = tf32_hi(v.x), hi[4 * c + 1] = tf32_hi(v.y), hi[4 * c + 2] = tf32_hi(v.z), hi[4 * c + 3] = tf32_hi(v.w);
            lo[4 * c] = v.x - hi[4 * c], lo[4 * c + 1] = v.y - hi[4 * c + 1], lo[4 * c + 2] = v.z - hi[4 * c + 2], lo[4 * c + 3] = v.w - hi[4 * c + 3];
        }
        if (p.three) {
            if (j >= 1) mbar_wait(bar_afree, (uint32_t)(j - 1) & 1u);  // MMAs of job j-1 have read the slot
        } else if (j >= 2) {
            mbar_wait(bar_afree + 8 * aslot, (uint32_t)((j >> 1) - 1) & 1u);  // MMAs of job j-2 have read this slot
        }
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kACol + aslot * 64u + (uint32_t)half * 16u;
        tmem_st16(ta, hi);
        tmem_st16(ta + 32u, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();  // the A block is in TMEM, the smem stage is drained
        if (warp == 0) {  // convergent after the barrier: one elected lane issues (see elect_one)
            if (j == 0) mbar_wait(bar_b, 0);
            tc_fence_after();
            const uint64_t b_hi = umma_desc(bsm + (uint32_t)kb * blk_bytes);
            const uint32_t a_hi = tmem_base + kACol + aslot * 64u, a_lo = a_hi + 32u;
            const int krem = p.K - kb * TC_BK;
            const int nks = krem >= TC_BK ? TC_BK / 8 : (krem + 7) / 8;
            if (elect_one()) {
                if (j + nst < total) issue_a(j + nst);
                if (nkb == 1) {  // one K block: corrections first, then the main products, ONE accumulator (as k_pw_tc's `single`)
                    const uint64_t b_lo = umma_desc(bsm + (uint32_t)kb * blk_bytes + (uint32_t)p.NC * 128u);
                    for (int k = 0; k < nks; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_tf32_ts(d_main, a_lo + 8u * k, b_hi + ko, idesc, k > 0 ? 1u : 0u);
                        umma_tf32_ts(d_main, a_hi + 8u * k, b_lo + ko, idesc, 1u);
                    }
                    for (int k = 0; k < nks; ++k) umma_tf32_ts(d_main, a_hi + 8u * k, b_hi + (uint64_t)(k * 2), idesc, 1u);
                } else if (nks == 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_tf32_ts(d_main, a_hi + 8u * k, b_hi + ko, idesc2, (kb > 0 || k > 0) ? 1u : 0u);  // main += hi.hi ; corr += hi.lo
                        umma_tf32_ts(d_corr, a_lo + 8u * k, b_hi + ko, idesc, 1u);                              // corr += lo.hi
                    }
                } else {
                    for (int k = 0; k < nks; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_tf32_ts(d_main, a_hi + 8u * k, b_hi + ko, idesc2, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_tf32_ts(d_corr, a_lo + 8u * k, b_hi + ko, idesc, 1u);
                    }
                }
                umma_commit(bar_afree + 8 * aslot);
                if (kb == nkb - 1) umma_commit(bar_acc);
            }
            __syncwarp();
        }
        if (kb == nkb - 1) {
            // ---- tile done: drain main + correction, epilogue, store this thread's half of the row ----
            mbar_wait(bar_acc, (uint32_t)(j / nkb) & 1u);
            tc_fence_after();
            const int grow = tile * TC_BM + row;
            const int ncols = p.NC >> 1;  // 16 or 32 columns per thread
            for (int c0 = half * ncols; c0 < (half + 1) * ncols; c0 += 16) {
                float v[16], c[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                tmem_ld16(taddr, v);
                if (nkb > 1) tmem_ld16(taddr + (uint32_t)p.NC, c);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) c[i] = 0.f;
                }
                tmem_ld_wait();
                if (grow < p.M) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n = c0 + 4 * g;
                        if (n < p.N) {
                            float4 o = make_float4(v[4 * g] + c[4 * g], v[4 * g + 1] + c[4 * g + 1], v[4 * g + 2] + c[4 * g + 2],
                                                   v[4 * g + 3] + c[4 * g + 3]);
                            if (EPI == EPI_IDAUP && p.NC <= 32) o = idaup_apply(o, lowv[g], n, lowq, p.N, p.ea, cst);  // one 16-column pass per thread
                            else o = apply_epi<EPI>(o, grow, n, p.N, p.ea);
                            st4(p.out + (size_t)grow * p.N + n, o);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();  // every accumulator read is done before the next tile's first MMA overwrites it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------
struct PwnLaunch {
    CUtensorMap tmA;
    PwnParams p;
    int epi = 0, grid = 0;
    size_t smem = 0;
};

// eligible: 3-pass, one column chunk of <= 64 columns, weight image small enough to sit beside a >= 3 stage A ring
// in a half (or a third) of an SM's shared memory
// Measured (tools/tc_shape_probe.py, same call): K=32,N=16: 179 -> 157 us; K=144,N=24: 159 -> 172 us; K=192,N=32: 53 -> 61 us --
// the CTA-wide barrier per K block costs more than the hand-off chain it removes once a tile has several K blocks, so
// only single-K-block layers (K <= 32: layer0 projection, FPN laterals up2/up3) take this kernel.
inline bool pwn_eligible(const TcLayer& L) {
    // up to three K blocks (K <= 96): with three CTAs per SM and the elect-based issue the role-free kernel now also wins on
    // layer1.0's projection (80 -> 74 us); from five K blocks on the weight image forces two CTAs and the per-K-block CTA
    // barrier loses again (K = 144: 123 -> 187 us).  CF_PWN_NKB overrides (development probe).
    int max_nkb = 3;
    if (const char* ev = getenv("CF_PWN_NKB")) max_nkb = atoi(ev);
    return L.nchunks == 1 && L.NC <= 64 && L.nkb <= max_nkb && L.img_bytes <= 49152;
}

inline int pwn_plan(PwTcState& st, int epi, const float* A, const float* Wkn, float* out, int M, int K, int N, EpiArgs ea, PwnLaunch* pl) {
    auto it = st.layers.find(Wkn);
    if (it == st.layers.end()) return fail(CF_EINVAL, "pwn_plan: weight matrix was not prepared");
    const TcLayer& L = it->second;
    if (!pwn_eligible(L)) return fail(CF_EINVAL, "pwn_plan: layer K=%d N=%d is not a narrow layer", K, N);
    int rc = tc_make_map(st, &pl->tmA, A, (uint64_t)M, (uint64_t)K, TC_BM);
    if (rc) return rc;
    PwnParams& p = pl->p;
    p.bimg = L.img;
    p.out = out;
    p.M = M;
    p.K = K;
    p.N = N;
    p.NC = L.NC;
    p.nkb = L.nkb;
    p.n_tiles = (M + TC_BM - 1) / TC_BM;
    p.b_bytes = (uint32_t)L.img_bytes;
    p.ea = ea;
    const uint32_t b_al = (p.b_bytes + 1023u) & ~1023u;
    // 256 TMEM columns per CTA (accumulator pair + two-slot A ring): two CTAs fill the SM's 512; or 128 columns and three CTAs.
    // Measured: layer0 projection (linear epilogue) 141 -> 115 us with three, the IDAUp laterals 74 -> 84 us (their epilogue
    // wants the registers / L1 of the larger share), so three CTAs only without the IDAUp epilogue.  CF_PWN_CTAS=2|3 overrides.
    p.three = (L.NC <= 32 && epi != EPI_IDAUP) ? 1 : 0;
    if (const char* ev = getenv("CF_PWN_CTAS")) p.three = (atoi(ev) == 3 && L.NC <= 32) ? 1 : 0;
    int ctas = p.three ? 3 : 2;
    int nst = ((int)(TC_SMEM_MAX / ctas) - 1024 - 2048 - (int)b_al) / TC_A_BYTES;
    if (p.three && nst < 3) {  // the weight image leaves no room for a ring in a third of the shared memory
        p.three = 0;
        ctas = 2;
        nst = ((int)(TC_SMEM_MAX / ctas) - 1024 - 2048 - (int)b_al) / TC_A_BYTES;
    }
    if (nst > 6) nst = 6;
    if (nst < 2) return fail(CF_EINVAL, "pwn_plan: does not fit shared memory");
    p.nst = nst;
    p.off_b = (uint32_t)nst * TC_A_BYTES;
    p.off_bars = p.off_b + b_al;
    pl->smem = (size_t)p.off_bars + 1024 + 1024;
    pl->epi = epi;
    const int want = ctas * st.sms;
    pl->grid = p.n_tiles < want ? p.n_tiles : want;
    return CF_OK;
}

template <int EPI>
inline cudaError_t pwn_launch_t(const PwnLaunch& pl, cudaStream_t s) {
    cudaError_t e = smem_optin((const void*)k_pwn<EPI>, (int)(TC_SMEM_MAX / 2));
    if (e != cudaSuccess) return e;
    return launch_pdl(k_pwn<EPI>, dim3(pl.grid), dim3(PWN_THREADS), pl.smem, s, pl.tmA, pl.p);
}

inline cudaError_t pwn_launch(const PwnLaunch& pl, cudaStream_t s) {
    switch (pl.epi) {
        case EPI_LINEAR: return pwn_launch_t<EPI_LINEAR>(pl, s);
        case EPI_SWISH: return pwn_launch_t<EPI_SWISH>(pl, s);
        case EPI_RESIDUAL: return pwn_launch_t<EPI_RESIDUAL>(pl, s);
        case EPI_BIAS_SWISH: return pwn_launch_t<EPI_BIAS_SWISH>(pl, s);
        default: return pwn_launch_t<EPI_IDAUP>(pl, s);
    }
}

}  // namespace cf
