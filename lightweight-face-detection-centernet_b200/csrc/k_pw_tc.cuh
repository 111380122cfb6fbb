// Point-wise (1x1) convolution on the 5th-generation tensor cores:
//     out[M,N] = epi( A[M,K] . W[K,N] ),  A = NHWC fp32 activations (pixels are GEMM rows).
//
// tcgen05.mma kind::tf32, M=128 per CTA, fp32 accumulators in TMEM, operands staged in shared
// memory: A by TMA (cp.async.bulk.tensor, SWIZZLE_128B, zero fill past K and past M), W as a
// pre-swizzled K-major image built once at cf_create (plain cp.async.bulk).
//
// Precision.  The reference is fp32 and its back-bone has no normalisation, so an 11-bit operand
// (one TF32 pass) breaks top-k parity.  kPasses == 3 therefore splits both operands
// into tf32 hi + tf32 lo parts and issues  A_lo.B_hi + A_hi.B_lo  into a correction accumulator and
// A_hi.B_hi into the main one (error ~2^-21 per product, fp32 class; see the MMA warp for why two).  W is split on the host; A is split in
// shared memory by four "splitter" warps between the TMA and the MMA stage (element-wise, so the
// swizzled layout is irrelevant to them).  kPasses == 1 is the throughput mode.
//
// Column chunks NC <= 96: the splitter warps hand the split operand to the tensor core through TENSOR MEMORY instead of
// shared memory: each splitter thread owns one row of the 128 x 32 block, reads it once from the TMA's swizzled image, and
// tcgen05.st's hi and lo halves into a TMEM ring (lane = row, column = k; four slots for NC <= 64, two for NC <= 96); the
// MMAs take A from TMEM ([a_tmem] operand form) and only the weight tile from shared memory, and the smem stage is
// 16 KB (no lo copy) -- which is what gives the wide stride-16/32 layers four pipeline stages instead of two.
//
// Warp roles (512 threads, one persistent CTA per SM):
//   warp 0      TMA producer          warp 1      MMA issuer            (whole warp walks the loop, elect.sync lane issues)
//   warp 2      TMEM allocator        warp 3      idle
//   warps 4-7   A splitters (3-pass)  warps 8-15  two epilogue groups taking alternate items:
//                                                 tcgen05.ld -> epilogue math -> swizzled smem -> TMA store (or row stores)
//
// The launch plan of a layer -- NC, A route, epilogue store kind, weights resident / one resident chunk per CTA / streamed,
// accumulator and staging ring depths -- is tc_plan's cost model overridden by tc_tuned_table (tools/tc_tune.py).
#pragma once
#include <cuda.h>

#include <map>
#include <vector>

#include "common.cuh"

namespace cf {

constexpr int TC_THREADS = 512;
constexpr int TC_BM = 128;           // rows per tile (UMMA M)
constexpr int TC_BK = 32;            // fp32 K elements per smem block = one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_STG_BYTES = 32 * 128;          // one epilogue staging buffer: 32 rows x 32 fp32
constexpr int TC_STG_BUFS = 2;                  // per epilogue warp (TcParams::stg_bufs: 2 or 4)
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_SMEM_MAX = 232448;             // 227 KB

struct TcLayer {        // per weight matrix, built once
    float* img = nullptr;   // device: [chunk][kb][hi NC x 128 B | lo NC x 128 B], SW128 K-major
    int K = 0, N = 0, NC = 0, nchunks = 0, nkb = 0;
    size_t img_bytes = 0;
};

struct PwTcState {
    void* encode = nullptr;  // cuTensorMapEncodeTiled
    int sms = 148;
    std::map<const float*, TcLayer> layers;  // keyed by the [K][N] device weight pointer
    float* heads_img = nullptr;  // tap-major tf32 hi/lo image of the collapsed head conv (k_heads_tc)
    unsigned long long* trace_buf = nullptr;  // development: k_mbf pipeline trace (CF_MBF_TRACE)
    std::map<const float*, float*> dw_imgs;  // per-chunk tap images of the fused blocks' depth-wise weights (k_mbf)
};

struct TcParams {
    const float* bimg;
    int M, K, N, NC, nchunks, nkb, n_items;
    int resident;    // 1: the whole B image lives in smem for the kernel's lifetime; 2: the image of ONE column chunk does --
                     //    the grid is a multiple of nchunks, so a CTA's items all share the chunk blockIdx.x % nchunks
    int direct;      // 0: swizzled smem staging + TMA store; 1: the epilogue stores rows straight from registers (the staging
                     //    buffers' 64 KB go to more pipeline stages); 2: staging tile + coalesced register stores (probe)
    float* out;      // [M][N] (direct stores)
    int nacc;        // TMEM accumulator stages (2 or 4); stage stride = acc_stride columns
    uint32_t acc_stride;
    int atmem;       // 1: the split A operand is handed to the tensor core in TMEM (tcgen05.st by the splitter warps)
                     //    instead of shared memory: NC <= 64 (accumulators in columns [0,256), four A slots above) or
                     //    NC <= 96 (two accumulator pairs in [0,384), two A slots above)
    uint32_t acol;   // first TMEM column of the A ring
    uint32_t amask;  // A ring slots - 1 (1 or 3)
    uint32_t ashift; // log2(A ring slots)
    int stg_bufs;    // staging buffers per epilogue warp (2 or 4)
    int single;      // 1 (3-pass, ONE K block): one accumulator -- the two correction products of every K step first, then the main
                     //    products (as k_mbf's expand: while the accumulator holds only corrections their per-MMA truncation is
                     //    2^-11 smaller) -- so the epilogue reads one accumulator and adds nothing.  The expand layers with K <= 32
                     //    are bound by their two epilogue groups' instruction chains (ncu + instruction count: ~230 per 32 columns)
    int dbg;         // development only (env CF_TC_DEBUG): 1 skip the A split, 2 skip the stores, 4 skip the MMAs
    int stages;
    uint32_t stage_bytes, a_bytes_stage, b_bytes_block;  // b_bytes_block = NC*128*(passes==3?2:1)
    uint32_t off_bres, off_stages, off_bars;             // smem offsets from the 1024-aligned base
    EpiArgs ea;
};

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps after ~4 s (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// (A suspend-time hint on try_wait -- CUTLASS passes 0x989680 -- was tried and measured SLOWER here: +2.6 % on the
// point-wise class; the default try_wait already parks the warp for a short, implementation-defined time.)
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {  // may park the warp for a short, bounded time
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// Bounded wait: a protocol bug traps after ~8 s (surfacing as a CUDA error) instead of hanging the GPU.  The first try is the common
// case; the retry loop is kept to a handful of instructions (a spinning warp shares its scheduler with working ones).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 128; ++i)
            if (mbar_try(bar, parity)) return;
        if (globaltimer_ns() - t0 > 8000000000ull) __trap();
    }
}
// One elected lane of a CONVERGENT warp.  tcgen05.mma / commit / TMA instructions are warp-uniform in SASS: issued under
// `if (lane == 0)` the compiler wraps every one of them in an ELECT + BRA.U.ANY loop with ~10 dependent uniform-datapath
// instructions around it (measured ~75 cycles per MMA in the single issuing thread); under `if (elect_one())` in
// convergent code they are emitted back to back.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)(1024u >> 4) << 32;         // stride byte offset
    d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::tf32: D fp32, A/B tf32 K-major, M=128, N=n
__host__ __device__ inline uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (lane = row, one tf32 per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// round-to-nearest split of an fp32 into tf32 hi (low 13 mantissa bits zero) + exact remainder
__host__ __device__ inline float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
#endif
}

template <int EPI>
__device__ __forceinline__ float4 tc_epi(float4 acc, int m, int n, int N, const EpiArgs& ea) {
    return apply_epi<EPI>(acc, m, n, N, ea);
}

// ---- the kernel -------------------------------------------------------------------------
template <int kPasses, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_pw_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmOut, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();

    // smem map: [staging 16 x 2 x 4 KB][resident B][stages][barriers]
    const uint32_t stg_base = base;
    const uint32_t bres = base + p.off_bres;
    const uint32_t stages0 = base + p.off_stages;
    const uint32_t bars = base + p.off_bars;
    const uint32_t bar_full = bars;                          // [stages]
    const uint32_t bar_ready = bars + 8 * TC_MAX_STAGES;     // [stages] splitters -> MMA
    const uint32_t bar_empty = bars + 16 * TC_MAX_STAGES;    // [stages]
    const uint32_t bar_tfull = bars + 24 * TC_MAX_STAGES;    // [4]
    const uint32_t bar_tempty = bar_tfull + 32;              // [4]
    const uint32_t bar_bres = bar_tempty + 32;               // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + p.off_bars + 24 * TC_MAX_STAGES + 80);
    const uint32_t bar_aready = bars + 24 * TC_MAX_STAGES + 96, bar_aempty = bar_aready + 32;  // [4] each: TMEM A ring (atmem)
    const uint32_t kATmemCol = p.acol;  // accumulators below, the A ring of (32 hi + 32 lo)-column slots from here

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_ready + 8 * s, 4);
            mbar_init(bar_empty + 8 * s, (p.atmem && p.resident) ? 4 : 1);  // atmem: the splitters free the smem stage
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(bar_aready + 8 * a, 4);
            mbar_init(bar_aempty + 8 * a, 1);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);
        }
        mbar_init(bar_bres, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.n_items;  // item = m_tile * nchunks + chunk
    const int nkb = p.nkb;

    if (warp == 0) {
        // ================= TMA producer: the whole warp walks the loop, one elected lane issues =================
        if (p.resident) {
            const uint32_t chunk_bytes = (uint32_t)nkb * p.b_bytes_block;
            const uint32_t total = p.resident == 2 ? chunk_bytes : (uint32_t)p.nchunks * chunk_bytes;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.bimg) +
                                 (p.resident == 2 ? (size_t)(blockIdx.x % (unsigned)p.nchunks) * chunk_bytes : (size_t)0);
            if (elect_one()) {
                mbar_expect_tx(bar_bres, total);
                for (uint32_t off = 0; off < total; off += 32768u) {
                    const uint32_t n = total - off < 32768u ? total - off : 32768u;
                    bulk_load(bres + off, src + off, n, bar_bres);
                }
            }
            __syncwarp();
        }
        pdl_wait();  // the weights above depend on nothing; A is the previous kernel's output
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int mt = item / p.nchunks, ch = item - mt * p.nchunks;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sa = stages0 + stage * p.stage_bytes;
                    const uint32_t tx = TC_A_BYTES + (p.resident ? 0u : p.b_bytes_block);
                    if (elect_one()) {
                        mbar_expect_tx(bar_full + 8 * stage, tx);
                        tma_load_2d(sa, &tmA, kb * TC_BK, mt * TC_BM, bar_full + 8 * stage);
                        if (!p.resident)
                            bulk_load(sa + p.a_bytes_stage,
                                      reinterpret_cast<const uint8_t*>(p.bimg) + (size_t)(ch * nkb + kb) * p.b_bytes_block,
                                      p.b_bytes_block, bar_full + 8 * stage);
                    }
                    __syncwarp();
                    if (++stage == p.stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp walks the loop, one elected lane issues =================
        {
            const uint32_t idesc = umma_idesc_tf32(p.NC);
            if (p.resident) mbar_wait(bar_bres, 0);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0, acnt = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int mt = item / p.nchunks, ch = item - mt * p.nchunks;
                (void)mt;
                const uint32_t as = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1u;
                mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
                tc_fence_after();
                // 3-pass: the two small cross terms go to their own accumulator (columns +NC).  The
                // tensor core truncates its fp32 accumulator once per MMA (measured -2^-24 per step,
                // tools/tc_accum_probe.py); keeping the 2K/8 correction steps out of the main sum
                // leaves it K/8 truncations instead of 3K/8, and the corrections' own truncation
                // is 2^-11 smaller.  The epilogue adds the two in fp32 (round-to-nearest).
                // The correction accumulator sits right behind the main one (columns [NC, 2NC)) and the weight image
                // stores the hi block right before the lo block, so  A_hi . [B_hi | B_lo]  is ONE MMA of width 2NC
                // filling both accumulators; only  A_lo . B_hi  needs a second one.
                const uint32_t d_tmem = tmem_base + as * p.acc_stride;  // atmem: accumulators below acol, A ring above
                const uint32_t d_corr = d_tmem + (uint32_t)p.NC;
                const uint32_t idesc2 = umma_idesc_tf32(2 * p.NC);
                for (int kb = 0; kb < nkb; ++kb) {
                    const int krem = p.K - kb * TC_BK;
                    const int nks = krem >= TC_BK ? TC_BK / 8 : (krem + 7) / 8;
                    const uint32_t first = kb > 0 ? 1u : 0u;
                    if (kPasses == 3 && p.atmem) {
                        const uint32_t aslot = acnt & p.amask;
                        mbar_wait(bar_aready + 8 * aslot, (acnt >> p.ashift) & 1u);
                        tc_fence_after();
                        const uint32_t sb = p.resident ? bres + (uint32_t)((p.resident == 2 ? 0 : ch * nkb) + kb) * p.b_bytes_block
                                                       : stages0 + stage * p.stage_bytes + p.a_bytes_stage;
                        const uint64_t b_hi = umma_desc(sb);
                        const uint32_t a_hi = tmem_base + kATmemCol + aslot * 64u, a_lo = a_hi + 32u;
                        if (elect_one()) {
                            if (p.single) {  // nkb == 1: corrections first, then the main products, one accumulator
                                const uint64_t b_lo = umma_desc(sb + (uint32_t)p.NC * 128u);
                                for (int k = 0; k < nks; ++k) {
                                    umma_tf32_ts(d_tmem, a_lo + 8u * k, b_hi + (uint64_t)(k * 2), idesc, k > 0 ? 1u : 0u);
                                    umma_tf32_ts(d_tmem, a_hi + 8u * k, b_lo + (uint64_t)(k * 2), idesc, 1u);
                                }
                                for (int k = 0; k < nks; ++k) umma_tf32_ts(d_tmem, a_hi + 8u * k, b_hi + (uint64_t)(k * 2), idesc, 1u);
                            } else if (nks == 4 && !(p.dbg & 4)) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma_tf32_ts(d_tmem, a_hi + 8u * k, b_hi + (uint64_t)(k * 2), idesc2, k > 0 ? 1u : first);  // main += hi.hi ; corr += hi.lo
                                    umma_tf32_ts(d_corr, a_lo + 8u * k, b_hi + (uint64_t)(k * 2), idesc, 1u);                   // corr += lo.hi
                                }
                            } else {
                                for (int k = 0; k < nks; ++k) {
                                    if ((p.dbg & 4) && !(kb == 0 && k == 0)) break;
                                    umma_tf32_ts(d_tmem, a_hi + 8u * k, b_hi + (uint64_t)(k * 2), idesc2, k > 0 ? 1u : first);
                                    umma_tf32_ts(d_corr, a_lo + 8u * k, b_hi + (uint64_t)(k * 2), idesc, 1u);
                                }
                            }
                            umma_commit(bar_aempty + 8 * aslot);
                            if (!p.resident) umma_commit(bar_empty + 8 * stage);
                            if (kb == nkb - 1) umma_commit(bar_tfull + 8 * as);
                        }
                        __syncwarp();
                        ++acnt;
                        if (++stage == p.stages) stage = 0, phase ^= 1;
                        continue;
                    }
                    mbar_wait((kPasses == 3 ? bar_ready : bar_full) + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = stages0 + stage * p.stage_bytes;
                    const uint32_t sb = p.resident ? bres + (uint32_t)((p.resident == 2 ? 0 : ch * nkb) + kb) * p.b_bytes_block
                                                   : sa + p.a_bytes_stage;
                    const uint64_t a_hi = umma_desc(sa), b_hi = umma_desc(sb);
                    const uint64_t a_lo = umma_desc(sa + TC_A_BYTES);
                    if (elect_one()) {
                        if (kPasses == 3 && p.single) {  // same order as the TMEM-A route: bit-identical results
                            const uint64_t b_lo = umma_desc(sb + (uint32_t)p.NC * 128u);
                            for (int k = 0; k < nks; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);
                                umma_tf32(d_tmem, a_lo + ko, b_hi + ko, idesc, k > 0 ? 1u : 0u);
                                umma_tf32(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                            }
                            for (int k = 0; k < nks; ++k) umma_tf32(d_tmem, a_hi + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc, 1u);
                        } else if (nks == 4 && !(p.dbg & 4)) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);  // +32 bytes along K, in 16-byte units
                                if (kPasses == 3) {
                                    umma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc2, k > 0 ? 1u : first);  // main += hi.hi ; corr += hi.lo
                                    umma_tf32(d_corr, a_lo + ko, b_hi + ko, idesc, 1u);                   // corr += lo.hi
                                } else {
                                    umma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc, k > 0 ? 1u : first);
                                }
                            }
                        } else {
                            for (int k = 0; k < nks; ++k) {
                                if ((p.dbg & 4) && !(kb == 0 && k == 0)) break;
                                const uint64_t ko = (uint64_t)(k * 2);
                                if (kPasses == 3) {
                                    umma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc2, k > 0 ? 1u : first);
                                    umma_tf32(d_corr, a_lo + ko, b_hi + ko, idesc, 1u);
                                } else {
                                    umma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc, k > 0 ? 1u : first);
                                }
                            }
                        }
                        umma_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
                        if (kb == nkb - 1) umma_commit(bar_tfull + 8 * as);
                    }
                    __syncwarp();
                    if (++stage == p.stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= A splitters (3-pass only) =================
        if (kPasses == 3) {
            const int t = threadIdx.x - 128;  // 0..127
            int stage = 0;
            uint32_t phase = 0, acnt = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                for (int kb = 0; kb < nkb; ++kb) {
                    if (p.atmem) {
                        // thread = one row of the 128 x 32 A block: read it from the TMA's swizzled image, split it and
                        // store both halves into this warp's TMEM lane quarter
                        const uint32_t aslot = acnt & p.amask;
                        mbar_wait(bar_full + 8 * stage, phase);
                        const int q = warp & 3, row = q * 32 + lane;
                        const uint8_t* ar = base_ptr + p.off_stages + stage * p.stage_bytes + row * 128;
                        float hi[32], lo[32];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (p.dbg & 1) break;  // development probe: garbage operand, no smem reads / split math
                            const float4 v = *reinterpret_cast<const float4*>(ar + ((j ^ (row & 7)) << 4));
                            hi[4 * j] = tf32_hi(v.x), hi[4 * j + 1] = tf32_hi(v.y), hi[4 * j + 2] = tf32_hi(v.z), hi[4 * j + 3] = tf32_hi(v.w);
                            lo[4 * j] = v.x - hi[4 * j], lo[4 * j + 1] = v.y - hi[4 * j + 1], lo[4 * j + 2] = v.z - hi[4 * j + 2],
                                   lo[4 * j + 3] = v.w - hi[4 * j + 3];
                        }
                        mbar_wait(bar_aempty + 8 * aslot, ((acnt >> p.ashift) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kATmemCol + aslot * 64u;
                        if (!(p.dbg & 8)) {  // development probe 8: no TMEM stores either
                            tmem_st32(ta, hi);
                            tmem_st32(ta + 32u, lo);
                            tmem_st_wait();
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(bar_aready + 8 * aslot);
                            if (p.resident) mbar_arrive(bar_empty + 8 * stage);  // the smem stage can be refilled already
                        }
                        ++acnt;
                        if (++stage == p.stages) stage = 0, phase ^= 1;
                        continue;
                    }
                    mbar_wait(bar_full + 8 * stage, phase);
                    float4* a = reinterpret_cast<float4*>(base_ptr + p.off_stages + stage * p.stage_bytes);
                    float4* l = a + TC_A_BYTES / 16;
#pragma unroll
                    for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
                        if (p.dbg & 1) break;
                        const float4 v = a[t + i * 128];
                        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                        a[t + i * 128] = h;
                        l[t + i * 128] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
                    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_ready + 8 * stage);
                    if (++stage == p.stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue: group g serves accumulator stage g =================
        pdl_wait();  // residual / low-res reads and the output stores touch activation memory
        const int g = (warp - 8) >> 2;
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const uint32_t stg = stg_base + (uint32_t)(warp - 8) * ((uint32_t)p.stg_bufs * TC_STG_BYTES);
        uint8_t* stg_ptr = base_ptr + (size_t)(warp - 8) * ((size_t)p.stg_bufs * TC_STG_BYTES);
        int buf = 0;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            if ((int)(it & 1u) != g) continue;  // the two groups take alternate items (any accumulator stage)
            const uint32_t as = it % (uint32_t)p.nacc;
            const int mt = item / p.nchunks, ch = item - mt * p.nchunks;
            const uint32_t aphase = (it / (uint32_t)p.nacc) & 1u;
            mbar_wait(bar_tfull + 8 * as, aphase);
            tc_fence_after();
            const int row = mt * TC_BM + q * 32 + lane;
            const int ncb = p.NC >> 5;
            for (int cb = 0; cb < ncb; ++cb) {
                const int col0 = ch * p.NC + cb * 32;
                if (col0 >= p.N) break;  // padded columns of the last chunk
                float v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (as * p.acc_stride + (uint32_t)(cb * 32));
                tmem_ld32(taddr, v);
                if (kPasses == 3 && !p.single) {
                    float c[32];
                    tmem_ld32(taddr + (uint32_t)p.NC, c);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += c[j];
                } else {
                    tmem_ld_wait();
                }
                if (cb == ncb - 1 || col0 + 32 >= p.N) {
                    // every TMEM read of this accumulator is done: hand it back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
                }
                if (p.dbg & 2) continue;
                if (p.direct == 1) {
                    // one thread = one output row: up to 128 contiguous bytes per 32-column block
                    if (row < p.M) {
                        float* orow = p.out + (size_t)row * p.N + col0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int n = col0 + 4 * j;
                            if (n < p.N) {
                                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                                o = tc_epi<EPI>(o, row, n, p.N, p.ea);
                                st4(orow + 4 * j, o);
                            }
                        }
                    }
                    continue;
                }
                if (p.direct != 2) {
                    if (elect_one()) {
                        if (p.stg_bufs == 4) bulk_wait_read<3>();
                        else bulk_wait_read<1>();
                    }  // the staging buffer we are about to overwrite (bulk groups
                                                                         // are per thread: the elected lane is the same one each time)
                }
                __syncwarp();
                uint8_t* sp = stg_ptr + buf * TC_STG_BYTES + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    const int n = col0 + 4 * j;
                    if (EPI != EPI_LINEAR && EPI != EPI_SWISH && EPI != EPI_SWISHQ) {
                        if (row < p.M && n < p.N) o = tc_epi<EPI>(o, row, n, p.N, p.ea);
                    } else {
                        o = tc_epi<EPI>(o, row, n, p.N, p.ea);
                    }
                    *reinterpret_cast<float4*>(sp + ((j ^ (lane & 7)) << 4)) = o;  // SWIZZLE_128B
                }
                if (p.direct == 2) {
                    // coalesced register stores out of the transposed tile: one STG.128 writes 4 rows x 128 contiguous bytes,
                    // and the warp does not have to wait for a TMA store to drain its staging buffer
                    __syncwarp();
                    const int row0 = mt * TC_BM + q * 32;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + (lane >> 3), c = lane & 7;
                        const float4 x = *reinterpret_cast<const float4*>(stg_ptr + buf * TC_STG_BYTES + r * 128 + ((c ^ (r & 7)) << 4));
                        if (row0 + r < p.M && col0 + 4 * c < p.N) st4(p.out + (size_t)(row0 + r) * p.N + col0 + 4 * c, x);
                    }
                    buf = (buf + 1) & (p.stg_bufs - 1);
                    continue;
                }
                fence_proxy_async();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&tmOut, stg + buf * TC_STG_BYTES, col0, mt * TC_BM + q * 32);  // clips rows >= M, cols >= N
                    bulk_commit();
                }
                buf = (buf + 1) & (p.stg_bufs - 1);
            }
        }
        __syncwarp();
        if (elect_one()) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ---- development probe: how fast can ONE thread per SM stream a [M][K] fp32 matrix through a TMA ring? ------------
// (tools/tma_probe.py; no consumer work at all: wait for a box, re-arm the slot, issue the next box)
__global__ void __launch_bounds__(32, 1) k_tma_probe(const __grid_constant__ CUtensorMap tmA, int n_tiles, int nkb, int stages,
                                                     int box_rows, uint32_t box_bytes) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + (uint32_t)stages * box_bytes;
    if (threadIdx.x != 0) return;
    for (int s = 0; s < stages; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
    const long long my = ((long long)n_tiles - 1 - blockIdx.x) / gridDim.x + 1;  // tiles of this CTA (grid <= n_tiles)
    const long long total = my * nkb;
    auto issue = [&](long long j) {
        const int s = (int)(j % stages);
        const long long tile = blockIdx.x + (j / nkb) * (long long)gridDim.x;
        mbar_expect_tx(bars + 8 * s, box_bytes);
        tma_load_2d(base + s * box_bytes, &tmA, (int)(j % nkb) * TC_BK, (int)tile * box_rows, bars + 8 * s);
    };
    for (long long j = 0; j < stages && j < total; ++j) issue(j);
    for (long long j = 0; j < total; ++j) {
        mbar_wait(bars + 8 * (int)(j % stages), (uint32_t)(j / stages) & 1u);
        if (j + stages < total) issue(j + stages);
    }
}

// ---- host side --------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int pw_tc_init(PwTcState& st, int device) {
    cudaDriverEntryPointQueryResult qr;
    void* fn = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    if (e != cudaSuccess || fn == nullptr) return fail(CF_ECUDA, "cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    st.encode = fn;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) st.sms = prop.multiProcessorCount;
    return CF_OK;
}

inline void pw_tc_destroy(PwTcState& st) {
    for (auto& kv : st.layers)
        if (kv.second.img) cudaFree(kv.second.img);
    st.layers.clear();
    for (auto& kv : st.dw_imgs)
        if (kv.second) cudaFree(kv.second);
    st.dw_imgs.clear();
    if (st.trace_buf) cudaFree(st.trace_buf);
    st.trace_buf = nullptr;
    if (st.heads_img) cudaFree(st.heads_img);
    st.heads_img = nullptr;
}

// fp32 [rows][cols] row-major tensor, box [box_rows][32 floats], SWIZZLE_128B, zero OOB fill
inline int tc_make_map(PwTcState& st, CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((PFN_encodeTiled)st.encode)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box,
                                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CF_ECUDA, "cuTensorMapEncodeTiled(rows=%llu, cols=%llu) failed with CUresult %d",
                                       (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return CF_OK;
}

// Per-layer plan overrides.  Defaults come from the cost model below; `tc_tuned_table` holds the (K, N) entries that a
// sweep on the device (tools/tc_tune.py, batch 32 @ 640x640) found faster than the model's choice; the CF_TC_* environment
// variables override both and exist only for that sweep.
struct TcTune {
    int nc = 0;       // column chunk width (0: cost model)
    int atmem = -1;   // A operand through TMEM (-1: NC <= 64)
    int direct = -1;  // epilogue stores rows straight from registers (-1: N <= 64)
    int nacc = 0;     // accumulator stages (0: as many as fit, 2 or 4)
    int pwn = -1;     // role-free kernel for eligible layers (-1: yes)
    int grid = 0;     // CTA count cap (0: one per SM)
    int rchunk = -1;  // per-CTA resident column chunk when the whole weight image does not fit (-1: yes)
    int stg = 0;      // staging buffers per epilogue warp (0: two)
};
struct TcTuneEntry {
    int K, N;
    TcTune t;
};
inline const TcTuneEntry* tc_tuned_table(int* n);  // defined at the end of this file

inline TcTune tc_tune_for(int K, int N, int passes) {
    TcTune t;
    if (passes == 3) {
        int n = 0;
        const TcTuneEntry* tab = tc_tuned_table(&n);
        for (int i = 0; i < n; ++i)
            if (tab[i].K == K && tab[i].N == N) t = tab[i].t;
    }
    if (const char* ev = getenv("CF_TC_TABLE")) {
        if (atoi(ev) == 0) t = TcTune();
    }
    if (const char* ev = getenv("CF_TC_NC")) t.nc = atoi(ev);
    if (const char* ev = getenv("CF_TC_ATMEM")) t.atmem = atoi(ev);
    if (const char* ev = getenv("CF_TC_DIRECT")) t.direct = atoi(ev);
    if (const char* ev = getenv("CF_TC_NACC")) t.nacc = atoi(ev);
    if (const char* ev = getenv("CF_PWN")) t.pwn = atoi(ev);
    if (const char* ev = getenv("CF_TC_GRID")) t.grid = atoi(ev);
    if (const char* ev = getenv("CF_TC_RCHUNK")) t.rchunk = atoi(ev);
    if (const char* ev = getenv("CF_TC_STG")) t.stg = atoi(ev);
    const int nc_max = passes == 3 ? 128 : 192;
    if (t.nc < 0 || t.nc % 32 != 0 || t.nc > nc_max) t.nc = 0;
    return t;
}

// Column-chunk width NC: a multiple of 32 (epilogue blocks).  One TMEM accumulator stage is 256
// columns; the 3-pass mode keeps two accumulators per stage (main + correction), so NC <= 128 there.
// Cost model: MMA/epilogue column work incl. padding + re-reading the A tile once per chunk.
inline void tc_choose_chunks(int K, int N, int passes, int* NC, int* nchunks) {
    const int n32 = (N + 31) / 32;
    const int max_blocks = passes == 3 ? 4 : 6;
    double best = 1e30;
    for (int blocks = 1; blocks <= max_blocks; ++blocks) {
        const int chunks = (n32 + blocks - 1) / blocks;
        const double cost = (double)chunks * blocks + (double)chunks * K / 64.0;
        if (cost < best - 1e-9) best = cost, *NC = blocks * 32, *nchunks = chunks;
    }
}

// Build (once per weight matrix) the tf32 hi/lo, K-major, 128B-swizzled image of W[K][N] (host copy `hw`).
inline int tc_prepare_layer(PwTcState& st, const float* key, const float* hw, int K, int N, int passes, int force_nc = 0) {
    if (st.layers.count(key)) return CF_OK;
    TcLayer L;
    L.K = K;
    L.N = N;
    tc_choose_chunks(K, N, passes, &L.NC, &L.nchunks);
    // One K block (K <= 32) runs with a single accumulator (TcParams::single), so a chunk may be up to 192 columns wide: the expand
    // layers 24 -> 144 and 32 -> 192 become ONE item per 128-row tile instead of two or three (A read once, a third of the hand-offs,
    // no padded columns).  CF_TC_WIDE=0 keeps the tuned narrow chunks.
    bool wide = passes == 3 && K <= TC_BK && !force_nc && N > 96 && N <= 192;
    if (const char* ev = getenv("CF_TC_WIDE")) wide = wide && atoi(ev) != 0;
    if (!force_nc) force_nc = tc_tune_for(K, N, passes).nc;
    if (wide) force_nc = (N + 31) / 32 * 32;
    if (force_nc) L.NC = force_nc, L.nchunks = (N + force_nc - 1) / force_nc;
    L.nkb = (K + TC_BK - 1) / TC_BK;
    const size_t blk = (size_t)L.NC * 128;  // bytes of one hi (or lo) block
    L.img_bytes = (size_t)L.nchunks * L.nkb * blk * 2;
    std::vector<float> img(L.img_bytes / 4, 0.f);
    for (int ch = 0; ch < L.nchunks; ++ch)
        for (int kb = 0; kb < L.nkb; ++kb) {
            float* hi = img.data() + ((size_t)(ch * L.nkb + kb) * 2) * (blk / 4);
            float* lo = hi + blk / 4;
            for (int r = 0; r < L.NC; ++r) {
                const int n = ch * L.NC + r;
                if (n >= N) continue;
                for (int kk = 0; kk < TC_BK; ++kk) {
                    const int k = kb * TC_BK + kk;
                    if (k >= K) continue;
                    const float w = hw[(size_t)k * N + n];
                    const float h = tf32_hi(w);
                    const int pos = r * 32 + (((kk >> 2) ^ (r & 7)) << 2) + (kk & 3);  // SWIZZLE_128B
                    hi[pos] = h;
                    lo[pos] = tf32_hi(w - h);
                }
            }
        }
    if (cudaMalloc((void**)&L.img, L.img_bytes) != cudaSuccess) return fail(CF_ECUDA, "tc_prepare_layer: cudaMalloc failed");
    if (cudaMemcpy(L.img, img.data(), L.img_bytes, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(CF_ECUDA, "tc_prepare_layer: cudaMemcpy failed");
    st.layers[key] = L;
    return CF_OK;
}

struct TcLaunch {  // everything a launch needs, resolved at plan-build time
    CUtensorMap tmA, tmOut;
    TcParams p;
    int grid = 0;
    size_t smem = 0;
    int passes = 3, epi = 0;
};

inline int tc_plan(PwTcState& st, int passes, int epi, const float* A, const float* Wkn, float* out, int M, int K, int N, EpiArgs ea,
                   TcLaunch* tl) {
    auto it = st.layers.find(Wkn);
    if (it == st.layers.end()) return fail(CF_EINVAL, "tc_plan: weight matrix was not prepared");
    const TcLayer& L = it->second;
    int rc;
    if ((rc = tc_make_map(st, &tl->tmA, A, (uint64_t)M, (uint64_t)K, TC_BM))) return rc;
    if ((rc = tc_make_map(st, &tl->tmOut, out, (uint64_t)M, (uint64_t)N, 32))) return rc;
    TcParams& p = tl->p;
    p.bimg = L.img;
    p.M = M;
    p.K = K;
    p.N = N;
    p.NC = L.NC;
    p.nchunks = L.nchunks;
    p.nkb = L.nkb;
    p.n_items = ((M + TC_BM - 1) / TC_BM) * L.nchunks;
    p.ea = ea;
    const uint32_t hl = passes == 3 ? 2u : 1u;
    // NOTE: the image always stores hi|lo pairs; one-pass mode addresses only the hi halves, so its
    // "block" stride is still the pair.
    p.b_bytes_block = (uint32_t)L.NC * 128u * 2u;
    // narrow layers: A operand through TMEM (both accumulator pairs fit columns [0,256), the A ring sits above them)
    const TcTune tune = tc_tune_for(K, N, passes);
    const bool wide = passes == 3 && L.nkb == 1 && L.NC > 96;  // single accumulator, one wide chunk (tc_prepare_layer)
    p.atmem = (passes == 3 && L.NC <= 64) ? 1 : 0;
    if (tune.atmem >= 0) p.atmem = (tune.atmem != 0 && passes == 3 && L.NC <= 96) ? 1 : 0;
    if (wide) p.atmem = 1;
    p.a_bytes_stage = p.atmem ? TC_A_BYTES : TC_A_BYTES * hl;
    // TMEM budget (512 columns): NC <= 64: accumulators [0,256) + four A slots; NC <= 96, or NC <= 64 with THREE accumulator
    // pairs (tune.nacc == 3: single-K-block layers, where the accumulator round trip bounds the item rate): [0,384) + two A slots
    const bool three = p.atmem && tune.nacc == 3 && L.NC <= 64;
    p.acol = (L.NC <= 64 && !three) ? 256u : 384u;
    p.amask = (L.NC <= 64 && !three) ? 3u : 1u;
    p.ashift = (L.NC <= 64 && !three) ? 2u : 1u;
    {   // accumulator ring: as many (main+correction) pairs as fit the accumulator columns, 2 or 4
        const uint32_t acc_cols = p.atmem ? p.acol : 512u, pair = ((passes == 3 && !wide) ? 2u : 1u) * (uint32_t)L.NC;
        p.nacc = (4u * pair <= acc_cols) ? 4 : 2;
        if (tune.nacc) p.nacc = tune.nacc == 4 && 4u * pair <= acc_cols ? 4 : 2;
        if (three && 3u * pair <= acc_cols) p.nacc = 3;
        p.acc_stride = p.nacc == 3 ? 128u : acc_cols / (uint32_t)p.nacc;
    }
    p.direct = (N <= 16) ? 1 : 0;  // since the elect-based issue the TMA-store epilogue wins from N = 24 up (r2r sweep)
    if (tune.direct >= 0) p.direct = tune.direct;  // 0 TMA stores, 1 row stores from registers, 2 transposed tile + coalesced stores
    p.out = out;
    p.single = (passes == 3 && L.nkb == 1) ? 1 : 0;
    p.dbg = 0;
    if (const char* ev = getenv("CF_TC_DEBUG")) p.dbg = atoi(ev);
    const uint32_t bar_bytes = 1024;
    const uint32_t b_total = (uint32_t)L.img_bytes;
    const uint32_t chunk_total = (uint32_t)L.nkb * p.b_bytes_block;
    uint32_t stg_bytes = 0, b_res = 0;
    int stages = 0;
    for (p.stg_bufs = tune.stg == 4 ? 4 : TC_STG_BUFS; p.stg_bufs >= TC_STG_BUFS; p.stg_bufs -= 2) {  // 4 staging buffers only if they fit
        stg_bytes = p.direct == 1 ? 0u : 8u * (uint32_t)p.stg_bufs * TC_STG_BYTES;  // 8 epilogue warps
        const uint32_t avail = TC_SMEM_MAX - 1024 /*alignment slack*/ - stg_bytes - bar_bytes;
        p.resident = (b_total <= 65536u && b_total + 3u * p.a_bytes_stage <= avail) ? 1 : 0;
        // Wide layers whose image does not fit: if one column chunk's image fits beside two A stages, pin each CTA to one
        // chunk (grid = a multiple of nchunks) and keep that chunk resident instead of streaming it again for every tile.
        b_res = p.resident ? b_total : 0u;
        if (!p.resident && tune.rchunk != 0 && L.nchunks > 1 && L.nchunks <= st.sms && chunk_total + 2u * p.a_bytes_stage <= avail) {
            p.resident = 2;
            b_res = chunk_total;
        }
        p.stage_bytes = p.a_bytes_stage + (p.resident ? 0u : p.b_bytes_block);
        stages = (int)((avail - b_res) / p.stage_bytes);
        if (stages >= 2) break;
    }
    if (p.stg_bufs < TC_STG_BUFS) p.stg_bufs = TC_STG_BUFS;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (stages < 2) return fail(CF_EINVAL, "tc_plan: K=%d N=%d does not fit the shared-memory pipeline", K, N);
    p.stages = stages;
    p.off_bres = stg_bytes;
    p.off_stages = stg_bytes + b_res;
    p.off_bars = p.off_stages + (uint32_t)stages * p.stage_bytes;
    tl->smem = (size_t)p.off_bars + bar_bytes + 1024;
    tl->grid = p.n_items < st.sms ? p.n_items : st.sms;
    if (p.resident == 2) tl->grid = tl->grid / L.nchunks * L.nchunks;  // n_items is a multiple of nchunks
    if (tune.grid > 0 && tune.grid < tl->grid) tl->grid = tune.grid;
    tl->passes = passes;
    // the one-K-block expand layers (24 -> 144: 115 -> 109 us, 32 -> 192: 44 -> 41) are MUFU-bound in the epilogue: one reciprocal
    // per four values there (a separate instantiation: as a run-time switch it cost every Swish layer 10 %); CF_TC_SWISHQ=0 = off
    // (CF_TC_SWISHQ=2: every 3-pass Swish layer -- probe)
    static const int swq = [] { const char* ev = getenv("CF_TC_SWISHQ"); return ev ? atoi(ev) : 1; }();
    tl->epi = (epi == EPI_SWISH && passes == 3 && ((p.single && swq == 1) || swq == 2)) ? (int)EPI_SWISHQ : epi;
    return CF_OK;
}

template <int kPasses, int EPI>
inline cudaError_t tc_launch_t(const TcLaunch& tl, cudaStream_t s) {
    cudaError_t e = smem_optin((const void*)k_pw_tc<kPasses, EPI>, (int)(TC_SMEM_MAX));
    if (e != cudaSuccess) return e;
    return launch_pdl(k_pw_tc<kPasses, EPI>, dim3(tl.grid), dim3(TC_THREADS), tl.smem, s, tl.tmA, tl.tmOut, tl.p);
}

inline cudaError_t tc_launch(const TcLaunch& tl, cudaStream_t s) {
#define CF_TC_CASE(P, E) \
    if (tl.passes == P && tl.epi == E) return tc_launch_t<P, E>(tl, s);
    CF_TC_CASE(3, EPI_LINEAR) CF_TC_CASE(3, EPI_SWISH) CF_TC_CASE(3, EPI_RESIDUAL) CF_TC_CASE(3, EPI_BIAS_SWISH) CF_TC_CASE(3, EPI_IDAUP) CF_TC_CASE(3, EPI_SWISHQ)
    CF_TC_CASE(1, EPI_LINEAR) CF_TC_CASE(1, EPI_SWISH) CF_TC_CASE(1, EPI_RESIDUAL) CF_TC_CASE(1, EPI_BIAS_SWISH) CF_TC_CASE(1, EPI_IDAUP)
#undef CF_TC_CASE
    return cudaErrorInvalidValue;
}

// (K, N) -> plan, from tools/tc_tune.py on a B200 at batch 32 @ 640x640 (profiles/r2_tc_tune.md).
inline const TcTuneEntry* tc_tuned_table(int* n) {
    auto mk = [](int nc, int atmem, int direct, int rchunk) {
        TcTune t;
        t.nc = nc, t.atmem = atmem, t.direct = direct, t.rchunk = rchunk;
        return t;
    };
    // us per launch, cost-model plan -> this plan (gpurun_out/r2e_tc_tune.jsonl, r2h_tc_tune.jsonl).  The wide stride-16/32
    // layers all want 96-column chunks with the A operand through TMEM: the shared-memory stage shrinks from 56-64 KB to
    // 40 KB, i.e. four pipeline stages instead of two (these kernels are hand-off-latency bound, not L2 or tensor bound).
    static const TcTuneEntry tab[] = {
        {24, 144, mk(64, 1, 0, -1)},    // layer1.1 / layer2.0 expand   145.1 -> 139.5
        {144, 24, mk(0, -1, 1, -1)},    // layer1.1 project (+res): row stores stay marginally ahead (115.7 vs 116.2)
        {64, 384, mk(96, 1, 0, 1)},     // layer3.1 / layer4.0 expand    26.7 -> 24.6
        {384, 96, mk(96, 1, 0, -1)},    // layer4.0 project              40.1 -> 30.7
        {96, 576, mk(96, 1, 0, -1)},    // layer4.1 / layer5.0 expand    60.8 -> 47.7 (-> 41.0 with the resident chunk)
        {576, 96, mk(96, 1, 0, -1)},    // layer4.1 project              64.1 -> 48.4
        {576, 160, mk(96, 1, 0, -1)},   // layer5.0 project              32.4 -> 28.8
        {160, 960, mk(96, 1, 0, 0)},    // layer5.1 / layer6.0 expand    31.8 -> 27.0
        {960, 160, mk(96, 1, 0, -1)},   // layer5.1 project              57.2 -> 48.7
        {960, 320, mk(96, 1, 0, 0)},    // layer6.0 project              77.1 -> 66.4
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}

}  // namespace cf
