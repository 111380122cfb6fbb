// k_mbf: one whole MBConv block (model/centernet.py:89-140) in one kernel, for the shallow stride-2/4/8 blocks:
//
//     X [B,Hi,Wi,Cin]  --expand 1x1 (tcgen05, 3xTF32) + Swish-->  E (hidden, input resolution, NEVER in HBM)
//                      --depth-wise KSxKS stride S + Swish-->     D (hidden, output resolution, NEVER in HBM)
//                      --project 1x1 (tcgen05, 3xTF32) [+ residual]-->  Y [B,Ho,Wo,Cout]
//
// The layer-wise engine writes and re-reads E and D (236 MB of the 385 MB it moves per 640x640 image belong to the four
// blocks layer1.0 .. layer2.1); here the block reads X and writes Y.  The earlier fused kernels of this library (k_expdw, k_mbx,
// k_dwp; profiles/r1_fused_kernels.md) each fused two of the three stages and lost to the layer-wise pair on SM-side
// execution: CTA-wide barriers between their phases, 8-16 resident warps, tiles sized by shared memory.  This kernel is a
// warp-specialised pipeline of small, identical JOBS with mbarrier hand-offs only (no __syncthreads in steady state):
//
//   job = (sub-tile, 32-channel chunk).  A sub-tile is STH x STW output pixels whose input halo
//   IH x IW = ((STH-1)S+KS) x ((STW-1)S+KS) is at most 128 pixels = ONE 128-row MMA block, so every stage of a job is one
//   fixed-size unit: one expand accumulator (32 TMEM columns), one E tile (17 KB).
//   A block = NSY x NSX sub-tiles (<= 128 output pixels) = the rows of ONE projection accumulator; jobs run chunk-major
//   inside a block, so chunk c of all its sub-tiles fills one D operand (hi/lo, K = 32) and the projection accumulates
//   over the chunks in TMEM.  A sub-tile's X box is loaded and split ONCE per block: its TMEM A slot is read by the expand
//   MMAs of every chunk.
//
//   warp 0        TMA producer: the X halo box of every (block, sub-tile) (zero fill outside the image = the reference's
//                 ZeroPad2d, :63-70; swish(0 . W) = 0 because the expand conv has no bias); the weight images once
//   warps 1, 2    expand issuers: the MMAs of a job as soon as its sub-tile is split and its team's accumulator is free;
//                 warp w serves the teams t % 2 == w (a tcgen05.mma costs its issuing thread ~55 cycles whatever its size)
//   warp 3        projection issuer: the MMAs of a (block, chunk) as soon as its D operand is complete (its own warp, so
//                 that neither MMA stream ever waits behind the other's barrier); also allocates / frees TMEM
//   warps 4-7     splitters: X row -> tf32 hi + lo -> TMEM A slot (tcgen05.st, lane = halo pixel), once per (block,
//                 sub-tile); after a block's splits they drain the projection accumulator of the block LAG = 2 blocks back
//                 (+ residual) and store Y
//   warps 8-      four or five compute teams of four warps, jobs round-robin.  A team takes its job through both
//                 element-wise phases: expand accumulator (TMEM, a warp = a lane quarter) -> Swish -> the team's private E
//                 tile in shared memory -> depth-wise taps -> Swish -> tf32 hi/lo rows of the D operand.  Teams are in
//                 different phases at any time, so the MUFU-bound drain of one overlaps the LDS/FMA-bound depth-wise phase
//                 of another on every scheduler; the only CTA-level synchronisation is two 128-thread named barriers per
//                 job inside a team.
//   Direct mode (EXP = false: depth-wise + projection from a hidden tensor in HBM): no expand roles, TMA writes the E tiles.
//
//   The development trace (clock64 of every hand-off, tools/mbf_trace.py) is a separate template instantiation: a warp of
//   this kernel is a latency-bound chain that pays ~10 cycles per instruction, and the run-time test in front of ~10 trace
//   sites per job cost 5 %.  History and measurements: profiles/r2_mbf.md.
//
// mbarrier waits are by phase PARITY, so a waiter must observe EVERY phase of a barrier it waits on (one that skips a phase
// finds the parity of a phase that has not begun "already complete" -- seen as a timing-dependent deadlock in an earlier
// version whose depth-wise teams shared E slots round-robin).  Hence every accumulator / E tile belongs to one team, every
// barrier has one fixed waiter role, and the D operand hand-back is per writer team or per slot (see DFREE_PER_TEAM below).
//
// 3xTF32: expand uses ONE accumulator, the two correction products first (while the accumulator is ~2^-11 of its final
// size their per-MMA accumulator truncation, tools/tc_accum_probe.py, is negligible), then the K/8 main products; the
// projection keeps k_pw_tc's main + correction pair because its chunks arrive over time.
#pragma once
#include "k_dwt.cuh"
#include "k_pwn.cuh"

// Swish of the compute teams with one reciprocal per four values (common.cuh, swish4q) where the teams' two Swish phases keep the
// special-function pipe busiest: layer1.0 (MUFU 67 %: 362 -> 346 us).  layer1.1 (MUFU 44 %) and layer0 measured flat (281 -> 284,
// 205 -> 204) and keep swish2.  CF_MBF_SWISHQ=0 = swish2 everywhere, for A/B builds.
#ifndef CF_MBF_SWISHQ
#define CF_MBF_SWISHQ 1
#endif

namespace cf {

template <int KS_, int S_, int CIN_, int STH_, int STW_, int NSY_, int NSX_, int XT_, int YT_, int TD_, int NX_, bool WDS_, bool EXP_ = true, int NT_ = 4, int ND_ = 2, bool TEPI_ = true>
struct MbfCfg {
    static constexpr int KS = KS_, S = S_, CIN = CIN_, STH = STH_, STW = STW_, NSY = NSY_, NSX = NSX_, XT = XT_, YT = YT_, TD = TD_;
    static constexpr int IH = (STH - 1) * S + KS, IW = (STW - 1) * S + KS, NPX = IH * IW;
    static constexpr int LO = (KS - S) / 2;  // model/centernet.py:68-70
    static constexpr int NSUB = NSY * NSX, SPX = STH * STW, DROWS = NSUB * SPX;
    static constexpr int KSTEPS = (CIN + 7) / 8;
    // EXP: the block's expand conv runs in this kernel (X = block input).  !EXP ("direct" mode): X is the HIDDEN tensor at input
    // resolution -- layer0 (t = 1, no expand conv) or a block whose expand conv stays a k_pw_tc launch -- and TMA writes its
    // 32-channel halo boxes straight into the teams' E tiles (two per team): the kernel is depth-wise + Swish + projection.
    static constexpr bool EXP = EXP_;
    static constexpr bool SWQ = CF_MBF_SWISHQ && EXP_ && S_ == 2;  // one-reciprocal Swish (see above)
    static constexpr int NT = NT_;  // compute teams (four warps each: one per TMEM lane quarter); 4 -> 80 registers per thread, 5 -> 72
    static constexpr int NWARPS = 8 + 4 * NT, THREADS = NWARPS * 32;
    static constexpr int NBX = STW / XT, NITEMS = (STH / YT) * NBX * 8;  // depth-wise items: (output block, float4 of channels)
    // ring depths: X boxes, TMEM A slots, expand accumulators (one per team), D operands, projection accumulators
    // (direct mode has the shared memory and the TMEM columns for one D operand and one projection accumulator per team: with two of
    // each, a slot's cycle -- depth-wise tail, projection, epilogue -- bounded the job rate of the one-chunk layer0)
    // Who drains a block's projection accumulator (+ residual, Y stores).  A one-job-per-chunk block (NSUB == 1) is nch jobs long,
    // shorter than the ~10-job pipeline behind the splitters: waiting for its projection on the splitter warps stalled every stage
    // (B2 trace: 4 200 cycles per block), and a single epilogue team bounded the one-chunk layer0 (1 000 cycles per job).  There the
    // COMPUTE TEAM that produced a block's last chunk drains it, one job later, when it has just observed the D hand-back (the
    // projection issuer is in order: that block's projection has retired).  Otherwise the splitters do it, one block behind.
    static constexpr int LAG = 2;  // splitter-side epilogues: blocks between a block's split and its epilogue
    static constexpr bool TEAM_EPI = EXP_ && NSUB == 1 && TEPI_;  // direct mode: measured slower on the teams (layer0 205 -> 249 us), the splitter warps are idle there
    static constexpr int NX = NX_, NE = NT, ND = EXP_ ? ND_ : NT, NP = EXP_ ? (TEAM_EPI ? 2 : LAG + 1) : 4;
    // TMEM A slots (the split block input).  A slot holds one sub-tile for ALL the chunks of its block: the splitters (and the X box
    // TMA) run once per (block, sub-tile), the expand issuer re-reads the slot with each chunk's weights and hands it back after the
    // last one.  Slots come in groups of NSUB (one block); NG groups rotate block by block.
    static constexpr int AW = KSTEPS <= 2 ? 32 : 64;  // columns per slot: hi | lo, 8 * KSTEPS each
    static constexpr int NA_AVAIL = (512 - NP * 64 - NE * 32) / AW;
    static constexpr int NG = EXP_ ? ((NA_AVAIL < 8 ? NA_AVAIL : 8) / NSUB < 4 ? (NA_AVAIL < 8 ? NA_AVAIL : 8) / NSUB : 4) : 1;
    static constexpr int NA = EXP_ ? NG * NSUB : 4;
    static constexpr bool WDS = WDS_;  // depth-wise taps resident in shared memory (else read through L1 from the chunk image)
    // D operand hand-back (projection retired -> the slot may be rewritten).  NSUB == 1 (an operand is one job): one barrier per
    // WRITER TEAM unless the teams map onto the slots one to one.  NSUB > 1: one barrier per SLOT; a team derives the parity it
    // waits for from the operand index, and may skip operands (NT does not divide ND * NSUB).  That is safe while a team's
    // consecutive jobs are at most ND operands apart (NT <= ND * NSUB): before its previous job it observed the retirement of
    // operand >= k - 2 ND, the issuer retires in order, so the slot's barrier is at most ONE phase behind the awaited one -- the
    // only distance at which a parity wait cannot alias.
    static constexpr bool DFREE_PER_TEAM = NSUB == 1 && ND % NT != 0;
    static constexpr int NDF = DFREE_PER_TEAM ? NT : ND;
    static_assert(NSUB == 1 || NT <= ND * NSUB, "D hand-back by slot: a team must not fall two phases behind a slot's barrier");
    static_assert(!EXP_ || NG >= 1, "TMEM: one block of A slots");
    static_assert(NITEMS <= 128, "one depth-wise item per thread of a team");
    // X box (written by TMA, read by the splitters): 128 pixel rows of 128 B, SWIZZLE_128B; a 16-channel input (layer1.0) takes
    // 64-byte rows, SWIZZLE_64B -- half the ring bytes per slot, which is what lets a deeper ring cover the TMA latency
    static constexpr uint32_t XROW = (EXP_ && CIN_ <= 16) ? 64 : 128, SLOT = 128 * XROW;
    // E slot: pixel rows at a 144-byte pitch instead of a swizzle -- the drain's stores (a lane = a pixel, 8 consecutive pixels
    // per quarter warp) and the depth-wise loads (8 lanes = the 128 bytes of one pixel) are both bank-conflict free, and every
    // window address is one base register + an immediate
    // (direct mode: TMA writes dense 128-byte rows, which the depth-wise loads read conflict free as well)
    static constexpr uint32_t EP = EXP ? 144 : 128, ESLOT = ((NPX * EP + 1023) / 1024) * 1024;
    static constexpr int NEB = EXP ? 1 : 2;      // E tiles per team
    static constexpr int NXB = EXP ? NX : NT * NEB;  // x_full / x_empty barriers: X ring slots, or one per E tile
    static constexpr uint32_t DHALF = ((DROWS + 7) / 8) * 1024u;    // the hi (or lo) half of one D operand
    static constexpr uint32_t PCOL = 0, ECOL = PCOL + NP * 64, ACOL = ECOL + NE * 32;  // TMEM columns
    static_assert(!EXP || NPX <= 128, "a sub-tile's halo is one MMA block");
    static_assert(DROWS <= 128, "a block's outputs are the rows of one projection accumulator");
    static_assert(STW % XT == 0 && STH % YT == 0, "output blocks tile the sub-tile");
    static_assert(TD_ >= 0, "legacy parameter");
    static_assert(EXP_ ? ACOL + NA * AW <= 512 : NP * 64 <= 512, "TMEM budget");
    static_assert(!EXP || (CIN % 8 == 0 && CIN <= 32), "expand K");
};

struct MbfParams {
    const float* we_img;  // [nch][hi 32 x 128 B | lo 32 x 128 B]   expand weights, K-major SWIZZLE_128B (tc_prepare_layer, NC = 32)
    const float* wp_img;  // [nch][hi 32 x 128 B | lo 32 x 128 B]   projection weights, one K block per chunk, N padded to 32
    const float* Wd;      // [KS*KS][hid]
    float* Y;             // [B][Ho][Wo][cout]
    const float* res;     // [B][Ho][Wo][cout] or NULL
    int B, Hi, Wi, Ho, Wo, hid, nch, cout;
    int blocks_x, blocks_y, n_blocks;
    const float* wd_img;  // [nch][KS*KS][32]   depth-wise taps per chunk, zero past hid
    uint32_t off_e, off_d, off_we, off_wp, off_wd, off_bars;  // the X ring starts at 0
    unsigned long long* trace;  // development only (env CF_MBF_TRACE=j0,nj): clock64 of the pipeline events of CTA 0, jobs [tr_j0, tr_j0 + tr_nj)
    int tr_j0, tr_nj;
    int dbg;  // development only (env CF_MBF_DEBUG; wrong results): 1 no Swish in the drain, 2 no depth-wise work, 4 no Y stores, 8 no split math, 16 no TMA loads, 32 no expand MMAs, 64 no projection MMAs
};

template <typename C, bool TRACE = false>
__global__ void __launch_bounds__(C::THREADS, 1) k_mbf(const __grid_constant__ CUtensorMap tmX, const MbfParams p) {
    constexpr int KS = C::KS, S = C::S, NX = C::NX, NA = C::NA, NE = C::NE, ND = C::ND, NP = C::NP, NT = C::NT, NDF = C::NDF;
    constexpr int NSUB = C::NSUB;
    constexpr uint32_t SLOT = C::SLOT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();

    // barriers (8 B each)
    const uint32_t bars = base + p.off_bars;
    constexpr int NXB = C::NXB;
    const uint32_t x_full = bars, x_empty = x_full + 8 * NXB;
    const uint32_t a_full = x_empty + 8 * NXB, a_empty = a_full + 8 * NA;
    const uint32_t e_full = a_empty + 8 * NA, e_empty = e_full + 8 * NE;
    const uint32_t d_full = e_empty + 8 * NE, d_free = d_full + 8 * ND;
    const uint32_t p_full = d_free + 8 * NDF, p_empty = p_full + 8 * NP;
    const uint32_t w_full = p_empty + 8 * NP;
    constexpr int NBARS = 2 * (NXB + NA + NE + NP) + ND + NDF + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + p.off_bars + 8 * NBARS + 8);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        for (int i = 0; i < NXB; ++i) mbar_init(x_full + 8 * i, 1), mbar_init(x_empty + 8 * i, 4);
        for (int i = 0; i < NA; ++i) mbar_init(a_full + 8 * i, 4), mbar_init(a_empty + 8 * i, 2);  // hand-back: one commit per expand issuer
        for (int i = 0; i < NE; ++i) mbar_init(e_full + 8 * i, 1), mbar_init(e_empty + 8 * i, 4);
        for (int i = 0; i < ND; ++i) mbar_init(d_full + 8 * i, NSUB * 4);
        for (int i = 0; i < NDF; ++i) mbar_init(d_free + 8 * i, 1);
        for (int i = 0; i < NP; ++i) mbar_init(p_full + 8 * i, 1), mbar_init(p_empty + 8 * i, 4);
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // development trace: event ev of job j (one lane of one warp per role writes; 32 slots per job)
    // (a compile-time switch: the seven-instruction test in front of every site cost the latency-bound warps ~5 % when it was a run-time one)
    auto TR = [&](int ev, int j) {
        if constexpr (!TRACE) return;
        if (p.trace && blockIdx.x == 0 && lane == 0 && (unsigned)(j - p.tr_j0) < (unsigned)p.tr_nj) p.trace[(size_t)(j - p.tr_j0) * 32 + ev] = (unsigned long long)clock64();
    };
    const int nch = p.nch;
    const int nblk = ((int)blockIdx.x < p.n_blocks) ? (p.n_blocks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int J = nblk * nch * NSUB;  // jobs of this CTA, in order (block i, chunk c, sub-tile s)
    const int Q = nblk * nch;         // D operands (block, chunk), NSUB jobs each

    // Every role walks the same job sequence with incremental ring cursors (slot, phase parity): no division in any loop.
    struct Ring {
        int slot = 0;
        uint32_t phase = 0;
        __device__ __forceinline__ void next(int n) {
            if (++slot == n) slot = 0, phase ^= 1u;
        }
    };

    // This CTA's blocks are blockIdx.x, + gridDim.x, ...: (bx, by, b) advance by a fixed carry-propagating step (no division per block)
    struct BlockWalk {
        int bx, by, b, sx, sy, sb, nx, ny;
        __device__ __forceinline__ void init(int first, int step, int nx_, int ny_) {
            nx = nx_, ny = ny_;
            bx = first % nx, by = (first / nx) % ny, b = first / (nx * ny);
            sx = step % nx, sy = (step / nx) % ny, sb = step / (nx * ny);
        }
        __device__ __forceinline__ void next() {
            bx += sx;
            if (bx >= nx) bx -= nx, ++by;
            by += sy;
            if (by >= ny) by -= ny, ++b;
            b += sb;
        }
    };

    // Projection epilogue of one block: this warp's lane quarter of accumulator slot `ps` (+ residual) -> Y, then the slot is handed back.
    auto epilogue_block = [&](int ps, uint32_t pphase, int bx, int by, int b, int row, uint32_t lane_base) {
        const int rs = row / C::SPX, rl = row - rs * C::SPX;  // accumulator row -> output pixel of the block
        const int rsy = rs / C::NSX, rsx = rs - rsy * C::NSX;
        const int yo = by * (C::NSY * C::STH) + rsy * C::STH + rl / C::STW, xo = bx * (C::NSX * C::STW) + rsx * C::STW + rl % C::STW;
        const bool valid = row < C::DROWS && yo < p.Ho && xo < p.Wo && !(p.dbg & 4);
        const size_t pix = ((size_t)(b * p.Ho + yo) * p.Wo + xo) * (size_t)p.cout;
        const uint32_t taddr = lane_base + C::PCOL + (uint32_t)ps * 64u;
        if (p.res) {
            // the residual does not depend on the accumulator: its (L2) latency runs under the wait below
            float4 rs4[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) rs4[h] = (valid && 4 * h < p.cout) ? ldcg4(p.res + pix + 4 * h) : make_float4(0, 0, 0, 0);
            mbar_wait(p_full + 8 * ps, pphase);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int c0 = 8 * g;
                if (c0 < p.cout) {  // warp-uniform
                    float v[8], cr[8];
                    tmem_ld8(taddr + (uint32_t)c0, v);
                    tmem_ld8(taddr + 32u + (uint32_t)c0, cr);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float4 r4 = rs4[2 * g + h];
                            st4(p.Y + pix + c0 + 4 * h, make_float4(v[4 * h] + cr[4 * h] + r4.x, v[4 * h + 1] + cr[4 * h + 1] + r4.y,
                                                                    v[4 * h + 2] + cr[4 * h + 2] + r4.z, v[4 * h + 3] + cr[4 * h + 3] + r4.w));
                        }
                    }
                }
            }
        } else {
            mbar_wait(p_full + 8 * ps, pphase);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 2; ++g) {  // 16 columns per TMEM round trip
                const int c0 = 16 * g;
                if (c0 < p.cout) {  // warp-uniform
                    float v[16], cr[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
                    tmem_ld16(taddr + 32u + (uint32_t)c0, cr);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            if (c0 + 4 * h < p.cout)
                                st4(p.Y + pix + c0 + 4 * h,
                                    make_float4(v[4 * h] + cr[4 * h], v[4 * h + 1] + cr[4 * h + 1], v[4 * h + 2] + cr[4 * h + 2], v[4 * h + 3] + cr[4 * h + 3]));
                    }
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_empty + 8 * ps);
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            const uint32_t wbytes = (uint32_t)nch * 8192u, dbytes = C::WDS ? (uint32_t)nch * (KS * KS * 128u) : 0u;
            mbar_expect_tx(w_full, (C::EXP ? 2u : 1u) * wbytes + dbytes);
            if (C::WDS) bulk_load(base + p.off_wd, p.wd_img, dbytes, w_full);
            for (uint32_t off = 0; off < wbytes; off += 32768u) {
                const uint32_t n = wbytes - off < 32768u ? wbytes - off : 32768u;
                if (C::EXP) bulk_load(base + p.off_we + off, reinterpret_cast<const uint8_t*>(p.we_img) + off, n, w_full);
                bulk_load(base + p.off_wp + off, reinterpret_cast<const uint8_t*>(p.wp_img) + off, n, w_full);
            }
        }
        __syncwarp();
        pdl_wait();  // X is the previous kernel's output
        Ring xr;
        int jt = 0;
        BlockWalk bw;
        bw.init((int)blockIdx.x, (int)gridDim.x, p.blocks_x, p.blocks_y);
        for (int i = 0; i < nblk; ++i, bw.next()) {
            const int bx = bw.bx, by = bw.by, b = bw.b;
            const int x00 = bx * (C::NSX * C::STW * S) - C::LO, y00 = by * (C::NSY * C::STH * S) - C::LO;
            for (int c = 0; c < (C::EXP ? 1 : nch); ++c) {  // expand mode: one box per (block, sub-tile), shared by the chunks
                if (C::EXP) jt = i * nch * NSUB;
#pragma unroll
                for (int s = 0; s < NSUB; ++s) {
                    const int sy = s / C::NSX, sx = s % C::NSX;  // compile-time after unrolling
                    mbar_wait(x_empty + 8 * xr.slot, xr.phase ^ 1u);
                    TR(0, jt);
                    if (elect_one()) {
                        if (p.dbg & 16) {
                            mbar_arrive(x_full + 8 * xr.slot);
                        } else {
                            mbar_expect_tx(x_full + 8 * xr.slot, (uint32_t)C::NPX * (C::EXP ? C::XROW : 128u));
                            // expand mode: the block input (all its channels) into the X ring; direct mode: chunk c of the hidden
                            // tensor into E tile xr.slot (the tiles of team t are t and t + NT: jobs go round-robin)
                            tma_load_4d(C::EXP ? base + xr.slot * SLOT : base + p.off_e + xr.slot * C::ESLOT, &tmX, C::EXP ? 0 : c * 32,
                                        x00 + sx * (C::STW * S), y00 + sy * (C::STH * S), b, x_full + 8 * xr.slot);
                        }
                    }
                    __syncwarp();
                    TR(1, jt);
                    ++jt;
                    xr.next(NXB);
                }
            }
        }
    } else if ((warp == 1 || warp == 2) && C::EXP) {
        // ================= expand issuers: one MMA group per job =================
        // A tcgen05.mma costs its issuing thread ~55 cycles whatever its size (trace: 6 MMAs + 2 commits = 370 cycles), so one
        // issuer bounded the job rate.  Two warps walk the same job sequence; warp w issues the jobs of the teams t % 2 == w (a
        // fixed issuer per team: it observes every phase of that team's e_empty barrier).  Both wait for every split sub-tile and
        // both commit its hand-back after the block's last chunk (a commit covers the thread's earlier MMAs; count 2).
        const int w = warp - 1;
        mbar_wait(w_full, 0);
        const uint32_t idesc32 = umma_idesc_tf32(32);
        Ring gr, er;  // gr = the A slot group of the block, er.slot = the team of job j
        int c = 0, s = 0;
        for (int j = 0; j < J; ++j) {
            if (w == 0) TR(24, j);
            const int aslot = gr.slot * NSUB + s;
            if (c == 0) mbar_wait(a_full + 8 * aslot, gr.phase);  // the split sub-tile: written once, read by every chunk
            const bool mine = (er.slot & 1) == w;
            if (w == 0) TR(5, j);
            if (mine) {
                mbar_wait(e_empty + 8 * er.slot, er.phase ^ 1u);
                TR(6, j);
            }
            tc_fence_after();
            const uint32_t wb = base + p.off_we + (uint32_t)c * 8192u;
            const uint64_t b_hi = umma_desc(wb), b_lo = umma_desc(wb + 4096u);
            const uint32_t a_hi = tmem_base + C::ACOL + (uint32_t)aslot * C::AW, a_lo = a_hi + C::AW / 2;
            const uint32_t d = tmem_base + C::ECOL + (uint32_t)er.slot * 32u;
            if (mine) TR(25, j);
            if (elect_one()) {
                if (mine) {
                    if (!(p.dbg & 32))
#pragma unroll
                    for (int k = 0; k < C::KSTEPS; ++k) {  // the small products first
                        umma_tf32_ts(d, a_lo + 8u * k, b_hi + (uint64_t)(k * 2), idesc32, k > 0 ? 1u : 0u);
                        umma_tf32_ts(d, a_hi + 8u * k, b_lo + (uint64_t)(k * 2), idesc32, 1u);
                    }
                    if (!(p.dbg & 32))
#pragma unroll
                    for (int k = 0; k < C::KSTEPS; ++k) umma_tf32_ts(d, a_hi + 8u * k, b_hi + (uint64_t)(k * 2), idesc32, 1u);
                    umma_commit(e_full + 8 * er.slot);
                }
                if (c == nch - 1) umma_commit(a_empty + 8 * aslot);
            }
            __syncwarp();
            if (mine) TR(7, j);
            er.next(NE);
            if (++s == NSUB) {
                s = 0;
                if (++c == nch) c = 0, gr.next(C::NG);
            }
        }
    } else if (warp == 3) {
        // ================= projection issuer (warp 3: scheduler 3 carries the lightest compute load, see the item rotation below): one MMA group per (block, chunk) =================
        mbar_wait(w_full, 0);
        const uint32_t idesc32 = umma_idesc_tf32(32), idesc64 = umma_idesc_tf32(64);
        Ring dr, pr;
        int c = 0, nxt = ND % NT;  // the team that writes D operand q + ND (per-team hand-back)
        for (int q = 0; q < Q; ++q) {
            mbar_wait(d_full + 8 * dr.slot, dr.phase);
            TR(15, q * NSUB + NSUB - 1);
            if (c == 0) mbar_wait(p_empty + 8 * pr.slot, pr.phase ^ 1u);
            TR(26, q * NSUB + NSUB - 1);
            tc_fence_after();
            const uint32_t dbase = base + p.off_d + (uint32_t)dr.slot * 2u * C::DHALF;
            const uint64_t a_hi = umma_desc(dbase), a_lo = umma_desc(dbase + C::DHALF);
            const uint64_t b_hi = umma_desc(base + p.off_wp + (uint32_t)c * 8192u);  // [hi 32 rows | lo 32 rows]
            const uint32_t d_main = tmem_base + C::PCOL + (uint32_t)pr.slot * 64u, d_corr = d_main + 32u;
            if (elect_one()) {
                if (!(p.dbg & 64))
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ko = (uint64_t)(k * 2);
                    umma_tf32(d_main, a_hi + ko, b_hi + ko, idesc64, (c > 0 || k > 0) ? 1u : 0u);  // main += hi.hi ; corr += hi.lo
                    umma_tf32(d_corr, a_lo + ko, b_hi + ko, idesc32, 1u);                            // corr += lo.hi
                }
                umma_commit(d_free + 8 * (C::DFREE_PER_TEAM ? nxt : dr.slot));
                if (c == nch - 1) umma_commit(p_full + 8 * pr.slot);
            }
            __syncwarp();
            TR(16, q * NSUB + NSUB - 1);
            dr.next(ND);
            if (++nxt == NT) nxt = 0;
            if (++c == nch) c = 0, pr.next(NP);
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= splitters; between blocks: projection epilogue of the previous block =================
        pdl_wait();  // residual reads and Y stores touch activation memory
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        Ring pr;
        BlockWalk ew;  // the epilogues run in block order
        ew.init((int)blockIdx.x, (int)gridDim.x, p.blocks_x, p.blocks_y);
        auto epilogue = [&](int) {
            epilogue_block(pr.slot, pr.phase, ew.bx, ew.by, ew.b, row, lane_base);
            ew.next();
            pr.next(NP);
        };
        Ring xr, gr;
        const int xswz = C::XROW == 64 ? (row >> 1) & 3 : row & 7;  // 16-byte chunk XOR of the TMA swizzle mode
        int jt = 0;
        if (!C::EXP && !C::TEAM_EPI) {  // direct mode: this team only drains the projection accumulators
            for (int i = 0; i < nblk; ++i) {
                if (q == 0) TR(17, (i + 1) * nch * NSUB - 1);
                epilogue(i);
                if (q == 0) TR(18, (i + 1) * nch * NSUB - 1);
            }
        }
        // Expand mode, epilogues on this team: the epilogue of block i - 1 runs after the splits of block i.
        int epi_done = 0;
        for (int i = 0; i < (C::EXP ? nblk : 0); ++i) {
            jt = i * nch * NSUB;
            for (int cs = 0; cs < NSUB; ++cs, ++jt) {  // one split per (block, sub-tile): every chunk's expand reads it
                const int aslot = gr.slot * NSUB + cs;
                mbar_wait(x_full + 8 * xr.slot, xr.phase);
                if (q == 0) TR(2, jt);
                const uint8_t* xrow = sm + (size_t)xr.slot * SLOT + (size_t)row * C::XROW;
                float hi[8 * C::KSTEPS], lo[8 * C::KSTEPS];
                if (!(p.dbg & 8))
#pragma unroll
                for (int g = 0; g < 2 * C::KSTEPS; ++g) {
                    const float4 v = *reinterpret_cast<const float4*>(xrow + ((g ^ xswz) << 4));  // SWIZZLE_128B image written by TMA
                    hi[4 * g] = tf32_hi(v.x), hi[4 * g + 1] = tf32_hi(v.y), hi[4 * g + 2] = tf32_hi(v.z), hi[4 * g + 3] = tf32_hi(v.w);
                    lo[4 * g] = v.x - hi[4 * g], lo[4 * g + 1] = v.y - hi[4 * g + 1], lo[4 * g + 2] = v.z - hi[4 * g + 2], lo[4 * g + 3] = v.w - hi[4 * g + 3];
                }
                mbar_wait(a_empty + 8 * aslot, gr.phase ^ 1u);
                if (q == 0) TR(3, jt);
                tc_fence_after();
                const uint32_t ta = lane_base + C::ACOL + (uint32_t)aslot * C::AW;
                if (C::KSTEPS == 4) {
                    tmem_st16(ta, hi), tmem_st16(ta + 16u, hi + 16);
                    tmem_st16(ta + 32u, lo), tmem_st16(ta + 48u, lo + 16);
                } else if (C::KSTEPS == 3) {
                    tmem_st16(ta, hi), tmem_st8(ta + 16u, hi + 16);
                    tmem_st16(ta + 32u, lo), tmem_st8(ta + 48u, lo + 16);
                } else {
                    tmem_st16(ta, hi);
                    tmem_st16(ta + 16u, lo);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(a_full + 8 * aslot);
                    mbar_arrive(x_empty + 8 * xr.slot);
                }
                if (q == 0) TR(4, jt);
                xr.next(NX);
            }
            gr.next(C::NG);
            jt = (i + 1) * nch * NSUB;
            if (!C::TEAM_EPI && i >= C::LAG) {
                if (q == 0) TR(17, jt - 1);
                epilogue(epi_done++);
                if (q == 0) TR(18, jt - 1);
            }
        }
        while (C::EXP && !C::TEAM_EPI && epi_done < nblk) epilogue(epi_done++);  // the tail: nothing left to split
    } else if (warp >= 8) {
        // ================= compute teams: expand accumulator -> Swish -> E tile -> taps -> Swish -> hi/lo rows of the D operand =================
        constexpr int NROW = (C::YT - 1) * S + KS, NCOL = (C::XT - 1) * S + KS;
        const int team = (warp - 8) >> 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        uint8_t* Es = sm + p.off_e + (size_t)team * C::ESLOT;  // direct mode: + NT * ESLOT for the second tile
        auto bar_team = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory"); };
        // depth-wise item of this thread (fixed): output block (by, bx) of the sub-tile, channels 4 * c4 .. + 3 of the chunk
        // A geometry with fewer than 128 items leaves warps without depth-wise work; a warp's scheduler is fixed (warp % 4, like its
        // TMEM lane quarter), so the items rotate by one warp per team: the idle warps of different teams sit on different schedulers.
        const int it = C::NITEMS < 128 ? (row + 32 * (team + 1)) & 127 : row;
        const int c4 = it & 7, blk = it >> 3;
        const int by = blk / C::NBX, bx = blk - by * C::NBX;
        const bool has_item = it < C::NITEMS;
        const uint8_t* eb = Es + ((C::YT * by * S) * C::IW + C::XT * bx * S) * (int)C::EP + c4 * 16;  // window origin: every load is this + an immediate
        Ring dr;            // D operand of the current job
        int s = 0, c = 0;   // sub-tile within the block, chunk
        uint32_t ephase = 0, fphase = (team >= ND) ? 1u : 0u;  // per-team hand-back: teams >= ND wait for a real retirement the first time
        BlockWalk tw;  // block of the current job, its projection accumulator (TEAM_EPI)
        tw.init((int)blockIdx.x, (int)gridDim.x, p.blocks_x, p.blocks_y);
        Ring tpr;
        // n jobs further in the CTA's job sequence (n <= NT), in closed form: the teams run this between two jobs, and a latency-bound
        // warp pays ~10 cycles per instruction -- the job-by-job walk cost ~1 000 cycles per job
        auto advance = [&](int n) {
            s += n;
            int dq = 0;  // D operands (block, chunk) passed
            while (s >= NSUB) s -= NSUB, ++dq;
            dr.slot += dq;
            while (dr.slot >= ND) dr.slot -= ND, dr.phase ^= 1u;
            c += dq;
            while (c >= nch) {
                c -= nch;
                if (C::TEAM_EPI) tw.next(), tpr.next(NP);
            }
        };
        advance(team);
        // TEAM_EPI: the block whose last chunk was this team's PREVIOUS job waits to be drained
        bool epi_pending = false;
        int e_ps = 0, e_bx = 0, e_by = 0, e_b = 0;
        uint32_t e_ph = 0;
        if (C::TEAM_EPI) pdl_wait();  // residual reads and Y stores touch activation memory
        if (C::WDS) mbar_wait(w_full, 0);  // the tap image
        uint32_t kjob = 0;  // this team's job counter (direct mode: E tile = kjob & 1, its barrier phase = (kjob >> 1) & 1)
        for (int j = team; j < J; j += NT, ++kjob) {
            const uint32_t xslot = (uint32_t)team + (kjob & 1u) * NT;
            const uint8_t* ebj = eb + (C::EXP ? 0u : (kjob & 1u) * NT * C::ESLOT);
            if (!C::EXP) {
                mbar_wait(x_full + 8 * xslot, (kjob >> 1) & 1u);  // TMA has written this job's halo box of the hidden tensor
                if (q == 0) TR(8, j);
            }
            // ---- drain: this warp's lane quarter of the accumulator ----
            if (C::EXP && q == 0) TR(27, j);
            if (C::EXP) mbar_wait(e_full + 8 * team, ephase);
            if (C::EXP && q == 0) TR(8, j);
            // probe the D hand-back now (non-blocking); the result is needed only after the depth-wise arithmetic
            const uint32_t fbar = d_free + 8 * (C::DFREE_PER_TEAM ? team : dr.slot), fpar = (C::DFREE_PER_TEAM ? fphase : dr.phase) ^ 1u;
            uint32_t d_ok;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(d_ok)
                : "r"(fbar), "r"(fpar)
                : "memory");
            if (C::EXP) {
                tc_fence_after();
                float v[32];
                tmem_ld32(lane_base + C::ECOL + (uint32_t)team * 32u, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(e_empty + 8 * team);  // the accumulator is in registers: the next job's MMAs may start
                ephase ^= 1u;
                if (q == 0) TR(9, j);
                if (q * 32 < C::NPX && !(p.dbg & 1)) {  // warp-uniform: a quarter past the halo tile has nothing to do
    #pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (C::SWQ) swish4q(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                        else swish2(v[4 * g], v[4 * g + 1]), swish2(v[4 * g + 2], v[4 * g + 3]);
                    }
                }
                if (q == 0) TR(19, j);
                bar_team();  // every warp of the team has finished the previous job's depth-wise reads of E
                if (q == 0) TR(10, j);
                if (row < C::NPX) {
                    uint8_t* erow = Es + (size_t)row * C::EP;
    #pragma unroll
                    for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(erow + 16 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                }
                bar_team();  // E complete
                if (q == 0) TR(11, j);
            }
            // ---- depth-wise: one item per thread, results stay in registers until the D slot is free ----
            float4 acc[C::YT][C::XT];
#pragma unroll
            for (int a = 0; a < C::YT; ++a)
#pragma unroll
                for (int b2 = 0; b2 < C::XT; ++b2) acc[a][b2] = make_float4(0, 0, 0, 0);
            if (has_item && !(p.dbg & 2)) {
                const uint8_t* wt = (C::WDS ? sm + p.off_wd : reinterpret_cast<const uint8_t*>(p.wd_img)) + (size_t)c * (KS * KS * 128) + c4 * 16;
#pragma unroll
                for (int rr = 0; rr < NROW; ++rr) {
                    float4 win[NCOL];
#pragma unroll
                    for (int cc = 0; cc < NCOL; ++cc) win[cc] = *reinterpret_cast<const float4*>(ebj + (rr * C::IW + cc) * (int)C::EP);
#pragma unroll
                    for (int dy = 0; dy < C::YT; ++dy) {
                        const int ky = rr - dy * S;
                        if (ky < 0 || ky >= KS) continue;
#pragma unroll
                        for (int kx = 0; kx < KS; ++kx) {
                            const float4 wv = C::WDS ? *reinterpret_cast<const float4*>(wt + (ky * KS + kx) * 128)
                                                     : ldg4(reinterpret_cast<const float*>(wt + (ky * KS + kx) * 128));
#pragma unroll
                            for (int dx = 0; dx < C::XT; ++dx) fma44p(acc[dy][dx], win[dx * S + kx], wv);
                        }
                    }
                }
            }
            if (!C::EXP) {  // the window loads have been consumed: TMA may refill this tile (job j + 2 NT)
                __syncwarp();
                if (lane == 0) mbar_arrive(x_empty + 8 * xslot);
            }
            if (q == 0) TR(12, j);
            if (!d_ok) mbar_wait(fbar, fpar);  // the projection that read the slot last (per team: this team's previous operand) has retired
            if (C::DFREE_PER_TEAM) fphase ^= 1u;
            if (q == 0) TR(13, j);
            if (has_item) {
                uint8_t* Dhi = sm + p.off_d + (size_t)dr.slot * 2u * C::DHALF;
                uint8_t* Dlo = Dhi + C::DHALF;
#pragma unroll
                for (int dy = 0; dy < C::YT; ++dy)
#pragma unroll
                    for (int dx = 0; dx < C::XT; ++dx) {
                        const int r = s * C::SPX + (C::YT * by + dy) * C::STW + C::XT * bx + dx;  // D operand row = output pixel of the block
                        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c4 ^ (r & 7)) << 4);
                        const float4 o = C::SWQ ? swish4qv(acc[dy][dx]) : swish4p(acc[dy][dx]);  // swish(0) = 0 keeps the padded channels zero
                        const float4 h = make_float4(tf32_hi(o.x), tf32_hi(o.y), tf32_hi(o.z), tf32_hi(o.w));
                        *reinterpret_cast<float4*>(Dhi + off) = h;
                        *reinterpret_cast<float4*>(Dlo + off) = make_float4(o.x - h.x, o.y - h.y, o.z - h.z, o.w - h.w);
                    }
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(d_full + 8 * dr.slot);
            if (q == 0) TR(14, j);
            if (C::TEAM_EPI) {
                // The D hand-back observed above means the in-order projection issuer has retired this team's previous job: if that
                // was a block's last chunk, its accumulator is complete -- drain it now, behind this job's D operand.
                if (epi_pending) {
                    if (q == 0) TR(17, j);
                    epilogue_block(e_ps, e_ph, e_bx, e_by, e_b, row, lane_base);
                    if (q == 0) TR(18, j);
                }
                epi_pending = c == nch - 1;
                e_ps = tpr.slot, e_ph = tpr.phase, e_bx = tw.bx, e_by = tw.by, e_b = tw.b;
            }
            if (q == 0) TR(28, j);
            advance(NT);
        }
        if (C::TEAM_EPI && epi_pending) epilogue_block(e_ps, e_ph, e_bx, e_by, e_b, row, lane_base);  // this team's last job closed a block
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------
// Geometry per block type (sub-tile, block, depth-wise item shape; the TD parameter is unused):
//   3x3 s2, Cin 16 (layer1.0): sub-tile 3 x 8 (halo 7 x 17 = 119 px), block 2 x 2 sub-tiles = 6 x 16 outputs, items of 1 x 2 outputs: 96 = 3 warps
//   3x3 s1, Cin 24 (layer1.1): sub-tile 8 x 10 (halo 10 x 12 = 120 px), block = the sub-tile, items of 1 x 5 outputs: 128 = 4 warps
// (the 5x5 blocks layer2.0 / layer2.1 have no configuration: 128-pixel halo sub-tiles recompute 1.9x / 2.5x of their expand work)
using MbfB1 = MbfCfg<3, 2, 16, 3, 8, 2, 2, 2, 1, 3, 4, true, true, 5>;
// (depth-wise items of 2 x 5 outputs instead of 1 x 5 -- 38 % fewer shared-memory loads per output, the change that took 20 % off
// k_dwt's 5x5 layers -- measured SLOWER here: layer1.1 279 -> 352 us, layer0 204 -> 254: half of a team's threads then walk a
// chain twice as long, and a team is bound by its per-thread latency chain, not by the shared-memory pipe)
using MbfB2 = MbfCfg<3, 1, 24, 8, 10, 1, 1, 5, 1, 4, 1, true, true, 4, 2, false>;
// direct mode (depth-wise + projection from the hidden tensor): 3x3 s1 with layer1.1's geometry
using MbfD31 = MbfCfg<3, 1, 32, 8, 10, 1, 1, 5, 1, 4, 1, true, false>;

struct MbfLaunch {
    CUtensorMap tmX;
    MbfParams p;
    int kind = 0, grid = 0;  // kind: 1 = MbfB1, 2 = MbfB2, 5 = MbfD31 (direct)
    size_t smem = 0;
};

inline int mbf_kind(int ks, int s, int cin) {
    if (ks == 3 && s == 2 && cin == 16) return 1;
    if (ks == 3 && s == 1 && cin == 24) return 2;
    return 0;
}
inline bool mbf_supported(int ks, int s, int cin, int hid, int cout) { return mbf_kind(ks, s, cin) != 0 && cout % 8 == 0 && cout <= 32 && hid % 4 == 0; }

template <typename C>
inline int mbf_plan_t(PwTcState& st, MbfLaunch* ml, const float* X, int B, int Hi, int Wi, int cin) {
    MbfParams& p = ml->p;
    // expand mode: X = block input, SWIZZLE_128B boxes (read by the splitters); direct mode: X = hidden tensor, plain boxes = E tiles
    int rc = xd_make_map(st, &ml->tmX, X, B, Hi, Wi, cin, C::IW, C::IH, /*swizzle=*/C::EXP ? (C::XROW == 64 ? 2 : 1) : 0, /*box_c=*/C::EXP ? (int)C::XROW / 4 : 32);
    if (rc) return rc;
    p.Ho = Hi / C::S, p.Wo = Wi / C::S;
    p.blocks_x = cdiv(p.Wo, C::NSX * C::STW);
    p.blocks_y = cdiv(p.Ho, C::NSY * C::STH);
    const long long nb = (long long)B * p.blocks_x * p.blocks_y;
    if (nb * p.nch * C::NSUB > 0x3fffffffLL) return fail(CF_EINVAL, "mbf_plan: too many jobs");
    p.n_blocks = (int)nb;
    p.off_e = C::EXP ? C::NX * C::SLOT : 0u;
    p.off_d = p.off_e + C::NT * C::NEB * C::ESLOT;
    p.off_we = p.off_d + C::ND * 2u * C::DHALF;
    p.off_wp = p.off_we + (C::EXP ? (uint32_t)p.nch * 8192u : 0u);
    p.off_wd = p.off_wp + (uint32_t)p.nch * 8192u;
    p.off_bars = (p.off_wd + (C::WDS ? (uint32_t)p.nch * (C::KS * C::KS * 128u) : 0u) + 127u) & ~127u;
    // the MMA reads 128 rows of every D half; rows past DROWS fall into whatever follows (unused accumulator lanes), which
    // must still be this CTA's shared memory: the weight images (>= 16 KB) follow the last half
    ml->smem = (size_t)p.off_bars + 1024 + 1024;
    if (ml->smem > (size_t)TC_SMEM_MAX) return fail(CF_EINVAL, "mbf_plan: %zu B of shared memory do not fit", ml->smem);
    ml->grid = p.n_blocks < st.sms ? p.n_blocks : st.sms;
    return CF_OK;
}

inline int mbf_trace_setup(PwTcState& st, MbfParams& p) {  // development: CF_MBF_TRACE="j0,nj", read back with cf_debug_mbf_trace
    p.trace = nullptr, p.tr_j0 = p.tr_nj = 0;
    if (const char* ev = getenv("CF_MBF_TRACE")) {
        if (sscanf(ev, "%d,%d", &p.tr_j0, &p.tr_nj) == 2 && p.tr_nj > 0 && p.tr_nj <= 4096) {
            if (!st.trace_buf && cudaMalloc((void**)&st.trace_buf, 4096 * 32 * 8) != cudaSuccess) return fail(CF_ECUDA, "mbf_plan: trace buffer");
            cudaMemset(st.trace_buf, 0, 4096 * 32 * 8);
            p.trace = st.trace_buf;
        }
    }
    return CF_OK;
}

inline int mbf_plan(PwTcState& st, int ks, int s, const float* X, const float* We, const float* Wd, const float* Wp, float* Y, const float* res,
                    int B, int Hi, int Wi, int cin, int hid, int cout, MbfLaunch* ml) {
    if (!mbf_supported(ks, s, cin, hid, cout)) return fail(CF_EINVAL, "mbf_plan: no fused kernel for k=%d s=%d cin=%d cout=%d", ks, s, cin, cout);
    auto ie = st.layers.find(We), ip = st.layers.find(Wp);
    if (ie == st.layers.end() || ie->second.NC != 32 || ie->second.nkb != 1)
        return fail(CF_EINVAL, "mbf_plan: expand weights were not prepared as 32-column images");
    if (ip == st.layers.end() || ip->second.NC != 32 || ip->second.nchunks != 1)
        return fail(CF_EINVAL, "mbf_plan: projection weights were not prepared as one 32-column image per K block");
    MbfParams& p = ml->p;
    auto id = st.dw_imgs.find(Wd);
    if (id == st.dw_imgs.end()) return fail(CF_EINVAL, "mbf_plan: depth-wise taps were not prepared as a chunk image");
    p.we_img = ie->second.img;
    p.wp_img = ip->second.img;
    p.wd_img = id->second;
    p.Wd = Wd;
    p.Y = Y;
    p.res = res;
    p.B = B, p.Hi = Hi, p.Wi = Wi, p.hid = hid, p.cout = cout;
    p.nch = (hid + 31) / 32;
    p.dbg = 0;
    if (const char* ev = getenv("CF_MBF_DEBUG")) p.dbg = atoi(ev);
    if (int rc = mbf_trace_setup(st, p)) return rc;
    ml->kind = mbf_kind(ks, s, cin);
    if (ml->kind == 1) return mbf_plan_t<MbfB1>(st, ml, X, B, Hi, Wi, cin);
    return mbf_plan_t<MbfB2>(st, ml, X, B, Hi, Wi, cin);
}

// Direct mode: depth-wise + Swish + projection (+ residual) of one block from its HIDDEN tensor E [B,Hi,Wi,hid] (layer0: the stem
// output; otherwise the expand conv's output).
inline bool mbf_direct_supported(int ks, int s, int hid, int cout) { return ks == 3 && s == 1 && cout % 8 == 0 && cout <= 32 && hid % 4 == 0; }
inline int mbf_plan_direct(PwTcState& st, int ks, int s, const float* E, const float* Wd, const float* Wp, float* Y, const float* res, int B, int Hi,
                           int Wi, int hid, int cout, MbfLaunch* ml) {
    if (!mbf_direct_supported(ks, s, hid, cout)) return fail(CF_EINVAL, "mbf_plan_direct: no kernel for k=%d s=%d hid=%d cout=%d", ks, s, hid, cout);
    auto ip = st.layers.find(Wp);
    if (ip == st.layers.end() || ip->second.NC != 32 || ip->second.nchunks != 1)
        return fail(CF_EINVAL, "mbf_plan_direct: projection weights were not prepared as one 32-column image per K block");
    auto id = st.dw_imgs.find(Wd);
    if (id == st.dw_imgs.end()) return fail(CF_EINVAL, "mbf_plan_direct: depth-wise taps were not prepared as a chunk image");
    MbfParams& p = ml->p;
    p.we_img = nullptr;
    p.wp_img = ip->second.img;
    p.wd_img = id->second;
    p.Wd = Wd;
    p.Y = Y;
    p.res = res;
    p.B = B, p.Hi = Hi, p.Wi = Wi, p.hid = hid, p.cout = cout;
    p.nch = (hid + 31) / 32;
    p.dbg = 0;
    if (const char* ev = getenv("CF_MBF_DEBUG")) p.dbg = atoi(ev);
    if (int rc = mbf_trace_setup(st, p)) return rc;
    ml->kind = 5;
    return mbf_plan_t<MbfD31>(st, ml, E, B, Hi, Wi, hid);
}

template <typename C>
inline cudaError_t mbf_launch_t(const MbfLaunch& ml, cudaStream_t s) {
    if (ml.p.trace) {  // development build of the same kernel with the clock64 trace compiled in
        cudaError_t e = smem_optin((const void*)k_mbf<C, true>, TC_SMEM_MAX);
        if (e != cudaSuccess) return e;
        return launch_pdl(k_mbf<C, true>, dim3(ml.grid), dim3(C::THREADS), ml.smem, s, ml.tmX, ml.p);
    }
    cudaError_t e = smem_optin((const void*)k_mbf<C, false>, TC_SMEM_MAX);
    if (e != cudaSuccess) return e;
    return launch_pdl(k_mbf<C, false>, dim3(ml.grid), dim3(C::THREADS), ml.smem, s, ml.tmX, ml.p);
}

inline cudaError_t mbf_launch(const MbfLaunch& ml, cudaStream_t s) {
    switch (ml.kind) {
        case 1: return mbf_launch_t<MbfB1>(ml, s);
        case 2: return mbf_launch_t<MbfB2>(ml, s);
        case 5: return mbf_launch_t<MbfD31>(ml, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cf
