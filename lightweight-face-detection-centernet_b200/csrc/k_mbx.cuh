// k_mbx: fused  expand 1x1 (tcgen05, 3xTF32) + Swish -> depth-wise KSxKS stride S + Swish  for the shallow MBConv
// blocks (Cin <= 32), second generation of k_expdw_tc (k_expdw.cuh), restructured after its ncu profile
// (profiles/r1_fused_kernels.md): with ONE 512-thread CTA per SM the MUFU-bound accumulator drain and the
// LDS-bound depth-wise phase alternated behind CTA-wide barriers and 45 % of all warp stalls were barrier waits.
// Here a CTA has 256 threads, a small tile (8x16 outputs at stride 1, 8x4 at stride 2: <= 256 halo pixels = two
// 128-row MMA blocks) and <= 110 KB of shared memory, so TWO CTAs are resident per SM and the hardware overlaps
// one CTA's Swish phase with the other's depth-wise phase.  Weights no longer live in shared memory: the
// expand weight images (8 KB of tf32 hi|lo per 32-channel chunk) stream through a two-slot cp.async.bulk ring,
// the depth-wise taps are read through L1.
//
// Per CTA, per tile:  TMA halo tile X (SWIZZLE_128B = the K-major A operand)  ->  split X = hi + lo in place  ->
// for every 32-channel chunk: [tcgen05.mma x (2 blocks x Cin/8 steps x 3)] -> TMEM -> +correction, Swish ->
// swizzled E tile in smem -> depth-wise from E -> D (global, 128 B per pixel and chunk).  The MMAs of chunk
// c+1 are issued before the depth-wise phase of chunk c; the TMA of the next tile is issued as soon as the
// last MMAs of this tile have retired.
#pragma once
#include "k_expdw.cuh"

namespace cf {

constexpr int MBX_THREADS = 256;

template <int KS, int S>
struct MbxGeom {
    // Tiles are sized so that the halo pixel count sits just under 256 = two 128-row MMA blocks = eight 32-row
    // TMEM lane groups, i.e. every epilogue warp drains two full groups (a 153-pixel tile left 3 of the 4
    // schedulers idle for half of the Swish phase): 3x3 s1: 12x16 (252 px), 5x5 s1: 8x16 (240), 3x3 s2: 8x6 (221),
    // 5x5 s2: 8x4 (209).
    static constexpr int TH = (S == 1 && KS == 3) ? 12 : 8;
    static constexpr int TW = S == 1 ? 16 : (KS == 3 ? 6 : 4);
    static constexpr int IH = (TH - 1) * S + KS, IW = (TW - 1) * S + KS;
    static constexpr int NPX = IH * IW;            // halo pixels
    static constexpr int NMB = (NPX + 127) / 128;  // 2
    static constexpr int LO = (KS - S) / 2;
    static constexpr int XBYTES = ((NPX * 128 + 1023) / 1024) * 1024;
    static constexpr int XREG = NMB * 16384;       // the A operand of the last M-block over-reads up to here
    // dw phase: output blocks per thread (x8 float4 lanes): 12x16 -> 2x3, 8x16 -> 2x2, stride 2 -> one output
    static constexpr int XT = S == 1 ? 2 : 1, YT = S == 1 ? (KS == 3 ? 3 : 2) : 1;
    // smem map (from the 1024-aligned base): X | Xlo | E | We ring (2 x 8 KB) | barriers
    static constexpr uint32_t OFF_XLO = XREG, OFF_E = 2 * XREG, OFF_WE = OFF_E + XBYTES, OFF_BARS = OFF_WE + 16384;
    static constexpr size_t SMEM = OFF_BARS + 64 + 1024;
};

struct MbxParams {
    XdParams x;
    const float* we_img;  // [chunk][hi 32 x 128 B | lo 32 x 128 B]
};

template <int KS, int S, int CIN>
__global__ void __launch_bounds__(MBX_THREADS, 2) k_mbx(const __grid_constant__ CUtensorMap tmX, const MbxParams P) {
    using G = MbxGeom<KS, S>;
    static_assert(G::NMB == 2 && G::NPX <= 256, "two M-blocks: 128 TMEM columns");
    const XdParams& p = P.x;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* X = sm;
    uint8_t* Xlo = sm + G::OFF_XLO;
    uint8_t* Es = sm + G::OFF_E;
    const uint32_t we_s = base + G::OFF_WE;
    const uint32_t bars = base + G::OFF_BARS;  // [0] X full, [1] MMA done, [2],[3] We slot full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + G::OFF_BARS + 32);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c4 = lane & 7, pg = lane >> 3;
    const int nch = (p.hid + 31) >> 5;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_items = ((int)blockIdx.x < p.n_items) ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_jobs = my_items * nch;  // job j = (tile j / nch, chunk j % nch)

    auto issue_tma = [&](int item) {  // thread 0 only
        fence_proxy_async();          // X was last written in place by generic-proxy stores
        const int tx = item % p.tiles_x;
        const int t2 = item / p.tiles_x;
        const int ty = t2 % p.tiles_y, b = t2 / p.tiles_y;
        mbar_expect_tx(bars, (uint32_t)G::NPX * 128u);
        tma_load_4d(base, &tmX, 0, tx * G::TW * S - G::LO, ty * G::TH * S - G::LO, b, bars);
    };
    auto issue_we = [&](int job) {  // thread 0 only: weight image of job's chunk -> ring slot job&1
        const int ch = job % nch;
        mbar_expect_tx(bars + 16 + 8 * (job & 1), 8192u);
        bulk_load(we_s + (job & 1) * 8192, P.we_img + (size_t)ch * 2048, 8192u, bars + 16 + 8 * (job & 1));
    };
    const uint32_t idesc = umma_idesc_tf32(32);
    auto issue_mma = [&](int job) {  // thread 0 only: both M-blocks of one 32-channel chunk
        mbar_wait(bars + 16 + 8 * (job & 1), (job >> 1) & 1u);
        tc_fence_after();
        const uint32_t wb = we_s + (job & 1) * 8192;
        const uint64_t b_hi = umma_desc(wb), b_lo = umma_desc(wb + 4096);
#pragma unroll
        for (int mb = 0; mb < G::NMB; ++mb) {
            const uint64_t a_hi = umma_desc(base + mb * 16384), a_lo = umma_desc(base + G::OFF_XLO + mb * 16384);
            const uint32_t d_main = tmem_base + mb * 64, d_corr = d_main + 32;
#pragma unroll
            for (int k = 0; k < (CIN + 7) / 8; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);
                umma_tf32(d_corr, a_lo + ko, b_hi + ko, idesc, k > 0);
                umma_tf32(d_corr, a_hi + ko, b_lo + ko, idesc, 1u);
                umma_tf32(d_main, a_hi + ko, b_hi + ko, idesc, k > 0);
            }
        }
        umma_commit(bars + 8);
    };

    if (tid == 0 && total_jobs > 0) {
        issue_tma(blockIdx.x);
        issue_we(0);
        if (total_jobs > 1) issue_we(1);
    }
    int job = 0;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        mbar_wait(bars, it & 1u);
        const int tx = item % p.tiles_x;
        const int t2 = item / p.tiles_x;
        const int ty = t2 % p.tiles_y, b = t2 / p.tiles_y;

        // split the halo tile once: X <- rn_tf32(X) in place, Xlo <- the exact remainder
        for (int i = tid; i < G::NPX * 8; i += MBX_THREADS) {
            float4* a = reinterpret_cast<float4*>(X) + i;
            const float4 v = *a;
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            *a = h;
            reinterpret_cast<float4*>(Xlo)[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) issue_mma(job);
        for (int ch = 0; ch < nch; ++ch, ++job) {
            const int cbase = ch * 32 + c4 * 4;
            const bool cvalid = cbase < p.hid;
            // ---- drain the accumulators: + correction, Swish, swizzled E rows ----
            mbar_wait(bars + 8, job & 1u);
            tc_fence_after();
            if (tid == 0) {
                // the MMAs of this job have retired: their weight slot is free, and after the tile's last chunk so is X
                if (job + 2 < total_jobs) issue_we(job + 2);
                if (ch == nch - 1 && item + (int)gridDim.x < p.n_items) issue_tma(item + gridDim.x);
            }
            {
                const int q = warp & 3, cg = warp >> 2;  // TMEM lane quarter, 16-column half of the chunk
#pragma unroll
                for (int mb = 0; mb < G::NMB; ++mb) {
                    if (mb * 128 + q * 32 >= G::NPX) break;  // warp-uniform
                    const int px = mb * 128 + q * 32 + lane;
                    float v[16], c[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * 64 + cg * 16);
                    tmem_ld8(taddr, v);
                    tmem_ld8(taddr + 8u, v + 8);
                    tmem_ld8(taddr + 32u, c);
                    tmem_ld8(taddr + 40u, c + 8);
                    tmem_ld_wait();
                    if (px < G::NPX) {
                        uint8_t* er = Es + px * 128;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 o = swish4(make_float4(v[4 * j] + c[4 * j], v[4 * j + 1] + c[4 * j + 1],
                                                                v[4 * j + 2] + c[4 * j + 2], v[4 * j + 3] + c[4 * j + 3]));
                            *reinterpret_cast<float4*>(er + (((4 * cg + j) ^ (px & 7)) << 4)) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();  // E complete; every TMEM read of this chunk is done
            if (tid == 0 && ch + 1 < nch) issue_mma(job + 1);  // overlaps the depth-wise phase below
            xd_dw_phase_g<G, KS, S, G::XT, G::YT, true, MBX_THREADS / 32>(Es, p.Wd, p, warp, pg, c4, cbase, cvalid, b, ty, tx);
            __syncthreads();  // E is rewritten by the next chunk
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------
struct MbxLaunch {
    CUtensorMap tmX;
    MbxParams p;
    int ks = 3, s = 1, cin = 16, grid = 0;
    size_t smem = 0;
};

template <int KS, int S, int CIN>
inline cudaError_t mbx_launch_t(const MbxLaunch& ml, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_mbx<KS, S, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MbxGeom<KS, S>::SMEM);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    k_mbx<KS, S, CIN><<<ml.grid, MBX_THREADS, MbxGeom<KS, S>::SMEM, st>>>(ml.tmX, ml.p);
    return cudaGetLastError();
}

inline cudaError_t mbx_launch(const MbxLaunch& ml, cudaStream_t st) {
    if (ml.ks == 3 && ml.s == 2 && ml.cin == 16) return mbx_launch_t<3, 2, 16>(ml, st);
    if (ml.ks == 3 && ml.s == 2 && ml.cin == 32) return mbx_launch_t<3, 2, 32>(ml, st);
    if (ml.ks == 3 && ml.s == 1 && ml.cin == 24) return mbx_launch_t<3, 1, 24>(ml, st);
    if (ml.ks == 5 && ml.s == 2 && ml.cin == 24) return mbx_launch_t<5, 2, 24>(ml, st);
    if (ml.ks == 5 && ml.s == 1 && ml.cin == 32) return mbx_launch_t<5, 1, 32>(ml, st);
    return cudaErrorInvalidValue;
}

template <int KS, int S>
inline void mbx_geom(int* th, int* tw, int* ih, int* iw) {
    using G = MbxGeom<KS, S>;
    *th = G::TH, *tw = G::TW, *ih = G::IH, *iw = G::IW;
}

inline int mbx_plan(PwTcState& st, int ks, int s, const float* X, const float* We, const float* Wd, float* D, int B, int Hi, int Wi,
                    int cin, int hid, MbxLaunch* ml) {
    if (!xd_supported(ks, s, cin)) return fail(CF_EINVAL, "mbx_plan: no fused kernel for k=%d s=%d cin=%d", ks, s, cin);
    auto it = st.layers.find(We);
    if (it == st.layers.end() || it->second.NC != 32 || it->second.nkb != 1)
        return fail(CF_EINVAL, "mbx_plan: expand weights were not prepared as 32-column tensor-core images");
    int th, tw, ih, iw;
    if (ks == 3 && s == 1) mbx_geom<3, 1>(&th, &tw, &ih, &iw);
    else if (ks == 3) mbx_geom<3, 2>(&th, &tw, &ih, &iw);
    else if (s == 1) mbx_geom<5, 1>(&th, &tw, &ih, &iw);
    else mbx_geom<5, 2>(&th, &tw, &ih, &iw);
    int rc = xd_make_map(st, &ml->tmX, X, B, Hi, Wi, cin, iw, ih);
    if (rc) return rc;
    XdParams& p = ml->p.x;
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
    ml->p.we_img = it->second.img;
    ml->ks = ks;
    ml->s = s;
    ml->cin = cin;
    ml->grid = p.n_items < 2 * st.sms ? p.n_items : 2 * st.sms;
    return CF_OK;
}

}  // namespace cf
