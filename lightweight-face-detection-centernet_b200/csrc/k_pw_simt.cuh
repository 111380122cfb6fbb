// Point-wise (1x1) convolution as an fp32 FFMA GEMM: out[M,N] = epi(A[M,K] . W[K,N]).
// This is the *validation engine* (CF_PW_SIMT): exact fp32 accumulation, used to cross-check
// the tcgen05 path on the device and as the first correct path.  M = B*H*W pixels (NHWC rows).
#pragma once
#include "common.cuh"

namespace cf {

// CTA tile 128 x BN, K chunks of 32, 256 threads; thread tile TM x 4 with
// BN=64: TM=8 (16x16 thread grid), BN=32: TM=4 (8x32 thread grid), BN=16: TM=2 (4x64).
template <int BN, int EPI>
__global__ void __launch_bounds__(256) k_pw_simt(const float* __restrict__ A, const float* __restrict__ Wkn,
                                                 float* __restrict__ out, int M, int K, int N, EpiArgs ea) {
    constexpr int BM = 128, BK = 32, LDA = BK + 4;
    constexpr int TXN = BN / 4;      // threads along N
    constexpr int TYN = 256 / TXN;   // threads along M
    constexpr int TM = BM / TYN;     // rows per thread
    __shared__ __align__(16) float As[BM * LDA];
    __shared__ __align__(16) float Ws[BK * BN];

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    float4 acc[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i] = make_float4(0, 0, 0, 0);

    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: 128 rows x 8 float4
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;
            const int r = idx >> 3, c4 = idx & 7;
            const int gm = m0 + r, gk = k0 + c4 * 4;
            float4 v = make_float4(0, 0, 0, 0);
            if (gm < M && gk < K) v = ldg4(A + (size_t)gm * K + gk);
            st4(As + r * LDA + c4 * 4, v);
        }
        // W tile: 32 rows x BN/4 float4
        for (int idx = tid; idx < BK * TXN; idx += 256) {
            const int kr = idx / TXN, n4 = idx % TXN;
            const int gk = k0 + kr, gn = n0 + n4 * 4;
            float4 v = make_float4(0, 0, 0, 0);
            if (gk < K && gn < N) v = ldg4(Wkn + (size_t)gk * N + gn);
            st4(Ws + kr * BN + n4 * 4, v);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float4 a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty * TM + i) * LDA + kk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(Ws + (kk + j) * BN + tx * 4);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const float av = j == 0 ? a[i].x : j == 1 ? a[i].y : j == 2 ? a[i].z : a[i].w;
                    fma4(acc[i], av, wv);
                }
            }
        }
        __syncthreads();
    }

    const int gn = n0 + tx * 4;
    if (gn >= N) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm < M) st4(out + (size_t)gm * N + gn, apply_epi<EPI>(acc[i], gm, gn, N, ea));
    }
}

}  // namespace cf
