// Development probe (not part of the product library): can a K-major SWIZZLE_128B tcgen05 shared-memory descriptor start at
// a row that is NOT a multiple of eight rows (1024 B) into a swizzled tile?
//
// Why: a 3x3 convolution over an NHWC halo tile [rows][cols][32 ch fp32 = 128 B] becomes nine K blocks of ONE GEMM if the A
// operand of tap (dy, dx) is simply "the same tile, starting (dy * cols + dx) pixels later" -- no im2col copy, no per-tap
// barrier (profiles/r1d_headroom.md, item 1: the heads kernel).  TMA writes the tile with the 128-byte swizzle keyed on the
// absolute shared-memory address (16-byte chunk c of 128-byte line r lands at chunk c ^ (r & 7)); the question is whether
// the tensor core applies the same address-keyed pattern when the descriptor's start address is base + s * 128 with
// s % 8 != 0, with the descriptor's base_offset field (bits 49-51) left 0 (mode 0) or set to (start >> 7) & 7 (mode 1).
//
// One CTA.  A tile: 256 lines of 128 B, line r holds A[r][k] = (r + 3k) & 255 (exact in tf32), written swizzled as TMA
// would.  B: 16 x 32, B[n][k] = (k == 2n), so D[m][n] = A[m + s][2n] = (m + s + 6n) & 255 for a correct read.  For every
// (shift s, mode) the kernel issues the four K-step MMAs of one 128 x 16 x 32 product and dumps D; the host counts
// mismatches and, where rows are wrong, reports which line was read instead.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/build/desc_shift_probe tools/desc_shift_probe.cu
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../lightweight-face-detection-centernet_b200/csrc/k_pw_tc.cuh"

using namespace cf;

constexpr int P_ROWS = 256, P_N = 16, P_MAXT = 64;

struct ProbeTests {
    int n;
    int shift[P_MAXT];
    int mode[P_MAXT];
};

__global__ void __launch_bounds__(128, 1) k_desc_shift_probe(const ProbeTests tests, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a_sm = base, b_sm = base + P_ROWS * 128, bar = b_sm + 2048;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + P_ROWS * 128 + 2048 + 16);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < P_ROWS * 32; i += 128) {
        const int r = i >> 5, k = i & 31;
        *reinterpret_cast<float*>(sm + r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4) = (float)((r + 3 * k) & 255);
    }
    for (int i = tid; i < P_N * 32; i += 128) {
        const int n = i >> 5, k = i & 31;
        *reinterpret_cast<float*>(sm + P_ROWS * 128 + n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4) = (k == 2 * n) ? 1.f : 0.f;
    }
    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_tf32(P_N);

    for (int t = 0; t < tests.n; ++t) {
        if (warp == 0) {
            const uint32_t start = a_sm + (uint32_t)tests.shift[t] * 128u;
            uint64_t a_desc = umma_desc(start);
            if (tests.mode[t] == 1) a_desc |= (uint64_t)((start >> 7) & 7u) << 49;  // base_offset
            const uint64_t b_desc = umma_desc(b_sm);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, k > 0 ? 1u : 0u);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, (uint32_t)t & 1u);
        tc_fence_after();
        float v[16];
        uint32_t* r = reinterpret_cast<uint32_t*>(v);
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
        tmem_ld_wait();
        for (int n = 0; n < P_N; ++n) out[((size_t)t * 128 + tid) * P_N + n] = v[n];
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
    }
}

// usage: desc_shift_probe <mode> [shift ...]   (one mode per process: a faulting descriptor form must not take the other down)
int main(int argc, char** argv) {
    ProbeTests T = {};
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int dflt[] = {0, 8, 16, 1, 2, 3, 4, 7, 9, 18, 19, 20, 36, 37, 38, 100};
    if (argc > 2) {
        for (int i = 2; i < argc && T.n < P_MAXT; ++i) T.shift[T.n] = atoi(argv[i]), T.mode[T.n] = mode, ++T.n;
    } else {
        for (int s : dflt) T.shift[T.n] = s, T.mode[T.n] = mode, ++T.n;
    }
    for (int t = 0; t < T.n; ++t)
        if (T.shift[t] < 0 || T.shift[t] > P_ROWS - 128) return printf("shift %d outside [0, %d]\n", T.shift[t], P_ROWS - 128), 1;
    float* d_out = nullptr;
    const size_t n_out = (size_t)T.n * 128 * P_N;
    if (cudaMalloc(&d_out, n_out * 4) != cudaSuccess) return printf("cudaMalloc failed\n"), 1;
    cudaMemset(d_out, 0xff, n_out * 4);
    const size_t smem = P_ROWS * 128 + 2048 + 64 + 1024;
    cudaFuncSetAttribute(k_desc_shift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_desc_shift_probe<<<1, 128, smem>>>(T, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return printf("kernel failed: %s\n", cudaGetErrorString(e)), 1;
    std::vector<float> h(n_out);
    cudaMemcpy(h.data(), d_out, n_out * 4, cudaMemcpyDeviceToHost);
    printf("| shift (128-byte lines) | base_offset field | wrong elements of 2048 | rows read (m = 0, 1, 7, 8, 127), from column 0 |\n|---|---|---|---|\n");
    for (int t = 0; t < T.n; ++t) {
        const int s = T.shift[t];
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < P_N; ++n)
                if (h[((size_t)t * 128 + m) * P_N + n] != (float)((m + s + 6 * n) & 255)) ++bad;
        printf("| %d | %s | %d |", s, T.mode[t] ? "(start >> 7) & 7" : "0", bad);
        const int ms[] = {0, 1, 7, 8, 127};
        for (int m : ms) printf(" %g", h[((size_t)t * 128 + m) * P_N]);  // column 0 = A[row read][0] = the line index that was read
        printf(" (expected");
        for (int m : ms) printf(" %d", (m + s) & 255);
        printf(") |\n");
    }
    cudaFree(d_out);
    return 0;
}
