// Shared helpers for the CenterFace B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/centerface_b200.h"

namespace cf {

// ---- thread-local error string ---------------------------------------------------------
inline std::string& err_slot() {
    static thread_local std::string s;
    return s;
}
inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err_slot() = buf;
    return code;
}

#define CF_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return cf::fail(CF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                              \
    } while (0)

#define CF_CHECK(cond, code, ...)                      \
    do {                                               \
        if (!(cond)) return cf::fail(code, __VA_ARGS__); \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel in to more than 48 KB of dynamic shared memory.  Function attributes are per DEVICE (context), and an engine may be
// created on any device of the process, from any host thread: remember per (kernel, device) what has been set.
inline cudaError_t smem_optin(const void* func, int bytes) {
    static std::mutex mu;
    static std::unordered_map<const void*, unsigned long long> done;  // kernel -> bit mask of devices
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> lk(mu);
    unsigned long long& m = done[func];
    if (m & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) m |= bit;
    return e;
}

// Kernel launch with the programmatic-stream-serialization attribute (CF_PDL=0 turns it off: plain stream order).
// Only kernels that call pdl_wait() before touching activation memory may be launched through this.
inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CF_PDL");
        v = e ? (atoi(e) != 0) : 1;
    }
    return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---- device math -----------------------------------------------------------------------
// Swish (model/centernet.py:39-40) x*sigmoid(x) = x * rcp(1 + 2^(-x*log2 e)): MUFU.EX2 + MUFU.RCP (~2 ulp each) and
// three FP32 ops.  The .ftz forms drop the denormal pre/post-scaling code the default intrinsics emit (4 extra
// instructions per element); denormal intermediates only occur for |x| > 87, where the result is x or -0 anyway.
// x -> -inf gives -0, x -> +inf gives x; no NaN is produced for finite x.
__device__ __forceinline__ float swishf(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return x * r;
}

// Packed fp32x2 arithmetic (sm_100: FMUL2 / FADD2 / FFMA2, one issue slot for two IEEE round-to-nearest results -- bit-identical to
// the scalar instructions).  The fused MBConv kernel is bound by issue slots and latency, not by the FMA pipe's flop rate.
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {  // c += a * b
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%0, %1};\n\tfma.rn.f32x2 c, a, b, c;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "+f"(c0), "+f"(c1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// swishf on two values: the same five operations per value, the three FP32 ones packed
__device__ __forceinline__ void swish2(float& x0, float& x1) {
    float t0, t1, e0, e1, r0, r1;
    mul2(t0, t1, x0, x1, -1.4426950408889634f, -1.4426950408889634f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    add2(e0, e1, e0, e1, 1.f, 1.f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(e0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(e1));
    mul2(x0, x1, x0, x1, r0, r1);
}
__device__ __forceinline__ float4 swish4p(float4 v) {
    swish2(v.x, v.y);
    swish2(v.z, v.w);
    return v;
}
// Swish on four values with ONE reciprocal: with d_i = 1 + 2^(-x_i log2 e), 1/d_0 = d_1 . (d_2 d_3) . rcp(d_0 d_1 d_2 d_3) and so on:
// 4 MUFU.EX2 + 1 MUFU.RCP instead of 4 + 4 (-37.5 % of the special-function work that bounds the fused MBConv kernels), the
// same twenty issue slots per four values as two swish2 + the clamps.  The exponent argument is clamped at 31, so that the
// product of four d stays below 2^127 (no inf . 0): sigmoid bottoms out at 2^-31 for x < -21.5, an absolute error below
// |x| . 4.7e-10 where the exact result is within 1e-8 of zero.  Five more roundings than swishf (a few 1e-7 relative at worst;
// tests/test_gpu_variants.py::test_swish_one_reciprocal_per_four holds both forms against fp64).  Used only where a Swish site
// is MUFU-bound and shallow: the fused layer1.0 kernel (362 -> 346 us) and the epilogue of the one-K-block expand layers
// (24 -> 144: 115 -> 109 us; 32 -> 192: 44 -> 41); the stem (producer-bound) and the depth-wise kernels measured flat or slower.
__device__ __forceinline__ void swish4q(float& x0, float& x1, float& x2, float& x3) {
    float t0, t1, t2, t3, d0, d1, d2, d3, p01, p23, r, r01, r23, i0, i1, i2, i3;
    mul2(t0, t1, x0, x1, -1.4426950408889634f, -1.4426950408889634f);
    mul2(t2, t3, x2, x3, -1.4426950408889634f, -1.4426950408889634f);
    t0 = fminf(t0, 31.f), t1 = fminf(t1, 31.f), t2 = fminf(t2, 31.f), t3 = fminf(t3, 31.f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d1) : "f"(t1));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d2) : "f"(t2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d3) : "f"(t3));
    add2(d0, d1, d0, d1, 1.f, 1.f);
    add2(d2, d3, d2, d3, 1.f, 1.f);
    mul2(p01, p23, d0, d2, d1, d3);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    mul2(r01, r23, p23, p01, r, r);  // 1 / (d0 d1), 1 / (d2 d3)
    mul2(i0, i1, d1, d0, r01, r01);
    mul2(i2, i3, d3, d2, r23, r23);
    mul2(x0, x1, x0, x1, i0, i1);
    mul2(x2, x3, x2, x3, i2, i3);
}
__device__ __forceinline__ float4 swish4qv(float4 v) {
    swish4q(v.x, v.y, v.z, v.w);
    return v;
}
__device__ __forceinline__ void fma44p(float4& acc, float4 a, float4 w) {
    fma2(acc.x, acc.y, a.x, a.y, w.x, w.y);
    fma2(acc.z, acc.w, a.z, a.w, w.z, w.w);
}

__device__ __forceinline__ float4 swish4(float4 v) {
    return make_float4(swishf(v.x), swishf(v.y), swishf(v.z), swishf(v.w));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// L2-only load for activations the PREVIOUS kernel of the stream wrote: under programmatic dependent launch this kernel
// may already be resident while they are produced, so they are never read through the non-coherent (.nc) path.
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- programmatic dependent launch ------------------------------------------------------
// Every kernel of the forward chain calls pdl_trigger() first (the next kernel of the stream may be scheduled as soon as
// all CTAs of this one have started) and pdl_wait() before its first access to activation memory (blocks until the
// previous kernel has completed and its writes are visible).  In between sits the prologue that depends on nothing:
// barrier init, TMEM allocation, tensor-map prefetch, weight staging.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ void fma4(float4& acc, float a, float4 w) {
    acc.x = fmaf(a, w.x, acc.x);
    acc.y = fmaf(a, w.y, acc.y);
    acc.z = fmaf(a, w.z, acc.z);
    acc.w = fmaf(a, w.w, acc.w);
}
__device__ __forceinline__ void fma44(float4& acc, float4 a, float4 w) {
    acc.x = fmaf(a.x, w.x, acc.x);
    acc.y = fmaf(a.y, w.y, acc.y);
    acc.z = fmaf(a.z, w.z, acc.z);
    acc.w = fmaf(a.w, w.w, acc.w);
}

// Epilogue kinds shared by the SIMT and tcgen05 point-wise GEMMs.
enum Epi : int {
    EPI_LINEAR = 0,      // project conv, model/centernet.py:118
    EPI_SWISH = 1,       // expand conv + Swish, :110
    EPI_RESIDUAL = 2,    // project + skip add, :137
    EPI_BIAS_SWISH = 3,  // conv_last: conv + folded BN + Swish, :178-184
    EPI_IDAUP = 4,       // IDAUp: relu(BN(lateral)) + relu(BN(convT2x2 dw(low))), :186-204
    EPI_SWISHQ = 5,      // EPI_SWISH with one reciprocal per four values (swish4q); chosen by k_pw_tc's plan for the one-K-block expand layers
};

struct EpiArgs {
    const float* res;   // [M,N]    residual (EPI_RESIDUAL)
    const float* bias;  // [N]      folded BN shift (BIAS_SWISH, IDAUP lateral)
    const float* low;   // [B,Ho/2,Wo/2,N] low-res input of the transposed conv (IDAUP)
    const float* su;    // [2][2][N] folded up-sample scale, sub-pixel major (IDAUP; transposed from the blob's [N][2][2] at cf_create)
    const float* tu;    // [N]      folded up-sample shift (IDAUP)
    int Ho, Wo;         // output map size (IDAUP: m -> (b,y,x))
};

// IDAUp epilogue in two halves, so that a kernel can fetch the low-resolution operand (an L2 read that does not depend on
// the GEMM) before it waits for its accumulator: idaup_low() -> the float4 of `low` under output row m, columns n..n+3,
// and the 2x2 sub-pixel q of that row; idaup_apply() -> relu(BN(convT(low))) + relu(BN(lateral)).
__device__ __forceinline__ const float* idaup_low_ptr(int m, int n, int N, const EpiArgs& ea, int* q) {
    const int x = m % ea.Wo;
    const int t = m / ea.Wo;
    const int y = t % ea.Ho;
    const int b = t / ea.Ho;
    const int Hl = ea.Ho >> 1, Wl = ea.Wo >> 1;
    *q = (y & 1) * 2 + (x & 1);
    return ea.low + ((size_t)(b * Hl + (y >> 1)) * Wl + (x >> 1)) * N + n;
}
// cst (optional): a shared-memory copy [bias N | tu N | su 4N] of the three constant vectors -- through L1 they are nine
// long-scoreboard loads at the very end of a tile's dependency chain
__device__ __forceinline__ float4 idaup_apply(float4 acc, float4 lo, int n, int q, int N, const EpiArgs& ea, const float* cst = nullptr) {
    const float4 bb = cst ? *reinterpret_cast<const float4*>(cst + n) : ldg4(ea.bias + n);
    const float4 tt = cst ? *reinterpret_cast<const float4*>(cst + N + n) : ldg4(ea.tu + n);
    const float4 ss = cst ? *reinterpret_cast<const float4*>(cst + 2 * N + q * N + n)
                          : ldg4(ea.su + q * N + n);  // [2x2 sub-pixel][channel]: one vector load instead of four scalar ones
    float4 o;
    o.x = fmaxf(fmaf(lo.x, ss.x, tt.x), 0.f) + fmaxf(acc.x + bb.x, 0.f);
    o.y = fmaxf(fmaf(lo.y, ss.y, tt.y), 0.f) + fmaxf(acc.y + bb.y, 0.f);
    o.z = fmaxf(fmaf(lo.z, ss.z, tt.z), 0.f) + fmaxf(acc.z + bb.z, 0.f);
    o.w = fmaxf(fmaf(lo.w, ss.w, tt.w), 0.f) + fmaxf(acc.w + bb.w, 0.f);
    return o;
}

template <int EPI>
__device__ __forceinline__ float4 apply_epi(float4 acc, int m, int n, int N, const EpiArgs& ea) {
    if (EPI == EPI_SWISH) return swish4p(acc);
    if (EPI == EPI_SWISHQ) return swish4qv(acc);
    if (EPI == EPI_RESIDUAL) {
        float4 r = ldcg4(ea.res + (size_t)m * N + n);
        return make_float4(acc.x + r.x, acc.y + r.y, acc.z + r.z, acc.w + r.w);
    }
    if (EPI == EPI_BIAS_SWISH) {
        float4 b = ldg4(ea.bias + n);
        return swish4p(make_float4(acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w));
    }
    if (EPI == EPI_IDAUP) {
        int q;
        const float4 lo = ldcg4(idaup_low_ptr(m, n, N, ea, &q));
        return idaup_apply(acc, lo, n, q, N, ea);
    }
    return acc;
}

}  // namespace cf
