// C-ABI implementation of the B200-native CenterFace engine (see include/centerface_b200.h).
//
// Path: EfficientNet.forward (model/centernet.py:263-280) -> sigmoid+clamp
// (centerface.py:43) -> decode (centerface_ext.py:52-82 | centerface.py:73-151 |
// eval_widerface.py:92-152).  Activations live in HBM as NHWC fp32 (pixels are GEMM rows, the
// channel axis is contiguous so every kernel moves float4s); see DESIGN.md for the layout.
//
// There is no CPU fallback in this file: every entry point either launches kernels on an
// sm_100 device or fails with an error code.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>

#include <functional>
#include <vector>

#include "common.cuh"
#include "k_conv.cuh"
#include "k_decode.cuh"
#include "k_pw_simt.cuh"
#include "k_pw_tc.cuh"
#include "k_dwt.cuh"
#include "k_pwn.cuh"
#include "k_mbf.cuh"
#include "k_heads_tc.cuh"
#include "k_stem_tc.cuh"
#include "net.hpp"

using namespace cf;

namespace {

enum WorkClass { CLS_ALL = 0, CLS_PW = 1, CLS_DW = 2, CLS_STEM = 3, CLS_HEADS = 4, CLS_DECODE = 5, CLS_FUSED = 6, CLS_COUNT = 7 };

inline bool engine_is_tc(int pw) { return pw != CF_PW_SIMT; }
inline int engine_passes(int pw) { return pw == CF_PW_TCGEN05_1P ? 1 : 3; }
// CF_PW_TCGEN05_MIXED: single TF32 pass for the point-wise convs of the stride-16/32 stages (layer3.1 .. layer6.0, i.e.
// every GEMM with K or N >= 384), 3xTF32 everywhere else.  Those stages are the precision-insensitive ones (SURVEY.md
// 7.3-2: quantising them to 11 bits moves the heat-map by 1e-4 .. 8e-6) and the tensor-bound ones.
inline int layer_passes(int pw, int K, int N) {
    if (pw == CF_PW_TCGEN05_MIXED) return (K >= 384 || N >= 384) ? 1 : 3;
    return engine_passes(pw);
}
// Whole-block fusion (k_mbf): the default tensor-core engine runs the blocks of this mask (bit i = block i) as one kernel each;
// CF_PW_TCGEN05_LAYERWISE is the same engine without it.  The mask holds the blocks whose fused kernel beats its three
// layer-wise launches on the device (profiles/r2_mbf.md); CF_MBF overrides it for A/B runs.
constexpr unsigned kMbfDefaultMask = 0x6u;  // layer1.0 (616 -> 460 us), layer1.1 (421 -> 382 us)
inline unsigned mbf_mask() {
    if (const char* ev = getenv("CF_MBF")) return (unsigned)strtoul(ev, nullptr, 0);
    return kMbfDefaultMask;
}
// Depth-wise + projection of a block as one kernel (k_mbf direct mode) where the whole-block fusion does not pay: the expand conv, if
// any, stays a k_pw_tc launch.  CF_MBD overrides the mask.
constexpr unsigned kMbdDefaultMask = 0x1u;  // layer0 (no expand conv: the whole block is one kernel): 214 us against 143 + 115
inline unsigned mbd_mask() {
    if (const char* ev = getenv("CF_MBD")) return (unsigned)strtoul(ev, nullptr, 0);
    return kMbdDefaultMask;
}
inline bool block_is_mbd(int pw, int i);
inline bool block_is_mbf(int pw, int i) {
    const MBBlock& b = kBlocks[i];
    return pw == CF_PW_TCGEN05 && b.t != 1 && ((mbf_mask() >> i) & 1u) && mbf_supported(b.k, b.s, b.cin, b.hid(), b.cout);
}

// Entry points that work on an engine's device switch to it for their own duration only: inside a PyTorch process a bare
// cudaSetDevice would silently move the calling thread's current device.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define CF_ON_DEVICE(dev)                                                                  \
    DeviceGuard _dg(dev);                                                                  \
    if (!_dg.ok) return cf::fail(CF_ECUDA, "cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(cudaGetLastError()))

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------
// The one exchange step of the path is an all-gather of the final box list.  The library does not link NCCL:
// the first cf_comm_* call binds the copy that is already in the process (PyTorch loads its own) or dlopens libnccl.so.2.
// Prototypes restated from nccl.h (NCCL 2.x ABI): ncclUniqueId is 128 opaque bytes, ncclFloat32 = 7, ncclSuccess = 0.
struct NcclId {
    char internal[128];
};
struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;  // ncclUniqueId is passed BY VALUE
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
inline NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy PyTorch (or the host program) already loaded
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
        api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
        api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
    });
    return api;
}

inline bool block_is_mbd(int pw, int i) {
    const MBBlock& b = kBlocks[i];
    return pw == CF_PW_TCGEN05 && !block_is_mbf(pw, i) && ((mbd_mask() >> i) & 1u) && mbf_direct_supported(b.k, b.s, b.hid(), b.cout);
}

struct Step {
    int cls;
    std::function<cudaError_t(cudaStream_t)> run;
};

struct EntrySpec {
    std::string name;
    uint64_t count;
};

// The packed-blob contract shared with weights.py::pack_weights (names, float counts, order).
std::vector<EntrySpec> expected_entries() {
    std::vector<EntrySpec> v;
    v.push_back({"stem.w", 27 * 32});  // [(ky*3+kx)*3+ci][co]
    v.push_back({"lut", 768});         // [c][256] u8 -> normalised fp32 (centerface.py:32-34)
    for (int i = 0; i < 12; ++i) {
        const MBBlock& b = kBlocks[i];
        const std::string p = "b" + std::to_string(i);
        if (b.t != 1) v.push_back({p + ".exp", (uint64_t)b.cin * b.hid()});  // [K=cin][N=hid]
        v.push_back({p + ".dw", (uint64_t)b.k * b.k * b.hid()});             // [k*k][hid]
        v.push_back({p + ".proj", (uint64_t)b.hid() * b.cout});              // [K=hid][N=cout]
    }
    v.push_back({"clast.w", 320 * 24});  // BN scale folded
    v.push_back({"clast.b", 24});
    const int skip_c[3] = {96, 32, 24};
    for (int j = 0; j < 3; ++j) {
        const std::string p = "up" + std::to_string(j + 1);
        v.push_back({p + ".w", (uint64_t)skip_c[j] * 24});  // lateral 1x1, BN scale folded
        v.push_back({p + ".b", 24});
        v.push_back({p + ".su", 24 * 4});  // [c][a][b] transposed-conv tap * BN scale
        v.push_back({p + ".tu", 24});
    }
    v.push_back({"heads.w", 216 * 16});  // [(ky*3+kx)*24+ci][16]
    v.push_back({"heads.b", 16});
    return v;
}

constexpr uint64_t kEntryAlignFloats = 32;  // 128-byte aligned entries (float4 / TMA friendly)
inline uint64_t round_up(uint64_t a, uint64_t b) { return (a + b - 1) / b * b; }

size_t blob_bytes() {
    const auto ents = expected_entries();
    uint64_t tab = sizeof(BlobHeader) + ents.size() * sizeof(BlobEntry);
    tab = round_up(tab, 128);
    uint64_t fl = 0;
    for (auto& e : ents) fl += round_up(e.count, kEntryAlignFloats);
    return (size_t)(tab + fl * 4);
}

}  // namespace

struct cf_engine {
    int device = 0;
    int max_batch = 0, max_h = 0, max_w = 0;
    int pw_engine = CF_PW_SIMT;
    float* d_w = nullptr;
    std::map<std::string, const float*> w;
    // activations (NHWC fp32)
    float* stem = nullptr;
    float* hidA = nullptr;  // expand output
    float* hidB = nullptr;  // depth-wise output
    float* blk[12] = {};
    float* clast = nullptr;
    float* up[3] = {};
    float* up_sut[3] = {};  // IDAUp scales transposed to [2x2 sub-pixel][24 channels]
    // heads (planar) + decode scratch
    float *hm = nullptr, *wh = nullptr, *lm = nullptr, *reg = nullptr, *hm_sig = nullptr, *peak = nullptr;
    // staging for the host entry points
    uint8_t* in_u8 = nullptr;
    float* o_dets = nullptr;
    float* o_lms = nullptr;
    int32_t* o_inds = nullptr;
    int32_t* o_counts = nullptr;
    size_t o_dets_floats = 0, o_lms_floats = 0, o_inds_n = 0;
    cudaStream_t stream = nullptr;  // used by the *_host entry points
    // last forward
    int B = 0, H = 0, W = 0, fmt = -1;
    const void* in = nullptr;
    std::vector<Step> plan;          // launch list of the current (input, format, shape)
    // The network launches of a plan replayed as ONE CUDA graph (kernel nodes with their programmatic-dependent-launch edges):
    // at batch 1 the forward is ~40 launches of a few microseconds each and the launch path, not the kernels, sets the latency.
    // Captured on the plan's second run (the first, eager one has opted every kernel in to its shared memory); CF_GRAPH=0 or a
    // failed capture falls back to eager launches.
    cudaGraphExec_t gexec = nullptr;
    int graph_state = 0;  // 0 not tried, 1 usable, -1 capture failed
    int plan_runs = 0, plan_net_launches = 0;
    struct CachedPlan {
        const void* in;
        int fmt, B, H, W;
        std::vector<Step> steps;
        cudaGraphExec_t gexec;
        int graph_state, plan_runs, plan_net_launches;
    };
    std::vector<CachedPlan> plan_cache;  // the pipelined host path alternates between two input buffers
    // pipelined host entry points: two input slots, copy stream, events
    uint8_t* in_slot[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    // device-side cv2.resize: source staging + coefficient tables of the last (source, network) size pair
    uint8_t* src_u8 = nullptr;
    size_t src_u8_bytes = 0;
    int32_t* rs_tab = nullptr;
    size_t rs_tab_ints = 0;
    ResizeTables rs;
    long long submitted = 0, waited = 0;
    long long launches = 0;
    void* comm = nullptr;      // ncclComm_t of cf_comm_init
    int comm_ranks = 1, comm_rank = 0;
    float* o_gather = nullptr;  // [ranks * max_batch, K <= 1024... sized at cf_comm_init for K = 100 .. 1024] gathered boxes
    size_t o_gather_floats = 0;
    // The exchange step runs on its own stream behind an event of the decode kernel, with per-slot send / receive buffers: the
    // all-gather synchronises the ranks, and on the compute stream every rank's next forward waited for the slowest rank of the
    // current step (end-to-end efficiency 0.95 at 8 GPUs).  CF_XCHG_STREAM=0 = the exchange on the compute stream.
    cudaStream_t xchg_stream = nullptr;
    cudaEvent_t ev_topk[2] = {nullptr, nullptr};
    float* o_send[2] = {nullptr, nullptr};    // [max_batch, 1024, 6] this rank's boxes of the submission in slot i
    float* o_gather2[2] = {nullptr, nullptr}; // [ranks * max_batch, 1024, 6] per slot
    PwTcState tc;  // tensor maps etc. of the tcgen05 engine
    StemW stem_w;  // host copy: the stem weights are passed to the kernel by value
    HeadsW heads_w;  // likewise the collapsed head conv
};

namespace {

int run_steps(cf_engine* e, int which, cudaStream_t s) {
    static const bool sync_each = getenv("CF_SYNC_EACH") != nullptr;  // development: attribute a device-side fault to its launch
    int idx = 0;
    for (auto& st : e->plan) {
        ++idx;
        if (which != CLS_ALL && st.cls != which) continue;
        if (st.cls == CLS_DECODE && which == CLS_ALL) continue;  // decode is launched by its own entry points
        cudaError_t err = st.run(s);
        if (err == cudaSuccess && sync_each) err = cudaStreamSynchronize(s);
        if (err != cudaSuccess) return fail(CF_ECUDA, "launch %d of the plan (class %d) failed: %s", idx - 1, st.cls, cudaGetErrorString(err));
        ++e->launches;
    }
    return CF_OK;
}

template <int EPI>
cudaError_t launch_pw_simt(const float* A, const float* Wkn, float* out, int M, int K, int N, EpiArgs ea,
                           cudaStream_t s) {
    if (N >= 64) {
        dim3 g(cdiv(M, 128), cdiv(N, 64));
        k_pw_simt<64, EPI><<<g, 256, 0, s>>>(A, Wkn, out, M, K, N, ea);
    } else if (N > 16) {
        dim3 g(cdiv(M, 128), cdiv(N, 32));
        k_pw_simt<32, EPI><<<g, 256, 0, s>>>(A, Wkn, out, M, K, N, ea);
    } else {
        dim3 g(cdiv(M, 128), 1);
        k_pw_simt<16, EPI><<<g, 256, 0, s>>>(A, Wkn, out, M, K, N, ea);
    }
    return cudaGetLastError();
}

cudaError_t launch_pw_simt_any(int epi, const float* A, const float* Wkn, float* out, int M, int K, int N, EpiArgs ea,
                               cudaStream_t s) {
    switch (epi) {
        case EPI_LINEAR: return launch_pw_simt<EPI_LINEAR>(A, Wkn, out, M, K, N, ea, s);
        case EPI_SWISH: return launch_pw_simt<EPI_SWISH>(A, Wkn, out, M, K, N, ea, s);
        case EPI_RESIDUAL: return launch_pw_simt<EPI_RESIDUAL>(A, Wkn, out, M, K, N, ea, s);
        case EPI_BIAS_SWISH: return launch_pw_simt<EPI_BIAS_SWISH>(A, Wkn, out, M, K, N, ea, s);
        default: return launch_pw_simt<EPI_IDAUP>(A, Wkn, out, M, K, N, ea, s);
    }
}

// One point-wise convolution of the plan: fp32 FFMA tiles (validation engine) or the tcgen05 kernel,
// whose tensor maps and shared-memory plan are resolved here, once per batch shape.
int make_pw_step(cf_engine* e, std::vector<Step>& P, int epi, const float* A, const float* Wkn, float* out, int M, int K,
                 int N, EpiArgs ea) {
    if (e->pw_engine == CF_PW_SIMT) {
        P.push_back({CLS_PW, [=](cudaStream_t s) { return launch_pw_simt_any(epi, A, Wkn, out, M, K, N, ea, s); }});
        return CF_OK;
    }
    const int passes = layer_passes(e->pw_engine, K, N);
    if (passes == 3) {  // narrow layers: the role-free kernel
        auto it = e->tc.layers.find(Wkn);
        if (it != e->tc.layers.end() && pwn_eligible(it->second) && tc_tune_for(K, N, passes).pwn != 0) {
            PwnLaunch pl;
            int rc = pwn_plan(e->tc, epi, A, Wkn, out, M, K, N, ea, &pl);
            if (rc) return rc;
            P.push_back({CLS_PW, [pl](cudaStream_t s) { return pwn_launch(pl, s); }});
            return CF_OK;
        }
    }
    TcLaunch tl;
    int rc = tc_plan(e->tc, passes, epi, A, Wkn, out, M, K, N, ea, &tl);
    if (rc) return rc;
    P.push_back({CLS_PW, [tl](cudaStream_t s) { return tc_launch(tl, s); }});
    return CF_OK;
}

template <int KS, int S>
cudaError_t launch_dw_t(const float* in, const float* w, float* out, int B, int Hi, int Wi, int C, int Ho, int Wo,
                        cudaStream_t s) {
    constexpr int XT = 4;
    const int tiles_x = cdiv(Wo, 4 * XT), tiles_y = cdiv(Ho, 8);
    const int ctas = B * tiles_x * tiles_y;
    int csplit = 1;  // aim for >= 8 CTAs per SM; keep >= 8 channel groups (32 channels) per CTA
    while (ctas * csplit < 148 * 8 && (C / 4) / (csplit * 2) >= 8) csplit *= 2;
    k_dw<KS, S, XT><<<ctas * csplit, 256, 0, s>>>(in, w, out, B, Hi, Wi, C, Ho, Wo, tiles_x, tiles_y, csplit);
    return cudaGetLastError();
}

template <int KS, int S, int VEC, int XT, int TY>
cudaError_t launch_dw3_t(const float* in, const float* w, float* out, int B, int Hi, int Wi, int C, int Ho, int Wo, int cvl,
                         cudaStream_t s) {
    const int cv = C / VEC;
    const int nstrip = cdiv(Wo, XT), ncvh = cv / cvl, ntile_y = cdiv(Ho, TY);
    const long long ntasks = (long long)B * ntile_y * ncvh * nstrip * cvl;
    k_dw3<KS, S, VEC, XT, TY><<<(unsigned)((ntasks + 255) / 256), 256, 0, s>>>(in, w, out, B, Hi, Wi, C, Ho, Wo, cvl, nstrip, ncvh,
                                                                                ntile_y, ntasks);
    return cudaGetLastError();
}

cudaError_t launch_dw(int ks, int st, const float* in, const float* w, float* out, int B, int Hi, int Wi, int C,
                      int Ho, int Wo, cudaStream_t s) {
    // register-tiled kernel when the channel vectors split into lane groups of >= 128 contiguous bytes
    const int vec = ks == 5 ? 2 : 4;
    if (C % vec == 0) {
        const int cv = C / vec;
        int cvl = 32;
        while (cvl > 1 && cv % cvl != 0) cvl >>= 1;
        if (cvl * vec * 4 >= 128) {
            if (ks == 3 && st == 1) return launch_dw3_t<3, 1, 4, 2, 4>(in, w, out, B, Hi, Wi, C, Ho, Wo, cvl, s);
            if (ks == 3 && st == 2) return launch_dw3_t<3, 2, 4, 2, 4>(in, w, out, B, Hi, Wi, C, Ho, Wo, cvl, s);
            if (ks == 5 && st == 1) return launch_dw3_t<5, 1, 2, 4, 4>(in, w, out, B, Hi, Wi, C, Ho, Wo, cvl, s);
            return launch_dw3_t<5, 2, 2, 4, 4>(in, w, out, B, Hi, Wi, C, Ho, Wo, cvl, s);
        }
    }
    if (ks == 3 && st == 1) return launch_dw_t<3, 1>(in, w, out, B, Hi, Wi, C, Ho, Wo, s);
    if (ks == 3 && st == 2) return launch_dw_t<3, 2>(in, w, out, B, Hi, Wi, C, Ho, Wo, s);
    if (ks == 5 && st == 1) return launch_dw_t<5, 1>(in, w, out, B, Hi, Wi, C, Ho, Wo, s);
    return launch_dw_t<5, 2>(in, w, out, B, Hi, Wi, C, Ho, Wo, s);
}

// Build the launch list of EfficientNet.forward (model/centernet.py:263-280) for one batch shape.
int build_plan(cf_engine* e, const void* input, int fmt, int B, int H, int W) {
    e->plan.clear();
    e->B = 0;
    int rc = CF_OK;
    auto& P = e->plan;
    const int H2 = H / 2, W2 = W / 2;
    // first_conv, :224
    {
        const StemW w = e->stem_w;
        const float* lut = e->w["lut"];
        float* out = e->stem;
        const long long thr = (long long)B * H2 * W2;
        const unsigned grid = (unsigned)((thr + 127) / 128);
        // The tensor-core engines run the stem on tcgen05 too (k_stem_tc2: warp-specialised, pipelined; 173 us per 32-image
        // batch against 199 us for the FFMA stem).  CF_STEM_TC=0 selects the FFMA stem, =1 the role-free first-generation
        // tcgen05 stem (parity-green, 255 us: its gather -> split -> MMA -> drain chain is serial per tile).
        const char* ev = getenv("CF_STEM_TC");
        const int stem_mode = ev ? atoi(ev) : 2;
        if (engine_is_tc(e->pw_engine) && stem_mode == 2) {
            StcParams sp{};
            int sgrid = 0;
            if ((rc = stc_plan(e->tc, e->w["stem.w"], input, lut, out, B, H, W, &sp, &sgrid))) return rc;
            const int sms = e->tc.sms;
            // u8 images whose rows are whole 32-bit words take the third-generation kernel (CF_STEM_TC=2 keeps the second)
            if (fmt == CF_IN_U8_HWC && !ev && stc3_supported(sp))
                P.push_back({CLS_STEM, [=](cudaStream_t s) { return stc3_launch(sp, sms, s); }});
            else if (fmt == CF_IN_U8_HWC)
                P.push_back({CLS_STEM, [=](cudaStream_t s) { return stc2_launch_t<1>(sp, sms, s); }});
            else
                P.push_back({CLS_STEM, [=](cudaStream_t s) { return stc2_launch_t<0>(sp, sms, s); }});
        } else if (engine_is_tc(e->pw_engine) && stem_mode == 1) {
            StcParams sp{};
            int sgrid = 0;
            if ((rc = stc_plan(e->tc, e->w["stem.w"], input, lut, out, B, H, W, &sp, &sgrid))) return rc;
            if (fmt == CF_IN_U8_HWC)
                P.push_back({CLS_STEM, [=](cudaStream_t s) { return stc_launch_t<1>(sp, sgrid, s); }});
            else
                P.push_back({CLS_STEM, [=](cudaStream_t s) { return stc_launch_t<0>(sp, sgrid, s); }});
        } else if (fmt == CF_IN_U8_HWC)
            P.push_back({CLS_STEM, [=](cudaStream_t s) {
                             return launch_pdl(k_stem<1>, dim3(grid), dim3(128), 0, s, input, w, lut, out, B, H, W);
                         }});
        else
            P.push_back({CLS_STEM, [=](cudaStream_t s) {
                             return launch_pdl(k_stem<0>, dim3(grid), dim3(128), 0, s, input, w, lut, out, B, H, W);
                         }});
    }
    // layer0..layer6, :225-235 -> MBConvBlock.forward :128-140
    const float* x = e->stem;
    int h = H2, wd = W2;
    for (int i = 0; i < 12; ++i) {
        const MBBlock& b = kBlocks[i];
        const std::string p = "b" + std::to_string(i);
        const int hid = b.hid();
        const float* dw_in = x;
        const int ho = h / b.s, wo = wd / b.s;
        if (block_is_mbf(e->pw_engine, i)) {
            // expand + Swish + depth-wise + Swish + projection (+ residual) in one kernel: neither hidden tensor reaches HBM
            MbfLaunch ml;
            if ((rc = mbf_plan(e->tc, b.k, b.s, x, e->w[p + ".exp"], e->w[p + ".dw"], e->w[p + ".proj"], e->blk[i], b.residual() ? x : nullptr, B, h,
                               wd, b.cin, hid, b.cout, &ml)))
                return rc;
            P.push_back({CLS_FUSED, [ml](cudaStream_t s) { return mbf_launch(ml, s); }});
            x = e->blk[i];
            h = ho;
            wd = wo;
            continue;
        }
        if (b.t != 1) {
            const float* wexp = e->w[p + ".exp"];
            float* o = e->hidA;
            const int M = B * h * wd, K = b.cin;
            if ((rc = make_pw_step(e, P, EPI_SWISH, x, wexp, o, M, K, hid, EpiArgs{}))) return rc;
            dw_in = e->hidA;
        }
        if (block_is_mbd(e->pw_engine, i)) {
            // depth-wise + Swish + projection (+ residual) in one kernel: the depth-wise output never reaches HBM
            MbfLaunch ml;
            if ((rc = mbf_plan_direct(e->tc, b.k, b.s, dw_in, e->w[p + ".dw"], e->w[p + ".proj"], e->blk[i], b.residual() ? x : nullptr, B, h, wd,
                                      hid, b.cout, &ml)))
                return rc;
            P.push_back({CLS_FUSED, [ml](cudaStream_t s) { return mbf_launch(ml, s); }});
            x = e->blk[i];
            h = ho;
            wd = wo;
            continue;
        }
        {
            const float* wdw = e->w[p + ".dw"];
            float* o = e->hidB;
            const int ks = b.k, st = b.s, hi = h, wi = wd;
            if (engine_is_tc(e->pw_engine) && dwt_supported(hid)) {  // TMA-fed depth-wise kernel
                DwtLaunch dl;
                if ((rc = dwt_plan(e->tc, ks, st, dw_in, wdw, o, B, hi, wi, hid, &dl))) return rc;
                P.push_back({CLS_DW, [dl](cudaStream_t s) { return dwt_launch(dl, s); }});
            } else {
                P.push_back({CLS_DW, [=](cudaStream_t s) { return launch_dw(ks, st, dw_in, wdw, o, B, hi, wi, hid, ho, wo, s); }});
            }
        }
        {
            const float* wpr = e->w[p + ".proj"];
            const float* a = e->hidB;
            float* o = e->blk[i];
            const int M = B * ho * wo, N = b.cout;
            EpiArgs ea{};
            int epi = EPI_LINEAR;
            if (b.residual()) {
                epi = EPI_RESIDUAL;
                ea.res = x;
            }
            if ((rc = make_pw_step(e, P, epi, a, wpr, o, M, hid, N, ea))) return rc;
        }
        x = e->blk[i];
        h = ho;
        wd = wo;
    }
    // conv_last, :236 (conv 1x1 + folded BN + Swish)
    {
        const float* a = e->blk[11];
        const float* wl = e->w["clast.w"];
        EpiArgs ea{};
        ea.bias = e->w["clast.b"];
        float* o = e->clast;
        const int M = B * h * wd;
        if ((rc = make_pw_step(e, P, EPI_BIAS_SWISH, a, wl, o, M, 320, 24, ea))) return rc;
    }
    // up1..up3, :237-239, :270-272 -> IDAUp.forward :200-204
    const float* low = e->clast;
    const int skip_blk[3] = {8, 4, 2};  // x4, x2, x1
    const int skip_c[3] = {96, 32, 24};
    for (int j = 0; j < 3; ++j) {
        h *= 2;
        wd *= 2;
        const std::string p = "up" + std::to_string(j + 1);
        EpiArgs ea{};
        ea.bias = e->w[p + ".b"];
        ea.low = low;
        ea.su = e->up_sut[j];
        ea.tu = e->w[p + ".tu"];
        ea.Ho = h;
        ea.Wo = wd;
        const float* a = e->blk[skip_blk[j]];
        const float* wl = e->w[p + ".w"];
        float* o = e->up[j];
        const int M = B * h * wd, K = skip_c[j];
        if ((rc = make_pw_step(e, P, EPI_IDAUP, a, wl, o, M, K, 24, ea))) return rc;
        low = e->up[j];
    }
    // heads, :240-261, :277-279 (+ sigmoid/clamp of centerface.py:43)
    if (engine_is_tc(e->pw_engine) && !getenv("CF_HEADS_FFMA")) {  // tap-shifted tcgen05 GEMM (CF_HEADS_FFMA: the fp32 FFMA kernel, for A/B runs)
        HeadsTcLaunch hl;
        if ((rc = heads_tc_plan(e->tc, e->up[2], e->heads_w.b, e->hm, e->wh, e->lm, e->reg, e->hm_sig, B, h, wd, &hl))) return rc;
        P.push_back({CLS_HEADS, [hl](cudaStream_t s) { return heads_tc_launch(hl, s); }});
    } else {
        const float* a = e->up[2];
        const int hh = h, ww = wd;
        P.push_back({CLS_HEADS, [=](cudaStream_t s) {
                         dim3 g(cdiv(ww, 32), cdiv(hh, 16), B);
                         return launch_pdl(k_heads, g, dim3(128), (size_t)HEADS_SMEM, s, a, e->heads_w, e->hm, e->wh, e->lm, e->reg, e->hm_sig,
                                           B, hh, ww);
                     }});
    }
    // path-C decode on the heads (entered through cf_decode_topk; listed for cf_replay_class)
    {
        const int hh = h, ww = wd;
        if (topk_fused(1, hh, ww)) {  // one launch: peak keep on the top-k kernel's shared copy of the map
            P.push_back({CLS_DECODE, [=](cudaStream_t s) {
                             return launch_topk(e->hm_sig, e->peak, e->wh, e->reg, B, hh, ww, 100, e->o_dets, e->o_inds, s);
                         }});
        } else {
            P.push_back({CLS_DECODE, [=](cudaStream_t s) {
                             const long long n = (long long)B * hh * ww;
                             return launch_pdl(k_peak_mask, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, (const float*)e->hm_sig, e->peak, B,
                                               hh, ww);
                         }});
            P.push_back({CLS_DECODE, [=](cudaStream_t s) {
                             return launch_pdl(k_topk<false>, dim3(B), dim3(1024), 0, s, (const float*)e->peak, (const float*)e->wh, (const float*)e->reg, hh, ww,
                                               100, e->o_dets, e->o_inds, 0, 1, 2);
                         }});
        }
    }
    e->in = input;
    e->fmt = fmt;
    e->B = B;
    e->H = H;
    e->W = W;
    return CF_OK;
}

int dalloc(float** p, size_t floats) {
    CF_CUDA(cudaMalloc((void**)p, floats * sizeof(float)));
    return CF_OK;
}

int check_decode_args(int batch, int h, int w) {
    CF_CHECK(batch > 0 && h > 0 && w > 0, CF_EINVAL, "decode: batch=%d h=%d w=%d must be positive", batch, h, w);
    CF_CHECK((long long)h * w < (1ll << 24), CF_EINVAL, "decode: map %dx%d too large (flat index must be < 2^24)", h, w);
    return CF_OK;
}

}  // namespace

extern "C" {

const char* cf_last_error(void) { return err_slot().c_str(); }
int cf_abi_version(void) { return CF_ABI_VERSION; }
size_t cf_weights_blob_bytes(void) { return blob_bytes(); }

int cf_create(const void* weights, size_t weights_bytes, int device, int max_batch, int max_h, int max_w,
              int pw_engine, cf_engine** out) {
    CF_CHECK(out != nullptr, CF_EINVAL, "cf_create: out is NULL");
    *out = nullptr;
    CF_CHECK(weights != nullptr, CF_EINVAL, "cf_create: weights is NULL");
    CF_CHECK(max_batch >= 1 && max_h >= 32 && max_w >= 32 && max_h % 32 == 0 && max_w % 32 == 0, CF_EINVAL,
             "cf_create: max_batch=%d max_h=%d max_w=%d (sizes must be positive multiples of 32)", max_batch, max_h, max_w);
    CF_CHECK(pw_engine == CF_PW_SIMT || pw_engine == CF_PW_TCGEN05 || pw_engine == CF_PW_TCGEN05_1P || pw_engine == CF_PW_TCGEN05_LAYERWISE ||
                 pw_engine == CF_PW_TCGEN05_MIXED,
             CF_EINVAL, "cf_create: unknown pw_engine %d", pw_engine);
    CF_CHECK(weights_bytes == blob_bytes(), CF_EWEIGHTS, "cf_create: blob is %zu bytes, expected %zu", weights_bytes, blob_bytes());
    Blob blob;
    std::string why;
    CF_CHECK(blob.parse(weights, weights_bytes, why), CF_EWEIGHTS, "cf_create: %s", why.c_str());

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return fail(CF_ENODEV, "cf_create: no CUDA device (%s)", cudaGetErrorString(ce));
    CF_CHECK(device >= 0 && device < ndev, CF_EINVAL, "cf_create: device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    CF_CUDA(cudaGetDeviceProperties(&prop, device));
    CF_CHECK(prop.major == 10, CF_ENODEV, "cf_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
             prop.major, prop.minor);
    CF_ON_DEVICE(device);

    cf_engine* e = new cf_engine();
    e->device = device;
    e->max_batch = max_batch;
    e->max_h = max_h;
    e->max_w = max_w;
    e->pw_engine = pw_engine;
    int rc = CF_OK;
    auto bail = [&](int code) {
        cf_destroy(e);
        return code;
    };

    // weights: upload the payload once, resolve every expected entry
    {
        BlobHeader hd;
        memcpy(&hd, weights, sizeof hd);
        CF_CUDA(cudaMalloc((void**)&e->d_w, hd.payload_floats * 4));
        CF_CUDA(cudaMemcpy(e->d_w, blob.payload, hd.payload_floats * 4, cudaMemcpyHostToDevice));
        for (auto& en : expected_entries()) {
            const float* hp = blob.get(en.name, en.count, why);
            if (!hp) return bail(fail(CF_EWEIGHTS, "cf_create: %s", why.c_str()));
            const uint64_t off = (uint64_t)(hp - blob.payload);
            if (off % kEntryAlignFloats != 0) return bail(fail(CF_EWEIGHTS, "cf_create: entry %s is not 128-byte aligned", en.name.c_str()));
            e->w[en.name] = e->d_w + off;
            if (en.name == "stem.w") memcpy(e->stem_w.w, hp, sizeof(e->stem_w.w));
            if (en.name == "heads.w") memcpy(e->heads_w.w, hp, sizeof(e->heads_w.w));
            if (en.name == "heads.b") memcpy(e->heads_w.b, hp, sizeof(e->heads_w.b));
        }
    }

    for (int j = 0; j < 3; ++j) {  // up{j+1}.su is [c][a][b] in the blob; the epilogue reads one float4 of channels per sub-pixel
        const float* hp = blob.get("up" + std::to_string(j + 1) + ".su", 24 * 4, why);
        if (!hp) return bail(fail(CF_EWEIGHTS, "cf_create: %s", why.c_str()));
        float t[4 * 24];
        for (int c = 0; c < 24; ++c)
            for (int q = 0; q < 4; ++q) t[q * 24 + c] = hp[c * 4 + q];
        if (cudaMalloc((void**)&e->up_sut[j], sizeof t) != cudaSuccess ||
            cudaMemcpy(e->up_sut[j], t, sizeof t, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(CF_ECUDA, "cf_create: IDAUp scale upload failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    const size_t Bm = (size_t)max_batch;
    const size_t px2 = (size_t)(max_h / 2) * (max_w / 2);
    // largest hidden tensor: layer1.0 expand output, 96 channels at stride 2
    size_t hid_max = 0, dwo_max = 0;
    {
        size_t h = max_h / 2, w = max_w / 2;
        for (int i = 0; i < 12; ++i) {
            const MBBlock& b = kBlocks[i];
            hid_max = std::max(hid_max, h * w * (size_t)b.hid());
            h /= b.s;
            w /= b.s;
            dwo_max = std::max(dwo_max, h * w * (size_t)b.hid());
        }
    }
    if ((rc = dalloc(&e->stem, Bm * px2 * 32))) return bail(rc);
    if ((rc = dalloc(&e->hidA, Bm * hid_max))) return bail(rc);
    if ((rc = dalloc(&e->hidB, Bm * dwo_max))) return bail(rc);
    {
        size_t h = max_h / 2, w = max_w / 2;
        for (int i = 0; i < 12; ++i) {
            h /= kBlocks[i].s;
            w /= kBlocks[i].s;
            if ((rc = dalloc(&e->blk[i], Bm * h * w * kBlocks[i].cout))) return bail(rc);
        }
        if ((rc = dalloc(&e->clast, Bm * h * w * 24))) return bail(rc);
        for (int j = 0; j < 3; ++j) {
            h *= 2;
            w *= 2;
            if ((rc = dalloc(&e->up[j], Bm * h * w * 24))) return bail(rc);
        }
        const size_t px4 = h * w;
        if ((rc = dalloc(&e->hm, Bm * px4))) return bail(rc);
        if ((rc = dalloc(&e->wh, Bm * px4 * 2))) return bail(rc);
        if ((rc = dalloc(&e->lm, Bm * px4 * 10))) return bail(rc);
        if ((rc = dalloc(&e->reg, Bm * px4 * 2))) return bail(rc);
        if ((rc = dalloc(&e->hm_sig, Bm * px4))) return bail(rc);
        if ((rc = dalloc(&e->peak, Bm * px4))) return bail(rc);
    }
    e->in_u8 = nullptr;
    e->o_dets_floats = Bm * (size_t)THRESH_MAX_CAP * 6;  // covers [B,K<=1024,6] and [B,cap<=4096,5]
    e->o_lms_floats = Bm * (size_t)THRESH_MAX_CAP * 10;
    e->o_inds_n = Bm * 1024;
    if ((rc = dalloc(&e->o_dets, e->o_dets_floats))) return bail(rc);
    if ((rc = dalloc(&e->o_lms, e->o_lms_floats))) return bail(rc);
    if (cudaMalloc((void**)&e->o_inds, e->o_inds_n * 4) != cudaSuccess ||
        cudaMalloc((void**)&e->o_counts, Bm * 4) != cudaSuccess ||
        cudaMalloc((void**)&e->in_slot[0], Bm * (size_t)max_h * max_w * 3) != cudaSuccess ||
        cudaMalloc((void**)&e->in_slot[1], Bm * (size_t)max_h * max_w * 3) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_copied[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_copied[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_done[1], cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(CF_ECUDA, "cf_create: staging allocation failed: %s", cudaGetErrorString(cudaGetLastError())));

    if (cudaFuncSetAttribute(k_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, HEADS_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_thresh_nms, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)thresh_smem_bytes(THRESH_MAX_CAP)) != cudaSuccess)
        return bail(fail(CF_ECUDA, "cf_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (pw_engine != CF_PW_SIMT) {
        if ((rc = pw_tc_init(e->tc, device))) return bail(rc);
        // tf32 hi/lo, K-major, 128B-swizzled images of every point-wise weight matrix
        auto prep = [&](const std::string& name, int K, int N) {
            const float* hp = blob.get(name, (uint64_t)K * N, why);
            return hp ? tc_prepare_layer(e->tc, e->w[name], hp, K, N, layer_passes(pw_engine, K, N)) : fail(CF_EWEIGHTS, "cf_create: %s", why.c_str());
        };
        for (int i = 0; i < 12 && !rc; ++i) {
            const MBBlock& b = kBlocks[i];
            if (block_is_mbf(pw_engine, i)) {  // 32-column expand chunks, one 32-column projection image per K block
                for (const char* part : {".exp", ".proj"}) {
                    const std::string nm = "b" + std::to_string(i) + part;
                    const bool ex = part[1] == 'e';
                    const int K = ex ? b.cin : b.hid(), N = ex ? b.hid() : b.cout;
                    const float* hp = blob.get(nm, (uint64_t)K * N, why);
                    if (!rc) rc = hp ? tc_prepare_layer(e->tc, e->w[nm], hp, K, N, 3, 32) : fail(CF_EWEIGHTS, "cf_create: %s", why.c_str());
                }
                if (!rc) {
                    const std::string nm = "b" + std::to_string(i) + ".dw";
                    const float* hp = blob.get(nm, (uint64_t)b.k * b.k * b.hid(), why);
                    rc = hp ? mbf_prepare_dw(e->tc, e->w[nm], hp, b.k * b.k, b.hid()) : fail(CF_EWEIGHTS, "cf_create: %s", why.c_str());
                }
                continue;
            }
            if (b.t != 1) rc = prep("b" + std::to_string(i) + ".exp", b.cin, b.hid());
            if (!rc && !block_is_mbd(pw_engine, i)) {  // the tap chunk image: k_dwt reads its taps from shared memory
                const std::string nd = "b" + std::to_string(i) + ".dw";
                const float* hd = blob.get(nd, (uint64_t)b.k * b.k * b.hid(), why);
                rc = hd ? mbf_prepare_dw(e->tc, e->w[nd], hd, b.k * b.k, b.hid()) : fail(CF_EWEIGHTS, "cf_create: %s", why.c_str());
            }
            if (!rc && block_is_mbd(pw_engine, i)) {  // one 32-column projection image per K block + the tap chunk image
                const std::string np = "b" + std::to_string(i) + ".proj", nd = "b" + std::to_string(i) + ".dw";
                const float* hp = blob.get(np, (uint64_t)b.hid() * b.cout, why);
                rc = hp ? tc_prepare_layer(e->tc, e->w[np], hp, b.hid(), b.cout, 3, 32) : fail(CF_EWEIGHTS, "cf_create: %s", why.c_str());
                const float* hd = blob.get(nd, (uint64_t)b.k * b.k * b.hid(), why);
                if (!rc) rc = hd ? mbf_prepare_dw(e->tc, e->w[nd], hd, b.k * b.k, b.hid()) : fail(CF_EWEIGHTS, "cf_create: %s", why.c_str());
                continue;
            }
            if (!rc) rc = prep("b" + std::to_string(i) + ".proj", b.hid(), b.cout);
        }
        if (!rc) rc = prep("clast.w", 320, 24);
        if (!rc) {  // the stem as a K = 27 (one K block), N = 32 GEMM: k_stem_tc
            const float* hp = blob.get("stem.w", 27 * 32, why);
            rc = hp ? tc_prepare_layer(e->tc, e->w["stem.w"], hp, 27, 32, 3, STC_NC) : fail(CF_EWEIGHTS, "cf_create: %s", why.c_str());
        }
        const int skip_c[3] = {96, 32, 24};
        for (int j = 0; j < 3 && !rc; ++j) rc = prep("up" + std::to_string(j + 1) + ".w", skip_c[j], 24);
        if (!rc) rc = heads_tc_prepare(e->tc, e->heads_w.w);
        if (rc) return bail(rc);
    }
    *out = e;
    return CF_OK;
}

int cf_destroy(cf_engine* e) {
    if (!e) return CF_OK;
    DeviceGuard _dg(e->device);
    pw_tc_destroy(e->tc);
    float* bufs[] = {e->d_w, e->stem, e->hidA, e->hidB, e->clast, e->up[0], e->up[1], e->up[2], e->hm, e->wh,
                     e->lm,  e->reg,  e->hm_sig, e->peak, e->o_dets, e->o_lms};
    for (float* p : bufs)
        if (p) cudaFree(p);
    for (float* p : e->blk)
        if (p) cudaFree(p);
    for (float* p : e->up_sut)
        if (p) cudaFree(p);
    if (e->o_inds) cudaFree(e->o_inds);
    if (e->o_counts) cudaFree(e->o_counts);
    for (int i = 0; i < 2; ++i) {
        if (e->in_slot[i]) cudaFree(e->in_slot[i]);
        if (e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]);
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
    }
    if (e->gexec) cudaGraphExecDestroy(e->gexec);
    for (auto& c : e->plan_cache)
        if (c.gexec) cudaGraphExecDestroy(c.gexec);
    if (e->comm && nccl_api().ok) nccl_api().CommDestroy(e->comm);
    if (e->o_gather) cudaFree(e->o_gather);
    for (int i = 0; i < 2; ++i) {
        if (e->o_send[i]) cudaFree(e->o_send[i]);
        if (e->o_gather2[i]) cudaFree(e->o_gather2[i]);
        if (e->ev_topk[i]) cudaEventDestroy(e->ev_topk[i]);
    }
    if (e->xchg_stream) cudaStreamDestroy(e->xchg_stream);
    if (e->src_u8) cudaFree(e->src_u8);
    if (e->rs_tab) cudaFree(e->rs_tab);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return CF_OK;
}

int cf_forward(cf_engine* e, const void* input, int in_format, int batch, int h, int w, void* stream) {
    CF_CHECK(e != nullptr && input != nullptr, CF_EINVAL, "cf_forward: NULL engine or input");
    CF_CHECK(in_format == CF_IN_F32_NCHW || in_format == CF_IN_U8_HWC, CF_EINVAL, "cf_forward: unknown in_format %d", in_format);
    CF_CHECK(batch >= 1 && batch <= e->max_batch, CF_ECAP, "cf_forward: batch %d outside [1,%d]", batch, e->max_batch);
    // EfficientNet needs H,W multiples of 32 (five stride-2 stages; centerface.py:69 guarantees it)
    CF_CHECK(h >= 32 && w >= 32 && h % 32 == 0 && w % 32 == 0, CF_EINVAL, "cf_forward: h=%d w=%d must be positive multiples of 32", h, w);
    CF_CHECK((size_t)h * w <= (size_t)e->max_h * e->max_w, CF_ECAP, "cf_forward: %dx%d exceeds the %dx%d the engine was created for", h, w,
             e->max_h, e->max_w);
    CF_ON_DEVICE(e->device);
    if (e->plan.empty() || e->B == 0 || e->in != input || e->fmt != in_format || e->B != batch || e->H != h || e->W != w) {
        // stash the current plan, then reuse a cached one or build a new one
        if (!e->plan.empty() && e->B != 0) {
            if (e->plan_cache.size() >= 4) {
                if (e->plan_cache.front().gexec) cudaGraphExecDestroy(e->plan_cache.front().gexec);
                e->plan_cache.erase(e->plan_cache.begin());
            }
            e->plan_cache.push_back({e->in, e->fmt, e->B, e->H, e->W, std::move(e->plan), e->gexec, e->graph_state, e->plan_runs, e->plan_net_launches});
        } else if (e->gexec) {
            cudaGraphExecDestroy(e->gexec);
        }
        e->plan.clear();
        e->gexec = nullptr, e->graph_state = 0, e->plan_runs = 0, e->plan_net_launches = 0;
        bool hit = false;
        for (size_t i = 0; i < e->plan_cache.size(); ++i) {
            auto& c = e->plan_cache[i];
            if (c.in == input && c.fmt == in_format && c.B == batch && c.H == h && c.W == w) {
                e->plan = std::move(c.steps);
                e->in = c.in, e->fmt = c.fmt, e->B = c.B, e->H = c.H, e->W = c.W;
                e->gexec = c.gexec, e->graph_state = c.graph_state, e->plan_runs = c.plan_runs, e->plan_net_launches = c.plan_net_launches;
                e->plan_cache.erase(e->plan_cache.begin() + i);
                hit = true;
                break;
            }
        }
        if (!hit) {
            int rc = build_plan(e, input, in_format, batch, h, w);
            if (rc) {
                e->plan.clear();
                return rc;
            }
        }
    }
    static const bool use_graph = !(getenv("CF_GRAPH") && atoi(getenv("CF_GRAPH")) == 0) && !getenv("CF_SYNC_EACH");
    if (use_graph && e->graph_state == 0 && e->plan_runs >= 1) {
        // capture on the engine's own stream (the caller's may be the legacy default stream, which cannot capture)
        e->graph_state = -1;
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            const long long before = e->launches;
            const int rc = run_steps(e, CLS_ALL, e->stream);
            e->plan_net_launches = (int)(e->launches - before);
            e->launches = before;
            const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
            if (rc == CF_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&e->gexec, g, 0) == cudaSuccess) e->graph_state = 1;
            if (g) cudaGraphDestroy(g);
        }
        cudaGetLastError();  // a failed capture leaves the sticky-free error state clean: eager launches take over
    }
    ++e->plan_runs;
    if (use_graph && e->graph_state == 1) {
        CF_CUDA(cudaGraphLaunch(e->gexec, (cudaStream_t)stream));
        e->launches += e->plan_net_launches;
        return CF_OK;
    }
    return run_steps(e, CLS_ALL, (cudaStream_t)stream);
}

int cf_heads(cf_engine* e, float** hm, float** wh, float** lm, float** reg, float** hm_sig) {
    CF_CHECK(e != nullptr, CF_EINVAL, "cf_heads: NULL engine");
    CF_CHECK(e->B > 0, CF_EINVAL, "cf_heads: no cf_forward has run on this engine");
    if (hm) *hm = e->hm;
    if (wh) *wh = e->wh;
    if (lm) *lm = e->lm;
    if (reg) *reg = e->reg;
    if (hm_sig) *hm_sig = e->hm_sig;
    return CF_OK;
}

int cf_tap(cf_engine* e, const char* name, float** ptr, int* h, int* w, int* c) {
    CF_CHECK(e != nullptr && name != nullptr && ptr != nullptr, CF_EINVAL, "cf_tap: NULL argument");
    CF_CHECK(e->B > 0, CF_EINVAL, "cf_tap: no cf_forward has run on this engine");
    const std::string n(name);
    int hh = e->H / 2, ww = e->W / 2, cc = 32;
    float* p = nullptr;
    if (n == "stem") p = e->stem;
    for (int i = 0; i < 12 && !p; ++i) {
        hh /= kBlocks[i].s;
        ww /= kBlocks[i].s;
        cc = kBlocks[i].cout;
        if (n == "layer" + std::to_string(kLayerOfBlock[i]) && i == kLastBlockOfLayer[kLayerOfBlock[i]]) p = e->blk[i];
        if (n == "block" + std::to_string(i)) p = e->blk[i];
    }
    if (!p) {
        hh = e->H / 32;
        ww = e->W / 32;
        cc = 24;
        if (n == "conv_last") p = e->clast;
        else if (n == "up1") p = e->up[0], hh *= 2, ww *= 2;
        else if (n == "up2") p = e->up[1], hh *= 4, ww *= 4;
        else if (n == "fpn" || n == "up3") p = e->up[2], hh *= 8, ww *= 8;
    }
    CF_CHECK(p != nullptr, CF_EINVAL, "cf_tap: unknown tap '%s'", name);
    *ptr = p;
    if (h) *h = hh;
    if (w) *w = ww;
    if (c) *c = cc;
    return CF_OK;
}

int cf_ctdet_decode_classes(const float* heat, const float* wh, const float* reg, int batch, int classes, int h, int w, int K,
                            int cat_spec_wh, float* out_dets, int32_t* out_inds, float* scratch, void* stream) {
    CF_CHECK(heat && wh && out_dets && scratch, CF_EINVAL, "cf_ctdet_decode: NULL pointer");
    int rc = check_decode_args(batch, h, w);
    if (rc) return rc;
    CF_CHECK(classes >= 1 && classes <= 1024 && (long long)classes * h * w <= (1ll << 30), CF_EINVAL, "cf_ctdet_decode: %d classes", classes);
    CF_CHECK(K >= 1 && K <= 1024 && K <= h * w, CF_EINVAL, "cf_ctdet_decode: K=%d outside [1,min(1024,h*w)]", K);
    cudaStream_t s = (cudaStream_t)stream;
    // _nms works per (image, class) plane (max_pool2d, centerface_ext.py:44-50): fused into the top-k kernel when the image's
    // classes*h*w keys fit its shared-memory cache, else k_peak_mask over the B*C planes into `scratch` first
    CF_CUDA(launch_topk(heat, scratch, wh, reg, batch, h, w, K, out_dets, out_inds, s, classes, cat_spec_wh ? 2 * classes : 2));
    return CF_OK;
}

int cf_ctdet_decode(const float* heat, const float* wh, const float* reg, int batch, int h, int w, int K,
                    float* out_dets, int32_t* out_inds, float* scratch, void* stream) {
    return cf_ctdet_decode_classes(heat, wh, reg, batch, 1, h, w, K, 0, out_dets, out_inds, scratch, stream);
}

int cf_decode_topk(cf_engine* e, int K, float* out_dets, int32_t* out_inds, void* stream) {
    CF_CHECK(e != nullptr, CF_EINVAL, "cf_decode_topk: NULL engine");
    CF_CHECK(e->B > 0, CF_EINVAL, "cf_decode_topk: no cf_forward has run on this engine");
    CF_ON_DEVICE(e->device);
    int rc = cf_ctdet_decode(e->hm_sig, e->wh, e->reg, e->B, e->H / 4, e->W / 4, K, out_dets, out_inds, e->peak, stream);
    if (rc == CF_OK) e->launches += topk_fused(1, e->H / 4, e->W / 4) ? 1 : 2;
    return rc;
}

int cf_nms(const float* boxes, const float* scores, int n, float nms_threshold, int32_t* keep, int32_t* count, void* scratch,
           size_t scratch_bytes, void* stream) {
    CF_CHECK(boxes && scores && keep && count && scratch, CF_EINVAL, "cf_nms: NULL pointer");
    CF_CHECK(n >= 0 && n <= (1 << 20), CF_EINVAL, "cf_nms: n=%d outside [0, 2^20]", n);
    CF_CHECK(scratch_bytes >= nms_scratch_bytes(n), CF_ECAP, "cf_nms: scratch needs %zu bytes", nms_scratch_bytes(n));
    k_nms_keep<<<1, 1024, 0, (cudaStream_t)stream>>>(boxes, scores, n, nms_threshold, (unsigned char*)scratch, keep, count);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

size_t cf_nms_scratch_bytes(int n) { return n < 0 ? 0 : nms_scratch_bytes(n); }

int cf_nms_host(int device, const float* boxes, const float* scores, int n, float nms_threshold, int32_t* keep, int32_t* count) {
    CF_CHECK(count != nullptr && n >= 0 && n <= (1 << 20), CF_EINVAL, "cf_nms_host: bad arguments");
    *count = 0;
    if (n == 0) return CF_OK;
    CF_CHECK(boxes && scores && keep, CF_EINVAL, "cf_nms_host: NULL pointer");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return fail(CF_ENODEV, "cf_nms_host: no CUDA device (%s)", cudaGetErrorString(ce));
    CF_CHECK(device >= 0 && device < ndev, CF_EINVAL, "cf_nms_host: device %d out of range (have %d)", device, ndev);
    CF_ON_DEVICE(device);
    const size_t sb = nms_scratch_bytes(n), bb = (size_t)n * 16, kb = (size_t)n * 4;
    unsigned char* d = nullptr;  // boxes | scores | keep | count | scratch
    const size_t off_s = bb, off_k = off_s + kb, off_c = off_k + kb, off_x = (off_c + 64 + 15) & ~(size_t)15;
    CF_CUDA(cudaMalloc((void**)&d, off_x + sb));
    int rc = CF_OK;
    if (cudaMemcpy(d, boxes, bb, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(d + off_s, scores, kb, cudaMemcpyHostToDevice) != cudaSuccess)
        rc = fail(CF_ECUDA, "cf_nms_host: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (!rc) rc = cf_nms((const float*)d, (const float*)(d + off_s), n, nms_threshold, (int32_t*)(d + off_k), (int32_t*)(d + off_c), d + off_x, sb, nullptr);
    if (!rc && (cudaMemcpy(count, d + off_c, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(keep, d + off_k, kb, cudaMemcpyDeviceToHost) != cudaSuccess))
        rc = fail(CF_ECUDA, "cf_nms_host: kernel or download failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    return rc;
}

int cf_ctdet_post_process(const float* dets, const double* trans, int batch, int K, float* out, void* stream) {
    CF_CHECK(dets && trans && out && batch > 0 && K > 0, CF_EINVAL, "cf_ctdet_post_process: bad arguments");
    k_affine_boxes<<<cdiv(batch * K, 256), 256, 0, (cudaStream_t)stream>>>(dets, trans, out, batch, K);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

int cf_decode_threshold(const float* hm_sig, const float* wh, const float* reg, const float* lm, int batch, int h,
                        int w, int variant, float threshold, float nms_threshold, int size_h, int size_w,
                        float scale_w, float scale_h, int cap, float* out_dets, float* out_lms,
                        int32_t* out_counts, void* stream) {
    CF_CHECK(hm_sig && wh && out_dets && out_counts, CF_EINVAL, "cf_decode_threshold: NULL pointer");
    CF_CHECK(variant == CF_DECODE_A || variant == CF_DECODE_B, CF_EINVAL, "cf_decode_threshold: unknown variant %d", variant);
    CF_CHECK(variant != CF_DECODE_B || reg != nullptr, CF_EINVAL, "cf_decode_threshold: variant B needs reg");
    int rc = check_decode_args(batch, h, w);
    if (rc) return rc;
    CF_CHECK(cap >= 1 && cap <= THRESH_MAX_CAP, CF_EINVAL, "cf_decode_threshold: cap=%d outside [1,%d]", cap, THRESH_MAX_CAP);
    int capP = 1;
    while (capP < cap) capP <<= 1;
    static thread_local int attr_dev = -1;
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        CF_CUDA(cudaFuncSetAttribute(k_thresh_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)thresh_smem_bytes(THRESH_MAX_CAP)));
        attr_dev = dev;
    }
    k_thresh_nms<<<batch, 1024, thresh_smem_bytes(capP), (cudaStream_t)stream>>>(
        hm_sig, wh, reg, lm, h, w, variant, threshold, nms_threshold, size_h, size_w, scale_w, scale_h, cap, capP, out_dets,
        out_lms, out_counts);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

// Enqueue one batch on the engine's own streams: H2D of the u8 images into input slot (n mod 2) on the
// copy stream (overlapping the previous batch's kernels), then network + decode + D2H on the compute
// stream.  At most two submissions are in flight; a third waits for the oldest.
namespace {
int host_wait_oldest(cf_engine* e) {
    if (e->waited >= e->submitted) return fail(CF_EINVAL, "cf_wait_host: nothing in flight");
    CF_CUDA(cudaEventSynchronize(e->ev_done[e->waited & 1]));
    ++e->waited;
    return CF_OK;
}
int host_stage_input(cf_engine* e, const uint8_t* images, size_t bytes, int* slot_out) {  // the caller is on the engine's device
    if (e->submitted - e->waited >= 2) {
        int rc = host_wait_oldest(e);
        if (rc) return rc;
    }
    const int slot = (int)(e->submitted & 1);
    // the slot's previous contents were last read by the stem kernel of submission n-2
    if (e->slot_used[slot]) CF_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_done[slot], 0));
    CF_CUDA(cudaMemcpyAsync(e->in_slot[slot], images, bytes, cudaMemcpyHostToDevice, e->copy_stream));
    CF_CUDA(cudaEventRecord(e->ev_copied[slot], e->copy_stream));
    CF_CUDA(cudaStreamWaitEvent(e->stream, e->ev_copied[slot], 0));
    e->slot_used[slot] = true;
    *slot_out = slot;
    return CF_OK;
}
}  // namespace

int cf_submit_topk_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int K, float* out_dets,
                        int32_t* out_inds) {
    CF_CHECK(e && images && out_dets, CF_EINVAL, "cf_submit_topk_host: NULL argument");
    CF_CHECK(batch >= 1 && batch <= e->max_batch, CF_ECAP, "cf_submit_topk_host: batch %d outside [1,%d]", batch, e->max_batch);
    CF_CHECK(K >= 1 && K <= 1024, CF_EINVAL, "cf_submit_topk_host: K=%d outside [1,1024]", K);
    CF_CHECK(h >= 32 && w >= 32 && h % 32 == 0 && w % 32 == 0 && (size_t)h * w <= (size_t)e->max_h * e->max_w, CF_EINVAL,
             "cf_submit_topk_host: bad size %dx%d", h, w);
    CF_ON_DEVICE(e->device);
    int slot = 0;
    int rc = host_stage_input(e, images, (size_t)batch * h * w * 3, &slot);
    if (rc) return rc;
    cudaStream_t s = e->stream;
    if ((rc = cf_forward(e, e->in_slot[slot], CF_IN_U8_HWC, batch, h, w, s))) return rc;
    if ((rc = cf_decode_topk(e, K, e->o_dets, e->o_inds, s))) return rc;
    CF_CUDA(cudaMemcpyAsync(out_dets, e->o_dets, (size_t)batch * K * 6 * 4, cudaMemcpyDeviceToHost, s));
    if (out_inds) CF_CUDA(cudaMemcpyAsync(out_inds, e->o_inds, (size_t)batch * K * 4, cudaMemcpyDeviceToHost, s));
    CF_CUDA(cudaEventRecord(e->ev_done[slot], s));
    ++e->submitted;
    return CF_OK;
}

int cf_comm_unique_id(void* id128) {
    CF_CHECK(id128 != nullptr, CF_EINVAL, "cf_comm_unique_id: NULL argument");
    NcclApi& n = nccl_api();
    CF_CHECK(n.ok, CF_ECUDA, "cf_comm_unique_id: libnccl.so.2 is not available in this process");
    const int r = n.GetUniqueId(reinterpret_cast<NcclId*>(id128));
    CF_CHECK(r == 0, CF_ECUDA, "ncclGetUniqueId failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    return CF_OK;
}

int cf_comm_init(cf_engine* e, int nranks, int rank, const void* id128) {
    CF_CHECK(e && id128 && nranks >= 1 && rank >= 0 && rank < nranks, CF_EINVAL, "cf_comm_init: bad arguments (nranks=%d rank=%d)", nranks, rank);
    CF_CHECK(e->comm == nullptr, CF_EINVAL, "cf_comm_init: the engine already has a communicator");
    NcclApi& n = nccl_api();
    CF_CHECK(n.ok, CF_ECUDA, "cf_comm_init: libnccl.so.2 is not available in this process");
    CF_ON_DEVICE(e->device);
    NcclId id;
    memcpy(id.internal, id128, 128);
    const int r = n.CommInitRank(&e->comm, nranks, id, rank);
    if (r != 0) {
        e->comm = nullptr;
        return fail(CF_ECUDA, "ncclCommInitRank(%d of %d) failed: %s", rank, nranks, n.GetErrorString ? n.GetErrorString(r) : "?");
    }
    e->comm_ranks = nranks, e->comm_rank = rank;
    e->o_gather_floats = (size_t)nranks * e->max_batch * 1024 * 6;  // K <= 1024
    CF_CUDA(cudaMalloc((void**)&e->o_gather, e->o_gather_floats * 4));
    const char* ev = getenv("CF_XCHG_STREAM");
    if (!ev || atoi(ev) != 0) {
        CF_CUDA(cudaStreamCreateWithFlags(&e->xchg_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CF_CUDA(cudaEventCreateWithFlags(&e->ev_topk[i], cudaEventDisableTiming));
            CF_CUDA(cudaMalloc((void**)&e->o_send[i], (size_t)e->max_batch * 1024 * 6 * 4));
            CF_CUDA(cudaMalloc((void**)&e->o_gather2[i], e->o_gather_floats * 4));
        }
    }
    return CF_OK;
}

int cf_submit_topk_gather_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int K, float* out_dets_all,
                               int32_t* out_inds) {
    CF_CHECK(e && images && out_dets_all, CF_EINVAL, "cf_submit_topk_gather_host: NULL argument");
    CF_CHECK(e->comm != nullptr, CF_EINVAL, "cf_submit_topk_gather_host: no communicator (cf_comm_init)");
    CF_CHECK(batch >= 1 && batch <= e->max_batch, CF_ECAP, "cf_submit_topk_gather_host: batch %d outside [1,%d]", batch, e->max_batch);
    CF_CHECK(K >= 1 && K <= 1024, CF_EINVAL, "cf_submit_topk_gather_host: K=%d outside [1,1024]", K);
    CF_CHECK(h >= 32 && w >= 32 && h % 32 == 0 && w % 32 == 0 && (size_t)h * w <= (size_t)e->max_h * e->max_w, CF_EINVAL,
             "cf_submit_topk_gather_host: bad size %dx%d", h, w);
    CF_ON_DEVICE(e->device);
    int slot = 0;
    int rc = host_stage_input(e, images, (size_t)batch * h * w * 3, &slot);
    if (rc) return rc;
    cudaStream_t s = e->stream;
    if ((rc = cf_forward(e, e->in_slot[slot], CF_IN_U8_HWC, batch, h, w, s))) return rc;
    const size_t cnt = (size_t)batch * K * 6;
    if (e->xchg_stream) {
        // the exchange step behind the decode kernel, by event (no host synchronisation), on its own stream: the compute stream
        // goes straight on to the next submission's forward.  Send / receive buffers are per slot, so neither the next top-k nor
        // the next gather can touch what this one still reads; NCCL calls keep their order (one stream) on every rank.
        if ((rc = cf_decode_topk(e, K, e->o_send[slot], e->o_inds, s))) return rc;
        if (out_inds) CF_CUDA(cudaMemcpyAsync(out_inds, e->o_inds, (size_t)batch * K * 4, cudaMemcpyDeviceToHost, s));
        CF_CUDA(cudaEventRecord(e->ev_topk[slot], s));
        CF_CUDA(cudaStreamWaitEvent(e->xchg_stream, e->ev_topk[slot], 0));
        const int r = nccl_api().AllGather(e->o_send[slot], e->o_gather2[slot], cnt, /*ncclFloat32*/ 7, e->comm, e->xchg_stream);
        CF_CHECK(r == 0, CF_ECUDA, "ncclAllGather failed: %s", nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?");
        CF_CUDA(cudaMemcpyAsync(out_dets_all, e->o_gather2[slot], cnt * e->comm_ranks * 4, cudaMemcpyDeviceToHost, e->xchg_stream));
        CF_CUDA(cudaEventRecord(e->ev_done[slot], e->xchg_stream));  // implies the forward that read the input slot has finished
        ++e->submitted;
        return CF_OK;
    }
    if ((rc = cf_decode_topk(e, K, e->o_dets, e->o_inds, s))) return rc;
    // the exchange step: every rank contributes its [batch,K,6] list, enqueued on the compute stream right behind the decode
    // kernel (no host synchronisation in between), then ONE device-to-host copy of the gathered list
    const int r = nccl_api().AllGather(e->o_dets, e->o_gather, cnt, /*ncclFloat32*/ 7, e->comm, s);
    CF_CHECK(r == 0, CF_ECUDA, "ncclAllGather failed: %s", nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?");
    CF_CUDA(cudaMemcpyAsync(out_dets_all, e->o_gather, cnt * e->comm_ranks * 4, cudaMemcpyDeviceToHost, s));
    if (out_inds) CF_CUDA(cudaMemcpyAsync(out_inds, e->o_inds, (size_t)batch * K * 4, cudaMemcpyDeviceToHost, s));
    CF_CUDA(cudaEventRecord(e->ev_done[slot], s));
    ++e->submitted;
    return CF_OK;
}

int cf_wait_host(cf_engine* e) {
    CF_CHECK(e != nullptr, CF_EINVAL, "cf_wait_host: NULL engine");
    CF_ON_DEVICE(e->device);
    return host_wait_oldest(e);
}

int cf_detect_topk_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int K, float* out_dets,
                        int32_t* out_inds) {
    int rc = cf_submit_topk_host(e, images, batch, h, w, K, out_dets, out_inds);
    if (rc) return rc;
    CF_ON_DEVICE(e->device);
    while (e->waited < e->submitted)
        if ((rc = host_wait_oldest(e))) return rc;
    return CF_OK;
}

int cf_detect_threshold_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int variant,
                             float threshold, float nms_threshold, float scale_w, float scale_h, int cap,
                             float* out_dets, float* out_lms, int32_t* out_counts) {
    CF_CHECK(e && images && out_dets && out_counts, CF_EINVAL, "cf_detect_threshold_host: NULL argument");
    CF_CHECK(batch >= 1 && batch <= e->max_batch, CF_ECAP, "cf_detect_threshold_host: batch %d outside [1,%d]", batch, e->max_batch);
    CF_CHECK(h >= 32 && w >= 32 && h % 32 == 0 && w % 32 == 0 && (size_t)h * w <= (size_t)e->max_h * e->max_w, CF_EINVAL,
             "cf_detect_threshold_host: bad size %dx%d", h, w);
    CF_ON_DEVICE(e->device);
    int slot = 0;
    int rc = host_stage_input(e, images, (size_t)batch * h * w * 3, &slot);
    if (rc) return rc;
    cudaStream_t s = e->stream;
    if ((rc = cf_forward(e, e->in_slot[slot], CF_IN_U8_HWC, batch, h, w, s))) return rc;
    // `size` as the reference passes it: (H',W') in centerface.py:51, fixed (640,640) in eval_widerface.py:88
    const int size_h = variant == CF_DECODE_B ? 640 : h, size_w = variant == CF_DECODE_B ? 640 : w;
    rc = cf_decode_threshold(e->hm_sig, e->wh, e->reg, e->lm, batch, h / 4, w / 4, variant, threshold, nms_threshold, size_h,
                             size_w, scale_w, scale_h, cap, e->o_dets, out_lms ? e->o_lms : nullptr, e->o_counts, s);
    if (rc) return rc;
    ++e->launches;
    CF_CUDA(cudaMemcpyAsync(out_counts, e->o_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, s));
    CF_CUDA(cudaMemcpyAsync(out_dets, e->o_dets, (size_t)batch * cap * 5 * 4, cudaMemcpyDeviceToHost, s));
    if (out_lms) CF_CUDA(cudaMemcpyAsync(out_lms, e->o_lms, (size_t)batch * cap * 10 * 4, cudaMemcpyDeviceToHost, s));
    CF_CUDA(cudaEventRecord(e->ev_done[slot], s));
    ++e->submitted;
    while (e->waited < e->submitted)
        if ((rc = host_wait_oldest(e))) return rc;
    return CF_OK;
}

int cf_resize_u8(const uint8_t* src, int batch, int sh, int sw, uint8_t* dst, int dh, int dw, const int32_t* tab, int area2,
                 void* stream) {
    CF_CHECK(src && dst && (tab || area2) && batch > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0, CF_EINVAL, "cf_resize_u8: bad arguments");
    const long long n = (long long)batch * dh * dw;
    k_resize_u8<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, tab, batch, sh, sw, dh, dw, area2);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

int cf_resize_tables(int sh, int sw, int dh, int dw, int32_t* tab, size_t tab_ints, int* area2) {
    CF_CHECK(sh > 0 && sw > 0 && dh > 0 && dw > 0 && tab && area2, CF_EINVAL, "cf_resize_tables: bad arguments");
    CF_CHECK(tab_ints >= (size_t)3 * dw + (size_t)4 * dh, CF_ECAP, "cf_resize_tables: table needs %zu ints", (size_t)3 * dw + (size_t)4 * dh);
    ResizeTables t;
    t.build(sh, sw, dh, dw);
    memcpy(tab, t.tab.data(), t.tab.size() * sizeof(int32_t));
    *area2 = t.area2;
    return CF_OK;
}

int cf_warp_affine_tables(const double* M, int dh, int dw, int32_t* tab, size_t tab_ints) {
    CF_CHECK(M && tab && dh > 0 && dw > 0, CF_EINVAL, "cf_warp_affine_tables: bad arguments");
    CF_CHECK(tab_ints >= (size_t)2 * dw + (size_t)2 * dh, CF_ECAP, "cf_warp_affine_tables: table needs %zu ints", (size_t)2 * dw + (size_t)2 * dh);
    // cv::warpAffine without WARP_INVERSE_MAP: invert the forward matrix in fp64, then AB_BITS = 10 fixed point
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double m0 = M[4] * D, m4 = M[0] * D, m1 = M[1] * (-D), m3 = M[3] * (-D);
    const double b1 = -m0 * M[2] - m1 * M[5], b2 = -m3 * M[2] - m4 * M[5];
    auto rnd = [](double v) -> int32_t {  // saturate_cast<int>(double) = cvRound (round half to even) + saturation
        const double r = nearbyint(v);
        return r >= 2147483647.0 ? INT32_MAX : r <= -2147483648.0 ? INT32_MIN : (int32_t)r;
    };
    for (int x = 0; x < dw; ++x) {
        tab[x] = rnd(m0 * x * 1024.0);
        tab[dw + x] = rnd(m3 * x * 1024.0);
    }
    for (int y = 0; y < dh; ++y) {
        tab[2 * dw + y] = rnd((m1 * y + b1) * 1024.0) + 16;
        tab[2 * dw + dh + y] = rnd((m4 * y + b2) * 1024.0) + 16;
    }
    return CF_OK;
}

int cf_warp_affine_u8(const uint8_t* src, int batch, int sh, int sw, uint8_t* dst, int dh, int dw, const int32_t* tab, void* stream) {
    CF_CHECK(src && dst && tab && batch > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0, CF_EINVAL, "cf_warp_affine_u8: bad arguments");
    const long long n = (long long)batch * dh * dw;
    k_warp_affine_u8<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, tab, batch, sh, sw, dh, dw);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

int cf_detect_image_host(cf_engine* e, const uint8_t* image, int h, int w, int net_h, int net_w, int variant, float threshold,
                         float nms_threshold, float scale_w, float scale_h, int cap, float* out_dets, float* out_lms,
                         int32_t* out_count) {
    CF_CHECK(e && image && out_dets && out_count, CF_EINVAL, "cf_detect_image_host: NULL argument");
    CF_CHECK(h > 0 && w > 0, CF_EINVAL, "cf_detect_image_host: bad source size %dx%d", h, w);
    CF_CHECK(net_h >= 32 && net_w >= 32 && net_h % 32 == 0 && net_w % 32 == 0 && (size_t)net_h * net_w <= (size_t)e->max_h * e->max_w,
             CF_EINVAL, "cf_detect_image_host: bad network size %dx%d", net_h, net_w);
    CF_ON_DEVICE(e->device);
    if (e->submitted - e->waited >= 2) {
        int rc0 = host_wait_oldest(e);
        if (rc0) return rc0;
    }
    cudaStream_t s = e->stream;
    const size_t src_bytes = (size_t)h * w * 3;
    if (src_bytes > e->src_u8_bytes) {
        CF_CUDA(cudaStreamSynchronize(s));
        if (e->src_u8) cudaFree(e->src_u8);
        e->src_u8 = nullptr, e->src_u8_bytes = 0;
        CF_CUDA(cudaMalloc((void**)&e->src_u8, src_bytes));
        e->src_u8_bytes = src_bytes;
    }
    if (e->rs.sh != h || e->rs.sw != w || e->rs.dh != net_h || e->rs.dw != net_w) {
        e->rs.build(h, w, net_h, net_w);
        if (e->rs.tab.size() > e->rs_tab_ints) {
            CF_CUDA(cudaStreamSynchronize(s));
            if (e->rs_tab) cudaFree(e->rs_tab);
            e->rs_tab = nullptr, e->rs_tab_ints = 0;
            CF_CUDA(cudaMalloc((void**)&e->rs_tab, e->rs.tab.size() * sizeof(int32_t)));
            e->rs_tab_ints = e->rs.tab.size();
        }
        CF_CUDA(cudaMemcpyAsync(e->rs_tab, e->rs.tab.data(), e->rs.tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    }
    const int slot = (int)(e->submitted & 1);
    CF_CUDA(cudaMemcpyAsync(e->src_u8, image, src_bytes, cudaMemcpyHostToDevice, s));
    int rc = cf_resize_u8(e->src_u8, 1, h, w, e->in_slot[slot], net_h, net_w, e->rs_tab, e->rs.area2, s);  // centerface.py:30
    if (rc) return rc;
    ++e->launches;
    e->slot_used[slot] = true;
    if ((rc = cf_forward(e, e->in_slot[slot], CF_IN_U8_HWC, 1, net_h, net_w, s))) return rc;
    const int size_h = variant == CF_DECODE_B ? 640 : net_h, size_w = variant == CF_DECODE_B ? 640 : net_w;
    rc = cf_decode_threshold(e->hm_sig, e->wh, e->reg, e->lm, 1, net_h / 4, net_w / 4, variant, threshold, nms_threshold, size_h, size_w,
                             scale_w, scale_h, cap, e->o_dets, out_lms ? e->o_lms : nullptr, e->o_counts, s);
    if (rc) return rc;
    ++e->launches;
    CF_CUDA(cudaMemcpyAsync(out_count, e->o_counts, 4, cudaMemcpyDeviceToHost, s));
    CF_CUDA(cudaMemcpyAsync(out_dets, e->o_dets, (size_t)cap * 5 * 4, cudaMemcpyDeviceToHost, s));
    if (out_lms) CF_CUDA(cudaMemcpyAsync(out_lms, e->o_lms, (size_t)cap * 10 * 4, cudaMemcpyDeviceToHost, s));
    CF_CUDA(cudaEventRecord(e->ev_done[slot], s));
    ++e->submitted;
    while (e->waited < e->submitted)
        if ((rc = host_wait_oldest(e))) return rc;
    return CF_OK;
}

int cf_debug_pw_gemm_time(int pw_engine, int epi, const float* dA, const float* hW, float* dOut, int M, int K, int N,
                          const float* dRes, void* stream, int iters, float* ms, char* desc, int desc_cap) {
    CF_CHECK(dA && hW && dOut && M > 0 && K > 0 && N > 0 && K % 4 == 0 && N % 4 == 0, CF_EINVAL, "cf_debug_pw_gemm: bad arguments");
    CF_CHECK(epi == EPI_LINEAR || epi == EPI_SWISH || (epi == EPI_RESIDUAL && dRes), CF_EINVAL, "cf_debug_pw_gemm: epi %d unsupported here", epi);
    CF_CHECK(iters >= 0 && (iters == 0 || ms != nullptr), CF_EINVAL, "cf_debug_pw_gemm: iters=%d needs ms", iters);
    cudaStream_t s = (cudaStream_t)stream;
    EpiArgs ea{};
    ea.res = dRes;
    float* dW = nullptr;
    CF_CUDA(cudaMalloc((void**)&dW, (size_t)K * N * 4));
    cudaError_t ce = cudaMemcpy(dW, hW, (size_t)K * N * 4, cudaMemcpyHostToDevice);
    int rc = ce == cudaSuccess ? CF_OK : fail(CF_ECUDA, "cf_debug_pw_gemm: %s", cudaGetErrorString(ce));
    PwTcState st;
    std::function<cudaError_t()> launch;
    char d[256] = "";
    TcLaunch tl;
    PwnLaunch pl;
    if (!rc && pw_engine == CF_PW_SIMT) {
        launch = [&]() { return launch_pw_simt_any(epi, dA, dW, dOut, M, K, N, ea, s); };
        snprintf(d, sizeof d, "simt");
    } else if (!rc) {
        int dev = 0;
        cudaGetDevice(&dev);
        const int passes = engine_passes(pw_engine);
        rc = pw_tc_init(st, dev);
        if (!rc) rc = tc_prepare_layer(st, dW, hW, K, N, passes);
        if (!rc && passes == 3 && pwn_eligible(st.layers[dW]) && tc_tune_for(K, N, passes).pwn != 0) {
            rc = pwn_plan(st, epi, dA, dW, dOut, M, K, N, ea, &pl);
            launch = [&]() { return pwn_launch(pl, s); };
            snprintf(d, sizeof d, "pwn NC=%d nst=%d grid=%d smem=%zu", pl.p.NC, pl.p.nst, pl.grid, pl.smem);
        } else if (!rc) {
            rc = tc_plan(st, passes, epi, dA, dW, dOut, M, K, N, ea, &tl);
            launch = [&]() { return tc_launch(tl, s); };
            snprintf(d, sizeof d, "tc NC=%d chunks=%d stages=%d atmem=%d direct=%d resident=%d nacc=%d grid=%d smem=%zu", tl.p.NC,
                     tl.p.nchunks, tl.p.stages, tl.p.atmem, tl.p.direct, tl.p.resident, tl.p.nacc, tl.grid, tl.smem);
        }
    }
    if (desc && desc_cap > 0) snprintf(desc, (size_t)desc_cap, "%s", d);
    if (!rc) {
        ce = launch();
        if (ce != cudaSuccess) rc = fail(CF_ECUDA, "cf_debug_pw_gemm: launch: %s", cudaGetErrorString(ce));
        ce = cudaStreamSynchronize(s);
        if (!rc && ce != cudaSuccess) rc = fail(CF_ECUDA, "cf_debug_pw_gemm: kernel: %s", cudaGetErrorString(ce));
    }
    if (!rc && iters > 0) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, s);
        for (int i = 0; i < iters && ce == cudaSuccess; ++i) ce = launch();
        cudaEventRecord(b, s);
        cudaError_t ce2 = cudaEventSynchronize(b);
        if (ce != cudaSuccess || ce2 != cudaSuccess)
            rc = fail(CF_ECUDA, "cf_debug_pw_gemm: timed loop: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ce2));
        float t = 0.f;
        cudaEventElapsedTime(&t, a, b);
        *ms = t / (float)iters;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    pw_tc_destroy(st);
    cudaFree(dW);
    return rc;
}

int cf_debug_pw_gemm(int pw_engine, int epi, const float* dA, const float* hW, float* dOut, int M, int K, int N,
                     const float* dRes, void* stream) {
    return cf_debug_pw_gemm_time(pw_engine, epi, dA, hW, dOut, M, K, N, dRes, stream, 0, nullptr, nullptr, 0);
}

namespace cf {
__global__ void k_debug_swish(const float4* x, float4* y, long long n4, int variant) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        y[i] = variant ? swish4qv(x[i]) : swish4p(x[i]);
}
}  // namespace cf

int cf_debug_swish(const float* dX, float* dY, long long n, int variant) {
    CF_CHECK(dX && dY && n > 0 && n % 4 == 0 && (variant == 0 || variant == 1), CF_EINVAL, "cf_debug_swish: bad arguments");
    k_debug_swish<<<(unsigned)std::min<long long>((n / 4 + 255) / 256, 4096), 256>>>(reinterpret_cast<const float4*>(dX), reinterpret_cast<float4*>(dY), n / 4, variant);
    CF_CUDA(cudaGetLastError());
    CF_CUDA(cudaDeviceSynchronize());
    return CF_OK;
}

int cf_debug_tma_stream(const float* dA, int M, int K, int stages, int box_rows, int ctas_per_sm, float* ms) {
    CF_CHECK(dA && M > 0 && K >= 4 && K % 4 == 0 && stages >= 1 && box_rows >= 8 && box_rows <= 256 && ms, CF_EINVAL, "cf_debug_tma_stream: bad arguments");
    PwTcState st;
    int dev = 0;
    cudaGetDevice(&dev);
    int rc = pw_tc_init(st, dev);
    if (rc) return rc;
    CUtensorMap tm;
    if ((rc = tc_make_map(st, &tm, dA, (uint64_t)M, (uint64_t)K, (uint32_t)box_rows))) return rc;
    const uint32_t box_bytes = (uint32_t)box_rows * 128u;
    const size_t smem = (size_t)stages * box_bytes + 1024 + 1024;
    CF_CHECK(smem <= (size_t)TC_SMEM_MAX / (size_t)ctas_per_sm, CF_EINVAL, "cf_debug_tma_stream: %zu B of smem do not fit", smem);
    CF_CUDA(cudaFuncSetAttribute(k_tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_tiles = (M + box_rows - 1) / box_rows, nkb = (K + TC_BK - 1) / TC_BK;
    const int grid = std::min(n_tiles, st.sms * ctas_per_sm);
    cudaEvent_t a, b;
    CF_CUDA(cudaEventCreate(&a));
    CF_CUDA(cudaEventCreate(&b));
    k_tma_probe<<<grid, 32, smem>>>(tm, n_tiles, nkb, stages, box_rows, box_bytes);  // warm
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k_tma_probe<<<grid, 32, smem>>>(tm, n_tiles, nkb, stages, box_rows, box_bytes);
    cudaEventRecord(b);
    cudaError_t ce = cudaEventSynchronize(b);
    float t = 0.f;
    cudaEventElapsedTime(&t, a, b);
    *ms = t / 5.f;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (ce != cudaSuccess) return fail(CF_ECUDA, "cf_debug_tma_stream: %s", cudaGetErrorString(ce));
    return CF_OK;
}

long long cf_launch_count(cf_engine* e) { return e ? e->launches : 0; }

unsigned cf_fused_block_mask(int pw_engine) {
    unsigned m = 0;
    for (int i = 0; i < 12; ++i)
        if (block_is_mbf(pw_engine, i)) m |= 1u << i;
    return m;
}

unsigned cf_dwp_block_mask(int pw_engine) {
    unsigned m = 0;
    for (int i = 0; i < 12; ++i)
        if (block_is_mbd(pw_engine, i)) m |= 1u << i;
    return m;
}

int cf_debug_mbf_trace(cf_engine* e, unsigned long long* out, int n_jobs) {
    CF_CHECK(e && out && n_jobs > 0 && n_jobs <= 4096, CF_EINVAL, "cf_debug_mbf_trace: bad arguments");
    CF_CHECK(e->tc.trace_buf != nullptr, CF_EINVAL, "cf_debug_mbf_trace: no trace was recorded (set CF_MBF_TRACE=j0,nj before the plan is built)");
    CF_CUDA(cudaDeviceSynchronize());
    CF_CUDA(cudaMemcpy(out, e->tc.trace_buf, (size_t)n_jobs * 32 * 8, cudaMemcpyDeviceToHost));
    return CF_OK;
}

int cf_work_model(int h, int w, int in_format, int pw_engine, int which, double* bytes, double* flops) {
    CF_CHECK(h >= 32 && w >= 32 && h % 32 == 0 && w % 32 == 0, CF_EINVAL, "cf_work_model: h=%d w=%d must be multiples of 32", h, w);
    CF_CHECK(which >= CLS_ALL && which < CLS_COUNT, CF_EINVAL, "cf_work_model: unknown class %d", which);
    double by[CLS_COUNT] = {}, fl[CLS_COUNT] = {};
    const double F = 4.0;  // fp32 storage
    double hh = h / 2, ww = w / 2;
    by[CLS_STEM] = (double)h * w * 3 * (in_format == CF_IN_U8_HWC ? 1.0 : F) + hh * ww * 32 * F;
    fl[CLS_STEM] = 2.0 * hh * ww * 32 * 27;
    for (int i = 0; i < 12; ++i) {
        const MBBlock& b = kBlocks[i];
        const double hid = b.hid();
        const double ho = hh / b.s, wo = ww / b.s;
        if (block_is_mbf(pw_engine, i)) {  // block input in, block output out (+ the residual re-read)
            by[CLS_FUSED] += (hh * ww * b.cin + ho * wo * b.cout * (b.residual() ? 2 : 1)) * F;
            fl[CLS_FUSED] += 2.0 * hh * ww * b.cin * hid + 2.0 * ho * wo * hid * b.k * b.k + 2.0 * ho * wo * hid * b.cout;
            hh = ho;
            ww = wo;
            continue;
        }
        if (b.t != 1) {
            by[CLS_PW] += hh * ww * (b.cin + hid) * F;
            fl[CLS_PW] += 2.0 * hh * ww * b.cin * hid;
        }
        if (block_is_mbd(pw_engine, i)) {  // hidden tensor in, block output out (+ the residual re-read)
            by[CLS_FUSED] += (hh * ww * hid + ho * wo * b.cout * (b.residual() ? 2 : 1)) * F;
            fl[CLS_FUSED] += 2.0 * ho * wo * hid * b.k * b.k + 2.0 * ho * wo * hid * b.cout;
            hh = ho;
            ww = wo;
            continue;
        }
        by[CLS_DW] += (hh * ww + ho * wo) * hid * F;
        fl[CLS_DW] += 2.0 * ho * wo * hid * b.k * b.k;
        by[CLS_PW] += ho * wo * (hid + b.cout + (b.residual() ? b.cout : 0)) * F;
        fl[CLS_PW] += 2.0 * ho * wo * hid * b.cout;
        hh = ho;
        ww = wo;
    }
    by[CLS_PW] += hh * ww * (320 + 24) * F;
    fl[CLS_PW] += 2.0 * hh * ww * 320 * 24;
    const int skip_c[3] = {96, 32, 24};
    for (int j = 0; j < 3; ++j) {
        const double lo = hh * ww;
        hh *= 2;
        ww *= 2;
        by[CLS_PW] += (hh * ww * (skip_c[j] + 24) + lo * 24) * F;
        fl[CLS_PW] += 2.0 * hh * ww * (skip_c[j] * 24 + 24);
    }
    by[CLS_HEADS] = hh * ww * (24 + 16) * F;                       // FPN map in; 15 head planes + hm_sig out
    fl[CLS_HEADS] = 2.0 * hh * ww * (4 * 24 * 24 * 9 + 24 * 15);   // reference graph: four 3x3 24->24 + 1x1s
    by[CLS_DECODE] = hh * ww * 5 * F + 100 * 6 * F;                // hm, wh, reg in; [K,6] out (centerface_ext.py:52-82)
    fl[CLS_DECODE] = 0;
    for (int c = CLS_PW; c < CLS_COUNT; ++c) {
        if (c == CLS_DECODE) continue;
        by[CLS_ALL] += by[c];
        fl[CLS_ALL] += fl[c];
    }
    if (bytes) *bytes = by[which];
    if (flops) *flops = fl[which];
    return CF_OK;
}

int cf_replay_class(cf_engine* e, int which, int iters, void* stream) {
    CF_CHECK(e != nullptr, CF_EINVAL, "cf_replay_class: NULL engine");
    CF_CHECK(!e->plan.empty(), CF_EINVAL, "cf_replay_class: no cf_forward has run on this engine");
    CF_CHECK(which >= CLS_ALL && which < CLS_COUNT && iters >= 1, CF_EINVAL, "cf_replay_class: which=%d iters=%d", which, iters);
    CF_ON_DEVICE(e->device);
    for (int i = 0; i < iters; ++i) {
        int rc = run_steps(e, which, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return CF_OK;
}

int cf_time_class(cf_engine* e, int which, int iters, void* stream, float* ms, int* launches) {
    CF_CHECK(e != nullptr && ms != nullptr, CF_EINVAL, "cf_time_class: NULL argument");
    CF_CHECK(!e->plan.empty(), CF_EINVAL, "cf_time_class: no cf_forward has run on this engine");
    CF_ON_DEVICE(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t a, b;
    CF_CUDA(cudaEventCreate(&a));
    CF_CUDA(cudaEventCreate(&b));
    const long long before = e->launches;
    int rc = cf_replay_class(e, which, 1, stream);  // warm
    const int per = (int)(e->launches - before);
    if (rc == CF_OK) {
        cudaEventRecord(a, s);
        rc = cf_replay_class(e, which, iters, stream);
        cudaEventRecord(b, s);
        cudaError_t ce = cudaEventSynchronize(b);
        if (rc == CF_OK && ce != cudaSuccess) rc = fail(CF_ECUDA, "cf_time_class: %s", cudaGetErrorString(ce));
        float t = 0.f;
        cudaEventElapsedTime(&t, a, b);
        *ms = t / (float)iters;
        if (launches) *launches = per;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return rc;
}

int cf_time_steps(cf_engine* e, int iters, void* stream, float* ms, int* cls, int cap, int* n_steps) {
    CF_CHECK(e != nullptr && ms != nullptr && n_steps != nullptr && iters >= 1, CF_EINVAL, "cf_time_steps: bad arguments");
    CF_CHECK(!e->plan.empty(), CF_EINVAL, "cf_time_steps: no cf_forward has run on this engine");
    const int n = (int)e->plan.size();
    *n_steps = n;
    CF_CHECK(cap >= n, CF_ECAP, "cf_time_steps: the plan has %d steps, room for %d", n, cap);
    CF_ON_DEVICE(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<cudaEvent_t> ev((size_t)n + 1);
    for (auto& x : ev) CF_CUDA(cudaEventCreate(&x));
    for (int i = 0; i < n; ++i) ms[i] = 0.f, cls ? cls[i] = e->plan[i].cls : 0;
    int rc = CF_OK;
    for (int it = 0; it <= iters && rc == CF_OK; ++it) {  // pass 0 warms up
        cudaEventRecord(ev[0], s);
        for (int i = 0; i < n; ++i) {
            cudaError_t err = e->plan[i].run(s);
            if (err != cudaSuccess) rc = fail(CF_ECUDA, "cf_time_steps: launch %d failed: %s", i, cudaGetErrorString(err));
            ++e->launches;
            cudaEventRecord(ev[i + 1], s);
        }
        cudaError_t ce = cudaEventSynchronize(ev[n]);
        if (rc == CF_OK && ce != cudaSuccess) rc = fail(CF_ECUDA, "cf_time_steps: %s", cudaGetErrorString(ce));
        if (it == 0 || rc != CF_OK) continue;
        for (int i = 0; i < n; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
            ms[i] += t / (float)iters;
        }
    }
    for (auto& x : ev) cudaEventDestroy(x);
    return rc;
}

}  // extern "C"
