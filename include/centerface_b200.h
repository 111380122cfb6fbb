/*
 * centerface_b200.h -- C ABI of the B200-native CenterFace inference engine.
 *
 * This is the drop-in boundary for ONE hot path of nvlong21/Lightweight-face-detection-CenterNet:
 *     image batch -> backbone + FPN + heads -> heat-map decode -> boxes.
 * The reference has no native code and therefore no FFI of its own; each entry point below
 * names the reference Python interface (file:line in the reference repository) it replaces.
 * The reference-side binding (a ctypes stub) is shown in INTEGRATION.md and shipped as
 * lightweight-face-detection-centernet_b200/_lib.py.
 *
 * Conventions
 *   - plain C types only; every pointer is a raw device or host address as documented;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - every call returns 0 on success or a negative CF_E* code; cf_last_error() returns a
 *     thread-local, human readable message for the most recent failure on this thread;
 *   - a handle owns all of its device buffers (sized at cf_create for max_batch/max_h/max_w);
 *     one handle per GPU / per host thread; calls on one handle are not re-entrant;
 *   - nothing here falls back to the CPU: without a usable sm_100 device cf_create fails.
 */
#ifndef CENTERFACE_B200_H_
#define CENTERFACE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CF_ABI_VERSION 1

/* error codes */
#define CF_OK 0
#define CF_EINVAL (-1)   /* bad argument / shape / state         */
#define CF_ECUDA (-2)    /* a CUDA runtime or driver call failed */
#define CF_EWEIGHTS (-3) /* malformed weight blob                */
#define CF_ENODEV (-4)   /* no sm_100 class device               */
#define CF_ECAP (-5)     /* a capacity given at cf_create was exceeded */

/* input formats for cf_forward */
#define CF_IN_F32_NCHW 0 /* normalised fp32 [B,3,H,W]: the tensor the reference feeds
                            EfficientNet.forward (model/centernet.py:263)               */
#define CF_IN_U8_HWC 1   /* raw BGR u8 [B,H,W,3] already at network size; the /255,
                            mean/std step of centerface.py:32-34 is fused into the stem */

/* GEMM engine for the point-wise (1x1) convolutions */
#define CF_PW_SIMT 0      /* fp32 FFMA tiles (validation engine)                          */
#define CF_PW_TCGEN05 1   /* tcgen05.mma kind::tf32, 3-pass split (fp32-class accuracy); the MBConv blocks of
                             cf_fused_block_mask() run as ONE kernel each (expand + Swish + depth-wise + Swish +
                             projection, neither hidden tensor reaches HBM).  The default.                   */
#define CF_PW_TCGEN05_1P 2 /* tcgen05.mma kind::tf32 single pass, layer-wise (throughput probe: fails the parity contract) */
#define CF_PW_TCGEN05_LAYERWISE 3 /* CF_PW_TCGEN05 with every block as three launches (expand, depth-wise, projection): the
                                     schedule the fused kernel is measured against                              */
#define CF_PW_TCGEN05_MIXED 6 /* CF_PW_TCGEN05_LAYERWISE with a single TF32 pass on the stride-16/32 stages (K or N >= 384) only */

/* decode variants for cf_decode_threshold */
#define CF_DECODE_A 0 /* centerface.py:73-109   : offsets unused, landmarks, clip to (H,W)   */
#define CF_DECODE_B 1 /* eval_widerface.py:92-110: swapped offsets, no landmarks             */

typedef struct cf_engine cf_engine; /* opaque */

/* ---- life cycle ----------------------------------------------------------------------
 * Replaces CenterFace.__init__ (centerface.py:16-27): efficientnet_b0() construction,
 * checkpoint load and .cuda().  `weights` is a HOST pointer to the packed fp32 blob written
 * by lightweight-face-detection-centernet_b200/weights.py::pack_weights (BatchNorms folded,
 * heads collapsed, kernel-ready layouts); it is copied to the device and may be freed after
 * the call.  `max_h`,`max_w` are network-input sizes (multiples of 32).                  */
int cf_create(const void* weights, size_t weights_bytes, int device, int max_batch, int max_h,
              int max_w, int pw_engine, cf_engine** out);
int cf_destroy(cf_engine* e);

/* thread-local message of the last failure on the calling thread ("" if none) */
const char* cf_last_error(void);
int cf_abi_version(void);
/* size in bytes the packed blob must have (so the packer and the engine cannot drift) */
size_t cf_weights_blob_bytes(void);

/* ---- network -------------------------------------------------------------------------
 * Replaces `self.net(img)[0]` (centerface.py:41) / `model(img_batch)[0]`
 * (eval_widerface.py:84) == EfficientNet.forward (model/centernet.py:263-280).
 * `input` is a DEVICE pointer in `in_format`.  Results stay in handle-owned device buffers,
 * fetched with cf_heads().  Asynchronous on `stream`.                                    */
int cf_forward(cf_engine* e, const void* input, int in_format, int batch, int h, int w, void* stream);

/* Device pointers to the head maps of the last cf_forward, fp32 planar (NCHW) like the
 * reference's output dict {'hm','wh','lm','reg'} (model/centernet.py:277-280):
 * hm [B,1,h/4,w/4] raw logits, wh [B,2,..], lm [B,10,..], reg [B,2,..]; hm_sig is
 * clamp(sigmoid(hm),1e-4,1-1e-4) (centerface.py:43, eval_widerface.py:85).  Any may be NULL. */
int cf_heads(cf_engine* e, float** hm, float** wh, float** lm, float** reg, float** hm_sig);

/* Debug/parity taps: device pointer + shape of an intermediate NHWC fp32 activation.
 * name in {"stem","layer0".."layer6","conv_last","fpn"}.                                  */
int cf_tap(cf_engine* e, const char* name, float** ptr, int* h, int* w, int* c);

/* ---- decode, path C ------------------------------------------------------------------
 * Replaces ctdet_decode (centerface_ext.py:52-82) with _nms/_topk/_gather_feat fused:
 * 3x3 peak keep -> per-image top-K (score desc, flat index asc) -> wh/reg gather -> boxes.
 * All pointers are DEVICE pointers: heat [B,1,h,w] post-sigmoid, wh/reg [B,2,h,w] (reg may
 * be NULL -> +0.5 centres), out_dets [B,K,6] = x1,y1,x2,y2,score,class(0) in output-map
 * units, out_inds [B,K] int32 flat indices (may be NULL).  `scratch` is a device buffer of
 * at least B*h*w floats (cf_decode_topk uses the engine's own).  K <= 1024 and K <= h*w. */
int cf_ctdet_decode(const float* heat, const float* wh, const float* reg, int batch, int h, int w,
                    int K, float* out_dets, int32_t* out_inds, float* scratch, void* stream);
/* The class-aware form of the same function (centerface_ext.py:11-27 with cat > 1, :72-77): heat [B,classes,h,w], top-K per
 * class then top-K of the classes*K candidates (= the image's global top-K; ties: lower class, then lower pixel index),
 * out_dets[..,5] = class, out_inds = pixel index inside the class plane.  cat_spec_wh != 0: wh is [B,2*classes,h,w] and a
 * detection reads its own class's (w,h) planes; else wh is [B,2,h,w].  `scratch`: B*classes*h*w floats.
 * classes = 1, cat_spec_wh = 0 is cf_ctdet_decode.                                                                         */
int cf_ctdet_decode_classes(const float* heat, const float* wh, const float* reg, int batch, int classes, int h, int w,
                            int K, int cat_spec_wh, float* out_dets, int32_t* out_inds, float* scratch, void* stream);
/* same, on the heads of the last cf_forward */
int cf_decode_topk(cf_engine* e, int K, float* out_dets, int32_t* out_inds, void* stream);

/* Replaces CenterFace.nms(boxes, scores, nms_thresh) (centerface.py:111-151; eval_widerface.py:112-152 is the same code): greedy
 * IoU suppression with "+1" areas in float32, visiting order = np.argsort(scores) reversed with ties in (index descending) order
 * (a stable sort; the reference's default argsort leaves the order of equal scores unspecified), suppress when ovr >= nms_threshold.
 * boxes [n,4] (x1,y1,x2,y2), scores [n]; keep[0..*count) = indices into boxes in keep order, exactly the list the reference returns.
 * cf_nms: DEVICE pointers, `scratch` of cf_nms_scratch_bytes(n) bytes, asynchronous on `stream`.
 * cf_nms_host: HOST pointers (what the reference's numpy caller holds), synchronous, on `device`.                              */
int cf_nms(const float* boxes, const float* scores, int n, float nms_threshold, int32_t* keep, int32_t* count, void* scratch,
           size_t scratch_bytes, void* stream);
size_t cf_nms_scratch_bytes(int n);
int cf_nms_host(int device, const float* boxes, const float* scores, int n, float nms_threshold, int32_t* keep, int32_t* count);

/* Replaces ctdet_post_process (utils/post_process.py:83-100) for the single face class: maps both corners of every
 * row of dets [B,K,6] (DEVICE, output-map units) through the per-image inverse affine trans [B,6] (DEVICE, fp64,
 * row-major 2x3 = get_affine_transform(c, s, 0, (w,h), inv=1), utils/image.py:27-61, built by the caller) and
 * writes out [B,K,5] = x1,y1,x2,y2,score (DEVICE fp32).                                                        */
int cf_ctdet_post_process(const float* dets, const double* trans, int batch, int K, float* out, void* stream);

/* ---- decode, paths A and B -----------------------------------------------------------
 * Replaces CenterFace.decode + nms (centerface.py:73-151) / eval_widerface.decode + nms
 * (eval_widerface.py:92-152) and, when scale_w/scale_h are non-zero, the float32 floor
 * division back to source pixels (centerface.py:55-58).
 * DEVICE pointers: hm_sig [B,1,h,w], wh/reg [B,2,h,w], lm [B,10,h,w] (variant A only);
 * out_dets [B,cap,5], out_lms [B,cap,10] (NULL for variant B), out_counts [B] int32 =
 * number of kept boxes; if more than `cap` pixels pass the threshold in some image its
 * count is set to -(number of candidates) and that image's rows are undefined.
 * size_h,size_w: the clipping size the reference passes as `size`.  cap <= 4096.          */
int cf_decode_threshold(const float* hm_sig, const float* wh, const float* reg, const float* lm,
                        int batch, int h, int w, int variant, float threshold, float nms_threshold,
                        int size_h, int size_w, float scale_w, float scale_h, int cap,
                        float* out_dets, float* out_lms, int32_t* out_counts, void* stream);

/* ---- end-to-end convenience (HOST buffers) -------------------------------------------
 * One call = H2D copy of a raw u8 BGR batch [B,h,w,3] (already at network size), network,
 * path-C decode, D2H copy of [B,K,6] boxes (+ optional [B,K] indices).  Synchronous.
 * This is the call the `e2e` figure of bench.py times.                                    */
int cf_detect_topk_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int K,
                        float* out_dets, int32_t* out_inds);

/* Pipelined form of cf_detect_topk_host for streams of batches: cf_submit_topk_host enqueues one batch
 * (H2D on a copy stream, overlapping the previous batch's kernels; network, decode and D2H on the compute
 * stream) and returns; at most two submissions are in flight (a third blocks on the oldest).
 * cf_wait_host blocks until the OLDEST in-flight submission has delivered its host outputs.  The host
 * buffers of a submission must stay valid (and should be pinned) until its wait returns.          */
int cf_submit_topk_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int K,
                        float* out_dets, int32_t* out_inds);
int cf_wait_host(cf_engine* e);

/* ---- multi-GPU: one process per GPU, one exchange step ---------------------------------------
 * Images are independent and the weights replicate, so the path shards with no data-path collective; the reference has no
 * counterpart (single device).  The only exchange is an all-gather of the final fixed-size box list:
 *   cf_comm_unique_id          rank 0 creates the 128-byte NCCL id; the caller broadcasts it (any side channel)
 *   cf_comm_init               every rank joins (ncclCommInitRank) -- NCCL is bound at run time (libnccl.so.2)
 *   cf_submit_topk_gather_host as cf_submit_topk_host, then ncclAllGather of the [batch,K,6] lists on the compute stream right
 *                              behind the decode kernel and ONE device-to-host copy of the gathered [nranks*batch,K,6] list
 *                              (rank r's boxes at rows r*batch ..); every rank must submit the same batch and K.
 *                              Completion: cf_wait_host.                                                                   */
int cf_comm_unique_id(void* id128);
int cf_comm_init(cf_engine* e, int nranks, int rank, const void* id128);
int cf_submit_topk_gather_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int K, float* out_dets_all,
                               int32_t* out_inds);

/* Same for the threshold paths: the body of CenterFace.__call__ after cv2.resize
 * (centerface.py:32-62) for variant A, or of get_detections (eval_widerface.py:83-89) for
 * variant B, on a HOST u8 BGR batch [B,h,w,3].  Host outputs as in cf_decode_threshold.    */
int cf_detect_threshold_host(cf_engine* e, const uint8_t* images, int batch, int h, int w, int variant,
                             float threshold, float nms_threshold, float scale_w, float scale_h, int cap,
                             float* out_dets, float* out_lms, int32_t* out_counts);

/* Validation hook: one point-wise convolution out[M,N] = epi(A[M,K].W[K,N]) outside any engine, with
 * the selected GEMM engine (CF_PW_*).  dA/dOut/dRes are DEVICE pointers, hW a HOST [K][N] matrix.
 * epi: 0 linear (model/centernet.py:118), 1 Swish (:110), 2 + residual dRes (:137).  Synchronous. */
int cf_debug_pw_gemm(int pw_engine, int epi, const float* dA, const float* hW, float* dOut, int M, int K, int N,
                     const float* dRes, void* stream);

/* Same call, then `iters` more launches bracketed by CUDA events on `stream`: *ms = mean kernel time; `desc` (may be
 * NULL) receives the resolved launch plan (kernel, column chunk, stages ...).  The CF_TC_* environment variables read at
 * plan time select plan variants.  (tools/tc_tune.py)                                                                 */
int cf_debug_pw_gemm_time(int pw_engine, int epi, const float* dA, const float* hW, float* dOut, int M, int K, int N,
                          const float* dRes, void* stream, int iters, float* ms, char* desc, int desc_cap);

/* Validation hook: y[i] = Swish(x[i]) (model/centernet.py:39-40) with the device formulation `variant` -- 0: ex2 + rcp per
 * value (every kernel's default), 1: one reciprocal per four values (swish4q: the fused layer1.0 kernel and the epilogue of
 * the one-K-block expand layers).  x, y are DEVICE arrays of n floats, n a multiple of 4.  Synchronous.                   */
int cf_debug_swish(const float* dX, float* dY, long long n, int variant);

/* Development probe: stream a device [M][K] fp32 matrix through a `stages`-deep TMA ring of box_rows x 32-float boxes
 * with one thread per CTA and no consumer work; *ms = mean kernel time.  (tools/tma_probe.py)            */
int cf_debug_tma_stream(const float* dA, int M, int K, int stages, int box_rows, int ctas_per_sm, float* ms);

/* ---- pre-processing ---------------------------------------------------------
 * cv2.resize(img, (W',H')) of centerface.py:30 (default INTER_LINEAR, 8UC3) on the device, bit-exact with OpenCV
 * 4.x (fixed-point bilinear; exact 2x decimation = INTER_AREA).  cf_resize_tables builds the per-column/per-row
 * index + weight table on the HOST (3*dw + 4*dh int32) with OpenCV's own arithmetic; cf_resize_u8 resizes a DEVICE
 * batch [B,sh,sw,3] -> [B,dh,dw,3] with a DEVICE copy of that table.                                    */
int cf_resize_tables(int sh, int sw, int dh, int dw, int32_t* tab, size_t tab_ints, int* area2);
int cf_resize_u8(const uint8_t* src, int batch, int sh, int sw, uint8_t* dst, int dh, int dw, const int32_t* tab, int area2,
                 void* stream);
/* cv2.warpAffine(img, M, (dw,dh), flags=cv2.INTER_LINEAR) on 8UC3 with the default constant-0 border: the letter-box step
 * of the reference's loader (dataset/dataset.py:130-134, trans_input from utils/image.py:27-61), bit-exact with OpenCV 4.x.
 * cf_warp_affine_tables builds, on the HOST and in fp64 like OpenCV, the integer tables of the inverse map from the FORWARD
 * 2x3 matrix M (row major, 6 doubles): adelta[dw] | bdelta[dw] | X0[dh] | Y0[dh] (2*dw + 2*dh int32);
 * cf_warp_affine_u8 warps a DEVICE batch [B,sh,sw,3] -> [B,dh,dw,3] (the same matrix for every image) with a DEVICE copy. */
int cf_warp_affine_tables(const double* M, int dh, int dw, int32_t* tab, size_t tab_ints);
int cf_warp_affine_u8(const uint8_t* src, int batch, int sh, int sw, uint8_t* dst, int dh, int dw, const int32_t* tab, void* stream);
/* The whole body of CenterFace.__call__ (centerface.py:29-62) for one HOST u8 BGR image [h,w,3] at its own size:
 * H2D, resize to (net_h,net_w), normalise, network, sigmoid/clamp, threshold decode (variant A/B), NMS, //scale,
 * D2H of out_dets [cap,5], out_lms [cap,10] (may be NULL) and out_count (negative = more than cap candidates).  */
int cf_detect_image_host(cf_engine* e, const uint8_t* image, int h, int w, int net_h, int net_w, int variant, float threshold,
                         float nms_threshold, float scale_w, float scale_h, int cap, float* out_dets, float* out_lms,
                         int32_t* out_count);

/* ---- instrumentation -----------------------------------------------------------------
 * Number of kernels this library launched on behalf of the handle since creation.        */
long long cf_launch_count(cf_engine* e);

/* Bit i set: MBConv block i (model/centernet.py:211-234, layer0 = block 0 .. layer6 = block 11) runs as one fused kernel under
 * engine `pw_engine` (the built-in mask, or CF_MBF from the environment). */
unsigned cf_fused_block_mask(int pw_engine);
/* Bit i set: depth-wise + Swish + projection (+ residual) of block i run as one kernel from the hidden tensor (the expand conv, if
 * the block has one, stays its own launch); CF_MBD overrides the built-in mask. */
unsigned cf_dwp_block_mask(int pw_engine);

/* Development only: the pipeline trace of the fused MBConv kernel (k_mbf) recorded during the last forward when the plan was built
 * with CF_MBF_TRACE=j0,nj in the environment: out[job][32] = clock64 of event 0..31 of CTA 0 (0 = not recorded). */
int cf_debug_mbf_trace(cf_engine* e, unsigned long long* out, int n_jobs);
/* Algorithmic bytes / flops of ONE image at (h,w) for kernel class `which` under engine `pw_engine`
 * (0 = network = 1+2+3+4+6, 1 = point-wise GEMMs, 2 = depth-wise, 3 = stem, 4 = heads,
 * 5 = path-C decode, 6 = fused MBConv blocks).  Bytes: every conv reads its un-padded input once and writes its output
 * once in the engine's fp32 storage (+ residual / low-res re-reads); weights (5 MB per LAUNCH,
 * not per image) are not counted.  Flops: 2*MAC of the REFERENCE graph (model/centernet.py),
 * i.e. the four un-collapsed heads.  in_format selects the stem's input bytes.            */
int cf_work_model(int h, int w, int in_format, int pw_engine, int which, double* bytes, double* flops);
/* Run only one layer class `iters` times on the current activations (for per-kernel
 * CUDA-event timing in bench.py). which as above.                                        */
int cf_replay_class(cf_engine* e, int which, int iters, void* stream);
/* cf_replay_class bracketed by CUDA events on `stream`; *ms = mean device time of ONE replay of
 * the class, *launches = kernels per replay.  Synchronises the stream.                    */
int cf_time_class(cf_engine* e, int which, int iters, void* stream, float* ms, int* launches);
/* Every launch of the current plan (network + path-C decode, in order) timed separately with CUDA events on `stream`,
 * mean over `iters` passes after one warm-up pass: ms[i], cls[i] (class as above, may be NULL) for i < *n_steps <= cap. */
int cf_time_steps(cf_engine* e, int iters, void* stream, float* ms, int* cls, int cap, int* n_steps);

#ifdef __cplusplus
}
#endif
#endif /* CENTERFACE_B200_H_ */
