/*
 * aznet_b200.h -- C ABI of libaznet_b200.so: the B200 (sm_100a) implementation of AZ-Net's
 * adaptive-search hot path.  One entry point per reference function on the path
 * (SURVEY.md section 8a); the "replaces" lines cite /root/reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all
 *     buffers (PyTorch tensors in the Python host layer) and sizes outputs for the worst case;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, never allocates, keeps no global state;
 *   - return value 0 = ok, nonzero = AZN_ERR_*; azn_last_error() gives the thread-local text;
 *   - region / ROI counts that are produced on the device stay on the device (int32 in HBM):
 *     kernels are launched for the capacity and read the live count themselves, so the level
 *     loop of the search never round-trips to the host (the reference does, once per level:
 *     caffe-fast-rcnn/python/caffe/pycaffe.py:90,95).
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef AZNET_B200_H_
#define AZNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZN_OK 0
#define AZN_ERR_INVALID 1   /* bad argument (shape, alignment, enum) */
#define AZN_ERR_CUDA 2      /* a CUDA runtime/driver call failed */
#define AZN_ERR_CAPACITY 3  /* a caller-provided buffer is too small for the request */

#define AZN_LAYOUT_NCHW 0   /* Caffe blob order: [n][c][h][w]; pooled output [r][c][ph][pw] */
#define AZN_LAYOUT_NHWC 1   /* channels-last:    [n][h][w][c]; pooled output [r][ph][pw][c] */
#define AZN_DTYPE_F32 0
#define AZN_DTYPE_BF16 1

#define AZN_ACT_NONE 0
#define AZN_ACT_RELU 1
#define AZN_ACT_AZ_HEAD 2   /* act_aux = nsub: sigmoid on columns [0,nsub) and column 5*nsub, identity elsewhere */
#define AZN_ACT_SOFTMAX_BBOX 3 /* act_aux = classes: softmax over columns [0,classes), identity on the rest */

typedef void *azn_stream_t;

const char *azn_version(void);
const char *azn_last_error(void);
/* 0 when the current device is sm_100 (B200); AZN_ERR_CUDA otherwise. */
int azn_check_device(void);
/* Programmatic dependent launch of the level-loop kernels (default on): kernel K+1 is scheduled while kernel K
 * drains and waits (griddepcontrol.wait) before it touches global memory.  0 = plain stream order (benchmarks). */
void azn_set_pdl(int on);
/* Cooperative launch of the persistent GEMM kernels (default on): azn_fc_forward / azn_conv3x3_forward spin on device-side
 * flags and a grid-wide barrier, so all of their CTAs (one per SM) must be co-resident.  With the cooperative launch
 * attribute the driver places the whole grid together -- waiting for SMs held by other kernels instead of
 * deadlocking -- or fails the launch (AZN_ERR_CUDA).  0 = plain / PDL launch, which requires the caller to keep the
 * GPU free of concurrent kernels (benchmarks, A/B). */
void azn_set_coop(int on);
/* Measurement utility (no reference counterpart): fills `bytes` of device memory with 16-byte streaming stores from a
 * grid of one CTA per SM x 8 -- the write-only HBM rate that a store-dominated kernel such as the ROI max-pool can be
 * compared with (the copy rate of MEASURED_PEAKS.json counts read + write bytes). */
int azn_hbm_write_probe(void *dst, size_t bytes, azn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ROI max pooling, forward.
 * replaces: ROIPoolingLayer<Dtype>::Forward_cpu / Forward_gpu
 *           caffe-fast-rcnn/src/caffe/layers/roi_pooling_layer.cpp:46-125, .cu:18-92
 *   feat      [n_img, C, H, W] (NCHW) or [n_img, H, W, C] (NHWC), f32 or bf16
 *   rois      f32 [R, 5] = (batch_index, x1, y1, x2, y2)
 *   n_rois    int32 device scalar with the live ROI count (<= R_cap), or NULL for R_cap
 *   out       same dtype as feat; [R, C, PH, PW] for NCHW, [R, PH, PW, C] for NHWC
 *   argmax    int32, out's shape, or NULL (NCHW only; h*W + w of the winner, -1 for empty bins)
 * Bit-exact with the reference for f32; bf16 in -> bf16 out is bit-exact with bf16(f32 result)
 * because max commutes with the monotone rounding.  NHWC needs C*sizeof(dtype) % 16 == 0.
 * A batch index outside [0, n_img) yields an all-zero row (the reference CPU path aborts,
 * roi_pooling_layer.cpp:66-67; its GPU path reads out of bounds).
 * The NCHW f32 path (Caffe's own blob layout) transposes the map into `workspace`
 * (azn_roi_pool_workspace_bytes).  For the other layouts the workspace is optional: with n_img > 1 it
 * holds the per-image ROI buckets of the shared-memory-staged kernel (taken for 7x7 pooling with a static
 * ROI count of >= 64 ROIs per image: one CTA stages a channel slice of one image's map in shared memory
 * and pools all of that image's ROIs out of it); without it the direct (L2-fed) kernel runs. */
size_t azn_roi_pool_workspace_bytes(int n_img, int C, int H, int W, int layout, int dtype, int R_cap);
/* Benchmark hook: 0 automatic kernel choice, 1 direct kernels only, 2 staged kernel whenever it applies; + 10 x the
 * staged kernel's pooling loop (1 generic loop nest, 2 fixed-height column reduces = default, 3 / 4 ROIs grouped by width:
 * kept for A/B, measured slower); + 100: the grouped loop with its stores disabled (profiling only); + 200: 64-byte
 * slices even where 128-byte ones fit; + 300: row bands of 128-byte slices for maps too tall for shared memory (A/B,
 * measured slower); + 400 / + 600: two-level map (slice + row-pair table; 600: with 64-byte slices); + 900: bin bounds from
 * a geometry pre-pass from 4096 ROIs (+ 800: never) -- both A/B variants with kernel instantiations of their own, measured
 * no faster. */
void azn_roi_pool_tune(int mode);
int azn_roi_pool_fwd(const void *feat, int n_img, int C, int H, int W, int layout, int dtype,
                     const float *rois, const int32_t *n_rois, int R_cap, int PH, int PW,
                     float spatial_scale, void *out, int32_t *argmax, void *workspace,
                     size_t workspace_bytes, azn_stream_t stream);
/* The same with the kernel choice as an ARGUMENT (0 automatic, 1 direct kernels only, 2 staged kernel whenever it
 * applies, 3 direct kernel with one CTA per ROI: for many SMALL ROIs with a device-side count, the deep levels of the
 * search -- NHWC only, same bits) instead of the process-wide azn_roi_pool_tune setting: what callers that run several
 * host threads use. */
int azn_roi_pool_fwd_ex(const void *feat, int n_img, int C, int H, int W, int layout, int dtype,
                        const float *rois, const int32_t *n_rois, int R_cap, int PH, int PW,
                        float spatial_scale, void *out, int32_t *argmax, void *workspace,
                        size_t workspace_bytes, int kernel_choice, azn_stream_t stream);

/* f32 NCHW -> bf16 NHWC feature-map conversion (layout the search engine keeps resident). */
int azn_nchw_f32_to_nhwc_bf16(const float *src, int n_img, int C, int H, int W, void *dst,
                              azn_stream_t stream);
/* The same from a bf16 NCHW map: what arrives when the host narrowed the batch before the upload (below). */
int azn_nchw_bf16_to_nhwc_bf16(const void *src, int n_img, int C, int H, int W, void *dst, azn_stream_t stream);

/* HOST helper (no device work): f32 -> bf16, round to nearest even, NaN -> 0x7fff -- bit for bit what the device
 * conversion above stores -- of n elements on `threads` worker threads (<= 0: all hardware threads).  The conv5_3
 * blobs the reference hands its 'fc' net are f32 (lib/detect/test.py:228-236, pycaffe.py:90); narrowing a batch on
 * the host halves the bytes that cross PCIe, the link that bounds the batched host-facing call
 * (aznet_b200/pipeline.py).  dst may be pinned memory; 32-byte aligned outputs are written with streaming stores. */
int azn_host_f32_to_bf16(const float *src, uint16_t *dst, size_t n, int threads);
int azn_host_threads(void);

/* ------------------------------------------------------------------------------------------
 * Fully connected layer on the tensor cores: out = act(A . W^T + bias).
 * replaces: InnerProductLayer<Dtype>::Forward_{cpu,gpu} (+ ReLU / Sigmoid / Softmax layers)
 *           caffe-fast-rcnn/src/caffe/layers/inner_product_layer.cpp:80-93, .cu:13-25,
 *           relu_layer.cu:10, sigmoid_layer.cu:11-15, softmax_layer.cu:14-70
 *   A     bf16 [M_cap, K] row-major (lda = K), W bf16 [N, K] row-major (Caffe's blob order),
 *   bias  f32 [N]; out bf16 or f32 [M_cap, ldo]
 *   m_live  int32 device scalar with the live row count, or NULL for M_cap
 * bf16 operands, fp32 accumulation in TMEM (tcgen05.mma kind::f16), TMA-fed.
 * K % 64 == 0, A and W 16-byte aligned.  `workspace` (azn_fc_workspace_bytes) holds the stream-K
 * partial accumulators and their flags; it must be ZEROED ONCE before its first use (every launch
 * leaves the flags clean) and must not be shared by launches on different streams.  The kernel is
 * persistent with one CTA per SM and its CTAs wait on each other: it needs the whole GPU. */
size_t azn_fc_workspace_bytes(int M_cap, int N, int K);
/* Tuning hook for benchmarks: force the split factor (0 = automatic), the fix-up mode (-1 automatic, 0 in-kernel,
 * 1 grid-wide finish phase) and the tile (0 automatic; 64 / 128 / 256 = tile width, + 1000 x the number of 128-row
 * accumulator halves: 1256 = 128 x 256 tiles, 2256 = 256 x 256) of subsequent azn_fc_forward calls. */
void azn_fc_tune(int parts, int finish_mode, int block_n);
/* Profiling hook: when non-NULL, every CTA of subsequent azn_fc_forward launches writes 8 int64 words
 * (globaltimer ns: start, setup done, first operands landed, last MMA issued, first accumulator ready,
 * epilogue done, end; and its number of work units) at device_buffer[8 * cta].  NULL switches it off. */
void azn_fc_trace(long long *device_buffer);
int azn_fc_forward(const void *A, const void *W, const float *bias, void *out, int out_dtype,
                   int ldo, int M_cap, const int32_t *m_live, int N, int K, int act, int act_aux,
                   void *workspace, size_t workspace_bytes, azn_stream_t stream);

/* The three output layers of the AZ-Net head in one small kernel (csrc/heads.cu):
 *   out[M, N] f32 = [sigmoid | identity | sigmoid](h7[M, K] . wh[N, K]^T + bias),  N = 5*nsub + 1 <= 64, K % 64 == 0
 * replaces: InnerProduct adj_score / adj_bbox / zoom_score + Sigmoid adj_prob / zoom_prob
 *           models/Pascal/VGG16/az-net/test_fc.prototxt:146-232 (inner_product_layer.cpp:80-93, sigmoid_layer.cpp:11-24)
 *   h7   bf16 [M_cap, K] = [int7_1 | int7_2] activations, wh bf16 [N, K] the block-diagonal fusion of the three
 *   weight blobs (rows: nsub adj_score, 4*nsub adj_bbox, 1 zoom_score), bias f32 [N], out f32 [M_cap, ldo];
 *   columns [0, nsub) and 5*nsub go through the sigmoid.  m_live: int32 device scalar with the live row count or NULL.
 * Same arithmetic as azn_fc_forward(..., AZN_ACT_AZ_HEAD, nsub) (bf16 operands, fp32 accumulation; the summation order
 * differs), as an ordinary -- not persistent, not cooperative -- grid of warp-level mma.sync tiles: a few
 * microseconds instead of the persistent kernel's fixed 13-18, and co-resident with another stream's GEMM. */
/* Benchmark hook: 1 = plain stream-ordered launch of azn_az_heads_forward instead of programmatic dependent launch (the
 * better choice with a single batch in flight, see csrc/heads.cu). */
void azn_az_heads_tune(int plain_launch);
int azn_az_heads_forward(const void *h7, const void *wh, const float *bias, float *out, int ldo, int M_cap,
                         const int32_t *m_live, int N, int K, int nsub, azn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * One level of the adaptive search for a batch of images, after the heads have run.
 * replaces: _bbox_pred, _clip_boxes, un-dedup, _unwrap_adj_pred (lib/detect/test.py:106-151,
 *           :171-187, :243-251), the zoom selection of im_propose (:380-389), divide_region +
 *           _sift_dup (lib/utils/div.pyx:15-88) and the feature-space dedup of the NEXT level's
 *           ROIs (lib/detect/test.py:61-97, :212-218).
 * The per-image state lives in an azn_search_state that the caller fills with device pointers. */
typedef struct azn_search_state {
    int32_t n_img;
    int32_t cap_regions;     /* per image, per level                                        */
    int32_t cap_children;    /* per image: children generated before _sift_dup              */
    int32_t cap_props;       /* per image: accumulated adjacent predictions (Y)             */
    int32_t nsub;            /* 11 sub-region priors (lib/detect/config.py:149-155)          */
    int32_t chunk;           /* cfg.SEAR.BATCH_SIZE: dedup is per chunk of this many regions */
    const int32_t *im_h;     /* [n_img] original image height / width (clip + root region)  */
    const int32_t *im_w;
    const double *im_scale;  /* [n_img] image -> network-input scale (_get_image_blob)       */
    double tz;               /* cfg.SEAR.Tz                                                  */
    double min_side;         /* cfg.SEAR.MIN_SIDE                                            */
    double eps;              /* cfg.EPS                                                      */
    double dedup;            /* cfg.DEDUP_BOXES (<= 0 disables the feature-space dedup)      */
    /* current level */
    double *regions;         /* [n_img, cap_regions, 4] B of this level, reference order     */
    int32_t *n_regions;      /* [n_img]                                                      */
    int32_t *inv;            /* [n_img, cap_regions] region -> unique ROI slot (inv_index)   */
    int32_t *rep;            /* [n_img, cap_regions] unique slot -> representative region    */
    int32_t *n_uniq;         /* [n_img]                                                      */
    int32_t *img_off;        /* [n_img + 1] exclusive scan of n_uniq = row offset in heads   */
    float *rois;             /* [n_img * cap_regions, 5] packed unique ROIs (pool input)     */
    int32_t *m_total;        /* [1] total unique ROIs of the level                           */
    /* next level (double buffer; the caller swaps the pointers between levels) */
    double *next_regions;
    int32_t *next_n_regions;
    /* scratch */
    double *children;        /* [n_img, cap_children, 4]                                     */
    int64_t *hashes;         /* [n_img, cap_children]                                        */
    int32_t *flags;          /* [n_img, cap_children]                                        */
    /* accumulated adjacent predictions */
    double *props;           /* [n_img, cap_props, 4]                                        */
    float *prop_scores;      /* [n_img, cap_props]                                           */
    int32_t *n_props;        /* [n_img]                                                      */
    int32_t *n_eval;         /* [n_img] regions evaluated so far (num_eval)                  */
    int32_t *depth;          /* [n_img] last level that ran (the k printed by im_propose)    */
    int32_t *status;         /* [1] sticky device-side error flag (capacity overflow)        */
    /* optional anchor-region history (NULL = off): every evaluated region with its zoom score, level by
     * level in the reference's order -- the `Bhis` of lib/detect/tune.py:270,298 (diagnostic im_propose) */
    double *hist_regions;    /* [n_img, cap_history, 4]                                      */
    float *hist_zoom;        /* [n_img, cap_history]                                         */
    int32_t *n_history;      /* [n_img]                                                      */
    int32_t cap_history;
} azn_search_state;

/* Level-1 setup: B = [[0, 0, W-1, H-1]] per image (lib/detect/test.py:355), its ROI, counters. */
int azn_search_init(const azn_search_state *st, azn_stream_t stream);
/* heads: zoom_prob f32 (row stride ld_zoom floats), adj_prob f32 [.,nsub] (ld_prob),
 * adj_bbox f32 [.,4*nsub] (ld_bbox), rows indexed by img_off[i] + unique slot.
 * level: 1-based k of the reference loop; flags:
 *   AZN_LEVEL_LAST        the subdivision of this level would be discarded (k == K-1): skip it;
 *   AZN_LEVEL_ROOT_PROPS  level 1 after azn_search_root: append the root's adjacent predictions (head rows
 *                         [0, n_img), one per image) and nothing else -- the subdivision already happened. */
#define AZN_LEVEL_LAST 1
#define AZN_LEVEL_ROOT_PROPS 2
/*   AZN_LEVEL_TUNE        the diagnostic search of lib/detect/tune.py:256-316: level 1 compares the root's
 *                         own zoom score with Tz = 0 (:278) instead of forcing it to 1.0 (test.py:383-384) */
#define AZN_LEVEL_TUNE 4
int azn_search_level(const azn_search_state *st, const float *zoom_prob, int ld_zoom,
                     const float *adj_prob, int ld_prob, const float *adj_bbox, int ld_bbox,
                     int level, int flags, azn_stream_t stream);
/* Levels 1 and 2 in one pass of the heads.  The reference forces zoom[0] = 1.0 at k == 1
 * (lib/detect/test.py:383-384), so the regions of level 2 = _sift_dup(divide_region(root)) do not depend on
 * the net's output for the root.  azn_search_root = azn_search_init + that subdivision + the level-2 dedup:
 * afterwards `rois` holds the n_img root ROIs in rows [0, n_img) followed by the packed unique level-2 ROIs
 * (img_off[i] >= n_img, m_total = n_img + sum n_uniq), and next_regions / next_n_regions hold level 2.
 * The caller runs the heads once over m_total rows, then azn_search_level(level 1, AZN_LEVEL_ROOT_PROPS) on
 * the unswapped state and azn_search_level(level 2, ...) on the swapped one: same Y, same order, one weight
 * pass less. */
int azn_search_root(const azn_search_state *st, azn_stream_t stream);

/* Final selection.  replaces: lib/detect/test.py:393-401.
 *   mode 0: top-`num_proposals` by score (ties: lower index first), mode 1: score >= tc in order.
 *   out_boxes f64 [n_img, cap_out, 4], out_scores f32 [n_img, cap_out], out_count i32 [n_img]. */
int azn_select_proposals(const azn_search_state *st, int mode, int num_proposals, double tc,
                         double *out_boxes, float *out_scores, int32_t *out_count, int cap_out,
                         azn_stream_t stream);

/* Collection of the final proposal lists of a run of batches on the device.
 * replaces: the `all_boxes[i] = ...` append of test_proposals (lib/detect/test.py:508-513); proposals.pkl is
 * written once after the loop (:533-539), so a multi-GPU job gathers the collection ONCE at the end.
 *   copies out_boxes f64 [n_img, cap_out, 4] / out_scores f32 [n_img, cap_out] / out_count i32 [n_img] into
 *   slot (state[0] % n_slots) of dst_* ([n_slots, ...] each) and increments state[0]; `state` = 2 uint32 words,
 *   zero-initialised by the caller (word 1 is scratch).  Same launch every step: replays from a CUDA graph. */
int azn_collect_proposals(const double *out_boxes, const float *out_scores, const int32_t *out_count, int n_img,
                          int cap_out, double *dst_boxes, float *dst_scores, int32_t *dst_counts, int n_slots,
                          uint32_t *state, azn_stream_t stream);

/* Peer window: device memory of the collecting rank (rank 0 writes proposals.pkl, lib/detect/test.py:533-539)
 * mapped into the other ranks' processes on the same box, so that azn_collect_proposals of rank r appends its
 * lists with ordinary stores over NVLink (dst_* = window pointers) -- replaces the end-of-job NCCL gather of
 * round 1 (no collective kernel, no SMs taken from the cooperative GEMMs); the job ends with one barrier.
 *   azn_peer_alloc: cudaMalloc + zero-fill + cudaIpcGetMemHandle -> *ptr, 64 opaque handle bytes (HOST buffer)
 *                   that the host side passes to the other processes (aznet_b200/dist.py: object broadcast).
 *   azn_peer_open:  cudaIpcOpenMemHandle (peer access enabled lazily) in ANOTHER process -> *ptr (device pointer
 *                   valid in the calling process); azn_peer_close unmaps it, azn_peer_free releases the owner's. */
int azn_peer_alloc(size_t bytes, void **ptr, unsigned char *handle64);
int azn_peer_open(const unsigned char *handle64, void **ptr);
int azn_peer_close(void *ptr);
int azn_peer_free(void *ptr);

/* Stand-alone pieces of the level kernel, exposed with the reference's own signatures.
 * azn_divide_region replaces utils.cython_div.divide_region / _sift_dup for ONE region set
 * (lib/utils/div.pyx:15-88): regions f64 [n,4] -> out f64 [cap_out,4], out_count[1]. */
int azn_divide_region(const double *regions, int n, double min_side, double *out, int32_t *out_count,
                      int cap_out, int sift_only, void *scratch, size_t scratch_bytes, azn_stream_t stream);
size_t azn_divide_region_scratch_bytes(int n);
/* azn_decode_boxes replaces _bbox_pred + _clip_boxes (lib/detect/test.py:106-151):
 * boxes f64 [n,4], deltas f32 [n, 4*ncol] -> out f64 [n, 4*ncol] clipped to (im_h, im_w);
 * im_h <= 0 or im_w <= 0 skips the clip (_bbox_pred alone). */
int azn_decode_boxes(const double *boxes, const float *deltas, int n, int ncol, double eps,
                     int im_h, int im_w, double *out, azn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Greedy NMS.  replaces: utils.cython_nms.nms (lib/utils/nms.pyx:17-68), called by apply_nms
 * (lib/detect/test.py:467-484).
 *   dets f32 [n,5] (x1,y1,x2,y2,score); thresh is a double, compared as (double)ovr >= thresh;
 *   keep int64 [n] kept indices in descending score order (ties: higher index first, i.e. the
 *   order of a stable ascending argsort reversed); keep_count int32 [1].
 * f32 arithmetic with one rounding per operation (no FMA contraction) => bit-exact keep lists. */
size_t azn_nms_workspace_bytes(int64_t n);
int azn_nms(const float *dets, int64_t n, double thresh, int64_t *keep, int32_t *keep_count,
            void *workspace, size_t workspace_bytes, azn_stream_t stream);
/* Many small independent problems in one launch (one per class per image in apply_nms).
 *   seg_off int32 [n_seg + 1] row offsets into dets; every segment <= AZN_NMS_SEG_MAX rows;
 *   keep int64 [total rows] segment-local indices written at seg_off[s]; keep_count int32 [n_seg]. */
#define AZN_NMS_SEG_MAX 1024
int azn_nms_batched(const float *dets, const int32_t *seg_off, int n_seg, double thresh,
                    int64_t *keep, int32_t *keep_count, azn_stream_t stream);

/* The same kernel over padded storage: segment s = rows [seg_off[s], seg_off[s] + seg_len[s]) of dets, both int32
 * [n_seg] on the device, every length <= max_len <= AZN_NMS_SEG_MAX (a longer segment reports keep_count -1).
 * This is how the batched detection step runs apply_nms over its [image, class, 100, 5] detections. */
/* Diagnostic / tuning hook: 0 = default (mask and greedy chain pipelined on two streams / SM partitions), 1 = one
 * stream, mask then chain, 2 = stop after the mask, 3 = stop after the sort.  Used by tools/microbench.py --nms-phases.
 * A/B flags added to the mode, each selecting the round-1 kernel of one stage: + 8 all-pairs rank sort for every n (default:
 * bucket sort above 2048 boxes), + 16 tile-by-tile greedy pass (default: block-wise fixed-point rounds), + 32 float32 mask
 * kernel for every threshold (default: packed-half screen + exact evaluation for thresh > 0), + 64 one persistent launch for
 * the whole greedy chain (measured slower than one launch per super-tile).  All variants return identical keep lists.
 * A time-out of the persistent variant's spin-waits reports keep_count = -1 instead of hanging. */
void azn_nms_tune(int mode);
int azn_nms_segments(const float *dets, const int32_t *seg_off, const int32_t *seg_len, int n_seg, int max_len,
                     double thresh, int64_t *keep, int32_t *keep_count, azn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The Fast R-CNN detection step, batched over images and classes (BASELINE config #3).
 * replaces: the host code of _frcnn_forward around the net call (lib/detect/test.py:259-318) and the
 *           per-class selection of test_net (:549-553, :608-651); the net itself is azn_roi_pool_fwd +
 *           3 x azn_fc_forward (fc6, fc7, cls_score|bbox_pred with AZN_ACT_SOFTMAX_BBOX), the NMS is
 *           azn_nms_segments.  All counts stay on the device. */
typedef struct azn_detect_state {
    int32_t n_img;
    int32_t cap_boxes;       /* proposals per image (capacity of boxes / inv / rep / hashes / flags) */
    int32_t num_classes;     /* C, background class 0 included                                  */
    int32_t max_per_image;   /* 100 (lib/detect/test.py:553)                                    */
    int32_t chunk;           /* cfg.SEAR.BATCH_SIZE: the dedup runs per chunk of this many boxes */
    int32_t ld_head;         /* floats per row of head_out                                      */
    const int32_t *im_h;     /* [n_img] original image sizes (clip)                             */
    const int32_t *im_w;
    const double *im_scale;  /* [n_img]                                                         */
    double eps;              /* cfg.EPS                                                         */
    double dedup;            /* cfg.DEDUP_BOXES (<= 0: no dedup)                                */
    const double *boxes;     /* [n_img, cap_boxes, 4] proposals in image coordinates            */
    const int32_t *n_boxes;  /* [n_img]                                                         */
    int32_t *inv;            /* [n_img, cap_boxes] proposal -> unique ROI slot                  */
    int32_t *rep;            /* [n_img, cap_boxes] unique slot -> representative proposal       */
    int32_t *n_uniq;         /* [n_img]                                                         */
    int32_t *img_off;        /* [n_img + 1] row offset of the image in the ROI blob / head_out  */
    float *rois;             /* [n_img * cap_boxes, 5] packed unique ROIs                       */
    int32_t *m_total;        /* [1]                                                             */
    int64_t *hashes;         /* scratch [n_img, cap_boxes]                                      */
    int32_t *flags;          /* scratch [n_img, cap_boxes]                                      */
    const float *head_out;   /* [M, ld_head]: cls_prob in columns [0, C), bbox_pred in [C, 5C)  */
    const float *thresh;     /* [C] running per-class thresholds (strict >), or NULL = -inf     */
    float *dets;             /* [n_img, C, max_per_image, 5] f32 (x1, y1, x2, y2, score), score-descending */
    float *top_scores;       /* [n_img, C, max_per_image] the same scores, padded with -inf     */
    int32_t *det_count;      /* [n_img, C]                                                      */
} azn_detect_state;

/* proposals -> [batch index, box * im_scale] float32 ROIs, deduplicated per image and per chunk in feature
 * space exactly like _frcnn_forward (np.round(rois * DEDUP_BOXES) hash, np.unique order), packed densely. */
int azn_detect_rois(const azn_detect_state *st, azn_stream_t stream);
/* After the head: for every (image, class j >= 1) the rows with cls_prob > thresh[j], the max_per_image
 * highest of them in descending score order (ties: lower row first), each decoded with _bbox_pred +
 * _clip_boxes from its representative's box and rounded to float32 like all_boxes[j][i] (:636). */
int azn_detect_select(const azn_detect_state *st, azn_stream_t stream);
/* thresh[j] of test_net after the whole image set: the max_per_set-th highest score among
 * top_scores[:, j, :det_count] if more than max_per_set were pushed, else -inf (thresh[0] = -inf).
 * top_scores / det_count may span more images than one azn_detect_select call (concatenate batches; on
 * several GPUs all-gather them first: every rank then computes identical thresholds). */
int azn_detect_thresholds(const float *top_scores, const int32_t *det_count, int n_images, int num_classes,
                          int max_per_image, long long max_per_set, float *thresh, azn_stream_t stream);
/* tune_thresh (lib/detect/tune.py:318-366): the zoom threshold that keeps max_per_set = num_images *
 * cfg.TRAIN.ANCHORS_PER_IMG anchor regions over the whole image set = the max_per_set-th highest zoom score
 * of all anchor histories (the reference's min-heap, order-independent), -inf if fewer were seen.
 * zoom [n_images, cap] f32 (azn_search_state.hist_zoom of one or several batches), counts [n_images]. */
int azn_tune_threshold(const float *zoom, const int32_t *counts, int n_images, int cap, long long max_per_set,
                       float *thresh, azn_stream_t stream);
/* Final `score > thresh[j]` filter (:646-651): shrinks det_count in place (rows are score-descending). */
int azn_detect_filter(const float *top_scores, int32_t *det_count, const float *thresh, int n_images,
                      int num_classes, int max_per_image, azn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The conv5_3 backbone (the first "next" row): VGG16 conv1_1 .. conv5_3 of
 * models/Pascal/VGG16/az-net/test.prototxt:16-384 and the image blob of lib/detect/test.py:27-59.
 * Every activation map is a ZERO-BORDERED channels-last grid [n_img, H+2, W+2, C] bf16: the border is the
 * convolution's padding and every producer below writes it, so buffers can be reused without memsets.
 *
 * azn_image_blob  replaces: _get_image_blob + im_list_to_blob (lib/detect/test.py:27-59, lib/utils/blob.py:13-29)
 *   images uint8 [n_img, H0, W0, 3] (BGR, cv2.imread order; one size per batch) -> (image - pixel_means) resized
 *   by im_scale with cv::resize's INTER_LINEAR arithmetic (float source) -> out bf16 [n_img, Hs+2, Ws+2, Cpad]
 *   (channels 0..2 = B, G, R, the rest zero) and, when blob_f32 != NULL, Caffe's own f32 NCHW 'data' blob
 *   [n_img, 3, Hs, Ws].  Hs/Ws = the reference's destination size, cvRound(H0 * im_scale), computed by the caller.
 * azn_conv3x3_forward  replaces: ConvolutionLayer::Forward (3x3, pad 1, stride 1) + in-place ReLU
 *   (caffe-fast-rcnn/src/caffe/layers/conv_layer.cpp, base_conv_layer.cpp forward_cpu_gemm/forward_cpu_bias)
 *   X bf16 [n_img, H+2, W+2, Cin] zero-bordered, Cin % 64 == 0; Wt bf16 [Cout, 9*Cin], column (ky*3+kx)*Cin + c
 *   = Caffe's weight[o][c][ky][kx]; bias f32 [Cout]; Y bf16 [n_img, H+2, W+2, Cout] (border written as zeros) or,
 *   with out_unpadded, [n_img, H, W, Cout] -- the map layout of azn_roi_pool_fwd.  Implicit GEMM on tcgen05:
 *   same kernel, workspace and constraints as azn_fc_forward (persistent, needs the whole GPU).
 * azn_maxpool2x2_forward  replaces: PoolingLayer MAX 2x2 / stride 2, ceil mode (pooling_layer.cpp:81-95,144-168)
 *   in [n_img, H+2, W+2, C] -> out [n_img, ceil(H/2)+2, ceil(W/2)+2, C], C % 8 == 0.
 * azn_nhwc_border: [n_img, H, W, C] <-> zero-bordered grid (to_padded = 1 / 0). */
int azn_image_blob(const uint8_t *images, int n_img, int H0, int W0, double im_scale, const double *pixel_means,
                   void *out_padded_nhwc, int Cpad, int Hs, int Ws, float *blob_f32, azn_stream_t stream);
int azn_conv3x3_forward(const void *X, const void *Wt, const float *bias, void *Y, int n_img, int H, int W,
                        int Cin, int Cout, int relu, int out_unpadded, void *workspace, size_t workspace_bytes,
                        azn_stream_t stream);
/* Tuning hook (A/B measurements, tests): reuse = 1 (default) loads ONE activation box per filter row and lets its three
 * taps read it through row-shifted shared-memory descriptors, 0 loads one box per tap (round 1); wide layers
 * (Cout >= 256) with Cin <= bn128_max_cin then run 256 x 128 tiles (default -1: all of them). */
void azn_conv_tune(int reuse, int bn128_max_cin);
/* conv1_1 (Cin = 3) without the 61 padding channels per tap: azn_patches3x3 gathers every pixel's 3x3 neighbourhood into
 * one K = Kp row -- entry (ky*3+kx)*Cin + c, the K order of the packed weights -- over the same zero-bordered grid
 * (in [n, H+2, W+2, Cs] bf16 with Cs >= Cin channels per pixel, e.g. azn_image_blob with Cpad = 8; out [n, H+2, W+2, Kp],
 * Kp % 64 == 0 for the GEMM), and azn_conv_patches_forward runs the convolution as ONE tap: Wt bf16 [Cout, Kp]. */
int azn_patches3x3(const void *in_padded, int n_img, int H, int W, int Cs, int Cin, void *out_padded, int Kp,
                   azn_stream_t stream);
/* The same layer in ONE kernel without the patch matrix (9 * Cin <= 32, Cs == 8, Cout == 64: conv1_1 of VGG16, test.prototxt:16-31):
 * warp-level mma.sync over 16-pixel tiles with the A fragments built straight from the network input; Wt bf16 [64, Kp] in the
 * K order above (Kp >= 32), out the zero-bordered [n, H+2, W+2, 64] bf16 grid with bias and (relu != 0) ReLU applied. */
int azn_conv3x3_direct_forward(const void *in_padded, int n_img, int H, int W, int Cs, int Cin, const void *wt, int Kp,
                               const float *bias, void *out_padded, int Cout, int relu, azn_stream_t stream);
int azn_conv_patches_forward(const void *Xp, const void *Wt, const float *bias, void *Y, int n_img, int H, int W,
                             int Kp, int Cout, int relu, int out_unpadded, void *workspace, size_t workspace_bytes,
                             azn_stream_t stream);
int azn_maxpool2x2_forward(const void *in_padded, int n_img, int H, int W, int C, void *out_padded,
                           azn_stream_t stream);
int azn_nhwc_border(const void *in, int n_img, int H, int W, int C, void *out, int to_padded, azn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The skip-layer detector head: models/COCO/VGG16_skip/frcnn/test_fc.prototxt:28-135.
 * azn_grn_concat_forward  replaces: GRNLayer::Forward (caffe-fast-rcnn/src/caffe/layers/grn_layer.cpp:27-56) of
 *   roi_norm3/4/5 + ConcatLayer (axis 1) + PowerLayer (scale 1000, power 1, shift 0).
 *   pooled[l] bf16 [rows, channels[l]]: the NHWC output of azn_roi_pool_fwd over map l seen as rows of pooled
 *   positions (rows = R * PH * PW); out bf16 [rows, ld_out], ld_out >= sum channels (columns past the sum are
 *   left untouched: zero them once when K must be padded for the GEMM):  out[r, off_l + c] = scale * (x / sqrt(sum_c x^2)),
 *   each source normalised on its own.  n_units (device, may be NULL) = live ROI count, rows_per_unit = PH * PW:
 *   only n_units * rows_per_unit rows are processed.  `pooled` and `channels` are HOST arrays of n_src <= 8 entries.
 *   The 1x1 convolution conv_pool5 (+ReLU) that follows is azn_fc_forward over the same rows (K = sum channels,
 *   N = 512): its output [rows, 512] is the [R, PH*PW*512] pooled-row matrix of fc6. */
/* azn_roi_pool_grn_fwd: ROI max-pool (bf16 NHWC map, the semantics of azn_roi_pool_fwd) with the GRN + concat
 * offset + Power scale fused into its epilogue: out bf16 [R * PH * PW, ld_out], columns [ch_off, ch_off + C) of row
 * (r, ph, pw) = grn_scale * pooled / sqrt(sum_c pooled^2).  Three calls (conv3_3 / conv4_3 / conv5_3 at 1/4, 1/8,
 * 1/16) fill the concat operand of conv_pool5 without the pooled intermediates ever reaching HBM.  C <= 1024. */
int azn_roi_pool_grn_fwd(const void *feat, int n_img, int C, int H, int W, const float *rois, const int32_t *n_rois,
                         int R_cap, int PH, int PW, float spatial_scale, float grn_scale, void *out, int ld_out,
                         int ch_off, azn_stream_t stream);
int azn_grn_concat_forward(const void *const *pooled, const int32_t *channels, int n_src, const int32_t *n_units,
                           long long rows_cap, int rows_per_unit, float scale, void *out, int ld_out, azn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AZNET_B200_H_ */
