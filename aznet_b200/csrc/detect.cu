// The Fast R-CNN detection step of the AZ-Net pipeline for sm_100a, batched over images and classes and
// device-resident: what lib/detect/test.py does on the host around `_frcnn_forward` and inside `test_net`.
//
//   azn_detect_rois        proposals -> ROI blob + feature-space dedup           test.py:259-285 (_frcnn_forward),
//                                                                                :61-97 (_get_rois_blob)
//   azn_detect_select      un-dedup (:309-312), per-class `score > thresh[j]`,   test.py:608-623, :636
//                          top-100 by score, _bbox_pred + _clip_boxes of the
//                          selected rows only (:106-151), float32 [box, score]
//   azn_detect_thresholds  the min-heap cap of test_net (:624-631): thresh[j] =  test.py:549-553, :624-631
//                          the max_per_set-th highest of all pushed scores
//   azn_detect_filter      final strict `> thresh[j]` filter                     test.py:646-651
//
// The reference pushes every image's top-100 scores of a class through a heap of max_per_set entries and
// raises thresh[j] to the heap minimum as it goes; early pruning never removes a value above the final
// threshold, so the final threshold is order-independent: the max_per_set-th highest pushed score if more
// than max_per_set were pushed, -inf otherwise.  That makes the step batchable (and shardable: all-gather the
// [images, C, 100] score tensor, every rank computes the same thresholds).
// Selection order: (score desc, original row asc) = np.argsort(-scores, kind='stable'); the reference's
// default sort is unstable, so parity is asserted on tie-free scores.
#include <math.h>
#include "common.cuh"
#include "box_common.cuh"

namespace {

constexpr int ROIS_THREADS = 512;
constexpr int SEL_THREADS_D = 128;
constexpr int TH_THREADS = 1024;

// ---- proposals -> ROI blob, per-chunk feature-space dedup (one CTA per image) --------------------------
__global__ void __launch_bounds__(ROIS_THREADS) detect_rois_kernel(azn_detect_state st) {
    pdl_enter();
    const int i = blockIdx.x, tid = threadIdx.x;
    const int cap = st.cap_boxes;
    int nN = st.n_boxes[i];
    nN = nN < 0 ? 0 : (nN > cap ? cap : nN);
    const double *boxes = st.boxes + (size_t)i * cap * 4;
    long long *hashes = (long long *)st.hashes + (size_t)i * cap;
    int *flags = st.flags + (size_t)i * cap;
    int *inv = st.inv + (size_t)i * cap, *rep = st.rep + (size_t)i * cap;
    if (st.dedup > 0.0) {
        const double scale = st.im_scale[i];
        const float fd = (float)st.dedup;
        const int chunk = st.chunk > 0 ? st.chunk : 0x7fffffff;
        for (int q = tid; q < nN; q += ROIS_THREADS) {
            float p[4];
            project_roi(boxes + (size_t)q * 4, scale, p);
            hashes[q] = feat_hash(p, fd) + ((long long)(q / chunk) << 50);   // np.unique runs per BATCH_SIZE chunk
        }
        __syncthreads();
        mark_first(hashes, flags, nN);
        __syncthreads();
        int nU = 0;
        for (int q0 = 0; q0 < nN; q0 += ROIS_THREADS) {
            const int q = q0 + tid;
            int first = 0;
            if (q < nN) {
                first = flags[q];
                const int slot = unique_slot(hashes, flags, nN, hashes[q]);
                inv[q] = slot;
                if (first) rep[slot] = q;
            }
            nU += __syncthreads_count(first);
        }
        if (tid == 0) st.n_uniq[i] = nU;
    } else {
        for (int q = tid; q < nN; q += ROIS_THREADS) inv[q] = q, rep[q] = q;
        if (tid == 0) st.n_uniq[i] = nN;
    }
}

__global__ void __launch_bounds__(256) detect_pack_kernel(azn_detect_state st) {
    pdl_enter();
    const int i = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_off;
    if (tid < 32) {
        int acc = 0;
        for (int j = tid; j < i; j += 32) acc += st.n_uniq[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (tid == 0) {
            s_off = acc;
            st.img_off[i] = acc;
            if (i == st.n_img - 1) {
                st.img_off[st.n_img] = acc + st.n_uniq[i];
                *st.m_total = acc + st.n_uniq[i];
            }
        }
    }
    __syncthreads();
    const int off = s_off, nU = st.n_uniq[i];
    const double *boxes = st.boxes + (size_t)i * st.cap_boxes * 4;
    const int *rep = st.rep + (size_t)i * st.cap_boxes;
    const double scale = st.im_scale[i];
    for (int u = tid; u < nU; u += blockDim.x) {
        float p[4];
        project_roi(boxes + (size_t)rep[u] * 4, scale, p);
        float *o = st.rois + (size_t)(off + u) * 5;
        o[0] = (float)i; o[1] = p[0]; o[2] = p[1]; o[3] = p[2]; o[4] = p[3];
    }
}

// ---- block-wide radix select: key of the `need`-th largest among the valid keys ---------------------
// key_at(c, key) -> valid.  Returns through shared memory: the k-th largest key and how many of the
// elements equal to it belong to the top `need0`.
template <int THREADS, typename KeyAt>
__device__ __forceinline__ void block_kth_key(KeyAt key_at, int n, unsigned need0, unsigned *s_hist, unsigned *s_pn,
                                              unsigned &kth, unsigned &need_eq) {
    unsigned prefix = 0, need = need0;
    const int tid = threadIdx.x;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = tid; b < 256; b += THREADS) s_hist[b] = 0;
        __syncthreads();
        const unsigned himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
        for (int c = tid; c < n; c += THREADS) {
            unsigned key;
            if (key_at(c, key) && (key & himask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            unsigned cnt[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { cnt[q] = s_hist[255 - (8 * tid + q)]; sum += cnt[q]; }
            const unsigned incl = (unsigned)warp_incl_scan((int)sum, tid), excl = incl - sum;
            if (excl < need && need <= incl) {
                unsigned acc = excl;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (acc + cnt[q] >= need) {
                        s_pn[0] = prefix | ((unsigned)(255 - (8 * tid + q)) << shift);
                        s_pn[1] = need - acc;
                        break;
                    }
                    acc += cnt[q];
                }
            }
        }
        __syncthreads();
        prefix = s_pn[0];
        need = s_pn[1];
        __syncthreads();
    }
    kth = prefix;
    need_eq = need;
}

// ---- per (class, image): threshold, top-max_per_image, decode the winners -----------------------------
__global__ void __launch_bounds__(SEL_THREADS_D) detect_select_kernel(azn_detect_state st) {
    pdl_enter();
    extern __shared__ float s_score[];                    // [cap_boxes] class-j scores of the image, then int cand[mpi]
    __shared__ int s_warp[33];
    __shared__ unsigned s_hist[256], s_pn[2];
    const int j = blockIdx.x + 1, i = blockIdx.y, tid = threadIdx.x;
    const int C = st.num_classes, mpi = st.max_per_image, cap = st.cap_boxes, ld = st.ld_head;
    int n = st.n_boxes[i];
    n = n < 0 ? 0 : (n > cap ? cap : n);
    int *s_cand = (int *)(s_score + cap);
    const int *inv = st.inv + (size_t)i * cap, *rep = st.rep + (size_t)i * cap;
    const int row0 = st.img_off[i];
    const float th = st.thresh ? st.thresh[j] : -INFINITY;
    const size_t slot = (size_t)i * C + j;
    float *dets = st.dets + slot * mpi * 5;
    float *tops = st.top_scores + slot * mpi;
    if (j == 1 && tid == 0) st.det_count[(size_t)i * C] = 0;           // all_boxes[0][i] stays []
    // scores[inv_index, j] > thresh[j]  (:310, :609); NaN compares false like numpy
    int nvalid = 0;
    for (int c0 = 0; c0 < n; c0 += SEL_THREADS_D) {
        const int c = c0 + tid;
        int v = 0;
        if (c < n) {
            const float s = st.head_out[(size_t)(row0 + inv[c]) * ld + j];
            s_score[c] = s;
            v = s > th ? 1 : 0;
        }
        nvalid += __syncthreads_count(v);
    }
    __syncthreads();
    const int k = nvalid < mpi ? nvalid : mpi;
    auto key_at = [&](int c, unsigned &key) {
        const float s = s_score[c];
        key = score_key(s);
        return s > th;
    };
    unsigned kth = 0, need_eq = 0;
    const bool cut = nvalid > mpi;
    if (cut) block_kth_key<SEL_THREADS_D>(key_at, n, (unsigned)mpi, s_hist, s_pn, kth, need_eq);
    // ordered compaction of the winners (row order), ties at the k-th score: lowest rows first
    int base = 0, eq_seen = 0;
    for (int c0 = 0; c0 < n; c0 += SEL_THREADS_D) {
        const int c = c0 + tid;
        int gt = 0, eq = 0;
        if (c < n) {
            unsigned key;
            const bool valid = key_at(c, key);
            gt = valid && (!cut || key > kth);
            eq = valid && cut && key == kth;
        }
        int tot_eq, total;
        const int eq_pos = eq_seen + block_excl_scan(eq, tot_eq, s_warp);
        const int take = gt || (eq && eq_pos < (int)need_eq);
        const int pos = base + block_excl_scan(take, total, s_warp);
        if (take && pos < mpi) s_cand[pos] = c;
        base += total;
        eq_seen += tot_eq;
    }
    __syncthreads();
    const int m = base < mpi ? base : mpi;                               // == k
    const double wmax = (double)st.im_w[i] - 1.0, hmax = (double)st.im_h[i] - 1.0;
    const double *boxes = st.boxes + (size_t)i * cap * 4;
    for (int a = tid; a < mpi; a += SEL_THREADS_D) {
        if (a >= m) { tops[a] = -INFINITY; continue; }
        const int ca = s_cand[a];
        const unsigned ka = score_key(s_score[ca]);
        int rank = 0;
        for (int b = 0; b < m; ++b) {
            const int cb = s_cand[b];
            const unsigned kb = score_key(s_score[cb]);
            rank += (kb > ka || (kb == ka && cb < ca)) ? 1 : 0;
        }
        const int u = inv[ca];
        const double *box = boxes + (size_t)rep[u] * 4;                  // boxes = boxes[index] (:282): representative
        const float *d = st.head_out + (size_t)(row0 + u) * ld + C + 4 * j;
        double o[4];
        decode_clip(box[0], box[1], box[2], box[3], d[0], d[1], d[2], d[3], st.eps, wmax, hmax, o);
        float *dst = dets + (size_t)rank * 5;                            // hstack((boxes f64, scores)).astype(float32) (:636)
        dst[0] = (float)o[0]; dst[1] = (float)o[1]; dst[2] = (float)o[2]; dst[3] = (float)o[3];
        dst[4] = s_score[ca];
        tops[rank] = s_score[ca];
    }
    if (tid == 0) st.det_count[slot] = m;
    (void)k;
}

// ---- per class: the max_per_set-th highest pushed score of the whole image set ----------------------
__global__ void __launch_bounds__(TH_THREADS)
detect_thresh_kernel(const float *__restrict__ top_scores, const int32_t *__restrict__ det_count, int n_images, int C,
                     int mpi, long long max_per_set, float *__restrict__ thresh, int has_background) {
    __shared__ unsigned s_hist[256], s_pn[2];
    __shared__ int s_warp[33];
    const int j = blockIdx.x, tid = threadIdx.x;
    if (j == 0 && has_background) {
        if (tid == 0) thresh[0] = -INFINITY;
        return;
    }
    long long total = 0;
    {
        int part = 0;
        for (int i = tid; i < n_images; i += TH_THREADS) part += det_count[(size_t)i * C + j];
        // n_images * mpi < 2^31 is checked by the host
        int tot;
        block_excl_scan(part, tot, s_warp);
        total = tot;
    }
    if (total <= max_per_set) {                                          // the heap never overflowed (:626)
        if (tid == 0) thresh[j] = -INFINITY;
        return;
    }
    const int n = n_images * mpi;
    auto key_at = [&](int c, unsigned &key) {
        const int i = c / mpi, t = c - i * mpi;
        if (t >= det_count[(size_t)i * C + j]) return false;
        key = score_key(top_scores[((size_t)i * C + j) * mpi + t]);
        return true;
    };
    unsigned kth, need_eq;
    block_kth_key<TH_THREADS>(key_at, n, (unsigned)max_per_set, s_hist, s_pn, kth, need_eq);
    if (tid == 0) {
        const unsigned b = (kth & 0x80000000u) ? (kth & 0x7fffffffu) : ~kth;   // inverse of score_key
        thresh[j] = __uint_as_float(b);
    }
}

// ---- final filter: rows with score > thresh[class]; rows are score-descending, so a prefix survives ----
__global__ void detect_filter_kernel(const float *__restrict__ top_scores, int32_t *__restrict__ det_count,
                                     const float *__restrict__ thresh, long n_slots, int C, int mpi) {
    const long s = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int j = (int)(s % C);
    const int cnt = det_count[s];
    const float th = thresh[j];
    int keep = 0;
    for (int t = 0; t < cnt; ++t) keep += top_scores[(size_t)s * mpi + t] > th ? 1 : 0;
    det_count[s] = j == 0 ? 0 : keep;
}

int check_detect(const azn_detect_state *st, bool need_head) {
    AZN_REQUIRE(st != nullptr, "detect: null state");
    AZN_REQUIRE(st->n_img > 0 && st->cap_boxes > 0 && st->num_classes >= 2 && st->max_per_image > 0,
                "detect: bad sizes n_img=%d cap_boxes=%d num_classes=%d max_per_image=%d", st->n_img, st->cap_boxes,
                st->num_classes, st->max_per_image);
    AZN_REQUIRE(st->im_h && st->im_w && st->im_scale && st->boxes && st->n_boxes && st->inv && st->rep && st->n_uniq &&
                    st->img_off && st->rois && st->m_total && st->hashes && st->flags,
                "detect: null pointer in state");
    if (need_head) {
        AZN_REQUIRE(st->head_out && st->dets && st->top_scores && st->det_count, "detect: null output pointer in state");
        AZN_REQUIRE(st->ld_head >= 5 * st->num_classes, "detect: ld_head %d < 5 * num_classes", st->ld_head);
        AZN_REQUIRE(st->cap_boxes <= 8192 && st->max_per_image <= 1024, "detect: cap_boxes <= 8192, max_per_image <= 1024");
    }
    return AZN_OK;
}

}  // namespace

extern "C" int azn_detect_rois(const azn_detect_state *st, azn_stream_t stream) {
    int rc = check_detect(st, false);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    AZN_CUDA(azn_launch_pdl(detect_rois_kernel, dim3(st->n_img), dim3(ROIS_THREADS), 0, s, *st));
    AZN_CUDA(azn_launch_pdl(detect_pack_kernel, dim3(st->n_img), dim3(256), 0, s, *st));
    return AZN_OK;
}

extern "C" int azn_detect_select(const azn_detect_state *st, azn_stream_t stream) {
    int rc = check_detect(st, true);
    if (rc) return rc;
    const size_t smem = (size_t)st->cap_boxes * sizeof(float) + (size_t)st->max_per_image * sizeof(int);
    AZN_CUDA(azn_launch_pdl(detect_select_kernel, dim3(st->num_classes - 1, st->n_img), dim3(SEL_THREADS_D), smem,
                            (cudaStream_t)stream, *st));
    return AZN_OK;
}

extern "C" int azn_detect_thresholds(const float *top_scores, const int32_t *det_count, int n_images, int num_classes,
                                     int max_per_image, long long max_per_set, float *thresh, azn_stream_t stream) {
    AZN_REQUIRE(top_scores && det_count && thresh, "azn_detect_thresholds: null pointer");
    AZN_REQUIRE(n_images > 0 && num_classes >= 2 && max_per_image > 0 && max_per_set > 0,
                "azn_detect_thresholds: bad sizes");
    AZN_REQUIRE((double)n_images * max_per_image < 2.0e9, "azn_detect_thresholds: too many scores for one launch");
    detect_thresh_kernel<<<num_classes, TH_THREADS, 0, (cudaStream_t)stream>>>(top_scores, det_count, n_images, num_classes,
                                                                             max_per_image, max_per_set, thresh, 1);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

// tune_thresh (lib/detect/tune.py:318-366): one "class" without a background column, scores = anchor zoom scores
extern "C" int azn_tune_threshold(const float *zoom, const int32_t *counts, int n_images, int cap, long long max_per_set,
                                  float *thresh, azn_stream_t stream) {
    AZN_REQUIRE(zoom && counts && thresh, "azn_tune_threshold: null pointer");
    AZN_REQUIRE(n_images > 0 && cap > 0 && max_per_set > 0, "azn_tune_threshold: bad sizes");
    AZN_REQUIRE((double)n_images * cap < 2.0e9, "azn_tune_threshold: too many scores for one launch");
    detect_thresh_kernel<<<1, TH_THREADS, 0, (cudaStream_t)stream>>>(zoom, counts, n_images, 1, cap, max_per_set, thresh, 0);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_detect_filter(const float *top_scores, int32_t *det_count, const float *thresh, int n_images,
                                 int num_classes, int max_per_image, azn_stream_t stream) {
    AZN_REQUIRE(top_scores && det_count && thresh && n_images > 0 && num_classes >= 2 && max_per_image > 0,
                "azn_detect_filter: bad argument");
    const long n_slots = (long)n_images * num_classes;
    detect_filter_kernel<<<(unsigned)((n_slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(top_scores, det_count, thresh,
                                                                                            n_slots, num_classes, max_per_image);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
