// Device helpers shared by the search-level kernels (search.cu) and the detection kernels (detect.cu):
// block scan, the reference's box decode, ROI projection, the feature-space dedup hash, the exact stable
// `np.unique` restatement and the monotone score key.  Semantics and citations: see search.cu's header.
#pragma once
#include "common.cuh"

namespace {

// exclusive block scan of one int per thread; returns the exclusive prefix, `total` = block sum.
__device__ __forceinline__ int block_excl_scan(int v, int &total, int *s_warp /* [33] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const int incl = warp_incl_scan(v, lane);
    __syncthreads();                       // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? s_warp[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        s_warp[lane] = wi - w;             // exclusive warp offsets
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[warp] + incl - v;
}

// ---- _bbox_pred + _clip_boxes for one (box, 4 deltas) ------------------------------------
template <bool CLIP = true>
__device__ __forceinline__ void decode_clip(const double bx1, const double by1, const double bx2, const double by2,
                                            const float dx, const float dy, const float dw, const float dh,
                                            const double eps, const double wmax, const double hmax, double out[4]) {
    const double w = __dadd_rn(__dsub_rn(bx2, bx1), eps);
    const double h = __dadd_rn(__dsub_rn(by2, by1), eps);
    const double cx = __dadd_rn(bx1, __dmul_rn(0.5, w));
    const double cy = __dadd_rn(by1, __dmul_rn(0.5, h));
    const double pcx = __dadd_rn(__dmul_rn((double)dx, w), cx);
    const double pcy = __dadd_rn(__dmul_rn((double)dy, h), cy);
    const double pw = __dmul_rn((double)expf(dw), w);        // np.exp runs in float32 (Q6)
    const double ph = __dmul_rn((double)expf(dh), h);
    double x1 = __dsub_rn(pcx, __dmul_rn(0.5, pw));
    double y1 = __dsub_rn(pcy, __dmul_rn(0.5, ph));
    double x2 = __dadd_rn(pcx, __dmul_rn(0.5, pw));
    double y2 = __dadd_rn(pcy, __dmul_rn(0.5, ph));
    if (!CLIP) { out[0] = x1; out[1] = y1; out[2] = x2; out[3] = y2; return; }
    out[0] = x1 > 0.0 ? x1 : 0.0;                            // np.maximum(., 0)
    out[1] = y1 > 0.0 ? y1 : 0.0;
    out[2] = x2 < wmax ? x2 : wmax;                          // np.minimum(., W-1)
    out[3] = y2 < hmax ? y2 : hmax;
}

__device__ __forceinline__ void project_roi(const double *__restrict__ r, const double scale, float out[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = __double2float_rn(__dmul_rn(r[k], scale));   // _project_im_rois + astype(float32)
}

// hash of the feature-space dedup: np.round(rois_f32 * DEDUP).dot([1,1e3,1e6,1e9,1e12]); the
// level column is 0 for single-scale testing.  float32 multiply + rint, exact int64 dot.
__device__ __forceinline__ long long feat_hash(const float roi[4], const float dedup) {
    const long long a = (long long)rintf(__fmul_rn(roi[0], dedup));
    const long long b = (long long)rintf(__fmul_rn(roi[1], dedup));
    const long long c = (long long)rintf(__fmul_rn(roi[2], dedup));
    const long long d = (long long)rintf(__fmul_rn(roi[3], dedup));
    return a * 1000LL + b * 1000000LL + c * 1000000000LL + d * 1000000000000LL;
}

// Stable unique by key over n elements held in global scratch (one CTA).
//   flags[c] <- 1 iff c is the first element with its key; slot[c] (may alias flags' storage
//   for non-representatives) is not stored: callers recompute through `unique_slot`.
__device__ __forceinline__ void mark_first(const long long *__restrict__ keys, int *__restrict__ flags, int n) {
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const long long k = keys[c];
        int first = 1;
        for (int j = 0; j < c; ++j)
            if (keys[j] == k) { first = 0; break; }
        flags[c] = first;
    }
}

__device__ __forceinline__ int unique_slot(const long long *__restrict__ keys, const int *__restrict__ flags, int n,
                                           long long k) {
    int slot = 0;
    for (int j = 0; j < n; ++j) slot += (flags[j] && keys[j] < k) ? 1 : 0;
    return slot;
}

__device__ __forceinline__ unsigned score_key(float s) {      // monotone float -> uint
    const unsigned b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}


}  // namespace
