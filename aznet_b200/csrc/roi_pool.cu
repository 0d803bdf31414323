// ROI max pooling, forward, for sm_100a.
//
// Reference semantics: ROIPoolingLayer<Dtype>::Forward_cpu,
// caffe-fast-rcnn/src/caffe/layers/roi_pooling_layer.cpp:46-125 (GPU twin .cu:18-92):
//   roi_{start,end} = round(coord * scale) (half away from zero), roi_h/w = max(end-start+1, 1),
//   bin = roi / pooled (f32), bin p covers [floor(p*bin), ceil((p+1)*bin)) + start clamped to the
//   map, empty bin -> 0, otherwise the maximum found by a strict `>` scan from -FLT_MAX.
//
// Two kernels:
//  (1) roi_pool_nhwc_kernel -- the production path.  Channels-last map, output [R, PH, PW, C].
//      One thread owns one 16-byte channel vector of one output bin; consecutive threads own
//      consecutive vectors, so global loads (one map cell = C contiguous channels) and the
//      stores are fully coalesced 16-byte accesses.  The map (<= a few MB per image) is served
//      by L2/L1; HBM traffic is the streaming write of the pooled rows.
//  (2) roi_pool_nchw_kernel -- Caffe blob layout in and out (+ optional argmax), the drop-in
//      for the layer itself.  One CTA stages the ROI window of a group of channels in shared
//      memory with row-coalesced loads, then every thread scans its bin out of shared memory
//      and the CTA writes its C_GROUP*PH*PW outputs as one contiguous run.
#include <float.h>
#include "common.cuh"

namespace {

struct RoiGeom {
    int b, start_w, start_h;
    float bin_h, bin_w;
};

__device__ __forceinline__ RoiGeom roi_geom(const float *__restrict__ roi, float scale, int PH, int PW) {
    RoiGeom g;
    g.b = (int)roi[0];
    g.start_w = (int)roundf(__fmul_rn(roi[1], scale));
    g.start_h = (int)roundf(__fmul_rn(roi[2], scale));
    int end_w = (int)roundf(__fmul_rn(roi[3], scale));
    int end_h = (int)roundf(__fmul_rn(roi[4], scale));
    int roi_h = max(end_h - g.start_h + 1, 1);
    int roi_w = max(end_w - g.start_w + 1, 1);
    g.bin_h = __fdiv_rn((float)roi_h, (float)PH);
    g.bin_w = __fdiv_rn((float)roi_w, (float)PW);
    return g;
}

__device__ __forceinline__ void bin_bounds(int p, float bin, int start, int limit, int &lo, int &hi) {
    lo = (int)floorf(__fmul_rn((float)p, bin));
    hi = (int)ceilf(__fmul_rn((float)(p + 1), bin));
    lo = min(max(lo + start, 0), limit);
    hi = min(max(hi + start, 0), limit);
}

// ---- 16-byte channel vectors -----------------------------------------------------------
struct VecF32 {
    float v[4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = -FLT_MAX;
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = 0.f;
    }
    __device__ __forceinline__ void take(const uint4 &q) {
        float x[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = x[i] > v[i] ? x[i] : v[i];
    }
    __device__ __forceinline__ uint4 pack() const {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
};

struct VecBF16 {   // 8 bf16 lanes, compared as the f32 values they denote
    float v[8];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = -INFINITY;   // bf16(-FLT_MAX) rounds to -inf
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    __device__ __forceinline__ void take(const uint4 &q) {
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
            v[2 * i] = lo > v[2 * i] ? lo : v[2 * i];
            v[2 * i + 1] = hi > v[2 * i + 1] ? hi : v[2 * i + 1];
        }
    }
    __device__ __forceinline__ uint4 pack() const {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            w[i] = (__float_as_uint(v[2 * i]) >> 16) | (__float_as_uint(v[2 * i + 1]) & 0xffff0000u);
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ uint4 ld_map(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
    // pooled rows are written once and read next by another kernel: keep them out of L1
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// blockDim = (LX lanes, SY bin slots).  L = 16-byte vectors per map cell (C * sizeof(T) / 16).
template <typename Vec>
__global__ void __launch_bounds__(256)
roi_pool_nhwc_kernel(const uint4 *__restrict__ feat, int n_img, int H, int W, int L,
                     const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                     int PH, int PW, float scale, uint4 *__restrict__ out) {
    const int R = n_rois ? min(*n_rois, R_cap) : R_cap;
    const unsigned bins = (unsigned)(PH * PW);
    const unsigned total_bins = (unsigned)R * bins;
    const size_t row_stride = (size_t)W * L;      // vectors per map row
    for (unsigned g = blockIdx.x * blockDim.y + threadIdx.y; g < total_bins; g += gridDim.x * blockDim.y) {
        const unsigned r = g / bins, bin = g - r * bins;
        const int ph = (int)(bin / (unsigned)PW), pw = (int)(bin - (unsigned)ph * PW);
        const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, PH, PW);
        int hs, he, ws, we;
        bin_bounds(ph, q.bin_h, q.start_h, H, hs, he);
        bin_bounds(pw, q.bin_w, q.start_w, W, ws, we);
        const bool bad_img = q.b < 0 || q.b >= n_img;
        const bool empty = bad_img || he <= hs || we <= ws;
        const uint4 *base = feat + (size_t)(bad_img ? 0 : q.b) * H * row_stride;
        for (int lane = threadIdx.x; lane < L; lane += blockDim.x) {
            Vec acc;
            if (empty) {
                acc.zero();
            } else {
                acc.init();
                for (int h = hs; h < he; ++h) {
                    const uint4 *p = base + (size_t)h * row_stride + (size_t)ws * L + lane;
#pragma unroll 4
                    for (int w = ws; w < we; ++w, p += L) acc.take(ld_map(p));
                }
            }
            st_stream(out + (size_t)g * L + lane, acc.pack());
        }
    }
}

// Caffe layout: f32 NCHW in, [R, C, PH, PW] out (+ argmax).  CTA = (roi, group of CG channels).
constexpr int CG = 8;
constexpr int NCHW_THREADS = 128;
constexpr int NCHW_SMEM_FLOATS = 10240;   // 40 KB window budget

__global__ void __launch_bounds__(NCHW_THREADS)
roi_pool_nchw_kernel(const float *__restrict__ feat, int n_img, int C, int H, int W,
                     const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                     int PH, int PW, float scale, float *__restrict__ out, int32_t *__restrict__ argmax) {
    extern __shared__ float win[];
    const int R = n_rois ? min(*n_rois, R_cap) : R_cap;
    const int groups = (C + CG - 1) / CG;
    const int bins = PH * PW;
    for (long item = blockIdx.x; item < (long)R * groups; item += gridDim.x) {
        const int r = (int)(item / groups), c0 = (int)(item - (long)r * groups) * CG;
        const int nc = min(CG, C - c0);
        const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, PH, PW);
        const bool bad_img = q.b < 0 || q.b >= n_img;
        int h0, h1, w0, w1, t0, t1;
        bin_bounds(0, q.bin_h, q.start_h, H, h0, t0);
        bin_bounds(PH - 1, q.bin_h, q.start_h, H, t1, h1);
        bin_bounds(0, q.bin_w, q.start_w, W, w0, t0);
        bin_bounds(PW - 1, q.bin_w, q.start_w, W, t1, w1);
        const int wh = max(h1 - h0, 0), ww = max(w1 - w0, 0);
        const bool staged = !bad_img && (long)wh * ww * nc <= NCHW_SMEM_FLOATS && wh * ww > 0;
        const float *img = feat + ((size_t)(bad_img ? 0 : q.b) * C + c0) * H * W;
        __syncthreads();   // previous item's readers are done with `win`
        if (staged) {
            const int rows = nc * wh;
            for (int row = threadIdx.x / 32; row < rows; row += NCHW_THREADS / 32) {
                const int c = row / wh, h = row - c * wh;
                const float *src = img + ((size_t)c * H + (h0 + h)) * W + w0;
                float *dst = win + (size_t)row * ww;
                for (int w = threadIdx.x % 32; w < ww; w += 32) dst[w] = __ldg(src + w);
            }
        }
        __syncthreads();
        for (int o = threadIdx.x; o < nc * bins; o += NCHW_THREADS) {
            const int c = o / bins, bin = o - c * bins;
            const int ph = bin / PW, pw = bin - ph * PW;
            int hs, he, ws, we;
            bin_bounds(ph, q.bin_h, q.start_h, H, hs, he);
            bin_bounds(pw, q.bin_w, q.start_w, W, ws, we);
            float best = -FLT_MAX;
            int best_i = -1;
            if (bad_img || he <= hs || we <= ws) {
                best = 0.f;
            } else if (staged) {
                const float *wc = win + (size_t)c * wh * ww;
                for (int h = hs; h < he; ++h)
                    for (int w = ws; w < we; ++w) {
                        float v = wc[(h - h0) * ww + (w - w0)];
                        if (v > best) { best = v; best_i = h * W + w; }
                    }
            } else {
                const float *gc = img + (size_t)c * H * W;
                for (int h = hs; h < he; ++h)
                    for (int w = ws; w < we; ++w) {
                        float v = __ldg(gc + (size_t)h * W + w);
                        if (v > best) { best = v; best_i = h * W + w; }
                    }
            }
            const size_t oi = ((size_t)r * C + c0) * bins + o;
            out[oi] = best;
            if (argmax) argmax[oi] = best_i;
        }
    }
}

// bf16 NCHW in/out reuses the f32 kernel's structure only for completeness of the dtype matrix:
// the search engine never uses it, so it is served by converting on the fly per element.
__global__ void __launch_bounds__(256)
roi_pool_nchw_bf16_kernel(const __nv_bfloat16 *__restrict__ feat, int n_img, int C, int H, int W,
                          const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                          int PH, int PW, float scale, __nv_bfloat16 *__restrict__ out) {
    const int R = n_rois ? min(*n_rois, R_cap) : R_cap;
    const int bins = PH * PW;
    const long total = (long)R * C * bins;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int bin = (int)(i % bins);
        const int c = (int)((i / bins) % C);
        const int r = (int)(i / ((long)bins * C));
        const int ph = bin / PW, pw = bin - ph * PW;
        const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, PH, PW);
        int hs, he, ws, we;
        bin_bounds(ph, q.bin_h, q.start_h, H, hs, he);
        bin_bounds(pw, q.bin_w, q.start_w, W, ws, we);
        const bool bad_img = q.b < 0 || q.b >= n_img;
        float best = -INFINITY;
        if (bad_img || he <= hs || we <= ws) best = 0.f;
        else {
            const __nv_bfloat16 *gc = feat + ((size_t)q.b * C + c) * H * W;
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    float v = __bfloat162float(gc[(size_t)h * W + w]);
                    if (v > best) best = v;
                }
        }
        out[i] = __float2bfloat16_rn(best);
    }
}

// f32 NCHW -> bf16 NHWC through a 32x33 shared tile (both sides coalesced).
__global__ void __launch_bounds__(256)
nchw_to_nhwc_bf16_kernel(const float *__restrict__ src, int C, int HW, __nv_bfloat16 *__restrict__ dst) {
    __shared__ float tile[32][33];
    const int img = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *s = src + (size_t)img * C * HW;
    __nv_bfloat16 *d = dst + (size_t)img * C * HW;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? s[(size_t)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int p = p0 + j, c = c0 + threadIdx.x;
        if (p < HW && c < C) d[(size_t)p * C + c] = __float2bfloat16_rn(tile[threadIdx.x][j]);
    }
}

}  // namespace

extern "C" int azn_roi_pool_fwd(const void *feat, int n_img, int C, int H, int W, int layout, int dtype,
                                const float *rois, const int32_t *n_rois, int R_cap, int PH, int PW,
                                float spatial_scale, void *out, int32_t *argmax, azn_stream_t stream) {
    if (R_cap == 0) return AZN_OK;
    AZN_REQUIRE(feat && rois && out, "azn_roi_pool_fwd: null pointer");
    AZN_REQUIRE(n_img > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && R_cap >= 0,
                "azn_roi_pool_fwd: bad shape n_img=%d C=%d H=%d W=%d PH=%d PW=%d R=%d", n_img, C, H, W, PH, PW, R_cap);
    AZN_REQUIRE(dtype == AZN_DTYPE_F32 || dtype == AZN_DTYPE_BF16, "azn_roi_pool_fwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const int sms = azn_num_sms();
    if (layout == AZN_LAYOUT_NHWC) {
        AZN_REQUIRE(argmax == nullptr, "azn_roi_pool_fwd: argmax is only produced for the NCHW layout");
        const int esize = dtype == AZN_DTYPE_F32 ? 4 : 2;
        AZN_REQUIRE((C * esize) % 16 == 0, "azn_roi_pool_fwd: NHWC needs C*sizeof(dtype) %% 16 == 0 (C=%d)", C);
        AZN_REQUIRE(((uintptr_t)feat % 16 == 0) && ((uintptr_t)out % 16 == 0), "azn_roi_pool_fwd: 16-byte alignment");
        const int L = C * esize / 16;
        AZN_REQUIRE((double)R_cap * PH * PW < 2.0e9, "azn_roi_pool_fwd: too many bins for one launch");
        const int LX = L < 128 ? L : 128;
        const int SY = 256 / LX > 0 ? 256 / LX : 1;
        dim3 block(LX, SY);
        const long total_bins = (long)R_cap * PH * PW;
        long blocks = (total_bins + SY - 1) / SY;
        const long max_blocks = (long)sms * 16;
        if (blocks > max_blocks) blocks = max_blocks;
        if (dtype == AZN_DTYPE_F32)
            roi_pool_nhwc_kernel<VecF32><<<(unsigned)blocks, block, 0, s>>>(
                (const uint4 *)feat, n_img, H, W, L, rois, n_rois, R_cap, PH, PW, spatial_scale, (uint4 *)out);
        else
            roi_pool_nhwc_kernel<VecBF16><<<(unsigned)blocks, block, 0, s>>>(
                (const uint4 *)feat, n_img, H, W, L, rois, n_rois, R_cap, PH, PW, spatial_scale, (uint4 *)out);
        AZN_LAUNCH_CHECK();
        return AZN_OK;
    }
    AZN_REQUIRE(layout == AZN_LAYOUT_NCHW, "azn_roi_pool_fwd: bad layout %d", layout);
    if (dtype == AZN_DTYPE_F32) {
        const long items = (long)R_cap * ((C + CG - 1) / CG);
        long blocks = items < (long)sms * 32 ? items : (long)sms * 32;
        roi_pool_nchw_kernel<<<(unsigned)blocks, NCHW_THREADS, NCHW_SMEM_FLOATS * sizeof(float), s>>>(
            (const float *)feat, n_img, C, H, W, rois, n_rois, R_cap, PH, PW, spatial_scale, (float *)out, argmax);
    } else {
        AZN_REQUIRE(argmax == nullptr, "azn_roi_pool_fwd: argmax needs f32");
        const long total = (long)R_cap * C * PH * PW;
        long blocks = (total + 255) / 256;
        if (blocks > (long)sms * 32) blocks = (long)sms * 32;
        roi_pool_nchw_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(
            (const __nv_bfloat16 *)feat, n_img, C, H, W, rois, n_rois, R_cap, PH, PW, spatial_scale,
            (__nv_bfloat16 *)out);
    }
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_nchw_f32_to_nhwc_bf16(const float *src, int n_img, int C, int H, int W, void *dst,
                                         azn_stream_t stream) {
    AZN_REQUIRE(src && dst && n_img > 0 && C > 0 && H > 0 && W > 0, "azn_nchw_f32_to_nhwc_bf16: bad argument");
    const int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, n_img), block(32, 8);
    nchw_to_nhwc_bf16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, C, HW, (__nv_bfloat16 *)dst);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
