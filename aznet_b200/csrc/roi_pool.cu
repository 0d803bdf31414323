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

// ---- 16-byte channel vectors: exact `v > best ? v : best` on 4 f32 / 8 bf16 lanes -------
struct OpsF32 {
    static __device__ __forceinline__ uint4 lowest() {
        const unsigned m = __float_as_uint(-FLT_MAX);
        return make_uint4(m, m, m, m);
    }
    static __device__ __forceinline__ unsigned pick(unsigned v, unsigned b) {
        return __uint_as_float(v) > __uint_as_float(b) ? v : b;
    }
    static __device__ __forceinline__ void take(uint4 &acc, const uint4 &q) {
        acc.x = pick(q.x, acc.x); acc.y = pick(q.y, acc.y); acc.z = pick(q.z, acc.z); acc.w = pick(q.w, acc.w);
    }
};

struct OpsBF16 {   // packed: HSET2.BF16.GT (mask) + LOP3 (select) per pair; gt is false on NaN and on +0 vs -0
    static __device__ __forceinline__ uint4 lowest() {
        return make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16(-FLT_MAX) rounds to -inf
    }
    static __device__ __forceinline__ unsigned pick(unsigned v, unsigned b) {
        const unsigned m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&v), *reinterpret_cast<const __nv_bfloat162 *>(&b));
        return (v & m) | (b & ~m);
    }
    static __device__ __forceinline__ void take(uint4 &acc, const uint4 &q) {
        acc.x = pick(q.x, acc.x); acc.y = pick(q.y, acc.y); acc.z = pick(q.z, acc.z); acc.w = pick(q.w, acc.w);
    }
};

__device__ __forceinline__ uint4 ld_map(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
    // pooled rows are written once and read next by another kernel: keep them out of L1
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// One CTA owns one (roi, ph) row of bins, one warp per bin of the row: lanes 0..PW-1 of the CTA derive the ROI
// geometry once per row (shared through shared memory), then every warp max-reduces its bin with its lanes
// spanning the channel vectors (NV 16-byte vectors per lane, 32*NV per pass), so every global access of a
// warp is a run of consecutive 16-byte vectors and large ROIs still spread over PW warps per row.
// L = vectors per map cell (C * sizeof(T) / 16).
constexpr int POOL_MAX_PW = 32;

template <typename Ops, int NV>
__global__ void __launch_bounds__(256)
roi_pool_nhwc_kernel(const uint4 *__restrict__ feat, int n_img, int H, int W, int L,
                     const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                     int PH, int PW, float scale, uint4 *__restrict__ out) {
    __shared__ int s_ws[2][POOL_MAX_PW], s_we[2][POOL_MAX_PW], s_row[2][3];
    const int R = n_rois ? min(*n_rois, R_cap) : R_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const unsigned items = (unsigned)R * (unsigned)PH;
    const size_t row_stride = (size_t)W * L;      // vectors per map row
    int buf = 0;
    for (unsigned it = blockIdx.x; it < items; it += gridDim.x, buf ^= 1) {
        const unsigned r = it / (unsigned)PH;
        const int ph = (int)(it - r * (unsigned)PH);
        if ((int)threadIdx.x < PW) {
            const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, PH, PW);
            int ws, we;
            bin_bounds(threadIdx.x, q.bin_w, q.start_w, W, ws, we);
            s_ws[buf][threadIdx.x] = ws;
            s_we[buf][threadIdx.x] = we;
            if (threadIdx.x == 0) {
                int hs, he;
                bin_bounds(ph, q.bin_h, q.start_h, H, hs, he);
                s_row[buf][0] = hs;
                s_row[buf][1] = he;
                s_row[buf][2] = (q.b < 0 || q.b >= n_img) ? -1 : q.b;
            }
        }
        __syncthreads();                          // double-buffered geometry: one barrier per item
        const int hs = s_row[buf][0], he = s_row[buf][1], b = s_row[buf][2];
        const uint4 *base = feat + (size_t)(b < 0 ? 0 : b) * H * row_stride;
        uint4 *orow = out + (size_t)it * PW * L;
        for (int pw = warp; pw < PW; pw += nwarps) {
            const int ws = s_ws[buf][pw], we = s_we[buf][pw];
            const bool empty = b < 0 || he <= hs || we <= ws;
            for (int v0 = lane; v0 < L; v0 += 32 * NV) {
                uint4 acc[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) acc[j] = empty ? make_uint4(0u, 0u, 0u, 0u) : Ops::lowest();
                if (!empty) {
                    for (int h = hs; h < he; ++h) {
                        const uint4 *p = base + (size_t)h * row_stride + (size_t)ws * L + v0;
#pragma unroll 4
                        for (int w = ws; w < we; ++w, p += L) {
#pragma unroll
                            for (int j = 0; j < NV; ++j)
                                if (v0 + 32 * j < L) Ops::take(acc[j], ld_map(p + 32 * j));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    if (v0 + 32 * j < L) st_stream(orow + (size_t)pw * L + v0 + 32 * j, acc[j]);
            }
        }
    }
}

// Caffe blob layout out: [R, C, PH, PW] f32 (+ argmax).  The map is read channels-last (the f32 NHWC copy
// that azn_roi_pool_fwd makes in its workspace), one CTA pools one ROI (x one channel chunk) into a
// [bins][CC+1] shared tile -- coalesced 16-byte map loads, lanes along channels -- and then streams the tile
// out transposed, so the CTA's global writes are one contiguous run of CC*PH*PW floats (the +1 padding
// makes the transposed shared reads conflict-free).
constexpr int CAFFE_THREADS = 256;

template <bool ARGMAX, int VEC>
__global__ void __launch_bounds__(CAFFE_THREADS)
roi_pool_caffe_kernel(const float *__restrict__ nhwc, int n_img, int C, int H, int W,
                      const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                      int PH, int PW, float scale, float *__restrict__ out, int32_t *__restrict__ argmax, int CC) {
    extern __shared__ float tile[];
    const int R = n_rois ? min(*n_rois, R_cap) : R_cap;
    const int bins = PH * PW, ld = CC + 1;
    int *atile = reinterpret_cast<int *>(tile + (size_t)bins * ld);
    const int nchunks = (C + CC - 1) / CC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long item = blockIdx.x; item < (long)R * nchunks; item += gridDim.x) {
        const int r = (int)(item / nchunks), c0 = (int)(item - (long)r * nchunks) * CC;
        const int cc = min(CC, C - c0);
        const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, PH, PW);
        const bool bad_img = q.b < 0 || q.b >= n_img;
        const float *img = nhwc + (size_t)(bad_img ? 0 : q.b) * H * W * C + c0;
        for (int bin = warp; bin < bins; bin += CAFFE_THREADS / 32) {
            const int ph = bin / PW, pw = bin - ph * PW;
            int hs, he, ws, we;
            bin_bounds(ph, q.bin_h, q.start_h, H, hs, he);
            bin_bounds(pw, q.bin_w, q.start_w, W, ws, we);
            const bool empty = bad_img || he <= hs || we <= ws;
            for (int c = lane * VEC; c < cc; c += 32 * VEC) {
                float best[VEC];
                int bi[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) { best[k] = empty ? 0.f : -FLT_MAX; bi[k] = -1; }
                if (!empty) {
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w) {
                            const float *p = img + ((size_t)h * W + w) * C + c;
                            float v[VEC];
                            if (VEC == 4) {
                                const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
                                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
                            } else {
                                v[0] = __ldg(p);
                            }
#pragma unroll
                            for (int k = 0; k < VEC; ++k)
                                if (v[k] > best[k]) { best[k] = v[k]; if (ARGMAX) bi[k] = h * W + w; }
                        }
                }
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    tile[(size_t)bin * ld + c + k] = best[k];
                    if (ARGMAX) atile[(size_t)bin * ld + c + k] = bi[k];
                }
            }
        }
        __syncthreads();
        const size_t obase = ((size_t)r * C + c0) * bins;
        for (int o = threadIdx.x; o < cc * bins; o += CAFFE_THREADS) {
            const int c = o / bins, bin = o - c * bins;
            out[obase + o] = tile[(size_t)bin * ld + c];
            if (ARGMAX) argmax[obase + o] = atile[(size_t)bin * ld + c];
        }
        __syncthreads();
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

// f32 NCHW -> bf16 / f32 NHWC through a 32x33 shared tile (both sides coalesced).
__device__ __forceinline__ void cvt_store(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void cvt_store(float *p, float v) { *p = v; }

template <typename TOut>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float *__restrict__ src, int C, int HW, TOut *__restrict__ dst) {
    __shared__ float tile[32][33];
    const int img = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *s = src + (size_t)img * C * HW;
    TOut *d = dst + (size_t)img * C * HW;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? s[(size_t)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int p = p0 + j, c = c0 + threadIdx.x;
        if (p < HW && c < C) cvt_store(d + (size_t)p * C + c, tile[threadIdx.x][j]);
    }
}

}  // namespace

extern "C" size_t azn_roi_pool_workspace_bytes(int n_img, int C, int H, int W, int layout, int dtype) {
    if (layout == AZN_LAYOUT_NCHW && dtype == AZN_DTYPE_F32) return (size_t)n_img * C * H * W * sizeof(float);
    return 0;
}

extern "C" int azn_roi_pool_fwd(const void *feat, int n_img, int C, int H, int W, int layout, int dtype,
                                const float *rois, const int32_t *n_rois, int R_cap, int PH, int PW,
                                float spatial_scale, void *out, int32_t *argmax, void *workspace,
                                size_t workspace_bytes, azn_stream_t stream) {
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
        AZN_REQUIRE((double)R_cap * PH < 2.0e9, "azn_roi_pool_fwd: too many ROI rows for one launch");
        AZN_REQUIRE(PW <= POOL_MAX_PW, "azn_roi_pool_fwd: pooled width %d > %d", PW, POOL_MAX_PW);
        const long items = (long)R_cap * PH;
        const int nwarps = PW < 8 ? PW : 8;
        long blocks = items;
        const long max_blocks = (long)sms * 16;
        if (blocks > max_blocks) blocks = max_blocks;
        const int nv = L >= 128 ? 4 : (L >= 64 ? 2 : 1);
        const uint4 *f = (const uint4 *)feat;
        uint4 *o = (uint4 *)out;
#define AZN_POOL_LAUNCH(OPS, NVV) \
        roi_pool_nhwc_kernel<OPS, NVV><<<(unsigned)blocks, nwarps * 32, 0, s>>>(f, n_img, H, W, L, rois, n_rois, R_cap, PH, PW, spatial_scale, o)
        if (dtype == AZN_DTYPE_F32) {
            if (nv == 4) AZN_POOL_LAUNCH(OpsF32, 4); else if (nv == 2) AZN_POOL_LAUNCH(OpsF32, 2); else AZN_POOL_LAUNCH(OpsF32, 1);
        } else {
            if (nv == 4) AZN_POOL_LAUNCH(OpsBF16, 4); else if (nv == 2) AZN_POOL_LAUNCH(OpsBF16, 2); else AZN_POOL_LAUNCH(OpsBF16, 1);
        }
#undef AZN_POOL_LAUNCH
        AZN_LAUNCH_CHECK();
        return AZN_OK;
    }
    AZN_REQUIRE(layout == AZN_LAYOUT_NCHW, "azn_roi_pool_fwd: bad layout %d", layout);
    if (dtype == AZN_DTYPE_F32) {
        const size_t need = azn_roi_pool_workspace_bytes(n_img, C, H, W, layout, dtype);
        if (!workspace || workspace_bytes < need) {
            azn_set_error("azn_roi_pool_fwd: the NCHW f32 path needs a %zu-byte workspace (got %zu)", need, workspace_bytes);
            return AZN_ERR_CAPACITY;
        }
        float *nhwc = (float *)workspace;
        const int HW = H * W;
        nchw_to_nhwc_kernel<float><<<dim3((HW + 31) / 32, (C + 31) / 32, n_img), dim3(32, 8), 0, s>>>((const float *)feat, C, HW, nhwc);
        AZN_LAUNCH_CHECK();
        const int bins = PH * PW;
        const int per = argmax ? 8 : 4;                        // bytes of shared tile per (bin, channel)
        int CC = (int)((200 * 1024) / ((size_t)bins * per)) - 1;
        AZN_REQUIRE(CC >= 32, "azn_roi_pool_fwd: pooled size %dx%d too large for the shared tile", PH, PW);
        CC = CC / 32 * 32;
        if (CC > C) CC = C;
        const size_t smem = (size_t)bins * (CC + 1) * per;
        const long items = (long)R_cap * ((C + CC - 1) / CC);
        long blocks = items < (long)sms * 4 ? items : (long)sms * 4;
        const bool vec = (C % 4 == 0) && (CC % 4 == 0);
#define AZN_CAFFE_LAUNCH(AM, V)                                                                                          \
        do {                                                                                                             \
            AZN_CUDA(cudaFuncSetAttribute(roi_pool_caffe_kernel<AM, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            roi_pool_caffe_kernel<AM, V><<<(unsigned)blocks, CAFFE_THREADS, smem, s>>>(                                  \
                nhwc, n_img, C, H, W, rois, n_rois, R_cap, PH, PW, spatial_scale, (float *)out, argmax, CC);            \
        } while (0)
        if (argmax) { if (vec) AZN_CAFFE_LAUNCH(true, 4); else AZN_CAFFE_LAUNCH(true, 1); }
        else        { if (vec) AZN_CAFFE_LAUNCH(false, 4); else AZN_CAFFE_LAUNCH(false, 1); }
#undef AZN_CAFFE_LAUNCH
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
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(src, C, HW, (__nv_bfloat16 *)dst);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
