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
#include <algorithm>
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
    typedef float tag;
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

struct OpsBF16 {
    typedef __nv_bfloat16 tag;
    // packed: HSET2.BF16.GT (mask) + LOP3 (select) per pair; gt is false on NaN and on +0 vs -0
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

// GRN = true (bf16 only, L <= 32 * NV so that one warp holds every channel of its bin): the skip-layer head's
// roi_norm + concat + Power epilogue (VGG16_skip test_fc.prototxt:39-110, grn_layer.cpp:27-56) fused into the pool --
// the warp sums the squares of its bin's channels (float32; a bf16 square is exact), and writes
// grn_scale * (x / sqrt(sum)) at channel-vector offset out_off of a row of out_ld vectors (the concat buffer).
__device__ __forceinline__ float sq_pair(unsigned u) {
    const float lo = __uint_as_float(u << 16), hi = __uint_as_float(u & 0xffff0000u);
    return __fadd_rn(__fmul_rn(lo, lo), __fmul_rn(hi, hi));
}
__device__ __forceinline__ unsigned grn_pair(unsigned u, float inv) {
    // inv = scale / norm, one IEEE division per pooled position.  The reference rounds scale * (x / norm) twice in
    // float32; x * inv differs from it by at most ~1.5 float32 ulp, far below the bf16 rounding (2^-9) of the operand.
    const float a = __fmul_rn(__uint_as_float(u << 16), inv);
    const float b = __fmul_rn(__uint_as_float(u & 0xffff0000u), inv);
    const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned *>(&r);
}

template <typename Ops, int NV, bool GRN = false>
__global__ void __launch_bounds__(256)
roi_pool_nhwc_kernel(const uint4 *__restrict__ feat, int n_img, int H, int W, int L,
                     const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                     int PH, int PW, float scale, uint4 *__restrict__ out, int out_ld = 0, int out_off = 0,
                     float grn_scale = 1.f) {
    pdl_enter();
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
        const int ld = GRN ? out_ld : L;
        uint4 *orow = out + (size_t)it * PW * ld + (GRN ? out_off : 0);
        for (int pw = warp; pw < PW; pw += nwarps) {
            const int ws = s_ws[buf][pw], we = s_we[buf][pw];
            const bool empty = b < 0 || he <= hs || we <= ws;
            // GRN: every lane runs exactly one pass (the warp-wide sum needs all 32 lanes; loads/stores stay guarded)
            const int vend = GRN ? 32 : L;
            for (int v0 = lane; v0 < vend; v0 += 32 * NV) {
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
                if (GRN) {
                    float ss = 0.f;
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        if (v0 + 32 * j < L)
                            ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(sq_pair(acc[j].x), sq_pair(acc[j].y)),
                                                         __fadd_rn(sq_pair(acc[j].z), sq_pair(acc[j].w))));
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, d));
                    const float inv = __fdiv_rn(grn_scale, sqrtf(ss));
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        acc[j].x = grn_pair(acc[j].x, inv); acc[j].y = grn_pair(acc[j].y, inv);
                        acc[j].z = grn_pair(acc[j].z, inv); acc[j].w = grn_pair(acc[j].w, inv);
                    }
                }
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    if (v0 + 32 * j < L) st_stream(orow + (size_t)pw * ld + v0 + 32 * j, acc[j]);
            }
        }
    }
}

// ---- (1') direct kernel, one CTA per ROI ---------------------------------------------------------------------------
// Same arithmetic as roi_pool_nhwc_kernel (the same roi_geom / bin_bounds / Ops::take in the same h-major, w-minor order
// per bin: the same bits), different mapping: a CTA owns a whole ROI, warp k pools bin row k -- its PW bins one after the
// other.  In kernel (1) the bin rows of a ROI are items of their own and land on different SMs; neighbouring bin rows
// overlap by a map row (a bin spans floor .. ceil of its bounds: 1.4 - 1.75 x the ROI's rows in total for the small ROIs
// of the deep search levels) and every one of those rows comes through L2 again.  Here the seven bin rows run side by
// side on ONE SM: a row shared by two of them is an L1 hit (or merges with the miss in flight) for the second, as the
// column shared by two neighbouring bins of a row already is in kernel (1).  One geometry pass and one barrier per ROI
// instead of per bin row.
template <typename Ops, int NV>
__global__ void __launch_bounds__(256)
roi_pool_nhwc_roi_kernel(const uint4 *__restrict__ feat, int n_img, int H, int W, int L,
                         const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                         int PH, int PW, float scale, uint4 *__restrict__ out) {
    pdl_enter();
    __shared__ int s_ws[2][POOL_MAX_PW], s_we[2][POOL_MAX_PW], s_hs[2][POOL_MAX_PW], s_he[2][POOL_MAX_PW], s_b[2];
    const int R = n_rois ? min(*n_rois, R_cap) : R_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t row_stride = (size_t)W * L;      // vectors per map row
    int buf = 0;
    for (int r = blockIdx.x; r < R; r += gridDim.x, buf ^= 1) {
        if ((int)threadIdx.x < PW + PH) {
            const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, PH, PW);
            int lo, hi;
            if ((int)threadIdx.x < PW) {
                bin_bounds(threadIdx.x, q.bin_w, q.start_w, W, lo, hi);
                s_ws[buf][threadIdx.x] = lo;
                s_we[buf][threadIdx.x] = hi;
                if (threadIdx.x == 0) s_b[buf] = (q.b < 0 || q.b >= n_img) ? -1 : q.b;
            } else {
                const int p = (int)threadIdx.x - PW;
                bin_bounds(p, q.bin_h, q.start_h, H, lo, hi);
                s_hs[buf][p] = lo;
                s_he[buf][p] = hi;
            }
        }
        __syncthreads();                          // double-buffered geometry: one barrier per ROI
        const int b = s_b[buf];
        const uint4 *base = feat + (size_t)(b < 0 ? 0 : b) * H * row_stride;
        for (int ph = warp; ph < PH; ph += nwarps) {
            const int hs = s_hs[buf][ph], he = s_he[buf][ph];
            uint4 *orow = out + ((size_t)r * PH + ph) * PW * L;
            for (int pw = 0; pw < PW; ++pw) {
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
}

// ---- (1b) staged kernel: the map slice lives in shared memory ---------------------------------
// With many ROIs per image the L2 -> SM traffic of kernel (1) (every bin re-reads its window, ~4 map
// cells per output vector) is what bounds it, not HBM.  Here one CTA stages a channel slice of ONE
// image -- all H*W cells x SV 16-byte vectors, <= 200 KB -- in shared memory with cp.async, then pools a
// whole chunk of that image's ROIs out of it: HBM/L2 see the slice once per CTA and the pooled rows
// once.  Work item = (image bucket, ROI chunk, slice); slices of one chunk are adjacent items so the 16-
// or 32-byte pieces of a pooled row are written at about the same time and merge in L2.
// ROI geometry is computed once per ROI per CTA (PH + PW threads per ROI, packed lo|hi<<16 bounds in
// shared memory, double-buffered so that one barrier per batch of ROIs suffices).
// MODE 0: out [R, P, P, C] (NHWC).  MODE 1: out [R, C, P, P] f32 (Caffe blob order), MODE 2: + argmax;
// both go through a shared [roi][c][bin] tile so that the global writes are contiguous runs of CC*P*P.
constexpr int ST_THREADS = 1024;
constexpr int ST_P = 7;                      // pooled size this kernel is specialised for
constexpr int ST_BINS = ST_P * ST_P;
constexpr int ST_RB_MAX = 64;                // ROIs per geometry batch (MODE 0)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

template <typename Ops, int SV, int MODE>
__global__ void __launch_bounds__(ST_THREADS, 1)
roi_pool_staged_kernel(const uint4 *__restrict__ feat, int n_img, int H, int W, int L,
                       const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                       const int32_t *__restrict__ bucket_off, const int32_t *__restrict__ perm,
                       float scale, void *__restrict__ out_v, int32_t *__restrict__ argmax,
                       int n_buckets, int nchunk, int nslices, int RB) {
    extern __shared__ uint4 s_dyn[];
    uint4 *s_map = s_dyn;                                   // [H*W][SV]
    __shared__ uint32_t s_hb[2][ST_RB_MAX][8], s_wb[2][ST_RB_MAX][8];
    __shared__ int s_r[2][ST_RB_MAX];
    constexpr int CC = SV * 4;                              // f32 channels per slice (MODE 1/2 are f32 only)
    float *s_tile = reinterpret_cast<float *>(s_map + (size_t)H * W * SV);       // MODE >= 1: [RB][CC][49]
    int *s_atile = reinterpret_cast<int *>(s_tile + (size_t)RB * CC * ST_BINS);  // MODE 2
    const int R = n_rois ? min(max(*n_rois, 0), R_cap) : R_cap;
    const int cells = H * W;
    const long n_items = (long)n_buckets * nchunk * nslices;
    int buf = 0;
    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int slice = (int)(item % nslices);
        const int chunk = (int)((item / nslices) % nchunk);
        const int bucket = (int)(item / ((long)nslices * nchunk));
        const int base = bucket_off ? bucket_off[bucket] : 0;
        const int cnt = bucket_off ? bucket_off[bucket + 1] - base : R;
        const int lo = base + (int)((long)cnt * chunk / nchunk), hi = base + (int)((long)cnt * (chunk + 1) / nchunk);
        if (hi <= lo) continue;
        const bool real = bucket < n_img;                   // the last bucket collects ROIs with a bad batch index
        const int c0v = slice * SV;
        if (real) {
            const uint4 *src = feat + (size_t)bucket * cells * L + c0v;
            const int n_vec = cells * SV;
            const int rot = (int)(((long)blockIdx.x * cells) / gridDim.x) * SV;     // see roi_pool_keys_kernel
            for (int i0 = threadIdx.x; i0 < n_vec; i0 += ST_THREADS) {
                const int i = i0 + rot < n_vec ? i0 + rot : i0 + rot - n_vec;
                const int cell = i / SV, j = i - cell * SV;
                if (c0v + j < L) cp_async16(s_map + i, src + (size_t)cell * L + j);
            }
        }
        cp_async_wait_all();
        __syncthreads();
        for (int r0 = lo; r0 < hi; r0 += RB, buf ^= 1) {
            const int nb = min(RB, hi - r0);
            if ((int)threadIdx.x < nb * 2 * ST_P) {
                const int rl = threadIdx.x / (2 * ST_P), k = threadIdx.x - rl * 2 * ST_P;
                const int r = perm ? perm[r0 + rl] : r0 + rl;
                const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, ST_P, ST_P);
                const bool ok = real && q.b == bucket;      // single-image call without buckets: b != 0 -> zeros
                int a = 0, b = 0;
                if (k < ST_P) {
                    if (ok) bin_bounds(k, q.bin_h, q.start_h, H, a, b);
                    s_hb[buf][rl][k] = (uint32_t)a | ((uint32_t)b << 16);
                } else {
                    if (ok) bin_bounds(k - ST_P, q.bin_w, q.start_w, W, a, b);
                    s_wb[buf][rl][k - ST_P] = (uint32_t)a | ((uint32_t)b << 16);
                }
                if (k == 0) s_r[buf][rl] = r;
            }
            __syncthreads();
            const int total = nb * ST_BINS * SV;
            for (int qi = threadIdx.x; qi < total; qi += ST_THREADS) {
                const int rl = qi / (ST_BINS * SV), rem = qi - rl * (ST_BINS * SV);
                const int bin = rem / SV, j = rem - bin * SV;
                const int ph = bin / ST_P, pw = bin - ph * ST_P;
                const uint32_t hb = s_hb[buf][rl][ph], wb = s_wb[buf][rl][pw];
                const int hs = hb & 0xffff, he = hb >> 16, ws = wb & 0xffff, we = wb >> 16;
                const bool empty = he <= hs || we <= ws;
                uint4 acc = empty ? make_uint4(0u, 0u, 0u, 0u) : Ops::lowest();
                int4 am = make_int4(-1, -1, -1, -1);
                if (!empty && c0v + j < L) {
                    for (int h = hs; h < he; ++h) {
                        const uint4 *p = s_map + ((size_t)h * W + ws) * SV + j;
#pragma unroll 2
                        for (int w = ws; w < we; ++w, p += SV) {
                            const uint4 v = *p;
                            if (MODE == 2) {
                                const int idx = h * W + w;
                                if (__uint_as_float(v.x) > __uint_as_float(acc.x)) { acc.x = v.x; am.x = idx; }
                                if (__uint_as_float(v.y) > __uint_as_float(acc.y)) { acc.y = v.y; am.y = idx; }
                                if (__uint_as_float(v.z) > __uint_as_float(acc.z)) { acc.z = v.z; am.z = idx; }
                                if (__uint_as_float(v.w) > __uint_as_float(acc.w)) { acc.w = v.w; am.w = idx; }
                            } else {
                                Ops::take(acc, v);
                            }
                        }
                    }
                }
                if (MODE == 0) {
                    if (c0v + j < L)
                        st_stream(reinterpret_cast<uint4 *>(out_v) + ((size_t)s_r[buf][rl] * ST_BINS + bin) * L + c0v + j, acc);
                } else {
                    float *t = s_tile + ((size_t)rl * CC + 4 * j) * ST_BINS + bin;
                    t[0] = __uint_as_float(acc.x); t[ST_BINS] = __uint_as_float(acc.y);
                    t[2 * ST_BINS] = __uint_as_float(acc.z); t[3 * ST_BINS] = __uint_as_float(acc.w);
                    if (MODE == 2) {
                        int *u = s_atile + ((size_t)rl * CC + 4 * j) * ST_BINS + bin;
                        u[0] = am.x; u[ST_BINS] = am.y; u[2 * ST_BINS] = am.z; u[3 * ST_BINS] = am.w;
                    }
                }
            }
            if (MODE != 0) {
                __syncthreads();
                const int C = L * 4;                         // f32 channels of the map
                const int cc = min(CC, C - slice * CC);
                const int run = cc * ST_BINS;                // contiguous floats per ROI of this slice
                for (int o = threadIdx.x; o < nb * run; o += ST_THREADS) {
                    const int rl = o / run, e = o - rl * run;
                    const size_t g = ((size_t)s_r[buf][rl] * C + (size_t)slice * CC) * ST_BINS + e;
                    reinterpret_cast<float *>(out_v)[g] = s_tile[(size_t)rl * CC * ST_BINS + e];
                    if (MODE == 2) argmax[g] = s_atile[(size_t)rl * CC * ST_BINS + e];
                }
                __syncthreads();                             // tile free for the next batch
            }
        }
        __syncthreads();                                     // every thread is done with the slice before it is replaced
    }
}

// ---- (1c) staged kernel, NHWC out, order-preserving integer keys --------------------------------
// Kernel (1b) is bound by instruction issue (compare + select per 16-bit pair, per-output index math), not
// by memory.  This variant maps every staged value ONCE to a signed integer key whose integer order is the
// float order (negative values: magnitude bits flipped), so that the pooling loop is VIMNMX3 -- a 3-input
// packed max, two map cells per instruction -- and maps the winner back.  NaN never wins in the reference
// (`v > best` is false): NaNs get the key of the initial value (-FLT_MAX / bf16 -inf).  The one thing a pure
// max cannot reproduce is the reference's first-wins rule between +0 and -0 when the maximum is zero, so a
// CTA whose slice contains a -0 anywhere pools that slice with the exact compare-select of (1b) instead.
// One warp owns one ROI at a time (geometry in registers, handed out by shuffles: no block barriers while
// pooling); its lanes span 32/SV bins x SV channel vectors per pass.
template <bool BF16> struct KeyOps;
template <> struct KeyOps<true> {
    typedef OpsBF16 Exact;
    static __device__ __forceinline__ unsigned flipmask(unsigned x) {     // 0x7fff in every negative half
        unsigned m;
        asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(x));           // sign-replicate bytes 1 and 3
        return m & 0x7fff7fffu;
    }
    static __device__ __forceinline__ bool neg_zero(unsigned x) { return __vcmpeq2(x, 0x80008000u) != 0u; }
    // a half with the sign bit set or a NaN pattern: anything that is not an ordinary value >= +0
    static __device__ __forceinline__ bool not_plain(unsigned x) { return ((x & 0x80008000u) | __vcmpgtu2(x & 0x7fff7fffu, 0x7f807f80u)) != 0u; }
    static __device__ __forceinline__ unsigned to_key(unsigned x) {
        const unsigned k = x ^ flipmask(x);
        const unsigned nan = __vcmpgtu2(x & 0x7fff7fffu, 0x7f807f80u);    // 0xffff in every NaN half
        return (k & ~nan) | (0x807f807fu & nan);                          // NaN -> key(-inf)
    }
    static __device__ __forceinline__ unsigned from_key(unsigned k) { return k ^ flipmask(k); }
    static __device__ __forceinline__ unsigned lowest() { return 0x807f807fu; }
    static __device__ __forceinline__ unsigned max3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
};
template <> struct KeyOps<false> {
    typedef OpsF32 Exact;
    static __device__ __forceinline__ bool neg_zero(unsigned x) { return x == 0x80000000u; }
    static __device__ __forceinline__ bool not_plain(unsigned x) { return x > 0x7f800000u; }        // sign bit set, or NaN
    static __device__ __forceinline__ unsigned to_key(unsigned x) {
        if ((x & 0x7fffffffu) > 0x7f800000u) return 0x80800000u;         // NaN -> key(-FLT_MAX)
        return x ^ (((unsigned)((int)x >> 31)) & 0x7fffffffu);
    }
    static __device__ __forceinline__ unsigned from_key(unsigned k) { return k ^ (((unsigned)((int)k >> 31)) & 0x7fffffffu); }
    static __device__ __forceinline__ unsigned lowest() { return 0x80800000u; }
    static __device__ __forceinline__ unsigned max3(unsigned a, unsigned b, unsigned c) { return (unsigned)__vimax3_s32((int)a, (int)b, (int)c); }
};

// Pooling of one ROI's bin rows with a FIXED number of map rows per column (variant 2 of the keys kernel).
// The generic loop nest of roi_pool_keys_kernel spends ~8 instructions of control flow per shared-memory load (the
// per-lane row count nh makes the row loop a real loop with a tail).  Here MH = the warp-wide maximum bin height
// is a template parameter and a lane whose bin is shorter re-reads its last row (max is idempotent), so that a
// column reduce is straight-line code: MH loads at precomputed row offsets and (MH+1)/2 packed max instructions.
// Column re-use between neighbouring bins is kept (see the generic path).
#ifndef AZN_POOL_PW_UNROLL
#define AZN_POOL_PW_UNROLL 7          // bins of a bin row per loop body of pool_bins_fixed (A/B: 1 = not unrolled)
#endif
constexpr int POOL_PW_UNROLL = AZN_POOL_PW_UNROLL;
template <typename K, int MH, bool PLAIN>
__device__ __forceinline__ void pool_bins_fixed(const uint4 *__restrict__ rowbase, int nh, int rot, int row_step, int sv, unsigned gb,
                                                bool phv, bool valid, uint4 *__restrict__ optr, int L) {
#define AZN_MX2(D, A) \
    (D).x = K::max3((D).x, (A).x, (A).x); (D).y = K::max3((D).y, (A).y, (A).y); (D).z = K::max3((D).z, (A).z, (A).z); (D).w = K::max3((D).w, (A).w, (A).w)
#define AZN_MX3(D, A, B) \
    (D).x = K::max3((D).x, (A).x, (B).x); (D).y = K::max3((D).y, (A).y, (B).y); (D).z = K::max3((D).z, (A).z, (B).z); (D).w = K::max3((D).w, (A).w, (B).w)
    // Row order of this lane: row (t + rot) mod nh at step t (any order will do: max is commutative, and a lane whose
    // bin is shorter than MH simply cycles).  `rot` = 1 for the second bin row of a quarter-warp when its first map row
    // lies an even distance from its partner's: the two 64-byte pieces of every LDS.128 then land in opposite bank halves
    // instead of the same one (17.8 M of 47.8 M shared-load wavefronts were such conflicts at R = 20 000).
    int roff[MH];
    const int nhe = max(nh, 1);
    int idx = rot < nhe ? rot : 0;
#pragma unroll
    for (int t = 0; t < MH; ++t) {
        roff[t] = idx * row_step;
        idx = idx + 1 < nhe ? idx + 1 : 0;
    }
    const unsigned a0 = PLAIN ? 0u : K::lowest();
    int w_cached = -1;                                       // warp-uniform
    uint4 cache = make_uint4(a0, a0, a0, a0);
#pragma unroll POOL_PW_UNROLL
    for (int pw = 0; pw < ST_P; ++pw) {
        const unsigned wb = __shfl_sync(0xffffffffu, gb, ST_P + pw);
        const int ws = wb & 0xffff, we = (int)(wb >> 16);    // warp-uniform
        uint4 acc = make_uint4(a0, a0, a0, a0);
        if (we > ws) {
            int w = ws;
            if (w == w_cached) { acc = cache; ++w; }         // the previous bin's last column, already reduced
#pragma unroll 1
            for (; w < we; ++w) {
                const uint4 *q = rowbase + w * sv;
                uint4 c = q[roff[0]];
                if (MH == 2) { const uint4 b = q[roff[1]]; AZN_MX2(c, b); }
                if (MH >= 3) { const uint4 b = q[roff[1]], d = q[roff[2]]; AZN_MX3(c, b, d); }
                if (MH == 4) { const uint4 b = q[roff[3]]; AZN_MX2(c, b); }
                if (MH >= 5) { const uint4 b = q[roff[3]], d = q[roff[4]]; AZN_MX3(c, b, d); }
                if (MH == 6) { const uint4 b = q[roff[5]]; AZN_MX2(c, b); }
                AZN_MX2(acc, c);
                cache = c;
            }
            w_cached = we - 1;
        }
        uint4 res = make_uint4(0u, 0u, 0u, 0u);
        if (we > ws && valid) res = PLAIN ? acc : make_uint4(K::from_key(acc.x), K::from_key(acc.y), K::from_key(acc.z), K::from_key(acc.w));
        if (phv) st_stream(optr + (size_t)pw * L, res);
    }
#undef AZN_MX2
#undef AZN_MX3
}

// Variant 3 of the keys kernel: fixed-height column reduces over a TWO-LEVEL map.  Behind the slice the CTA keeps a
// second table, pair[p][w] = max(map[2p][w], map[2p+1][w]) (half the slice again: 38x63 cells x 64 bytes + 19x63 x 64
// = 229 824 of the 232 448 bytes a CTA may have), and a bin's rows [hs, he) are covered by at most one odd single row
// in front, the aligned pairs inside and one even single row behind: 1 / 2 / 2 / 3 / 3 / 4 shared-memory loads per
// (bin row, map column) for bins of 1 .. 6 rows instead of 1 .. 6 -- max is associative and idempotent, so the result
// is the same bit pattern.  The big ROIs are where the loads are (a tenth of the microbench's boxes, half of the
// reads), and they are the ones with 4-6 rows per bin.  ME = the warp-wide maximum entry count; a lane with fewer
// entries cycles through its own (as in pool_bins_fixed).  `base` = the slice at this lane's channel vector,
// `pair_off` = offset of the pair table in uint4 units (a multiple of 2 cells: bank parity = row + column there too).
template <typename K, int ME, bool PLAIN>
__device__ __forceinline__ void pool_bins_pairs(const uint4 *__restrict__ base, int hs, int nh, int rot, int row_step, int pair_off, int sv,
                                                unsigned gb, bool phv, bool valid, uint4 *__restrict__ optr, int L) {
#define AZN_MX2(D, A) \
    (D).x = K::max3((D).x, (A).x, (A).x); (D).y = K::max3((D).y, (A).y, (A).y); (D).z = K::max3((D).z, (A).z, (A).z); (D).w = K::max3((D).w, (A).w, (A).w)
#define AZN_MX3(D, A, B) \
    (D).x = K::max3((D).x, (A).x, (B).x); (D).y = K::max3((D).y, (A).y, (B).y); (D).z = K::max3((D).z, (A).z, (B).z); (D).w = K::max3((D).w, (A).w, (B).w)
    const int he = hs + max(nh, 1);
    const int odd_s = hs & 1, odd_e = he & 1;
    const int np = (he - odd_e - hs - odd_s) >> 1;           // aligned pairs inside (>= 0 for nh >= 1)
    const int ne = odd_s + np + odd_e;                       // entries of this lane: 1 .. ME
    const int off_s = hs * row_step, off_e = (he - 1) * row_step;
    const int off_p = pair_off + (((hs + odd_s) >> 1) - odd_s) * row_step;     // + idx * row_step for entry idx in [odd_s, odd_s + np)
    int eoff[ME];
    int idx = rot < ne ? rot : 0;
#pragma unroll
    for (int t = 0; t < ME; ++t) {
        eoff[t] = idx < odd_s ? off_s : (idx < odd_s + np ? off_p + idx * row_step : off_e);
        idx = idx + 1 < ne ? idx + 1 : 0;
    }
    const unsigned a0 = PLAIN ? 0u : K::lowest();
    int w_cached = -1;                                       // warp-uniform
    uint4 cache = make_uint4(a0, a0, a0, a0);
#pragma unroll
    for (int pw = 0; pw < ST_P; ++pw) {
        const unsigned wb = __shfl_sync(0xffffffffu, gb, ST_P + pw);
        const int ws = wb & 0xffff, we = (int)(wb >> 16);    // warp-uniform
        uint4 acc = make_uint4(a0, a0, a0, a0);
        if (we > ws) {
            int w = ws;
            if (w == w_cached) { acc = cache; ++w; }         // the previous bin's last column, already reduced
#pragma unroll 1
            for (; w < we; ++w) {
                const uint4 *q = base + w * sv;
                uint4 c = q[eoff[0]];
                if (ME == 2) { const uint4 b = q[eoff[1]]; AZN_MX2(c, b); }
                if (ME >= 3) { const uint4 b = q[eoff[1]], d = q[eoff[2]]; AZN_MX3(c, b, d); }
                if (ME == 4) { const uint4 b = q[eoff[3]]; AZN_MX2(c, b); }
                AZN_MX2(acc, c);
                cache = c;
            }
            w_cached = we - 1;
        }
        uint4 res = make_uint4(0u, 0u, 0u, 0u);
        if (we > ws && valid) res = PLAIN ? acc : make_uint4(K::from_key(acc.x), K::from_key(acc.y), K::from_key(acc.z), K::from_key(acc.w));
        if (phv) st_stream(optr + (size_t)pw * L, res);
    }
#undef AZN_MX2
#undef AZN_MX3
}

// ---- (1d) grouped pooling: one warp = 32/SV ROIs of the SAME width in map cells -----------------------------------
// The per-ROI loop nest above spends ~630 warp instructions per ROI-slice of which ~140 are loads, maxima and stores:
// every lane recomputes the ROI geometry (16 times per ROI, once per slice), a warp walks ONE ROI with its lanes on the
// seven bin rows -- so every map row on a bin boundary is read by two lane groups -- and the column loop is a real
// loop per bin.  Here a pre-pass (roi_group_kernel) sorts the ROIs of an image by their width in cells and packs them
// into groups of G = 32/SV; a warp pools one group, lane group g = ROI g, and walks the seven bin rows one after the
// other.  What that buys:
//   * the bins' column ranges depend on the ROI width alone (floor(p*w/7), ceil((p+1)*w/7): a 64 x 7 table built once
//     per CTA with the reference's float formulas), so the column sweep of a group is ONE warp-uniform program:
//     load column, fold, and where a bin ends store it and restart the accumulator from nothing or from that column
//     (consecutive bins overlap by at most one column) -- every column is reduced once, no per-bin loop, no geometry;
//   * the bins' row ranges come from the same table indexed by the ROI height (per lane); the tallest bin of the
//     group fixes the template parameter MH as before, shorter bins cycle through their rows;
//   * the pre-pass orders a class by height and alternates the parity of (start_h + start_w): the two ROIs that share a
//     quarter-warp of an LDS.128 then read cells of opposite parity, i.e. opposite bank halves (odd row pitch).
// ROIs the tables do not cover (clipped by the map border, wider or taller than 64 cells, bad batch index) stay on the
// per-ROI path through the old bucket list.
constexpr int GP_MAXDIM = 64;            // ROI sides, in cells, of the grouped path
constexpr int GP_MAX_IMG = 6;            // images per call (the pre-pass keeps a 33 KB histogram per image in shared memory)
constexpr int GP_CLS = GP_MAXDIM;        // width classes per image
constexpr int GP_PITCH = 2 * GP_MAXDIM + 1;   // histogram cells per class: 2 parities x 64 heights (+1: bank spread)

// floor(p * bin) | ceil((p + 1) * bin) << 8 of a ROI side of rs cells, before the start offset and the clamp
__device__ __forceinline__ unsigned gp_bin_rel(int rs, int p) {
    const float bin = __fdiv_rn((float)rs, (float)ST_P);
    const int lo = (int)floorf(__fmul_rn((float)p, bin));
    const int hi = (int)ceilf(__fmul_rn((float)(p + 1), bin));
    return (unsigned)lo | ((unsigned)hi << 8);
}
__device__ __forceinline__ int gp_extent(int rs) { return (int)(gp_bin_rel(rs, ST_P - 1) >> 8); }   // cells a side of rs really spans

struct GpTables {
    unsigned short rel[GP_MAXDIM][8];                       // [rs - 1][p]: lo | hi << 8
    // [rs - 1]: the column program of a side of rs cells, 4 bits per column: bins that end at this column (0..7) |
    // 8 if the bin after them starts at this very column (consecutive bins overlap by at most one column; bins that
    // end together with a later one necessarily start at it)
    unsigned long long code[GP_MAXDIM][4];
};

__device__ __forceinline__ void gp_build_tables(GpTables &t) {
    for (int i = threadIdx.x; i < GP_MAXDIM * 8; i += blockDim.x)
        t.rel[i / 8][i % 8] = (unsigned short)(i % 8 < ST_P ? gp_bin_rel(i / 8 + 1, i % 8) : 0u);
    __syncthreads();
    if (threadIdx.x < GP_MAXDIM) {
        const int rs = threadIdx.x + 1;
        const unsigned short *r = t.rel[rs - 1];
        const int ncol = r[ST_P - 1] >> 8;
        unsigned long long word[4] = {0ull, 0ull, 0ull, 0ull};
        int p = 0;
        for (int w = 0; w < ncol && w < 64; ++w) {
            unsigned n = 0, carry = 0;
            while (p < ST_P && (r[p] >> 8) - 1 == w) {
                carry = (p + 1 < ST_P && (r[p + 1] & 0xff) == w) ? 8u : 0u;
                ++n;
                ++p;
            }
            word[w >> 4] |= (unsigned long long)(n | carry) << (4 * (w & 15));
        }
        for (int k = 0; k < 4; ++k) t.code[rs - 1][k] = word[k];
    }
    __syncthreads();
}

__device__ __forceinline__ uint4 lds128(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned mov_fence(unsigned x) {   // a copy the compiler cannot fold away
    unsigned y;
    asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// One bin row of a group: `arow` = shared byte address of (first map row of the lane's bin, first column of its ROI,
// its channel vector), nh rows of row_bytes, MH = the group's tallest bin in this bin row (0: any height, row loop).
// The loop is latency-bound unless loads run ahead (32 warps per SM, a handful of instructions per load): the next
// column's rows are requested as soon as the current column is reduced, before it is folded into the bins, and a
// stored accumulator is copied first, so that the restart of the accumulator does not wait for the store to drain.
template <typename K, int SV, int MH, bool PLAIN>
__device__ __forceinline__ void pool_group_row(const GpTables &t, unsigned arow, int nh, int mh, unsigned row_bytes, int rw,
                                               uint4 *__restrict__ optr, int L, bool store) {
#define AZN_MX2(D, A) \
    (D).x = K::max3((D).x, (A).x, (A).x); (D).y = K::max3((D).y, (A).y, (A).y); (D).z = K::max3((D).z, (A).z, (A).z); (D).w = K::max3((D).w, (A).w, (A).w)
#define AZN_MX3(D, A, B) \
    (D).x = K::max3((D).x, (A).x, (B).x); (D).y = K::max3((D).y, (A).y, (B).y); (D).z = K::max3((D).z, (A).z, (B).z); (D).w = K::max3((D).w, (A).w, (B).w)
    constexpr unsigned COL = SV * 16;
    // row r of the lane at step k: k mod nh (max is idempotent: a lane whose bin is shorter cycles through its rows)
    unsigned ra0 = arow, ra1 = arow, ra2 = arow, ra3 = arow, ra4 = arow, ra5 = arow;
    if (MH >= 2) {
        int idx = 1 < nh ? 1 : 0;
        ra1 = arow + (unsigned)idx * row_bytes; idx = idx + 1 < nh ? idx + 1 : 0;
        if (MH >= 3) { ra2 = arow + (unsigned)idx * row_bytes; idx = idx + 1 < nh ? idx + 1 : 0; }
        if (MH >= 4) { ra3 = arow + (unsigned)idx * row_bytes; idx = idx + 1 < nh ? idx + 1 : 0; }
        if (MH >= 5) { ra4 = arow + (unsigned)idx * row_bytes; idx = idx + 1 < nh ? idx + 1 : 0; }
        if (MH >= 6) { ra5 = arow + (unsigned)idx * row_bytes; }
    }
    const unsigned a0 = PLAIN ? 0u : K::lowest();
    uint4 v0, v1, v2, v3, v4, v5;
    v1 = v2 = v3 = v4 = v5 = make_uint4(a0, a0, a0, a0);
#define AZN_LOAD_COL()                                                                            \
    do {                                                                                          \
        if (MH == 0) {       /* tall bins: a row loop, rows past the lane's own bin re-read its last */ \
            unsigned a = ra0;                                                                     \
            v0 = lds128(a);                                                                       \
            for (int k = 1; k < mh; ++k) {                                                        \
                if (k < nh) a += row_bytes;                                                       \
                const uint4 b = lds128(a);                                                        \
                AZN_MX2(v0, b);                                                                   \
            }                                                                                     \
            ra0 += COL;                                                                           \
        } else {                                                                                  \
            v0 = lds128(ra0); ra0 += COL;                                                         \
            if (MH >= 2) { v1 = lds128(ra1); ra1 += COL; }                                        \
            if (MH >= 3) { v2 = lds128(ra2); ra2 += COL; }                                        \
            if (MH >= 4) { v3 = lds128(ra3); ra3 += COL; }                                        \
            if (MH >= 5) { v4 = lds128(ra4); ra4 += COL; }                                        \
            if (MH >= 6) { v5 = lds128(ra5); ra5 += COL; }                                        \
        }                                                                                         \
    } while (0)
    uint4 acc = make_uint4(a0, a0, a0, a0);
#define AZN_EMIT()                                                                                \
    do {                                                                                          \
        uint4 r;                                                                                  \
        if (PLAIN) r = make_uint4(mov_fence(acc.x), mov_fence(acc.y), mov_fence(acc.z), mov_fence(acc.w)); \
        else r = make_uint4(K::from_key(acc.x), K::from_key(acc.y), K::from_key(acc.z), K::from_key(acc.w)); \
        if (store) st_stream(optr, r);                                                            \
        optr += L;                                                                                \
    } while (0)
    const unsigned long long *cw = t.code[rw - 1];
    const int ncol = t.rel[rw - 1][ST_P - 1] >> 8;
    unsigned long long word = cw[0];
    int w = 0;
    AZN_LOAD_COL();
#pragma unroll 1
    for (;;) {
        uint4 c = v0;                                        // the column's maximum over the lane's rows
        if (MH == 2) { AZN_MX2(c, v1); }
        if (MH >= 3) { AZN_MX3(c, v1, v2); }
        if (MH == 4) { AZN_MX2(c, v3); }
        if (MH >= 5) { AZN_MX3(c, v3, v4); }
        if (MH == 6) { AZN_MX2(c, v5); }
        ++w;
        if (w < ncol) AZN_LOAD_COL();                        // in flight while this column is folded and stored
        AZN_MX2(acc, c);
        const unsigned code = (unsigned)word & 15u;
        word >>= 4;
        unsigned n = code & 7u;                              // bins that end at this column
        if (n) {
#pragma unroll 1
            for (; n > 1u; --n) { AZN_EMIT(); acc = c; }
            AZN_EMIT();
            acc = (code & 8u) ? c : make_uint4(a0, a0, a0, a0);
        }
        if (w >= ncol) break;
        if ((w & 15) == 0) word = cw[w >> 4];
    }
#undef AZN_EMIT
#undef AZN_LOAD_COL
#undef AZN_MX2
#undef AZN_MX3
}

// ---- (1e) row bands: 128-byte slices for maps whose full height does not fit shared memory at SV = 8 ------------------
// A slice of SV = 8 vectors (64 bf16 / 32 f32 channels = one 128-byte line per cell) pools 15 % faster than SV = 4 on the
// same map (30x50: 0.69 vs 0.60 of the HBM peak): every output piece is a full line (no reliance on L2 merging the 64-byte
// halves written by two CTAs at different times) and the per-ROI geometry is paid half as often.  The default-cfg 38x63 map
// needs 306 KB at SV = 8, so the map is cut into `nb` overlapping bands of `rows` rows (27 for 63 columns); a pre-pass
// (roi_band_kernel) sends every ROI to the band that holds all of its rows -- bucket = image * nb + band -- and the ROIs
// that fit no band (taller than a band minus the band step: ~8 % of the microbench's boxes) to a second launch of the
// SV = 4 kernel over the whole map.  Buckets differ in size, so the pre-pass also cuts them into chunks of about equal
// ROI counts (the chunk table) and the CTAs draw (chunk, slice) items from a device-side counter.
// MEASURED (R = 20 000, 38x63, bf16): 0.351 ms banded vs 0.272 ms for the plain SV = 4 kernel (f32: 0.593 vs 0.519; R = 2000:
// 0.077 vs 0.045).  The ROIs no band holds are the tall ones -- a tenth of the boxes but about half of the shared-memory
// reads -- so the launch that keeps the 64-byte slices still carries half of the work, now behind the banded launch
// instead of beside it, and the pre-pass + second launch cost ~30 us.  Bit-exact, kept behind azn_roi_pool_tune(322).
constexpr int BAND_MAX = 8;                 // bands per image
constexpr int BAND_TAB_MAX = 1024;          // chunk table entries
constexpr int BAND_MAX_BUCKETS = 2048;      // n_img * nb (the pre-pass scans the buckets serially)
struct BandPlan {
    const int4 *ctab;                       // [ctrl[1]]: (bucket, first, last + 1 in perm, -); nullptr: no bands
    int32_t *ctrl;                          // [0] item counter (zeroed by the pre-pass), [1] chunks in the table
    int nb, rows;
};
__host__ __device__ __forceinline__ int band_row0(int k, int nb, int H, int rows) {
    return nb > 1 ? (int)(((long)k * (H - rows)) / (nb - 1)) : 0;
}

// EXT selects the experimental paths at compile time -- 0: none (the default kernel carries none of their code: with the
// first two compiled in, the 64-register build spilled 80 bytes and ran 3 % slower), 1: width-grouped ROIs (1d), 2: row bands
// (1e), 3: the two-level map (pool_bins_pairs), 4: bin bounds from the records of the geometry pre-pass.  3 and 4 lived in
// the default instantiation for a while: no spills to speak of (8 bytes at SV = 8), but the same-box A/B of three builds
// (tools/runs/gpu_r2bo.sh, profiles/r2_pool_regression_ab.txt) showed the default path 5 % (38x63) to 22 % (30x50, SV = 8)
// slower at R = 20 000 with their code present -- the 64-register schedule of the hot loops is that fragile.
template <bool BF16, int SV, int EXT = 0>
__global__ void __launch_bounds__(ST_THREADS, 1)
roi_pool_keys_kernel(const uint4 *__restrict__ feat, int n_img, int H, int W, int L,
                     const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap,
                     const int32_t *__restrict__ bucket_off, const int32_t *__restrict__ perm,
                     float scale, uint4 *__restrict__ out, int n_buckets, int nchunk, int nslices, int variant,
                     const int32_t *__restrict__ goff, const uint2 *__restrict__ gdesc, BandPlan bp,
                     const uint4 *__restrict__ geom = nullptr) {
    // geom != nullptr: the ROIs' bin bounds come from the records of roi_geom_kernel (the launch before this one; this
    // kernel is then its programmatic dependent: the slice is staged while the pre-pass runs).
    // gdesc != nullptr: the grouped path (1d).  goff[b] .. goff[b + 1] = image b's groups in gdesc (G slots of
    // (ROI index or -1, start_w | start_h << 8 | roi_h << 16 | roi_w << 24) each); bucket_off / perm then list only
    // the ROIs the grouped path does not cover.  The kernel is launched as a programmatic dependent of the pre-pass.
    typedef KeyOps<BF16> K;
    typedef typename K::Exact Exact;
    extern __shared__ uint4 s_dyn[];
    uint4 *s_map = s_dyn;                                   // [H*W][SV]
    __shared__ int s_negzero;
    __shared__ int s_next, s_gnext;
    __shared__ GpTables s_gp;
    constexpr int PHG = 32 / SV;                            // bin rows a warp covers at once
    constexpr int G = 32 / SV;                              // ROIs per group (1d)
    constexpr int NG = (ST_P + PHG - 1) / PHG;              // groups of bin rows per ROI
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane % SV, bsub = lane / SV;
    const int R = n_rois ? min(max(*n_rois, 0), R_cap) : R_cap;
    const int cells = H * W;
    const bool banded = EXT == 2 && bp.ctab != nullptr;
    const int SH = banded ? bp.rows : H;                    // map rows a CTA stages
    __shared__ long s_item;
    const long n_items = banded ? (long)bp.ctrl[1] * nslices : (long)n_buckets * nchunk * nslices;
    // Shared-memory row pitch in cells: ODD.  A quarter-warp of an LDS.128 serves two bin rows (2 x 4 vectors of 64
    // bytes); the two 64-byte pieces share their banks exactly when their cell indices have the same parity, and with an
    // even width the parity of cell (h, w) does not depend on h at all: every pair of distinct rows conflicts (the
    // 30x50 map).  With an odd pitch rows alternate, and the row rotation of pool_bins_fixed takes care of the rest.
    const int Wp = W | 1;
    const int cells_p = SH * Wp;
    const int row_step = Wp * SV;
    // every CTA starts its sweep over the map at a different cell: CTAs that all walk the same cells in
    // the same order queue up on the same few L2 slices (measured: ~35 us per staging instead of ~3)
    auto stage_slice = [&](int img, int row0, int c0v) {
        const uint4 *src = feat + ((size_t)img * cells + (size_t)row0 * W) * L + c0v;
        const int n_vec = cells_p * SV;
        const int rot = (int)(((long)blockIdx.x * cells_p) / gridDim.x) * SV;
        for (int i0 = threadIdx.x; i0 < n_vec; i0 += ST_THREADS) {
            const int i = i0 + rot < n_vec ? i0 + rot : i0 + rot - n_vec;
            const int cell = i / SV, jj = i - cell * SV;
            const int h = cell / Wp, w = cell - h * Wp;
            if (w < W && c0v + jj < L) cp_async16(s_map + i, src + (size_t)(h * W + w) * L + jj);
            else s_map[i] = make_uint4(0u, 0u, 0u, 0u);              // channel padding of the last slice / the pitch column
        }
    };
    long prestaged = -1;
    if (EXT == 1 && gdesc) {
        // The map does not depend on the pre-pass: build the tables and stage the first item's slice while it runs.
        gp_build_tables(s_gp);
        const long item = blockIdx.x;
        const int bucket = (int)(item / ((long)nslices * nchunk));
        if (item < n_items && bucket < n_img) { stage_slice(bucket, 0, (int)(item % nslices) * SV); prestaged = item; }
        pdl_grid_wait();
    }
    // the next item: dealt round-robin, or (banded: items of unequal size) drawn from the device-side counter
    auto next_item = [&](long item) -> long {
        if (!banded) return item + gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) s_item = (long)gridDim.x + atomicAdd(&bp.ctrl[0], 1);
        __syncthreads();
        return s_item;
    };
    for (long item = blockIdx.x; item < n_items; item = next_item(item)) {
        const int slice = (int)(item % nslices);
        int chunk, bucket, lo, hi;
        if (banded) {
            const int4 e = bp.ctab[item / nslices];
            chunk = 0; bucket = e.x; lo = e.y; hi = e.z;
        } else {
            chunk = (int)((item / nslices) % nchunk);
            bucket = (int)(item / ((long)nslices * nchunk));
            const int base = bucket_off ? bucket_off[bucket] : 0;
            const int cnt = bucket_off ? bucket_off[bucket + 1] - base : R;
            lo = base + (int)((long)cnt * chunk / nchunk);
            hi = base + (int)((long)cnt * (chunk + 1) / nchunk);
        }
        const bool real = bucket < (banded ? n_img * bp.nb : n_img);
        const int img = banded ? bucket / bp.nb : bucket;                       // image of a real bucket
        const int row0 = banded && real ? band_row0(bucket % bp.nb, bp.nb, H, SH) : 0;   // first map row in shared memory
        // this chunk's share of the image's groups: every nchunk-th one, so that all chunks get the same mix of sizes
        const int g_lo = EXT == 1 && gdesc && real ? goff[bucket] : 0, g_hi = EXT == 1 && gdesc && real ? goff[bucket + 1] : 0;
        const int n_grp = g_hi - g_lo > chunk ? (g_hi - g_lo - chunk + nchunk - 1) / nchunk : 0;
        if (hi <= lo && n_grp == 0) {
            if (prestaged == item) { cp_async_wait_all(); __syncthreads(); }     // nobody may still be writing the slice
            continue;
        }
        const int c0v = slice * SV;
        if (threadIdx.x == 0) { s_negzero = 0; s_next = lo + ST_THREADS / 32; s_gnext = ST_THREADS / 32; }
        const bool jv = c0v + j < L;
#ifdef AZN_POOL_TRACE
        const long long t0 = clock64();
#endif
        if (real && prestaged != item) stage_slice(img, row0, c0v);
        if (EXT == 4 && geom != nullptr) pdl_grid_wait();    // the records are complete from here on (a no-op after the first time)
        cp_async_wait_all();
        __syncthreads();
#ifdef AZN_POOL_TRACE
        const long long t2 = clock64();
#endif
        // A slice of ordinary values >= +0 (every post-ReLU map: the conv5_3 of this path) needs no keys at all: the
        // bit patterns of non-negative floats already order like signed integers, so the slice is pooled as staged,
        // from an accumulator of +0, with no transform before and none after.
        bool plain = false;
        if (real) {
            bool np = false;
            for (int i = threadIdx.x; i < cells_p * SV; i += ST_THREADS) {
                const uint4 v = s_map[i];
                np |= K::not_plain(v.x) | K::not_plain(v.y) | K::not_plain(v.z) | K::not_plain(v.w);
            }
            plain = __syncthreads_or(np ? 1 : 0) == 0;
        }
        if (real && !plain) {                                // raw bits -> keys in place, looking for a -0 on the way
            bool nz = false;
            for (int i = threadIdx.x; i < cells_p * SV; i += ST_THREADS) {
                uint4 v = s_map[i];
                nz |= K::neg_zero(v.x) | K::neg_zero(v.y) | K::neg_zero(v.z) | K::neg_zero(v.w);
                v.x = K::to_key(v.x); v.y = K::to_key(v.y); v.z = K::to_key(v.z); v.w = K::to_key(v.w);
                s_map[i] = v;
            }
            if (nz) s_negzero = 1;
        }
        __syncthreads();
        const bool exact = s_negzero != 0;
        if (exact) {                                         // rare: bring the raw slice back for the exact path
            const uint4 *src = feat + ((size_t)img * cells + (size_t)row0 * W) * L + c0v;
            for (int i = threadIdx.x; i < cells_p * SV; i += ST_THREADS) {
                const int cell = i / SV, jj = i - cell * SV;
                const int h = cell / Wp, w = cell - h * Wp;
                if (w < W && c0v + jj < L) cp_async16(s_map + i, src + (size_t)(h * W + w) * L + jj);
                else s_map[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            cp_async_wait_all();
            __syncthreads();
        }
        // variant 3: the pair table behind the slice (see pool_bins_pairs); rows 2p, 2p+1 of the staged (key) slice
        const bool pairs = EXT == 3 && (variant & 0x10000) != 0 && !exact && real;
        const int pair_off = ((cells_p + 1) & ~1) * SV;       // uint4 units
        if (pairs) {
            const int n_pv = (SH >> 1) * row_step;
            uint4 *s_pair = s_map + pair_off;
            for (int i = threadIdx.x; i < n_pv; i += ST_THREADS) {
                const int pr = i / row_step, rem = i - pr * row_step;
                const uint4 a = s_map[2 * pr * row_step + rem], b = s_map[(2 * pr + 1) * row_step + rem];
                s_pair[i] = make_uint4(K::max3(a.x, b.x, b.x), K::max3(a.y, b.y, b.y), K::max3(a.z, b.z, b.z), K::max3(a.w, b.w, b.w));
            }
            __syncthreads();
        }
#ifdef AZN_POOL_TRACE
        const long long t3 = clock64();
#endif
        // the per-ROI path: one warp, one ROI, lanes on the bin rows
        auto pool_one = [&](const int r, const unsigned gw) {
            unsigned gb = 0;                                 // lanes 0..6: packed h bounds of bin row `lane`; 7..13: w bounds
            int mh;                                          // the tallest bin of the ROI (warp-uniform): selects the fixed-height variant
            if (EXT == 4 && geom != nullptr) {               // lane k < 16 holds word k of the ROI's record
                const bool ok = real && (int)__shfl_sync(0xffffffffu, gw, 14) == img;
                if (ok && lane < 2 * ST_P) gb = gw;
                mh = ok ? (int)__shfl_sync(0xffffffffu, gw, 15) : 0;
            } else {
                const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, ST_P, ST_P);
                const bool ok = real && q.b == img;
                if (ok && lane < 2 * ST_P) {
                    int a, b;
                    if (lane < ST_P) bin_bounds(lane, q.bin_h, q.start_h, H, a, b);
                    else bin_bounds(lane - ST_P, q.bin_w, q.start_w, W, a, b);
                    gb = (unsigned)a | ((unsigned)b << 16);
                }
                mh = __reduce_max_sync(0xffffffffu, lane < ST_P ? (int)(gb >> 16) - (int)(gb & 0xffff) : 0);
            }
            // A pass pools one COLUMN of bins (fixed pw): every lane of the warp then has the same [ws, we), so the
            // inner loops are warp-uniform; lanes differ only in their bin row (ph = lane / SV) and channel vector.
#pragma unroll 1
            for (int rg = 0; rg < NG; ++rg) {
                const int ph = rg * PHG + bsub;
                const bool phv = ph < ST_P && jv;
                const unsigned hb = __shfl_sync(0xffffffffu, gb, min(ph, ST_P - 1));
                const int hs = hb & 0xffff, nh = phv ? (int)(hb >> 16) - hs : 0;
                uint4 *optr = out + ((size_t)r * ST_BINS + (size_t)min(ph, ST_P - 1) * ST_P) * L + c0v + j;
                if (pairs && mh <= 6) {                              // warp-uniform: two-level map, <= 4 loads per (bin row, column)
                    const int nh_rows = (int)(hb >> 16) - hs;
                    const int hs_c = min(max(hs - row0, 0), SH - 1), nh_c = max(min(nh_rows, SH - hs_c), 1);
                    const int he_c = hs_c + nh_c;
                    const int ne_l = (hs_c & 1) + (he_c & 1) + ((he_c - (he_c & 1) - hs_c - (hs_c & 1)) >> 1);
                    const int me = __reduce_max_sync(0xffffffffu, ph < ST_P ? ne_l : 1);
                    const int rot = 0;
#define AZN_PAIR(MEE)                                                                                                     \
    do {                                                                                                                  \
        if (plain) pool_bins_pairs<K, MEE, true>(s_map + j, hs_c, nh_c, rot, row_step, pair_off, SV, gb, phv, nh > 0, optr, L);   \
        else pool_bins_pairs<K, MEE, false>(s_map + j, hs_c, nh_c, rot, row_step, pair_off, SV, gb, phv, nh > 0, optr, L);        \
    } while (0)
                    switch (me) {
                    case 0: case 1: AZN_PAIR(1); break;
                    case 2: AZN_PAIR(2); break;
                    case 3: AZN_PAIR(3); break;
                    default: AZN_PAIR(4); break;
                    }
#undef AZN_PAIR
                    continue;
                }
                if ((variant & 255) == 2 && !exact && mh <= 6) {     // warp-uniform: fixed-height straight-line column reduces
                    const uint4 *rb = s_map + (size_t)min(max(hs - row0, 0), SH - 1) * row_step + j;
                    // partner = the other bin row of this lane's quarter-warp (lanes 4k..4k+7 hold bin rows 2k', 2k'+1
                    // when SV == 4; with SV == 8 a quarter-warp is one bin row and there is nothing to rotate)
                    // The idle lane group of the last pass (bin row 7 of 8) mirrors bin row 6 -- same rows, same
                    // order: a broadcast, not a second set of wavefronts.
                    const int nh_rows = (int)(hb >> 16) - hs;
                    int rot = 0;
                    if (SV <= 4) {
                        const int hs_partner = __shfl_xor_sync(0xffffffffu, hs, SV);
                        rot = (SV == 4 && (bsub & 1) && ph < ST_P && nh_rows >= 2 && hs_partner != hs && ((hs_partner - hs) & 1) == 0) ? 1 : 0;
                    }
#define AZN_FIX(MHH)                                                                                   \
    do {                                                                                               \
        if (plain) pool_bins_fixed<K, MHH, true>(rb, nh_rows, rot, row_step, SV, gb, phv, nh > 0, optr, L);     \
        else pool_bins_fixed<K, MHH, false>(rb, nh_rows, rot, row_step, SV, gb, phv, nh > 0, optr, L);          \
    } while (0)
                    switch (mh) {
                    case 0: case 1: AZN_FIX(1); break;
                    case 2: AZN_FIX(2); break;
                    case 3: AZN_FIX(3); break;
                    case 4: AZN_FIX(4); break;
                    case 5: AZN_FIX(5); break;
                    default: AZN_FIX(6); break;
                    }
#undef AZN_FIX
                    continue;
                }
                const uint4 *rowbase = s_map + (size_t)min(max(hs - row0, 0), SH - 1) * row_step + j;   // (a bin with rows lies inside the band)
                // Column re-use.  Consecutive bins of a row overlap by at most one map column -- bin p+1 starts at
                // floor((p+1) b) >= ceil((p+1) b) - 1, the last column of bin p, and clamping keeps that order; for ROIs
                // narrower than 7 cells several bins even share all their columns -- so the maximum over the lane's bin rows
                // of the LAST column visited is kept in registers and re-used by the next bin: every (bin row, map column) pair
                // of the ROI is read from shared memory exactly once (roi_w + 1 columns instead of roi_w + 7: 40 % fewer reads
                // for the average ROI, 3x fewer for small ones).  New columns are reduced two at a time: four independent
                // shared-memory loads in flight per lane.
                const unsigned a0 = plain ? 0u : K::lowest();
                int w_cached = -1;                           // warp-uniform
                uint4 cache = make_uint4(a0, a0, a0, a0);
#pragma unroll 1
                for (int pw = 0; pw < ST_P; ++pw, optr += L) {
                    const unsigned wb = __shfl_sync(0xffffffffu, gb, ST_P + pw);
                    const int ws = wb & 0xffff, we = (int)(wb >> 16);                // warp-uniform
                    uint4 res = make_uint4(0u, 0u, 0u, 0u);
                    if (we > ws) {
                        if (!exact) {
                            uint4 acc = make_uint4(a0, a0, a0, a0);
#define AZN_MAX3(D, A, B)                                                                        \
    (D).x = K::max3((D).x, (A).x, (B).x); (D).y = K::max3((D).y, (A).y, (B).y);                  \
    (D).z = K::max3((D).z, (A).z, (B).z); (D).w = K::max3((D).w, (A).w, (B).w)
                            int w = ws;
                            if (w == w_cached) { acc = cache; ++w; }         // the previous bin's last column, already reduced
#pragma unroll 1
                            for (; w + 1 < we; w += 2) {
                                uint4 c0 = make_uint4(a0, a0, a0, a0), c1 = c0;
                                const uint4 *q = rowbase + w * SV;
                                int t = 0;
#pragma unroll 1
                                for (; t + 1 < nh; t += 2, q += 2 * row_step) {
                                    const uint4 a = q[0], b = q[SV], c = q[row_step], d = q[row_step + SV];
                                    AZN_MAX3(c0, a, c); AZN_MAX3(c1, b, d);
                                }
                                if (t < nh) { const uint4 a = q[0], b = q[SV]; AZN_MAX3(c0, a, a); AZN_MAX3(c1, b, b); }
                                AZN_MAX3(acc, c0, c1);
                                cache = c1;
                            }
                            if (w < we) {
                                uint4 c0 = make_uint4(a0, a0, a0, a0);
                                const uint4 *q = rowbase + w * SV;
                                int t = 0;
#pragma unroll 1
                                for (; t + 1 < nh; t += 2, q += 2 * row_step) { const uint4 a = q[0], b = q[row_step]; AZN_MAX3(c0, a, b); }
                                if (t < nh) { const uint4 a = q[0]; AZN_MAX3(c0, a, a); }
                                AZN_MAX3(acc, c0, c0);
                                cache = c0;
                            }
                            w_cached = we - 1;
#undef AZN_MAX3
                            if (nh > 0) res = plain ? acc : make_uint4(K::from_key(acc.x), K::from_key(acc.y), K::from_key(acc.z), K::from_key(acc.w));
                        } else if (nh > 0) {
                            const int nw = we - ws;
                            const uint4 *p = rowbase + ws * SV;
                            res = Exact::lowest();
#pragma unroll 1
                            for (int t = 0; t < nh; ++t, p += row_step)
#pragma unroll 1
                                for (int w = 0; w < nw; ++w) Exact::take(res, p[w * SV]);
                        }
                    }
                    if (phv) st_stream(optr, res);
                }
            }
        };
        // (1d) the image's groups, largest ROIs first (the pre-pass sorts by ascending width; a big group handed out last
        // would be the item's tail)
        if (EXT == 1 && n_grp > 0) {
            // Work unit = one bin row of one group, handed out in order: the seven bin rows of a group are pooled by seven
            // warps at about the same time, and -- more important -- the CTAs of the other slices reach the same unit within
            // a few microseconds.  A pooled row of C channels is assembled in L2 from the 64-byte pieces of nslices CTAs; with
            // whole groups as units (7 x longer) the CTAs drifted apart by more than the L2 residency of a line, the halves
            // of a 128-byte line went to DRAM separately, and the kernel ran at the rate of scattered 64-byte writes
            // (2.6 TB/s, tools/probes/tlb_probe.cu) instead of the streaming rate.
            const int g = lane / SV;
            const unsigned smap_a = (unsigned)__cvta_generic_to_shared(s_map);
            const unsigned row_bytes = (unsigned)row_step * 16u;
            const int n_units = n_grp * ST_P;
            // the next unit's index and descriptor are fetched while the current one is pooled
            int u = warp;
            uint2 e = make_uint2(0u, 0u);
            if (u < n_units) e = gdesc[(size_t)(g_hi - 1 - (chunk + (u / ST_P) * nchunk)) * G + g];
            while (u < n_units) {
                int u2 = 0;
                if (lane == 0) u2 = atomicAdd(&s_gnext, 1);
                u2 = __shfl_sync(0xffffffffu, u2, 0);
                uint2 e2 = make_uint2(0u, 0u);
                if (u2 < n_units) e2 = gdesc[(size_t)(g_hi - 1 - (chunk + (u2 / ST_P) * nchunk)) * G + g];
                const int ph = u % ST_P;
                if (!exact) {
                    const int dbg = variant >> 8;
                    const int idx = (int)e.x;
                    const int sw = e.y & 0xff, sh = (e.y >> 8) & 0xff, rh = (e.y >> 16) & 0xff;
                    const int rw = __shfl_sync(0xffffffffu, (int)(e.y >> 24), 0);        // the same for the whole group
                    const bool store = idx >= 0 && jv && dbg != 1;
                    const unsigned hb = s_gp.rel[rh - 1][ph];
                    const int hl = hb & 0xff, nh = (int)(hb >> 8) - hl;
                    const int mh = __reduce_max_sync(0xffffffffu, nh);
                    const unsigned arow = smap_a + (unsigned)(((sh + hl) * Wp + sw) * SV + j) * 16u;
                    uint4 *optr = out + ((size_t)max(idx, 0) * ST_BINS + (size_t)ph * ST_P) * L + c0v + j;
#define AZN_GRP(MHH)                                                                                      \
    do {                                                                                                  \
        if (plain) pool_group_row<K, SV, MHH, true>(s_gp, arow, nh, mh, row_bytes, rw, optr, L, store);    \
        else pool_group_row<K, SV, MHH, false>(s_gp, arow, nh, mh, row_bytes, rw, optr, L, store);         \
    } while (0)
                    switch (mh) {
                    case 1: AZN_GRP(1); break;
                    case 2: AZN_GRP(2); break;
                    case 3: AZN_GRP(3); break;
                    case 4: AZN_GRP(4); break;
                    case 5: AZN_GRP(5); break;
                    case 6: AZN_GRP(6); break;
                    default: AZN_GRP(0); break;
                    }
#undef AZN_GRP
                } else if (ph == 0) {                        // a slice with a -0: the exact per-ROI path, ROI by ROI
                    for (int gg = 0; gg < G; ++gg) {
                        const int idx = __shfl_sync(0xffffffffu, (int)e.x, gg * SV);
                        if (idx >= 0) pool_one(idx, 0u);
                    }
                }
                u = u2;
                e = e2;
            }
        }
        // ROIs are handed out dynamically (shared cursor): their cost varies by an order of magnitude with
        // their size, and a static deal leaves a tail of one big ROI per item.
        // (with the geometry pre-pass: the NEXT ROI's record is fetched while the current one is pooled)
        int ri = lo + warp;
        if (EXT == 4 && geom != nullptr) {
            auto fetch = [&](const int at) -> unsigned {
                if (at >= hi || lane >= 16) return 0u;
                return reinterpret_cast<const unsigned *>(geom)[(size_t)(perm ? perm[at] : at) * 16 + lane];
            };
            unsigned gw = fetch(ri);
            while (ri < hi) {
                int nxt = 0;
                if (lane == 0) nxt = atomicAdd(&s_next, 1);
                nxt = __shfl_sync(0xffffffffu, nxt, 0);
                const unsigned gw_next = fetch(nxt);
                pool_one(perm ? perm[ri] : ri, gw);
                ri = nxt;
                gw = gw_next;
            }
        } else {
            while (ri < hi) {
                pool_one(perm ? perm[ri] : ri, 0u);
                int nxt = 0;
                if (lane == 0) nxt = atomicAdd(&s_next, 1);
                ri = __shfl_sync(0xffffffffu, nxt, 0);
            }
        }
#ifdef AZN_POOL_TRACE
        const long long t4 = clock64();
#endif
        __syncthreads();                                     // every warp is done with the slice before it is replaced
#ifdef AZN_POOL_TRACE
        if ((blockIdx.x == 0 || blockIdx.x == 77) && (threadIdx.x == 0 || threadIdx.x == 1023))
            printf("pool trace cta %d t %d item %ld rois %d: stage+detect %lld transform %lld pool(own) %lld tail %lld cycles\n", blockIdx.x, threadIdx.x, item,
                   hi - lo, t2 - t0, t3 - t2, t4 - t3, clock64() - t4);
#endif
    }
}

// ROIs grouped by image for the staged kernel: bucket b in [0, n_img) = batch index b, bucket n_img = bad
// index.  One CTA: shared-memory histogram, scan, scatter (order inside a bucket is irrelevant: every ROI's
// output row depends on that ROI alone).  off[n_img + 2], perm[R].
constexpr int BUCKET_MAX_IMG = 8190;
__global__ void __launch_bounds__(1024)
roi_bucket_kernel(const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap, int n_img,
                  int32_t *__restrict__ off, int32_t *__restrict__ perm) {
    extern __shared__ int s_cnt[];                           // [n_img + 2] counts -> offsets, then cursors
    __shared__ int s_part[1024];
    const int R = n_rois ? min(max(*n_rois, 0), R_cap) : R_cap;
    const int nb = n_img + 1;
    for (int i = threadIdx.x; i < nb + 1; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int b = (int)rois[(size_t)r * 5];
        atomicAdd(&s_cnt[(b < 0 || b >= n_img) ? n_img : b], 1);
    }
    __syncthreads();
    // exclusive scan of s_cnt[0..nb): each thread owns a run of `per` consecutive buckets
    const int per = (nb + blockDim.x - 1) / blockDim.x;
    const int b0 = min((int)threadIdx.x * per, nb), b1 = min(b0 + per, nb);
    int sum = 0;
    for (int b = b0; b < b1; ++b) sum += s_cnt[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {                                  // warp 0 scans the 1024 partial sums, 32 per lane
        int run = 0;
        for (int k = 0; k < 32; ++k) run += s_part[threadIdx.x * 32 + k];
        const int incl = warp_incl_scan(run, threadIdx.x);
        int pre = incl - run;
        for (int k = 0; k < 32; ++k) { const int t = s_part[threadIdx.x * 32 + k]; s_part[threadIdx.x * 32 + k] = pre; pre += t; }
    }
    __syncthreads();
    int pre = s_part[threadIdx.x];
    for (int b = b0; b < b1; ++b) { const int t = s_cnt[b]; s_cnt[b] = pre; off[b] = pre; pre += t; }
    if (threadIdx.x == 0) off[nb] = R;
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int b = (int)rois[(size_t)r * 5];
        perm[atomicAdd(&s_cnt[(b < 0 || b >= n_img) ? n_img : b], 1)] = r;
    }
}

// Geometry pre-pass of the default keys kernel.  The pooling kernel meets every ROI once per channel slice (16 x for
// C = 512 bf16), and deriving the ROI's bin bounds -- four roundf, two IEEE divisions, 14 x floorf / ceilf / clamp --
// was a quarter of its warp instructions (135 of ~630 per ROI-slice, F2I alone 8 % of the stall samples).  Here they are
// computed ONCE per ROI, by the very same device functions (roi_geom, bin_bounds), into a 64-byte record:
//   words 0..6 packed row bounds of the bin rows (lo | hi << 16), 7..13 packed column bounds, 14 batch index
//   (0x7fffffff when it is not an integer in [0, n_img)), 15 the tallest bin of the ROI in rows.
// A warp of the pooling kernel fetches a record with one coalesced 64-byte load -- the NEXT ROI's while it pools the
// current one.
__global__ void __launch_bounds__(256)
roi_geom_kernel(const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap, int n_img, int H, int W, float scale,
                uint4 *__restrict__ geom) {
    const int R = n_rois ? min(max(*n_rois, 0), R_cap) : R_cap;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();                                 // the pooling kernel may start staging its slice
    if (r >= R) return;
    const RoiGeom q = roi_geom(rois + (size_t)r * 5, scale, ST_P, ST_P);
    unsigned wd[16];
    int mh = 0;
#pragma unroll
    for (int p = 0; p < ST_P; ++p) {
        int a, b;
        bin_bounds(p, q.bin_h, q.start_h, H, a, b);
        wd[p] = (unsigned)a | ((unsigned)b << 16);
        mh = max(mh, b - a);
        bin_bounds(p, q.bin_w, q.start_w, W, a, b);
        wd[ST_P + p] = (unsigned)a | ((unsigned)b << 16);
    }
    wd[14] = (q.b >= 0 && q.b < n_img) ? (unsigned)q.b : 0x7fffffffu;
    wd[15] = (unsigned)mh;
#pragma unroll
    for (int k = 0; k < 4; ++k) geom[(size_t)r * 4 + k] = make_uint4(wd[4 * k], wd[4 * k + 1], wd[4 * k + 2], wd[4 * k + 3]);
}

// Pre-pass of the banded path (1e).  One CTA.  Every ROI goes to set 1 -- bucket image * nb + band of the band that holds
// all of its rows (the one whose middle is nearest when several do), bucket n_img * nb for a bad batch index -- or, when no
// band holds it, to set 2 (bucket = image), the ROIs of the whole-map launch.  Counting sort into perm1 / perm2 (order
// inside a bucket is irrelevant), offsets off1[n_img * nb + 2] / off2[n_img + 2], and the chunk table of set 1: every
// non-empty bucket is cut into round(T * share) >= 1 chunks of equal ROI counts.
__device__ __forceinline__ int band_classify(const float *__restrict__ roi, float scale, int n_img, int H, int nb, int rows, int &bucket) {
    const RoiGeom q = roi_geom(roi, scale, ST_P, ST_P);
    if (q.b < 0 || q.b >= n_img) { bucket = n_img * nb; return 1; }
    int lo, hi, t;
    bin_bounds(0, q.bin_h, q.start_h, H, lo, t);              // bins are monotone: the ROI's rows are [lo of bin 0, hi of bin 6)
    bin_bounds(ST_P - 1, q.bin_h, q.start_h, H, t, hi);
    if (hi <= lo) { bucket = q.b * nb; return 1; }           // no rows at all (outside the map): any band will do
    int best = -1, best_d = 0;
    for (int k = 0; k < nb; ++k) {
        const int r0 = band_row0(k, nb, H, rows);
        if (lo < r0 || hi > r0 + rows) continue;
        const int d = abs(2 * r0 + rows - (lo + hi));
        if (best < 0 || d < best_d) { best = k; best_d = d; }
    }
    if (best < 0) { bucket = q.b; return 2; }
    bucket = q.b * nb + best;
    return 1;
}

__global__ void __launch_bounds__(1024)
roi_band_kernel(const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap, int n_img, int H, float scale,
                int nb, int rows, int T, int32_t *__restrict__ ctrl, int4 *__restrict__ ctab, int32_t *__restrict__ off1,
                int32_t *__restrict__ perm1, int32_t *__restrict__ off2, int32_t *__restrict__ perm2) {
    extern __shared__ int s_bc[];                            // [n1 + 1] set 1 | [n_img + 2] set 2: counts -> offsets -> cursors
    const int n1 = n_img * nb + 1;                           // buckets of set 1 (the last one: bad batch index)
    int *c1 = s_bc, *c2 = s_bc + n1 + 1;
    const int R = n_rois ? min(max(*n_rois, 0), R_cap) : R_cap;
    for (int i = threadIdx.x; i < n1 + 1 + n_img + 2; i += blockDim.x) s_bc[i] = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        int b;
        if (band_classify(rois + (size_t)r * 5, scale, n_img, H, nb, rows, b) == 1) atomicAdd(&c1[b], 1);
        else atomicAdd(&c2[b], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int b = 0; b < n1; ++b) total += c1[b];
        int run = 0, nch = 0;
        for (int b = 0; b < n1; ++b) {
            const int cnt = c1[b];
            off1[b] = run;
            c1[b] = run;
            if (cnt > 0) {
                int t = (int)(((long)T * cnt + total / 2) / total);
                t = max(1, min(min(t, cnt), BAND_TAB_MAX - nch - (n1 - 1 - b)));      // leave an entry for every later bucket
                for (int c = 0; c < t; ++c)
                    ctab[nch++] = make_int4(b, run + (int)((long)cnt * c / t), run + (int)((long)cnt * (c + 1) / t), 0);
            }
            run += cnt;
        }
        off1[n1] = run;
        ctrl[0] = 0;
        ctrl[1] = nch;
    }
    if (threadIdx.x == 32) {
        int run = 0;
        for (int b = 0; b <= n_img; ++b) { const int cnt = c2[b]; off2[b] = run; c2[b] = run; run += cnt; }
        off2[n_img + 1] = run;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        int b;
        if (band_classify(rois + (size_t)r * 5, scale, n_img, H, nb, rows, b) == 1) perm1[atomicAdd(&c1[b], 1)] = r;
        else perm2[atomicAdd(&c2[b], 1)] = r;
    }
}

// Pre-pass of the grouped path (1d).  One CTA: every ROI is classified -- regular (the tables cover it: sides of at
// most 64 cells, not clipped by the map border) or not -- and the regular ones are counting-sorted by (image, width,
// parity of start_h + start_w, height) into gdesc, each (image, width) class padded to a multiple of G slots; inside a
// class the two parity sequences (each ascending in height) are zipped, so that neighbouring slots -- the two ROIs of a
// quarter-warp -- have opposite parity and similar heights.  The rest goes to the per-ROI lists ioff / iperm (bucket
// n_img = bad batch index), in the format of roi_bucket_kernel.
__device__ __forceinline__ int gp_classify(const float *__restrict__ roi, float scale, int n_img, int H, int W, int &bucket,
                                           int &cell, unsigned &geom) {
    const int b = (int)roi[0];
    const int sw = (int)roundf(__fmul_rn(roi[1], scale)), sh = (int)roundf(__fmul_rn(roi[2], scale));
    const int ew = (int)roundf(__fmul_rn(roi[3], scale)), eh = (int)roundf(__fmul_rn(roi[4], scale));
    if (b < 0 || b >= n_img) { bucket = n_img; return 2; }
    bucket = b;
    const long rw = max((long)ew - sw + 1, 1L), rh = max((long)eh - sh + 1, 1L);
    if (rw > GP_MAXDIM || rh > GP_MAXDIM || sw < 0 || sh < 0) return 1;
    const int xw = gp_extent((int)rw), xh = gp_extent((int)rh);
    if (xw > GP_MAXDIM || xh > GP_MAXDIM || sw + xw > W || sh + xh > H) return 1;
    cell = (b * GP_CLS + (int)rw - 1) * GP_PITCH + ((sw + sh) & 1) * GP_MAXDIM + (int)rh - 1;
    geom = (unsigned)sw | ((unsigned)sh << 8) | ((unsigned)rh << 16) | ((unsigned)rw << 24);
    return 0;
}

__global__ void __launch_bounds__(1024)
roi_group_kernel(const float *__restrict__ rois, const int32_t *__restrict__ n_rois, int R_cap, int n_img, int H, int W,
                 float scale, int G, int32_t *__restrict__ goff, int32_t *__restrict__ ioff, int32_t *__restrict__ iperm,
                 uint2 *__restrict__ gdesc) {
    pdl_launch_dependents();                                 // the pooling kernel may start staging its first slice
    extern __shared__ int s_gh[];
    __shared__ int s_wtot[33];
    const int ncls = n_img * GP_CLS;
    int *hist = s_gh, *n0 = hist + ncls * GP_PITCH, *n1 = n0 + ncls, *cbase = n1 + ncls, *icnt = cbase + ncls;   // icnt[n_img + 2]
    const int R = n_rois ? min(max(*n_rois, 0), R_cap) : R_cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < ncls * GP_PITCH + 3 * ncls + n_img + 2; i += blockDim.x) s_gh[i] = 0;
    __syncthreads();
    for (int r = tid; r < R; r += blockDim.x) {
        int bucket, cell = 0;
        unsigned geom;
        if (gp_classify(rois + (size_t)r * 5, scale, n_img, H, W, bucket, cell, geom) == 0) atomicAdd(&hist[cell], 1);
        else atomicAdd(&icnt[bucket], 1);
    }
    __syncthreads();
    // per class: totals of the two parity sequences, counts -> exclusive offsets inside each sequence
    int slots = 0;
    if (tid < ncls) {
        int *h = hist + tid * GP_PITCH;
        int a = 0, b = 0;
        for (int k = 0; k < GP_MAXDIM; ++k) { const int t = h[k]; h[k] = a; a += t; }
        for (int k = 0; k < GP_MAXDIM; ++k) { const int t = h[GP_MAXDIM + k]; h[GP_MAXDIM + k] = b; b += t; }
        n0[tid] = a;
        n1[tid] = b;
        slots = (a + b + G - 1) / G * G;
    }
    // exclusive scan of the padded class sizes (ncls <= 1024: one class per thread)
    int incl = warp_incl_scan(slots, lane);
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int t = s_wtot[lane];
        const int ti = warp_incl_scan(t, lane);
        s_wtot[lane] = ti - t;
        if (lane == 31) s_wtot[32] = ti;
    }
    __syncthreads();
    const int my_base = s_wtot[warp] + incl - slots;
    if (tid < ncls) {
        cbase[tid] = my_base;
        if (tid % GP_CLS == 0) goff[tid / GP_CLS] = my_base / G;
        const unsigned pad = ((unsigned)(tid % GP_CLS + 1) << 24) | (1u << 16);      // a 1-row ROI of this width at cell (0, 0)
        for (int k = n0[tid] + n1[tid]; k < slots; ++k) gdesc[my_base + k] = make_uint2(0xffffffffu, pad);
    }
    if (tid == 0) {
        goff[n_img] = s_wtot[32] / G;
        int run = 0;
        for (int b = 0; b <= n_img; ++b) { const int t = icnt[b]; ioff[b] = run; icnt[b] = run; run += t; }
        ioff[n_img + 1] = run;
    }
    __syncthreads();
    for (int r = tid; r < R; r += blockDim.x) {
        int bucket, cell = 0;
        unsigned geom;
        if (gp_classify(rois + (size_t)r * 5, scale, n_img, H, W, bucket, cell, geom) == 0) {
            const int cls = cell / GP_PITCH, par = (cell - cls * GP_PITCH) / GP_MAXDIM;
            const int i = atomicAdd(&hist[cell], 1);         // rank inside the class's parity sequence
            const int m = min(n0[cls], n1[cls]);
            gdesc[cbase[cls] + (i < m ? 2 * i + par : m + i)] = make_uint2((unsigned)r, geom);
        } else {
            iperm[atomicAdd(&icnt[bucket], 1)] = r;
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

__device__ __forceinline__ float cvt_load(const float *p) { return *p; }
__device__ __forceinline__ float cvt_load(const __nv_bfloat16 *p) { return __bfloat162float(*p); }   // exact; the store rounds it back

template <typename TOut, typename TIn = float>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const TIn *__restrict__ src, int C, int HW, TOut *__restrict__ dst) {
    __shared__ float tile[32][33];
    const int img = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const TIn *s = src + (size_t)img * C * HW;
    TOut *d = dst + (size_t)img * C * HW;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? cvt_load(s + (size_t)c * HW + p) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int p = p0 + j, c = c0 + threadIdx.x;
        if (p < HW && c < C) cvt_store(d + (size_t)p * C + c, tile[threadIdx.x][j]);
    }
}

}  // namespace

// workspace = [transposed f32 map (NCHW f32 only)] [bucket offsets n_img + 2] [bucket permutation R_cap]
static size_t map_ws_bytes(int n_img, int C, int H, int W, int layout, int dtype) {
    if (layout == AZN_LAYOUT_NCHW && dtype == AZN_DTYPE_F32) return ((size_t)n_img * C * H * W * sizeof(float) + 255) / 256 * 256;
    return 0;
}
static size_t bucket_ws_bytes(int n_img, int R_cap) {
    if (n_img <= 1 || n_img > BUCKET_MAX_IMG) return 0;
    return ((size_t)(n_img + 2) + (size_t)(R_cap > 0 ? R_cap : 0)) * sizeof(int32_t);
}

// grouped path (NHWC): [goff n_img + 1] [ioff n_img + 2] [iperm R_cap] | [gdesc (R_cap + 7 pads per width class) x 8 bytes]
static size_t group_ints(int n_img, int R_cap) { return ((size_t)(2 * n_img + 3) + (size_t)(R_cap > 0 ? R_cap : 0) + 1) / 2 * 2; }
static size_t group_ws_bytes(int n_img, int R_cap) {
    if (n_img < 1 || n_img > GP_MAX_IMG) return 0;
    return group_ints(n_img, R_cap) * sizeof(int32_t) + ((size_t)(R_cap > 0 ? R_cap : 0) + (size_t)n_img * GP_CLS * 7) * sizeof(uint2);
}
// geometry records of the default staged path (roi_geom_kernel), at the same place as the grouped / banded workspaces
static size_t geom_ws_bytes(int R_cap) { return (size_t)(R_cap > 0 ? R_cap : 0) * 64; }
static size_t bucket_ws_aligned(int n_img, int R_cap) { return (bucket_ws_bytes(n_img, R_cap) + 15) / 16 * 16; }
// banded path (NHWC), at the same place as the grouped path's (the two exclude each other):
// [ctab BAND_TAB_MAX x int4] [ctrl 4] [off1 n_img * BAND_MAX + 2] [off2 n_img + 2] [perm1 R_cap] [perm2 R_cap]
static size_t band_ws_bytes(int n_img, int R_cap) {
    if (n_img < 1 || (long)n_img * BAND_MAX > BAND_MAX_BUCKETS) return 0;
    return (size_t)BAND_TAB_MAX * sizeof(int4) + ((size_t)4 + (size_t)n_img * BAND_MAX + 2 + n_img + 2 + 2 * (size_t)(R_cap > 0 ? R_cap : 0)) * sizeof(int32_t);
}

extern "C" size_t azn_roi_pool_workspace_bytes(int n_img, int C, int H, int W, int layout, int dtype, int R_cap) {
    if (layout == AZN_LAYOUT_NCHW && dtype != AZN_DTYPE_F32) return 0;
    if (layout == AZN_LAYOUT_NHWC)
        return bucket_ws_aligned(n_img, R_cap) + std::max(std::max(group_ws_bytes(n_img, R_cap), band_ws_bytes(n_img, R_cap)), geom_ws_bytes(R_cap));
    return map_ws_bytes(n_img, C, H, W, layout, dtype) + bucket_ws_bytes(n_img, R_cap);
}

namespace {
int g_pool_mode = 0;           // azn_roi_pool_tune: 0 automatic, 1 direct kernels only, 2 staged whenever possible
int g_pool_debug = 0;
int g_pool_variant = 2;        // keys kernel: 1 generic loop nest, 2 fixed-height column reduces (default), 3 grouped by ROI width (1d)
                               // when an image has enough ROIs, 4 grouped whenever the path applies (azn_roi_pool_tune(mode + 10 *
                               // variant)).  The grouped path is kept for A/B only: measured SLOWER than variant 2 at every size
                               // (profiles/README.md, negative results of round 2)

constexpr size_t ST_SMEM_BUDGET = 216 * 1024;      // dynamic shared memory the staged kernel may ask for
constexpr size_t ST_SMEM_MAX = 227 * 1024 - 256;   // ... the keys kernel with its pair table (EXT = 3; its static shared memory is < 64 bytes)
constexpr bool POOL_GEOM_AUTO = false;             // default route from 4096 ROIs: with (true) or without the geometry pre-pass

// Staged-kernel launch (see roi_pool_staged_kernel).  `nhwc` is the channels-last map; returns AZN_OK after a
// launch, or 1 when the configuration is outside the staged kernel's domain (caller takes the direct kernel).
template <typename Ops, int MODE>
int launch_staged(const void *nhwc, int n_img, int H, int W, int L, const float *rois, const int32_t *n_rois, int R_cap,
                  float scale, void *out, int32_t *argmax, int32_t *bucket_ws, cudaStream_t s, void *group_ws = nullptr,
                  size_t group_bytes = 0) {
    const size_t cells = (size_t)H * (MODE == 0 ? (W | 1) : W);     // the keys kernel pads its row pitch to an odd cell count
    int sv = 0, rb = ST_RB_MAX;
    for (int cand = (MODE == 0 ? 8 : 4); cand >= 2; cand >>= 1) {
        const size_t map_bytes = cells * cand * 16;
        if (map_bytes > ST_SMEM_BUDGET) continue;
        if (MODE == 0 && (g_pool_debug == 2 || g_pool_debug == 6) && cand == 8) continue;   // A/B: 64-byte slices where 128-byte ones would fit
        if (MODE != 0) {
            const size_t per = (size_t)cand * 4 * ST_BINS * 4 * (MODE == 2 ? 2 : 1);
            const size_t fit = (ST_SMEM_BUDGET - map_bytes) / per;
            if (fit < 4) continue;
            rb = (int)(fit < (size_t)ST_RB_MAX ? fit : ST_RB_MAX);
        }
        sv = cand;
        break;
    }
    if (sv == 0) return 1;
    const int sms = azn_num_sms();
    // (1e) row bands: the map does not fit at SV = 8, a band of >= 12 rows does, and the workspace for the pre-pass is there
    if (MODE == 0 && sv < 8 && L >= 8 && g_pool_variant == 2 && g_pool_debug == 3 && H <= 65535) {      // opt-in: measured slower (below)
        const int rows = (int)(ST_SMEM_BUDGET / ((size_t)(W | 1) * 8 * 16));
        int nb = rows >= 12 && rows < H ? 1 + (H - rows + std::max(1, rows / 4) - 1) / std::max(1, rows / 4) : 0;
        if (nb > BAND_MAX) nb = 0;                           // a map this tall against a band this short: most ROIs would fit no band
        const size_t need = band_ws_bytes(n_img, R_cap);
        if (nb >= 2 && (long)n_img * nb <= BAND_MAX_BUCKETS && (long)n_img * nb + 1 < BAND_TAB_MAX / 2 && group_ws && need > 0 &&
            group_bytes >= need && ((uintptr_t)group_ws % 16 == 0)) {
            int4 *ctab = (int4 *)group_ws;
            int32_t *ctrl = (int32_t *)(ctab + BAND_TAB_MAX);
            int32_t *off1 = ctrl + 4, *off2 = off1 + (size_t)n_img * BAND_MAX + 2;
            int32_t *perm1 = off2 + n_img + 2, *perm2 = perm1 + (R_cap > 0 ? R_cap : 0);
            const int ns8 = (L + 7) / 8;
            // chunks of the banded launch: the cost model below with all buckets together (the pre-pass shares them out)
            long T = 1;
            double best_cost = 0.0;
            for (int rounds = 1; rounds <= 4; ++rounds) {
                long c = (long)rounds * sms / ns8;
                if (c < 1) c = 1;
                const long it = c * ns8;
                const double cost = (double)((it + sms - 1) / sms) * (27.0 + (double)R_cap / (double)c);
                if (rounds == 1 || cost < best_cost * 0.97) { best_cost = cost; T = c; }
            }
            T = std::min<long>(T, BAND_TAB_MAX - ((long)n_img * nb + 1));
            const size_t psm = ((size_t)n_img * nb + 2 + n_img + 2) * sizeof(int);
            AZN_REQUIRE(psm <= 40 * 1024, "azn_roi_pool_fwd: too many images for the banded pre-pass");
            roi_band_kernel<<<1, 1024, psm, s>>>(rois, n_rois, R_cap, n_img, H, scale, nb, rows, (int)T, ctrl, ctab, off1, perm1, off2, perm2);
            AZN_LAUNCH_CHECK();
            constexpr bool kBf16 = sizeof(typename Ops::tag) == 2;
            static bool attr8 = false;
            if (!attr8) {
                AZN_CUDA(cudaFuncSetAttribute(roi_pool_keys_kernel<kBf16, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM_BUDGET));
                attr8 = true;
            }
            BandPlan bp;
            bp.ctab = ctab; bp.ctrl = ctrl; bp.nb = nb; bp.rows = rows;
            const size_t smem8 = (size_t)rows * (W | 1) * 8 * 16;
            roi_pool_keys_kernel<kBf16, 8, 2><<<sms, ST_THREADS, smem8, s>>>((const uint4 *)nhwc, n_img, H, W, L, rois, n_rois, R_cap, off1, perm1,
                                                                          scale, (uint4 *)out, n_img * nb + 1, 1, ns8, 2, nullptr, nullptr, bp);
            AZN_LAUNCH_CHECK();
            // the ROIs no band holds: the whole map at the slice width that fits, buckets = images (set 2 of the pre-pass);
            // their number is only known on the device -- chunks sized for an eighth of the ROIs
            const int nsl = (L + sv - 1) / sv;
            const long pairs2 = (long)n_img * nsl;
            long c2 = std::max<long>(1, std::min<long>(sms / pairs2, (long)(R_cap / 8 / n_img / 64)));
            const long items2 = (long)(n_img + 1) * c2 * nsl;
            const unsigned grid2 = (unsigned)std::min<long>(items2, sms);
            const size_t smem2 = cells * sv * 16;
            BandPlan none;
            none.ctab = nullptr; none.ctrl = nullptr; none.nb = 1; none.rows = H;
#define AZN_REST_LAUNCH(SVV)                                                                                             \
    do {                                                                                                                 \
        static bool attr_set = false;                                                                                    \
        if (!attr_set) {                                                                                                 \
            AZN_CUDA(cudaFuncSetAttribute(roi_pool_keys_kernel<kBf16, SVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ST_SMEM_BUDGET));                                                         \
            attr_set = true;                                                                                             \
        }                                                                                                                \
        roi_pool_keys_kernel<kBf16, SVV><<<grid2, ST_THREADS, smem2, s>>>(                                               \
            (const uint4 *)nhwc, n_img, H, W, L, rois, n_rois, R_cap, off2, perm2, scale, (uint4 *)out, n_img + 1,       \
            (int)c2, nsl, 2, nullptr, nullptr, none);                                                                    \
    } while (0)
            if (sv == 4) AZN_REST_LAUNCH(4); else AZN_REST_LAUNCH(2);
#undef AZN_REST_LAUNCH
            AZN_LAUNCH_CHECK();
            return AZN_OK;
        }
    }
    const int nslices = (L + sv - 1) / sv;
    int n_buckets = n_img > 1 ? n_img + 1 : 1;
    int32_t *off = nullptr, *perm = nullptr;
    // (1d) grouped by ROI width: needs the pre-pass workspace, few images, and enough ROIs per image to fill the classes
    int32_t *goff = nullptr;
    uint2 *gdesc = nullptr;
    const bool grouped = MODE == 0 && (sv == 4 || sv == 8) && g_pool_variant >= 3 && n_img <= GP_MAX_IMG && H <= 255 && W <= 255 &&
                         group_ws && group_bytes >= group_ws_bytes(n_img, R_cap) && ((uintptr_t)group_ws % 8 == 0) &&
                         (g_pool_variant == 4 || (long)R_cap >= 1024L * n_img);
    if (grouped) {
        goff = (int32_t *)group_ws;
        off = goff + n_img + 1;
        perm = off + n_img + 2;
        gdesc = (uint2 *)(goff + group_ints(n_img, R_cap));
        n_buckets = n_img + 1;
        const size_t sm = ((size_t)n_img * GP_CLS * (GP_PITCH + 3) + n_img + 2) * sizeof(int);
        static size_t attr_sm = 0;
        if (sm > attr_sm) {
            AZN_CUDA(cudaFuncSetAttribute(roi_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            attr_sm = sm;
        }
        roi_group_kernel<<<1, 1024, sm, s>>>(rois, n_rois, R_cap, n_img, H, W, scale, 32 / sv, goff, off, perm, gdesc);
        AZN_LAUNCH_CHECK();
    } else if (n_img > 1) {
        off = bucket_ws;
        perm = bucket_ws + n_img + 2;
        roi_bucket_kernel<<<1, 1024, (size_t)(n_img + 2) * sizeof(int), s>>>(rois, n_rois, R_cap, n_img, off, perm);
        AZN_LAUNCH_CHECK();
    }
    // (bucket, chunk, slice) items are dealt round-robin to one CTA per SM.  Cost model in units of one ROI-slice
    // of pooling (measured ~300 cycles): an item costs its staging (~27 units: load + key transform of the slice)
    // plus its ROIs; the launch costs rounds(items) x the item.  Pick the chunk count that minimises it.
    const long pairs = (long)(grouped ? n_img : n_buckets) * nslices;   // (the grouped path's extra bucket holds bad indices only)
    const double rois_per_bucket = (double)R_cap / (double)n_img;
    long nchunk = 1;
    double best_cost = 0.0;
    for (int rounds = 1; rounds <= 4; ++rounds) {
        long c = (long)rounds * sms / pairs;
        if (c < 1) c = 1;
        const long it = c * pairs;
        const double cost = (double)((it + sms - 1) / sms) * (27.0 + rois_per_bucket / (double)c);
        if (rounds == 1 || cost < best_cost * 0.97) { best_cost = cost; nchunk = c; }
    }
    const long items = (long)n_buckets * nchunk * nslices;
    const unsigned grid = (unsigned)(items < sms ? items : sms);
    BandPlan no_bands;
    no_bands.ctab = nullptr; no_bands.ctrl = nullptr; no_bands.nb = 1; no_bands.rows = H;
    size_t smem = cells * sv * 16;
    if (MODE != 0) smem += (size_t)rb * sv * 4 * ST_BINS * 4 * (MODE == 2 ? 2 : 1);
    // variant 3 (pool_bins_pairs): the pair table behind the slice, when slice + table fit a CTA's shared memory.
    // OPT-IN, azn_roi_pool_tune(422) (and (622): with 64-byte slices where 128-byte ones would fit).  MEASURED (R = 20 000,
    // bf16): 38x63 0.2805 ms vs 0.2751 without the table; 30x50 at SV = 4 0.234 vs 0.262 (and 0.244 for the default SV = 8
    // there; f32 0.459 vs 0.451).  The table does cut the ideal shared-load wavefronts by a quarter (30.1 M -> 22.5 M,
    // profiles/r2_ncu_pool_pairs.csv), but the two bin rows of a quarter-warp no longer walk rows of alternating bank
    // parity -- a single odd row, aligned pairs, a single even row -- so the row rotation that keeps the plain kernel
    // at 7.8 M conflict wavefronts has nothing to hold on to: 25.7 M with the table, data pipe 88 % instead of 78 %.
    const size_t smem_pairs = (((cells + 1) & ~(size_t)1) + (size_t)(H / 2) * (W | 1)) * sv * 16;
    const bool use_pairs = MODE == 0 && !grouped && g_pool_variant == 2 && (g_pool_debug == 4 || g_pool_debug == 6) && H >= 2 &&
                       smem_pairs <= ST_SMEM_MAX;
    if (use_pairs) smem = smem_pairs;
    // geometry pre-pass (roi_geom_kernel + the EXT = 4 instantiation): needs the 64 bytes per ROI of workspace.
    // azn_roi_pool_tune(9xx) switches it on, (8xx) off; POOL_GEOM_AUTO says what the default does from 4096 ROIs.
    // MEASURED while its code sat in the default instantiation (same box, bf16, 38x63): R = 20 000 0.2875 vs 0.2948 ms,
    // R = 8000 0.127 vs 0.131, R = 2000 0.047 vs 0.046 (the extra launch) -- a quarter fewer warp instructions bought 2.5 %
    // there, but that build as a whole was 5-22 % behind the one without the code (see the note on EXT above).
    uint4 *geom = nullptr;
    const bool want_geom = g_pool_debug == 9 || (POOL_GEOM_AUTO && g_pool_debug != 8);
    if (MODE == 0 && !grouped && !use_pairs && want_geom && group_ws && group_bytes >= geom_ws_bytes(R_cap) && R_cap >= 4096 &&
        ((uintptr_t)group_ws % 16 == 0)) {
        geom = (uint4 *)group_ws;
        roi_geom_kernel<<<(unsigned)((R_cap + 255) / 256), 256, 0, s>>>(rois, n_rois, R_cap, n_img, H, W, scale, geom);
        AZN_LAUNCH_CHECK();
    }
#define AZN_ST_LAUNCH(SVV)                                                                                               \
    do {                                                                                                                 \
        static bool attr_set = false;            /* once per instantiation: the call costs tens of microseconds */     \
        if (!attr_set) {                                                                                                 \
            AZN_CUDA(cudaFuncSetAttribute(roi_pool_staged_kernel<Ops, SVV, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ST_SMEM_BUDGET));                                                         \
            attr_set = true;                                                                                             \
        }                                                                                                                \
        roi_pool_staged_kernel<Ops, SVV, MODE><<<grid, ST_THREADS, smem, s>>>(                                           \
            (const uint4 *)nhwc, n_img, H, W, L, rois, n_rois, R_cap, off, perm, scale, out, argmax, n_buckets,          \
            (int)nchunk, nslices, rb);                                                                                   \
    } while (0)
#define AZN_KEY_LAUNCH(SVV)                                                                                              \
    do {                                                                                                                 \
        constexpr bool kBf16 = sizeof(typename Ops::tag) == 2;                                                           \
        static bool attr_set = false;                                                                                    \
        if (!attr_set) {                                                                                                 \
            AZN_CUDA(cudaFuncSetAttribute(roi_pool_keys_kernel<kBf16, SVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ST_SMEM_BUDGET));                                                         \
            AZN_CUDA(cudaFuncSetAttribute(roi_pool_keys_kernel<kBf16, SVV, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ST_SMEM_BUDGET));                                                         \
            AZN_CUDA(cudaFuncSetAttribute(roi_pool_keys_kernel<kBf16, SVV, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ST_SMEM_MAX));                                                            \
            AZN_CUDA(cudaFuncSetAttribute(roi_pool_keys_kernel<kBf16, SVV, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ST_SMEM_BUDGET));                                                         \
            attr_set = true;                                                                                             \
        }                                                                                                                \
        if (grouped)                                                                                                     \
            AZN_CUDA(azn_launch_pdl(roi_pool_keys_kernel<kBf16, SVV, 1>, dim3(grid), dim3(ST_THREADS), smem, s,          \
                                    (const uint4 *)nhwc, n_img, H, W, L, rois, n_rois, R_cap, off, perm, scale,          \
                                    (uint4 *)out, n_buckets, (int)nchunk, nslices, 2 + (g_pool_debug << 8), goff, gdesc, \
                                    no_bands, (const uint4 *)nullptr));                                                  \
        else if (geom)                                                                                                   \
            AZN_CUDA(azn_launch_pdl(roi_pool_keys_kernel<kBf16, SVV, 4>, dim3(grid), dim3(ST_THREADS), smem, s,          \
                                    (const uint4 *)nhwc, n_img, H, W, L, rois, n_rois, R_cap, off, perm, scale,          \
                                    (uint4 *)out, n_buckets, (int)nchunk, nslices,                                       \
                                    (g_pool_variant >= 2 ? 2 : 1), nullptr, nullptr, no_bands, (const uint4 *)geom));    \
        else if (use_pairs)                                                                                              \
            roi_pool_keys_kernel<kBf16, SVV, 3><<<grid, ST_THREADS, smem, s>>>(                                          \
                (const uint4 *)nhwc, n_img, H, W, L, rois, n_rois, R_cap, off, perm, scale, (uint4 *)out, n_buckets,     \
                (int)nchunk, nslices, (g_pool_variant >= 2 ? 2 : 1) | 0x10000, nullptr, nullptr, no_bands);              \
        else                                                                                                             \
            roi_pool_keys_kernel<kBf16, SVV><<<grid, ST_THREADS, smem, s>>>(                                             \
                (const uint4 *)nhwc, n_img, H, W, L, rois, n_rois, R_cap, off, perm, scale, (uint4 *)out, n_buckets,     \
                (int)nchunk, nslices, (g_pool_variant >= 2 ? 2 : 1), nullptr, nullptr, no_bands);                        \
    } while (0)
    if (MODE == 0) {
        if (sv == 8) AZN_KEY_LAUNCH(8); else if (sv == 4) AZN_KEY_LAUNCH(4); else AZN_KEY_LAUNCH(2);
    } else {
        if (sv == 8) return 1;
        else if (sv == 4) AZN_ST_LAUNCH(4);
        else AZN_ST_LAUNCH(2);
    }
#undef AZN_KEY_LAUNCH
#undef AZN_ST_LAUNCH
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

// The staged kernel pays one slice load per (image, chunk): worth it when an image has many ROIs.
bool want_staged(int mode, int n_img, int H, int W, int PH, int PW, const int32_t *n_rois, int R_cap, bool have_bucket_ws) {
    if (mode == 1 || mode == 3) return false;
    if (PH != ST_P || PW != ST_P || H > 65535 || W > 65535) return false;
    if (n_img > 1 && !have_bucket_ws) return false;
    if (mode == 2) return true;
    return n_rois == nullptr && (long)R_cap >= 64L * n_img;
}
}  // namespace

extern "C" void azn_roi_pool_tune(int mode) {
    g_pool_debug = mode / 100;
    mode %= 100;
    g_pool_mode = mode % 10;
    if (mode >= 10) g_pool_variant = mode / 10;          // 12 / 22: staged with the generic / fixed-height pooling loop
}

extern "C" int azn_roi_pool_fwd(const void *feat, int n_img, int C, int H, int W, int layout, int dtype,
                                const float *rois, const int32_t *n_rois, int R_cap, int PH, int PW,
                                float spatial_scale, void *out, int32_t *argmax, void *workspace,
                                size_t workspace_bytes, azn_stream_t stream) {
    return azn_roi_pool_fwd_ex(feat, n_img, C, H, W, layout, dtype, rois, n_rois, R_cap, PH, PW, spatial_scale, out, argmax,
                               workspace, workspace_bytes, g_pool_mode, stream);
}

extern "C" int azn_roi_pool_fwd_ex(const void *feat, int n_img, int C, int H, int W, int layout, int dtype,
                                   const float *rois, const int32_t *n_rois, int R_cap, int PH, int PW,
                                   float spatial_scale, void *out, int32_t *argmax, void *workspace,
                                   size_t workspace_bytes, int kernel_choice, azn_stream_t stream) {
    if (R_cap == 0) return AZN_OK;
    AZN_REQUIRE(kernel_choice >= 0 && kernel_choice <= 3, "azn_roi_pool_fwd_ex: kernel_choice must be 0 (auto), 1 (direct), 2 (staged) or 3 (direct, one CTA per ROI)");
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
        const bool bucket_ok = n_img <= 1 || (workspace && bucket_ws_bytes(n_img, R_cap) > 0 && workspace_bytes >= bucket_ws_bytes(n_img, R_cap));
        if (want_staged(kernel_choice, n_img, H, W, PH, PW, n_rois, R_cap, bucket_ok)) {
            const size_t g_at = bucket_ws_aligned(n_img, R_cap);
            void *gws = workspace && workspace_bytes > g_at ? (void *)((char *)workspace + g_at) : nullptr;
            const size_t gbytes = gws ? workspace_bytes - g_at : 0;
            const int rc = dtype == AZN_DTYPE_F32
                ? launch_staged<OpsF32, 0>(feat, n_img, H, W, L, rois, n_rois, R_cap, spatial_scale, out, nullptr, (int32_t *)workspace, s, gws, gbytes)
                : launch_staged<OpsBF16, 0>(feat, n_img, H, W, L, rois, n_rois, R_cap, spatial_scale, out, nullptr, (int32_t *)workspace, s, gws, gbytes);
            if (rc != 1) return rc;
        }
        AZN_REQUIRE((double)R_cap * PH < 2.0e9, "azn_roi_pool_fwd: too many ROI rows for one launch");
        AZN_REQUIRE(PW <= POOL_MAX_PW, "azn_roi_pool_fwd: pooled width %d > %d", PW, POOL_MAX_PW);
        const long items = (long)R_cap * PH;
        const int nwarps = PW < 8 ? PW : 8;
        long blocks = items;
        const long max_blocks = (long)sms * 16;     // (32 per SM measured the same: level 2's 4032 unequal items 0.0505 vs 0.0509 ms)
        if (blocks > max_blocks) blocks = max_blocks;
        const int nv = L >= 128 ? 4 : (L >= 64 ? 2 : 1);
        const uint4 *f = (const uint4 *)feat;
        uint4 *o = (uint4 *)out;
        // kernel_choice 3 (or azn_roi_pool_tune(5xx)): one CTA per ROI, one warp per bin row (roi_pool_nhwc_roi_kernel) -- for
        // MANY SMALL ROIs (the deep levels of the search: level 5 0.0325 -> 0.0254 ms, level 4 0.0242 -> 0.0227 per 64 images);
        // with few large ROIs the seven long serial bin rows of a CTA lose to the bin-row items of kernel (1) (level 2:
        // 0.051 -> 0.135 ms), and the library cannot see the ROI sizes from the host: the caller says which it has.
        const bool per_roi = (kernel_choice == 3 || g_pool_debug == 5) && PH <= POOL_MAX_PW && PW + PH <= 32;
        const int nwarps_r = PH < 8 ? PH : 8;
        const long blocks_r = std::min<long>((long)R_cap, (long)sms * 16);
#define AZN_POOL_LAUNCH(OPS, NVV) \
        do { \
            if (per_roi) \
                AZN_CUDA(azn_launch_pdl(roi_pool_nhwc_roi_kernel<OPS, NVV>, dim3((unsigned)blocks_r), dim3(nwarps_r * 32), 0, s, f, n_img, H, W, L, rois, n_rois, R_cap, PH, PW, spatial_scale, o)); \
            else \
                AZN_CUDA(azn_launch_pdl(roi_pool_nhwc_kernel<OPS, NVV>, dim3((unsigned)blocks), dim3(nwarps * 32), 0, s, f, n_img, H, W, L, rois, n_rois, R_cap, PH, PW, spatial_scale, o, 0, 0, 1.f)); \
        } while (0)
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
        const size_t need = map_ws_bytes(n_img, C, H, W, layout, dtype);
        if (!workspace || workspace_bytes < need) {
            azn_set_error("azn_roi_pool_fwd: the NCHW f32 path needs a %zu-byte workspace (got %zu)", need, workspace_bytes);
            return AZN_ERR_CAPACITY;
        }
        float *nhwc = (float *)workspace;
        const int HW = H * W;
        nchw_to_nhwc_kernel<float><<<dim3((HW + 31) / 32, (C + 31) / 32, n_img), dim3(32, 8), 0, s>>>((const float *)feat, C, HW, nhwc);
        AZN_LAUNCH_CHECK();
        {
            int32_t *bws = (int32_t *)((char *)workspace + need);
            const bool bucket_ok = n_img <= 1 || (bucket_ws_bytes(n_img, R_cap) > 0 && workspace_bytes >= need + bucket_ws_bytes(n_img, R_cap));
            if (C % 4 == 0 && want_staged(kernel_choice, n_img, H, W, PH, PW, n_rois, R_cap, bucket_ok)) {
                const int rc = argmax
                    ? launch_staged<OpsF32, 2>(nhwc, n_img, H, W, C / 4, rois, n_rois, R_cap, spatial_scale, out, argmax, bws, s)
                    : launch_staged<OpsF32, 1>(nhwc, n_img, H, W, C / 4, rois, n_rois, R_cap, spatial_scale, out, nullptr, bws, s);
                if (rc != 1) return rc;
            }
        }
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
            static size_t attr_smem = 0;         /* raise the limit only when it grows: the call is slow */              \
            if (smem > attr_smem) {                                                                                      \
                AZN_CUDA(cudaFuncSetAttribute(roi_pool_caffe_kernel<AM, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                attr_smem = smem;                                                                                        \
            }                                                                                                            \
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

// ROI max-pool + GRN + concat offset + Power scale in one kernel (the skip-layer head).
extern "C" int azn_roi_pool_grn_fwd(const void *feat, int n_img, int C, int H, int W, const float *rois, const int32_t *n_rois,
                                    int R_cap, int PH, int PW, float spatial_scale, float grn_scale, void *out, int ld_out,
                                    int ch_off, azn_stream_t stream) {
    if (R_cap == 0) return AZN_OK;
    AZN_REQUIRE(feat && rois && out, "azn_roi_pool_grn_fwd: null pointer");
    AZN_REQUIRE(n_img > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && R_cap > 0,
                "azn_roi_pool_grn_fwd: bad shape n_img=%d C=%d H=%d W=%d PH=%d PW=%d R=%d", n_img, C, H, W, PH, PW, R_cap);
    AZN_REQUIRE(C % 8 == 0 && C <= 1024, "azn_roi_pool_grn_fwd: C=%d must be a multiple of 8 and <= 1024 (one warp holds a bin's channels)", C);
    AZN_REQUIRE(ld_out % 8 == 0 && ch_off % 8 == 0 && ch_off >= 0 && ch_off + C <= ld_out,
                "azn_roi_pool_grn_fwd: ld_out=%d / ch_off=%d must be multiples of 8 with ch_off + C <= ld_out", ld_out, ch_off);
    AZN_REQUIRE(((uintptr_t)feat % 16 == 0) && ((uintptr_t)out % 16 == 0), "azn_roi_pool_grn_fwd: 16-byte alignment");
    AZN_REQUIRE((double)R_cap * PH < 2.0e9 && PW <= POOL_MAX_PW, "azn_roi_pool_grn_fwd: too many ROI rows / pooled width > %d", POOL_MAX_PW);
    const int L = C / 8;
    const long items = (long)R_cap * PH;
    const int nwarps = PW < 8 ? PW : 8;
    long blocks = items;
    const long max_blocks = (long)azn_num_sms() * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    const int nv = L > 64 ? 4 : (L > 32 ? 2 : 1);
    cudaStream_t s = (cudaStream_t)stream;
    const uint4 *f = (const uint4 *)feat;
    uint4 *o = (uint4 *)out;
#define AZN_GRN_LAUNCH(NVV) \
    AZN_CUDA(azn_launch_pdl(roi_pool_nhwc_kernel<OpsBF16, NVV, true>, dim3((unsigned)blocks), dim3(nwarps * 32), 0, s, f, n_img, H, W, L, \
                            rois, n_rois, R_cap, PH, PW, spatial_scale, o, ld_out / 8, ch_off / 8, grn_scale))
    if (nv == 4) AZN_GRN_LAUNCH(4); else if (nv == 2) AZN_GRN_LAUNCH(2); else AZN_GRN_LAUNCH(1);
#undef AZN_GRN_LAUNCH
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

extern "C" int azn_nchw_bf16_to_nhwc_bf16(const void *src, int n_img, int C, int H, int W, void *dst, azn_stream_t stream) {
    AZN_REQUIRE(src && dst && n_img > 0 && C > 0 && H > 0 && W > 0, "azn_nchw_bf16_to_nhwc_bf16: bad argument");
    const int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, n_img), block(32, 8);
    nchw_to_nhwc_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)src, C, HW, (__nv_bfloat16 *)dst);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
