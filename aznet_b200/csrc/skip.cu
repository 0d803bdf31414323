// The skip-layer detector head for sm_100a: what sits between the three ROI pools and fc6 in
// models/COCO/VGG16_skip/frcnn/test_fc.prototxt:28-135 --
//
//   roi_norm{3,4,5}  GRN      per pooled position, x / sqrt(sum_c x^2)     caffe-fast-rcnn/src/caffe/layers/grn_layer.cpp:27-56
//   concat5          Concat   along channels: [conv3_3 | conv4_3 | conv5_3]  test_fc.prototxt:88-99
//   scale5           Power    y = 1000 * x (power 1, shift 0)               test_fc.prototxt:100-110
//   conv_pool5       1x1 Convolution 1280 -> 512 + ReLU                    test_fc.prototxt:111-142
//
// The pooled rows of azn_roi_pool_fwd (NHWC) are [R, 49, C]: position-major, channel-minor, so one pooled
// position IS one contiguous run of C channels.  GRN + concat + scale is therefore one pass over rows of
// C3 + C4 + C5 channels (azn_grn_concat_forward below: a warp per position, HBM-bound), and the 1x1 convolution
// is azn_fc_forward over M = 49 R rows, K = 1280, N = 512 whose output [49 R, 512] is, without any copy, the
// [R, 49 * 512] pooled-row matrix that fc6 consumes.
//
// Arithmetic: the sum of squares runs in float32 (each bf16 square is exact in float32), norm = sqrtf(sum),
// y = x * (1000 / norm) -- one IEEE division per position instead of one per element; the reference's
// 1000 * (x / norm) differs by at most ~1.5 float32 ulp, which the single bf16 rounding of the tensor-core operand
// (2^-9) hides but for rare ties.  A position whose channels are all zero divides 0 by 0 exactly like the
// reference (NaN): the layer has no epsilon.
#include "common.cuh"

namespace {

constexpr int GRN_MAX_SRC = 8;
constexpr int GRN_MAX_VPL = 4;          // 16-byte vectors per lane per source: C <= 1024 channels per source

struct GrnArgs {
    const uint4 *src[GRN_MAX_SRC];      // bf16 [rows, C_l]
    int vec[GRN_MAX_SRC];               // C_l / 8
    int off[GRN_MAX_SRC];               // first output vector of source l in a row
    int n_src;
    int out_vec;                        // output row stride in 16-byte vectors (>= sum C_l / 8)
};

__device__ __forceinline__ float bf_lo(unsigned u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(unsigned u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ float sq_pair(unsigned u) {
    const float lo = bf_lo(u), hi = bf_hi(u);
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

__global__ void __launch_bounds__(256)
grn_concat_kernel(GrnArgs a, const int32_t *__restrict__ n_units, long rows_cap, int rows_per_unit, float scale,
                  uint4 *__restrict__ out) {
    pdl_enter();
    long rows = rows_cap;
    if (n_units) {
        const long live = (long)max(*n_units, 0) * rows_per_unit;
        rows = live < rows_cap ? live : rows_cap;
    }
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long row = warp0; row < rows; row += nwarps) {
        for (int l = 0; l < a.n_src; ++l) {
            const int nv = a.vec[l];
            const uint4 *p = a.src[l] + row * nv;
            uint4 v[GRN_MAX_VPL];
            float ss = 0.f;
#pragma unroll
            for (int q = 0; q < GRN_MAX_VPL; ++q) {
                const int i = lane + 32 * q;
                v[q] = make_uint4(0u, 0u, 0u, 0u);
                if (i < nv) v[q] = __ldg(p + i);
                // the same summation order as the fused epilogue of azn_roi_pool_grn_fwd (roi_pool.cu): bit-identical results
                if (i < nv) ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(sq_pair(v[q].x), sq_pair(v[q].y)), __fadd_rn(sq_pair(v[q].z), sq_pair(v[q].w))));
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, d));
            const float inv = __fdiv_rn(scale, sqrtf(ss));
            uint4 *o = out + row * a.out_vec + a.off[l];
#pragma unroll
            for (int q = 0; q < GRN_MAX_VPL; ++q) {
                const int i = lane + 32 * q;
                if (i < nv) {
                    uint4 r;
                    r.x = grn_pair(v[q].x, inv); r.y = grn_pair(v[q].y, inv);
                    r.z = grn_pair(v[q].z, inv); r.w = grn_pair(v[q].w, inv);
                    o[i] = r;
                }
            }
        }
    }
}

}  // namespace

extern "C" int azn_grn_concat_forward(const void *const *pooled, const int32_t *channels, int n_src, const int32_t *n_units,
                                      long long rows_cap, int rows_per_unit, float scale, void *out, int ld_out, azn_stream_t stream) {
    AZN_REQUIRE(pooled && channels && out, "azn_grn_concat_forward: null pointer");
    AZN_REQUIRE(n_src >= 1 && n_src <= GRN_MAX_SRC, "azn_grn_concat_forward: 1..%d sources (got %d)", GRN_MAX_SRC, n_src);
    AZN_REQUIRE(rows_cap >= 0 && rows_per_unit >= 1, "azn_grn_concat_forward: bad row counts");
    if (rows_cap == 0) return AZN_OK;
    GrnArgs a = {};
    a.n_src = n_src;
    int off = 0;
    for (int l = 0; l < n_src; ++l) {
        AZN_REQUIRE(pooled[l] != nullptr && ((uintptr_t)pooled[l] % 16) == 0, "azn_grn_concat_forward: source %d null or unaligned", l);
        AZN_REQUIRE(channels[l] > 0 && channels[l] % 8 == 0 && channels[l] <= 8 * 32 * GRN_MAX_VPL,
                    "azn_grn_concat_forward: channels[%d] = %d must be a multiple of 8 and <= %d", l, channels[l], 8 * 32 * GRN_MAX_VPL);
        a.src[l] = (const uint4 *)pooled[l];
        a.vec[l] = channels[l] / 8;
        a.off[l] = off;
        off += a.vec[l];
    }
    AZN_REQUIRE(ld_out % 8 == 0 && ld_out >= 8 * off, "azn_grn_concat_forward: ld_out = %d must be a multiple of 8 and >= %d", ld_out, 8 * off);
    a.out_vec = ld_out / 8;
    AZN_REQUIRE(((uintptr_t)out % 16) == 0, "azn_grn_concat_forward: out must be 16-byte aligned");
    const long warps_needed = (long)rows_cap;
    long blocks = (warps_needed + 7) / 8;
    const long max_blocks = (long)azn_num_sms() * 8;           // 8 CTAs of 256 threads per SM: one full wave
    if (blocks > max_blocks) blocks = max_blocks;
    AZN_CUDA(azn_launch_pdl(grn_concat_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a, n_units,
                            (long)rows_cap, rows_per_unit, scale, (uint4 *)out));
    return AZN_OK;
}
