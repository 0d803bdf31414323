// The three output layers of the AZ-Net head in one small kernel:
//
//   out[M, 5*nsub+1] = [sigmoid(adj_score) | adj_bbox | sigmoid(zoom_score)] = act(h7[M, K] . Wh[N, K]^T + bias)
//
// replaces: InnerProduct adj_score (1024 -> 11), adj_bbox (1024 -> 44), zoom_score (256 -> 1) and the Sigmoid layers
// adj_prob / zoom_prob of models/Pascal/VGG16/az-net/test_fc.prototxt:146-232 (inner_product_layer.cpp:80-93,
// sigmoid_layer.cpp:11-24).  Wh is the block-diagonal fusion of the three weight blobs over the concatenated
// [int7_1 | int7_2] activations (aznet_b200/engine.py::AZHeadWeights), N = 56, K = 1280.
//
// Why not azn_fc_forward: a 56-column product is 0.2 GFLOP per level -- the persistent tcgen05 kernel spends its
// 13-18 us on fixed costs (TMEM allocation, tensor-map fetch, a 20-k-block serial loop on <= 12 of 148 SMs, teardown)
// at 1-2 % tensor-pipe activity, and as a cooperative, one-CTA-per-SM launch it keeps every SM from the other stream's
// kernels while it runs.  This kernel is an ordinary grid: one CTA per 16 rows (M = 1494 -> 94 CTAs), its four warps
// split K four ways; every warp runs warp-level mma.sync.m16n8k16 (bf16 -> f32) over its quarter of K with the A and
// B fragments loaded straight from global memory / L2 two k-steps ahead (Wh is 143 KB and shared by every CTA; an
// activation byte is used exactly once) -- no shared-memory staging, no barrier in the loop, so the serial chain
// is 20 k-steps instead of 80.  The four partial tiles meet in 18 KB of shared memory (fixed order: deterministic), then
// bias + sigmoid.  It is short, and it is co-resident with the big GEMM's CTAs, so with two batches in flight it hides
// behind the other batch's int6.  (A first version -- 64 rows per CTA, Wh staged chunk by chunk through cp.async --
// measured 30 us: 24 CTAs, each waiting out an L2 round trip per chunk, 20 chunks in a row.)
#include "common.cuh"

namespace {

constexpr int HD_WARPS = 4;                // K is split over the warps of a CTA
constexpr int HD_ROWS = 16;                // rows per CTA: one m16 tile
constexpr int HD_NMAX = 64;                // columns
constexpr int HD_PITCH = HD_NMAX + 8;      // floats per row of a partial tile in shared memory
constexpr int HD_AHEAD = 2;                // k-steps of fragments in flight

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// sigmoid_layer.cpp:11-13 (`1. / (1. + exp(-x))`): float exp, IEEE float division -- the same formula as the GEMM
// epilogue's AZN_ACT_AZ_HEAD (csrc/gemm.cu), <= 1 ulp from the reference's float-exp / double-divide evaluation.
__device__ __forceinline__ float sigmoid_caffe(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

int g_heads_plain = 0;

template <int NT>      // column tiles of 8: N <= 8 * NT
__global__ void __launch_bounds__(32 * HD_WARPS)
az_heads_kernel(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ W, const float *__restrict__ bias,
                float *__restrict__ out, int ldo, int M_cap, const int32_t *__restrict__ m_live_ptr, int N, int K, int nsub) {
    __shared__ float red[HD_WARPS][HD_ROWS][HD_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    pdl_enter();
    const int m = m_live_ptr ? min(max(*m_live_ptr, 0), M_cap) : M_cap;
    const int row0 = blockIdx.x * HD_ROWS;
    if (row0 >= m) return;
    // fragment sources (mma.m16n8k16 layouts): A regs {row g | row g+8} x {k 2t.. | k 2t+8..}; B regs of column tile nt
    // {k 2t.. | k 2t+8..} of Wh row nt*8 + g.  Rows past the live count / past N are clamped: computed, never stored.
    const int steps = K / (16 * HD_WARPS);                    // k-steps of this warp
    const int kw = warp * steps * 8;                          // first 32-bit word of this warp's K range
    const uint32_t *a_lo = reinterpret_cast<const uint32_t *>(A + (size_t)min(row0 + g, m - 1) * K) + kw + t;
    const uint32_t *a_hi = reinterpret_cast<const uint32_t *>(A + (size_t)min(row0 + g + 8, m - 1) * K) + kw + t;
    const uint32_t *b_row[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b_row[nt] = reinterpret_cast<const uint32_t *>(W + (size_t)min(nt * 8 + g, N - 1) * K) + kw + t;
    uint32_t fa[HD_AHEAD + 1][4], fb[HD_AHEAD + 1][NT][2];
    auto load = [&](int ks, int slot) {
        const int w0 = ks * 8;                                // 16 bf16 = 8 words per k-step
        fa[slot][0] = __ldg(a_lo + w0); fa[slot][1] = __ldg(a_hi + w0);
        fa[slot][2] = __ldg(a_lo + w0 + 4); fa[slot][3] = __ldg(a_hi + w0 + 4);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { fb[slot][nt][0] = __ldg(b_row[nt] + w0); fb[slot][nt][1] = __ldg(b_row[nt] + w0 + 4); }
    };
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int p = 0; p < HD_AHEAD; ++p)
        if (p < steps) load(p, p);
#pragma unroll 1
    for (int ks0 = 0; ks0 < steps; ks0 += HD_AHEAD + 1) {     // unrolled by the ring size: slot indices are constants
#pragma unroll
        for (int u = 0; u <= HD_AHEAD; ++u) {
            const int ks = ks0 + u;
            if (ks < steps) {
                if (ks + HD_AHEAD < steps) load(ks + HD_AHEAD, (u + HD_AHEAD) % (HD_AHEAD + 1));
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], fa[u], fb[u][nt][0], fb[u][nt][1]);
            }
        }
    }
    // the four K-quarters meet in shared memory: c0,c1 -> row g, columns 2t, 2t+1; c2,c3 -> row g+8
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        *reinterpret_cast<float2 *>(&red[warp][g][nt * 8 + 2 * t]) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2 *>(&red[warp][g + 8][nt * 8 + 2 * t]) = make_float2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
    for (int e = tid; e < HD_ROWS * N; e += 32 * HD_WARPS) {
        const int r = e / N, col = e - r * N;
        if (row0 + r >= m) continue;
        float v = red[0][r][col];
#pragma unroll
        for (int w = 1; w < HD_WARPS; ++w) v = __fadd_rn(v, red[w][r][col]);
        v = __fadd_rn(v, __ldg(bias + col));
        if (col < nsub || col == 5 * nsub) v = sigmoid_caffe(v);
        out[(size_t)(row0 + r) * ldo + col] = v;
    }
}

}  // namespace

extern "C" void azn_az_heads_tune(int plain_launch) { g_heads_plain = plain_launch ? 1 : 0; }

extern "C" int azn_az_heads_forward(const void *h7, const void *wh, const float *bias, float *out, int ldo, int M_cap,
                                    const int32_t *m_live, int N, int K, int nsub, azn_stream_t stream) {
    AZN_REQUIRE(h7 && wh && bias && out, "azn_az_heads_forward: null pointer");
    AZN_REQUIRE(M_cap > 0 && N > 0 && N <= HD_NMAX && K > 0 && K % (16 * HD_WARPS) == 0, "azn_az_heads_forward: bad shape M=%d N=%d (<= %d) K=%d (multiple of %d)",
                M_cap, N, HD_NMAX, K, 16 * HD_WARPS);
    AZN_REQUIRE(nsub > 0 && 5 * nsub + 1 <= N && ldo >= N, "azn_az_heads_forward: needs 5*nsub+1 <= N <= ldo (nsub=%d N=%d ldo=%d)", nsub, N, ldo);
    AZN_REQUIRE(((uintptr_t)h7 % 16 == 0) && ((uintptr_t)wh % 16 == 0), "azn_az_heads_forward: h7 and wh must be 16-byte aligned");
    const dim3 grid((M_cap + HD_ROWS - 1) / HD_ROWS), block(32 * HD_WARPS);
    cudaStream_t s = (cudaStream_t)stream;
    const __nv_bfloat16 *a = (const __nv_bfloat16 *)h7, *w = (const __nv_bfloat16 *)wh;
    // Launch mode (azn_az_heads_tune): with ONE batch in flight the programmatic-dependent launch of this kernel
    // measured 70 us per step slower than a plain stream-ordered launch (1.034 vs 0.964 ms: its up to 2256 CTAs become
    // resident behind the persistent int7 GEMM and its own dependents behind them); with several batches in flight it is
    // the faster one (0.775 vs 0.785 ms at four streams).  Default: PDL (the multi-stream configuration).
    if (g_heads_plain) {
        if (N <= 56) az_heads_kernel<7><<<grid, block, 0, s>>>(a, w, bias, out, ldo, M_cap, m_live, N, K, nsub);
        else az_heads_kernel<8><<<grid, block, 0, s>>>(a, w, bias, out, ldo, M_cap, m_live, N, K, nsub);
        AZN_LAUNCH_CHECK();
        return AZN_OK;
    }
    if (N <= 56) AZN_CUDA(azn_launch_pdl(az_heads_kernel<7>, grid, block, 0, s, a, w, bias, out, ldo, M_cap, m_live, N, K, nsub));
    else AZN_CUDA(azn_launch_pdl(az_heads_kernel<8>, grid, block, 0, s, a, w, bias, out, ldo, M_cap, m_live, N, K, nsub));
    return AZN_OK;
}
