// Greedy non-maximum suppression for sm_100a, bit-exact with utils.cython_nms.nms
// (lib/utils/nms.pyx:17-68) on the same float32 detections.
//
// Reference semantics restated:
//   areas = (x2 - x1 + 1) * (y2 - y1 + 1)                      f32, one rounding per op   (:24)
//   order = scores.argsort()[::-1]                              descending score           (:25)
//   for i in order (skipping suppressed): keep i; for every later j:
//       w = max(0, min(x2i,x2j) - max(x1i,x1j) + 1), h likewise; inter = w*h
//       ovr = inter / (area_i + area_j - inter)                 f32                        (:57-64)
//       suppress j iff (double)ovr >= thresh                                               (:65-66)
// Every f32 operation below is an explicit round-to-nearest intrinsic so nvcc cannot contract
// a multiply-add into an FMA (x86-64 gcc, which built the reference, has no FMA by default).
//
// Large single problem (azn_nms), four launches on one stream:
//   1. nms_rank_kernel   -- rank sort by (score desc, index desc) on a 2-D grid: every thread counts
//      nms_scatter_kernel   the detections of one score tile that precede its own (shared memory),
//                           partial counts are added atomically, then boxes move to sorted position.
//   2. nms_mask_kernel   -- 64x64 IoU tiles of the upper triangle -> one 64-bit suppression word
//                           per (row, column tile); exact intersection pre-test, division only
//                           for intersecting pairs.
//   3. nms_scan_kernel   -- one CTA walks the column tiles in order: it gathers (pull) the column
//                           word of every row kept so far and OR-reduces it with warp shuffles,
//                           warp 0 resolves the 64x64 diagonal block over the surviving rows with
//                           register-resident words, and the kept original indices are appended to
//                           `keep` (on-device compaction; the count never visits the host).
// Many small problems (azn_nms_batched, one per class per image in apply_nms,
// lib/detect/test.py:467-484): one warp per problem, boxes in shared memory, the suppression
// state in per-lane registers exchanged with ballots.
#include "common.cuh"

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2) {
    float w = __fadd_rn(__fsub_rn(x2, x1), 1.f);
    float h = __fadd_rn(__fsub_rn(y2, y1), 1.f);
    return __fmul_rn(w, h);
}

// true iff box j must be suppressed by kept box i
__device__ __forceinline__ bool suppresses(const float4 &a, float area_a, const float4 &b, float area_b, double thresh) {
    float xx1 = a.x >= b.x ? a.x : b.x;
    float yy1 = a.y >= b.y ? a.y : b.y;
    float xx2 = a.z <= b.z ? a.z : b.z;
    float yy2 = a.w <= b.w ? a.w : b.w;
    float w = __fadd_rn(__fsub_rn(xx2, xx1), 1.f);
    float h = __fadd_rn(__fsub_rn(yy2, yy1), 1.f);
    w = 0.f >= w ? 0.f : w;
    h = 0.f >= h ? 0.f : h;
    float inter = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    float ovr = __fdiv_rn(inter, uni);
    return (double)ovr >= thresh;
}

// (score desc, index desc): does detection j come before detection i ?
__device__ __forceinline__ bool precedes(float sj, int j, float si, int i) {
    return sj > si || (sj == si && j > i);
}

constexpr int RANK_THREADS = 256;
constexpr int RANK_TILE = 2048;

// grid = (ceil(n/256), ceil(n/RANK_TILE)): block (bx, by) counts, for its 256 detections, how many of the
// by-th tile of detections precede each of them, and adds the partial count to rank[i].
__global__ void __launch_bounds__(RANK_THREADS)
nms_rank_kernel(const float *__restrict__ dets, int n, int *__restrict__ rank) {
    __shared__ float s_tile[RANK_TILE];
    const int i = blockIdx.x * RANK_THREADS + threadIdx.x;
    const int t0 = blockIdx.y * RANK_TILE, tn = min(RANK_TILE, n - t0);
    for (int k = threadIdx.x; k < tn; k += RANK_THREADS) s_tile[k] = dets[(size_t)(t0 + k) * 5 + 4];
    __syncthreads();
    if (i >= n) return;
    const float si = dets[(size_t)i * 5 + 4];
    int cnt = 0;
#pragma unroll 8
    for (int k = 0; k < tn; ++k) cnt += precedes(s_tile[k], t0 + k, si, i) ? 1 : 0;
    if (cnt) atomicAdd(rank + i, cnt);
}

__global__ void __launch_bounds__(256)
nms_scatter_kernel(const float *__restrict__ dets, int n, const int *__restrict__ rank, float4 *__restrict__ boxes,
                   float *__restrict__ areas, int *__restrict__ order) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float *d = dets + (size_t)i * 5;
    const float4 b = make_float4(d[0], d[1], d[2], d[3]);
    const int r = rank[i];
    boxes[r] = b;
    areas[r] = box_area(b.x, b.y, b.z, b.w);
    order[r] = i;
}

// true iff the clamped intersection of a and b is non-empty (w > 0 and h > 0 with the reference's f32 ops)
__device__ __forceinline__ bool intersects(const float4 &a, const float4 &b) {
    float xx1 = a.x >= b.x ? a.x : b.x;
    float yy1 = a.y >= b.y ? a.y : b.y;
    float xx2 = a.z <= b.z ? a.z : b.z;
    float yy2 = a.w <= b.w ? a.w : b.w;
    float w = __fadd_rn(__fsub_rn(xx2, xx1), 1.f);
    float h = __fadd_rn(__fsub_rn(yy2, yy1), 1.f);
    return w > 0.f && h > 0.f;
}

// grid = (col_tiles, row_tiles), 64 threads; only tiles with col >= row do work.  Two phases per thread:
// a cheap exact intersection test over the 64 columns, then the full IoU (IEEE division + double compare)
// only for the columns that intersect -- with an empty intersection the reference computes ovr = +-0 or
// NaN, which is >= thresh for no positive threshold, so those pairs can never be suppressed.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4 *__restrict__ boxes, const float *__restrict__ areas, int n, double thresh,
                u64 *__restrict__ mask, int col_tiles) {
    const int rt = blockIdx.y, ct = blockIdx.x;
    if (ct < rt) return;
    __shared__ float4 cb[64];
    __shared__ float ca[64];
    const int cn = min(64, n - ct * 64);
    if ((int)threadIdx.x < cn) {
        cb[threadIdx.x] = boxes[ct * 64 + threadIdx.x];
        ca[threadIdx.x] = areas[ct * 64 + threadIdx.x];
    }
    __syncthreads();
    const int row = rt * 64 + threadIdx.x;
    if (row >= n) return;
    const float4 rb = boxes[row];
    const float ra = areas[row];
    const int start = (rt == ct) ? threadIdx.x + 1 : 0;
    u64 cand = 0;
    if (thresh > 0.0) {
        for (int k = start; k < cn; ++k) cand |= (u64)(intersects(rb, cb[k]) ? 1 : 0) << k;
    } else {
        cand = (cn == 64 ? ~0ull : ((1ull << cn) - 1ull)) & (start >= 64 ? 0ull : (~0ull << start));
    }
    u64 bits = 0;
    while (cand) {
        const int k = __ffsll((long long)cand) - 1;
        cand &= cand - 1;
        if (suppresses(rb, ra, cb[k], ca[k], thresh)) bits |= 1ull << k;
    }
    mask[(size_t)row * col_tiles + ct] = bits;
}

constexpr int SCAN_THREADS = 1024;

// One CTA walks the column tiles in score order.  For tile b it PULLS the suppression word of column tile b
// from every row kept so far (parallel gather + OR-reduction), warp 0 resolves the 64x64 diagonal block by
// iterating over the surviving rows only, and the kept rows are appended to `keep` / `kept_rows`.
__global__ void __launch_bounds__(SCAN_THREADS)
nms_scan_kernel(const u64 *__restrict__ mask, const int *__restrict__ order, int n, int col_tiles,
                int *__restrict__ kept_rows, int64_t *__restrict__ keep, int32_t *__restrict__ keep_count) {
    __shared__ u64 s_or[SCAN_THREADS / 32];
    __shared__ int s_nkept;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_nkept = 0;
    __syncthreads();
    for (int b = 0; b < col_tiles; ++b) {
        const int row0 = b * 64;
        const int nkept = s_nkept;
        // diagonal words first (independent of the gather below)
        u64 d_lo = 0, d_hi = 0;
        if (warp == 0) {
            const int r_lo = row0 + lane, r_hi = row0 + 32 + lane;
            d_lo = r_lo < n ? mask[(size_t)r_lo * col_tiles + b] : 0ull;
            d_hi = r_hi < n ? mask[(size_t)r_hi * col_tiles + b] : 0ull;
        }
        u64 acc = 0;
        int j = tid;
        for (; j + 3 * SCAN_THREADS < nkept; j += 4 * SCAN_THREADS) {
            const int r0 = __ldcg(kept_rows + j), r1 = __ldcg(kept_rows + j + SCAN_THREADS), r2 = __ldcg(kept_rows + j + 2 * SCAN_THREADS),
                      r3 = __ldcg(kept_rows + j + 3 * SCAN_THREADS);
            const u64 m0 = mask[(size_t)r0 * col_tiles + b], m1 = mask[(size_t)r1 * col_tiles + b],
                      m2 = mask[(size_t)r2 * col_tiles + b], m3 = mask[(size_t)r3 * col_tiles + b];
            acc |= m0 | m1 | m2 | m3;
        }
        for (; j < nkept; j += SCAN_THREADS) acc |= mask[(size_t)__ldcg(kept_rows + j) * col_tiles + b];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) s_or[warp] = acc;
        __syncthreads();
        if (warp == 0) {
            u64 cur = s_or[lane];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cur |= __shfl_xor_sync(0xffffffffu, cur, d);
            const int rows = min(64, n - row0);
            u64 alive = ~cur & (rows == 64 ? ~0ull : ((1ull << rows) - 1ull));
            u64 kept = 0;
            while (alive) {                                   // warp-uniform: every lane holds the same `alive`
                const int r = __ffsll((long long)alive) - 1;
                kept |= 1ull << r;
                const u64 d = __shfl_sync(0xffffffffu, r < 32 ? d_lo : d_hi, r & 31);
                alive &= ~d;
                alive &= ~(1ull << r);
            }
            const int r_lo = row0 + lane, r_hi = row0 + 32 + lane;
            if ((kept >> lane) & 1ull) {
                const int pos = nkept + __popcll(kept & ((1ull << lane) - 1ull));
                keep[pos] = order[r_lo];
                kept_rows[pos] = r_lo;
            }
            if ((kept >> (lane + 32)) & 1ull) {
                const int pos = nkept + __popcll(kept & ((1ull << (lane + 32)) - 1ull));
                keep[pos] = order[r_hi];
                kept_rows[pos] = r_hi;
            }
            if (lane == 0) s_nkept = nkept + __popcll(kept);
        }
        __syncthreads();      // kept_rows / s_nkept visible to the whole CTA (global writes by the same CTA)
    }
    if (tid == 0) *keep_count = s_nkept;
}

// One warp per segment.  Shared memory per warp: 6 * max_seg floats/ints.
constexpr int BATCH_WARPS = 4;

__global__ void __launch_bounds__(BATCH_WARPS * 32)
nms_batched_kernel(const float *__restrict__ dets, const int32_t *__restrict__ seg_off, int n_seg, double thresh,
                   int max_seg, int64_t *__restrict__ keep, int32_t *__restrict__ keep_count) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sx1 = smem + (size_t)warp * 6 * max_seg;
    float *sy1 = sx1 + max_seg, *sx2 = sy1 + max_seg, *sy2 = sx2 + max_seg, *sar = sy2 + max_seg;
    int *sid = (int *)(sar + max_seg);
    for (int seg = blockIdx.x * BATCH_WARPS + warp; seg < n_seg; seg += gridDim.x * BATCH_WARPS) {
        const int o = seg_off[seg], n = seg_off[seg + 1] - o;
        if (n > max_seg || n < 0) {
            if (lane == 0) keep_count[seg] = -1;
            continue;
        }
        const float *d = dets + (size_t)o * 5;
        // rank sort straight from global memory (n is small), scatter to shared
        for (int i = lane; i < n; i += 32) {
            const float si = d[(size_t)i * 5 + 4];
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += precedes(d[(size_t)j * 5 + 4], j, si, i) ? 1 : 0;
            const float x1 = d[(size_t)i * 5], y1 = d[(size_t)i * 5 + 1], x2 = d[(size_t)i * 5 + 2], y2 = d[(size_t)i * 5 + 3];
            sx1[rank] = x1; sy1[rank] = y1; sx2[rank] = x2; sy2[rank] = y2;
            sar[rank] = box_area(x1, y1, x2, y2);
            sid[rank] = i;
        }
        __syncwarp();
        // lane l owns sorted positions l, l+32, ...; `dead` bit k <-> position l + 32k (max_seg <= 1024)
        unsigned dead = 0;
        int nkept = 0;
        for (int i = 0; i < n; ++i) {
            const unsigned owner_dead = __shfl_sync(0xffffffffu, dead, i & 31);
            if ((owner_dead >> (i >> 5)) & 1u) continue;
            if (lane == 0) keep[o + nkept] = sid[i];
            ++nkept;
            const float4 bi = make_float4(sx1[i], sy1[i], sx2[i], sy2[i]);
            const float ai = sar[i];
            for (int j = lane + ((i + 1 - lane + 31) & ~31); j < n; j += 32) {   // first j > i owned by lane
                const float4 bj = make_float4(sx1[j], sy1[j], sx2[j], sy2[j]);
                if (suppresses(bi, ai, bj, sar[j], thresh)) dead |= 1u << (j >> 5);
            }
        }
        if (lane == 0) keep_count[seg] = nkept;
        __syncwarp();
    }
}

struct NmsWorkspace {
    float4 *boxes;
    float *areas;
    int *order, *rank, *kept_rows;
    u64 *mask;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline NmsWorkspace carve(void *ws, int64_t n, int col_tiles) {
    char *p = (char *)ws;
    NmsWorkspace w;
    w.boxes = (float4 *)p;  p += align_up(sizeof(float4) * n, 256);
    w.areas = (float *)p;   p += align_up(sizeof(float) * n, 256);
    w.order = (int *)p;     p += align_up(sizeof(int) * n, 256);
    w.rank = (int *)p;      p += align_up(sizeof(int) * n, 256);
    w.kept_rows = (int *)p; p += align_up(sizeof(int) * n, 256);
    w.mask = (u64 *)p;
    (void)col_tiles;
    return w;
}

}  // namespace

extern "C" size_t azn_nms_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    const size_t ct = (size_t)((n + 63) / 64);
    return align_up(sizeof(float4) * n, 256) + align_up(sizeof(float) * n, 256) + 3 * align_up(sizeof(int) * n, 256) +
           align_up(sizeof(u64) * (size_t)n * ct, 256);
}

extern "C" int azn_nms(const float *dets, int64_t n, double thresh, int64_t *keep, int32_t *keep_count,
                       void *workspace, size_t workspace_bytes, azn_stream_t stream) {
    AZN_REQUIRE(keep_count != nullptr, "azn_nms: keep_count is null");
    AZN_REQUIRE(n >= 0 && n <= 1000000, "azn_nms: n=%lld out of range", (long long)n);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        AZN_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), s));
        return AZN_OK;
    }
    AZN_REQUIRE(dets && keep && workspace, "azn_nms: null pointer");
    if (workspace_bytes < azn_nms_workspace_bytes(n)) {
        azn_set_error("azn_nms: workspace %zu < %zu bytes", workspace_bytes, azn_nms_workspace_bytes(n));
        return AZN_ERR_CAPACITY;
    }
    const int col_tiles = (int)((n + 63) / 64);
    NmsWorkspace w = carve(workspace, n, col_tiles);
    AZN_CUDA(cudaMemsetAsync(w.rank, 0, sizeof(int) * n, s));
    nms_rank_kernel<<<dim3((unsigned)((n + RANK_THREADS - 1) / RANK_THREADS), (unsigned)((n + RANK_TILE - 1) / RANK_TILE)),
                      RANK_THREADS, 0, s>>>(dets, (int)n, w.rank);
    AZN_LAUNCH_CHECK();
    nms_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dets, (int)n, w.rank, w.boxes, w.areas, w.order);
    AZN_LAUNCH_CHECK();
    nms_mask_kernel<<<dim3(col_tiles, col_tiles), 64, 0, s>>>(w.boxes, w.areas, (int)n, thresh, w.mask, col_tiles);
    AZN_LAUNCH_CHECK();
    nms_scan_kernel<<<1, SCAN_THREADS, 0, s>>>(w.mask, w.order, (int)n, col_tiles, w.kept_rows, keep, keep_count);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_nms_batched(const float *dets, const int32_t *seg_off, int n_seg, double thresh,
                               int64_t *keep, int32_t *keep_count, azn_stream_t stream) {
    AZN_REQUIRE(n_seg >= 0, "azn_nms_batched: n_seg < 0");
    if (n_seg == 0) return AZN_OK;
    AZN_REQUIRE(dets && seg_off && keep && keep_count, "azn_nms_batched: null pointer");
    const int max_seg = AZN_NMS_SEG_MAX;
    const size_t smem = (size_t)BATCH_WARPS * 6 * max_seg * sizeof(float);   // 96 KB
    static bool attr_set = false;
    if (!attr_set) {
        AZN_CUDA(cudaFuncSetAttribute(nms_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int blocks = (n_seg + BATCH_WARPS - 1) / BATCH_WARPS;
    const int cap = azn_num_sms() * 2;
    if (blocks > cap) blocks = cap;
    nms_batched_kernel<<<blocks, BATCH_WARPS * 32, smem, (cudaStream_t)stream>>>(dets, seg_off, n_seg, thresh, max_seg,
                                                                              keep, keep_count);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
