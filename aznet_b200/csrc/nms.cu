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
// Large single problem (azn_nms).  The mask is ONE launch, column-major over the triangle (cheapest columns first),
// publishing per-super-column tile counters; the greedy pass runs concurrently on an internal stream (its own 16-SM
// green-context partition when the driver has them) and launch s polls the counters of super-columns s and s+1, so
// the serial chain runs alongside the mask and ends ~25 us after it.  Kernels (round 2; the round-1 kernels they replace
// stay selectable through azn_nms_tune for A/B):
//   1. n <= 2048: nms_rank1_kernel -- eight threads count the detections that precede theirs (score desc, index desc)
//      among the keys staged in shared memory; the box moves to sorted position.
//      n > 2048: nms_bucket_kernel + nms_bucket_rank_kernel -- one CTA spreads the keys over 2048 buckets that are linear
//      in the uint key (histogram, scan, shared-memory-staged scatter), then every detection is ranked inside its own
//      bucket only: the same exact order for a few dozen compares per box instead of n.
//   2. nms_mask2_kernel  -- 64x64 tiles of the upper triangle -> one 64-bit suppression word per (row, column tile).
//      A packed-half screen with directed rounding (two columns per instruction; intersection + area ratio; can only err
//      towards "candidate") leaves ~3 % of the pairs, which the warp gathers into one list and evaluates 32 at a time
//      with the exact float32 tests and the IEEE division of the reference.  Tiles of a diagonal super-block are also
//      written transposed (sup_t) for the greedy pass.  (thresh <= 0: nms_mask_kernel, the float32 version.)
//   3. nms_block_kernel  -- the greedy pass, blocked like a triangular solve: one PDL launch per super-tile of 1024 boxes.
//      CTA 0: thread j owns row j and keeps in registers the 16 words that say which rows of the super-tile suppress it;
//      the kept set K and the undecided set U live in shared memory and the 1024 rows are decided together in fixed-point
//      rounds (a row with a kept suppressor is removed, a row with no undecided suppressor left is kept); kept original
//      indices are appended to `keep` (on-device compaction; the count never visits the host) and the kept rows are pushed
//      into the `removed` words of the next super-tile's columns.  The other CTAs of the launch push the kept rows of ALL
//      EARLIER super-tiles into the NEXT super-tile's columns (left-looking bulk update, atomicOr).
//      No CTA ever waits for another inside a launch.
// Many small problems (azn_nms_batched, one per class per image in apply_nms,
// lib/detect/test.py:467-484): one warp per problem, boxes in shared memory, the suppression
// state in per-lane registers exchanged with ballots.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2) {
    float w = __fadd_rn(__fsub_rn(x2, x1), 1.f);
    float h = __fadd_rn(__fsub_rn(y2, y1), 1.f);
    return __fmul_rn(w, h);
}

// true iff box j must be suppressed by kept box i
// (double)ovr >= thresh for a float ovr  <=>  ovr >= F, F = the smallest float whose double value is >= thresh: the
// reference's double compare (nms.pyx:65-66) without a conversion and a 64-bit compare per pair.
__device__ __forceinline__ float thresh_as_float(double t) {
    float f = (float)t;                                   // round to nearest
    if ((double)f < t) f = nextafterf(f, INFINITY);
    return f;
}

__device__ __forceinline__ bool suppresses(const float4 &a, float area_a, const float4 &b, float area_b, float thresh_f) {
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
    return ovr >= thresh_f;
}

// (score desc, index desc): does detection j come before detection i ?
__device__ __forceinline__ bool precedes(float sj, int j, float si, int i) {
    return sj > si || (sj == si && j > i);
}

constexpr int RANK_THREADS = 256;
constexpr int RANK_TILE = 2048;

// monotone float -> uint key of the reference's comparison (`>` on float32; -0 == +0; NaN scores are not ordered by
// numpy either and are out of contract)
__device__ __forceinline__ unsigned rank_key(float s) {
    const unsigned b = __float_as_uint(s + 0.f);                // -0 -> +0
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// grid = (ceil(n/256), ceil(n/RANK_TILE)): block (bx, by) counts, for its 256 detections, how many of the
// by-th tile of detections precede each of them, and adds the partial count to rank[i].
// j precedes i  <=>  key_j > key_i, or key_j == key_i and j > i: one unsigned compare per pair -- `>=` for the
// elements behind i, `>` for those before it -- on keys staged once per tile, four per 16-byte shared load.
// With a single score tile (n <= RANK_TILE) the count is already the final rank: the thread moves its box to sorted
// position itself and the scatter launch is skipped (small problems are launch-latency bound).
__global__ void __launch_bounds__(RANK_THREADS)
nms_rank_kernel(const float *__restrict__ dets, int n, int *__restrict__ rank, float4 *__restrict__ boxes,
                float *__restrict__ areas, int *__restrict__ order, uint4 *__restrict__ zero, int n_zero) {
    __shared__ __align__(16) unsigned s_key[RANK_TILE];
    // the per-call state of the later kernels (removed | kept_bits | nkept | row_done | ctl: a few KB) is zeroed here instead of
    // by a memset of its own -- small problems are bound by the number of host-side launches
    if (blockIdx.x == 0 && blockIdx.y == 0)
        for (int k = threadIdx.x; k < n_zero; k += RANK_THREADS) zero[k] = make_uint4(0u, 0u, 0u, 0u);
    const int i = blockIdx.x * RANK_THREADS + threadIdx.x;
    const int t0 = blockIdx.y * RANK_TILE, tn = min(RANK_TILE, n - t0);
    for (int k = threadIdx.x; k < RANK_TILE; k += RANK_THREADS)
        s_key[k] = k < tn ? rank_key(dets[(size_t)(t0 + k) * 5 + 4]) : 0u;       // padding: key 0 precedes nothing (see below)
    __syncthreads();
    if (i >= n) return;
    const unsigned ki = rank_key(dets[(size_t)i * 5 + 4]);
    // elements with index < split compare with `>`, the others with `>=` (equal keys: the higher index comes first);
    // `a >= ki` is `a > ki - 1` and ki >= 1 for every real score (key 0 would be the NaN pattern 0xffffffff), and the
    // zero padding never counts: 0 > x is false for unsigned x.
    const int split = min(max(i + 1 - t0, 0), RANK_TILE);      // local index of the first element behind i
    const unsigned kge = ki - 1u;
    int cnt = 0;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(s_key);
    const int q_split = split >> 2;
#pragma unroll 4
    for (int q = 0; q < q_split; ++q) {
        const uint4 v = s4[q];
        cnt += (v.x > ki) + (v.y > ki) + (v.z > ki) + (v.w > ki);
    }
    if (q_split < RANK_TILE / 4) {                              // the group that straddles the split
        const uint4 v = s4[q_split];
        const int r = split & 3;
        cnt += (v.x > (r > 0 ? ki : kge)) + (v.y > (r > 1 ? ki : kge)) + (v.z > (r > 2 ? ki : kge)) + (v.w > kge);
    }
#pragma unroll 4
    for (int q = q_split + 1; q < RANK_TILE / 4; ++q) {
        const uint4 v = s4[q];
        cnt += (v.x > kge) + (v.y > kge) + (v.z > kge) + (v.w > kge);
    }
    // element i itself (local index split - 1 when it lies in this tile) was compared with `>`: not counted
    if (gridDim.y == 1) {
        const float *d = dets + (size_t)i * 5;
        const float4 b = make_float4(d[0], d[1], d[2], d[3]);
        boxes[cnt] = b;
        areas[cnt] = box_area(b.x, b.y, b.z, b.w);
        order[cnt] = i;
    } else if (cnt) {
        atomicAdd(rank + i, cnt);
    }
}

// Single-tile problems (n <= RANK_TILE), round 2: the kernel above gives one thread all 2048 compares of its detection and
// runs n / 256 = 8 CTAs -- 20 us at n = 2000, two fifths of the whole call.  Here eight threads share a detection (256 keys
// each, four per 16-byte shared load, combined by shuffles) and a CTA takes 32 detections: 63 CTAs instead of 8, each thread a
// few hundred instructions.  Same rule: j precedes i  <=>  key_j > key_i, or key_j == key_i and j > i.
constexpr int RANK1_SUB = 8;                         // threads per detection
constexpr int RANK1_DETS = RANK_THREADS / RANK1_SUB; // detections per CTA
__global__ void __launch_bounds__(RANK_THREADS)
nms_rank1_kernel(const float *__restrict__ dets, int n, float4 *__restrict__ boxes, float *__restrict__ areas, int *__restrict__ order,
                 uint4 *__restrict__ zero, int n_zero) {
    __shared__ __align__(16) unsigned s_key[RANK_TILE];
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < n_zero; k += RANK_THREADS) zero[k] = make_uint4(0u, 0u, 0u, 0u);      // see nms_rank_kernel
    for (int k = threadIdx.x; k < RANK_TILE; k += RANK_THREADS) s_key[k] = k < n ? rank_key(dets[(size_t)k * 5 + 4]) : 0u;
    __syncthreads();
    const int det = blockIdx.x * RANK1_DETS + threadIdx.x / RANK1_SUB, sub = threadIdx.x % RANK1_SUB;
    const int i = min(det, n - 1);                                // lanes past the end compute along (the shuffles need them)
    const unsigned ki = s_key[i], kge = ki - 1u;                  // `a >= ki` is `a > ki - 1`; ki >= 1 for every real score
    constexpr int PER = RANK_TILE / RANK1_SUB / 4;                // 16-byte groups per thread
    const uint4 *s4 = reinterpret_cast<const uint4 *>(s_key) + sub * PER;
    const int j0 = sub * PER * 4;
    int cnt = 0;
#pragma unroll 8
    for (int q = 0; q < PER; ++q) {
        const uint4 v = s4[q];
        const int j = j0 + 4 * q;                                 // elements j .. j + 3: `>` up to and including i itself, `>=` behind it
        cnt += (v.x > (j <= i ? ki : kge)) + (v.y > (j + 1 <= i ? ki : kge)) + (v.z > (j + 2 <= i ? ki : kge)) + (v.w > (j + 3 <= i ? ki : kge));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    if (sub == 0 && det < n) {
        const float *d = dets + (size_t)det * 5;
        const float4 b = make_float4(d[0], d[1], d[2], d[3]);
        boxes[cnt] = b;
        areas[cnt] = box_area(b.x, b.y, b.z, b.w);
        order[cnt] = det;
    }
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

// ---- bucket sort for large n ------------------------------------------------------------------------------------
// The rank kernel above compares every pair: n^2 = 4e8 key compares at n = 20 000, 60-70 us -- a fifth of the whole call.
// For n > RANK_TILE the order is found in two steps instead.  nms_bucket_kernel (one CTA) spreads the keys over BKT
// buckets that are linear in the uint key between the smallest and the largest key present, highest keys first
// (shared-memory histogram, scan, scatter of (key, index) pairs in bucket order); nms_bucket_rank_kernel then ranks every
// detection inside its own bucket only -- a few dozen compares with the same rule as above (key desc, index desc) -- and
// moves its box to sorted position.  The result is the exact order of the rank kernel; the cost depends on the score
// distribution only through the largest bucket (all scores equal: one bucket, the n^2 loop again).
constexpr int BKT = 2048;
constexpr int BKT_THREADS = 1024;
constexpr int BKT_MIN_N = RANK_TILE + 1;     // everything the single-tile rank kernel does not take

// ITEMS > 0: n <= ITEMS * BKT_THREADS and every thread keeps its keys in registers -- the scores are read from global
// memory once (strided 20-byte records: the one pass costs ~5 us on one SM) instead of once per phase; ITEMS == 0: any n.
template <int ITEMS>
__global__ void __launch_bounds__(BKT_THREADS)
nms_bucket_kernel(const float *__restrict__ dets, int n, uint2 *__restrict__ bpair, int *__restrict__ boff, int staged,
                  uint4 *__restrict__ zero, int n_zero) {
    for (int k = threadIdx.x; k < n_zero; k += BKT_THREADS) zero[k] = make_uint4(0u, 0u, 0u, 0u);      // see nms_rank_kernel
    // staged: dynamic shared memory holds n (key, index) pairs -- the scatter into bucket order goes there and leaves the
    // SM as one coalesced copy (40 000 scattered 4-byte stores from a single SM cost ~20 us, more than everything else)
    extern __shared__ uint2 s_stage[];
    __shared__ int s_hist[BKT];                               // counts -> exclusive offsets -> cursors
    __shared__ unsigned s_min[BKT_THREADS / 32], s_max[BKT_THREADS / 32];
    __shared__ int s_wsum[BKT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned kmin = 0xffffffffu, kmax = 0u;
    unsigned key[ITEMS > 0 ? ITEMS : 1];
    if (ITEMS > 0) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int i = tid + j * BKT_THREADS;
            key[j] = i < n ? rank_key(dets[(size_t)i * 5 + 4]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
            if (tid + j * BKT_THREADS < n) { kmin = min(kmin, key[j]); kmax = max(kmax, key[j]); }
    } else {
        for (int i = tid; i < n; i += BKT_THREADS) {
            const unsigned k = rank_key(dets[(size_t)i * 5 + 4]);
            kmin = min(kmin, k);
            kmax = max(kmax, k);
        }
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) { s_min[warp] = kmin; s_max[warp] = kmax; }
    for (int b = tid; b < BKT; b += BKT_THREADS) s_hist[b] = 0;
    __syncthreads();
    kmin = s_min[lane];                                       // 32 warps: one partial per lane
    kmax = s_max[lane];
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    int shift = 0;
    while (((kmax - kmin) >> shift) >= (unsigned)BKT) ++shift;           // <= 21 steps: (kmax - kmin) >> shift < BKT
    if (ITEMS > 0) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
            if (tid + j * BKT_THREADS < n) atomicAdd(&s_hist[(kmax - key[j]) >> shift], 1);
    } else {
        for (int i = tid; i < n; i += BKT_THREADS) atomicAdd(&s_hist[(kmax - rank_key(dets[(size_t)i * 5 + 4])) >> shift], 1);
    }
    __syncthreads();
    // exclusive scan of the BKT counts: two consecutive buckets per thread
    const int c0 = s_hist[2 * tid], c1 = s_hist[2 * tid + 1];
    const int incl = warp_incl_scan(c0 + c1, lane);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int t = s_wsum[lane];
        const int ti = warp_incl_scan(t, lane);
        s_wsum[lane] = ti - t;
    }
    __syncthreads();
    const int base = s_wsum[warp] + incl - (c0 + c1);
    s_hist[2 * tid] = base;
    s_hist[2 * tid + 1] = base + c0;
    boff[2 * tid] = base;
    boff[2 * tid + 1] = base + c0;
    if (tid == 0) { boff[BKT] = n; boff[BKT + 1] = shift; boff[BKT + 2] = (int)kmax; }
    __syncthreads();
    uint2 *dst = staged ? s_stage : bpair;
    if (ITEMS > 0) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int i = tid + j * BKT_THREADS;
            if (i < n) dst[atomicAdd(&s_hist[(kmax - key[j]) >> shift], 1)] = make_uint2(key[j], (unsigned)i);
        }
    } else {
        for (int i = tid; i < n; i += BKT_THREADS) {
            const unsigned k = rank_key(dets[(size_t)i * 5 + 4]);
            dst[atomicAdd(&s_hist[(kmax - k) >> shift], 1)] = make_uint2(k, (unsigned)i);
        }
    }
    if (staged) {
        __syncthreads();
        for (int e = tid; e < n; e += BKT_THREADS) bpair[e] = s_stage[e];
    }
}

__global__ void __launch_bounds__(256)
nms_bucket_rank_kernel(const float *__restrict__ dets, int n, const uint2 *__restrict__ bpair, const int *__restrict__ boff,
                       float4 *__restrict__ boxes, float *__restrict__ areas, int *__restrict__ order) {
    pdl_enter();
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= n) return;
    const uint2 me = bpair[e];
    const unsigned k = me.x;
    const int i = (int)me.y;
    const int shift = boff[BKT + 1];
    const unsigned kmax = (unsigned)boff[BKT + 2];
    const int b = (int)((kmax - k) >> shift);
    const int lo = boff[b], hi = boff[b + 1];
    int cnt = 0;
#pragma unroll 4
    for (int q = lo; q < hi; ++q) {                            // neighbouring threads share their bucket: broadcast loads
        const uint2 p = bpair[q];
        cnt += (int)(p.x > k) | ((int)(p.x == k) & (int)((int)p.y > i));
    }
    const int r = lo + cnt;
    const float *d = dets + (size_t)i * 5;
    const float4 bx = make_float4(d[0], d[1], d[2], d[3]);
    boxes[r] = bx;
    areas[r] = box_area(bx.x, bx.y, bx.z, bx.w);
    order[r] = i;
}

// Is the clamped intersection of a and b non-empty, i.e. w > 0 and h > 0 with the reference's float32 operations
// w = (min(a.x2, b.x2) - max(a.x1, b.x1)) + 1 ?  Two exact simplifications:
//   (1) fl(fl(d) + 1) > 0  <=>  fl(d) > -1: floats just above -1 are spaced 2^-24, so fl(d) + 1 is then an exactly
//       representable positive number, and fl(d) <= -1 gives a sum <= 0;
//   (2) rounding is monotone, so fl(min(a2, b2) - max(a1, b1)) is the minimum of the four fl(x2 - x1) combinations:
//       the test is  fl(a2 - b1) > -1 && fl(b2 - a1) > -1  for the pair, and  fl(a2 - a1) > -1, fl(b2 - b1) > -1  per box
//       (box_valid: checked once per box, an invalid column box is poisoned with x1 = +inf).
// Four subtractions and four compares per pair instead of four min/max, four add/sub and two compares plus guards.
__device__ __forceinline__ bool box_valid(const float4 &b) {
    return __fsub_rn(b.z, b.x) > -1.f && __fsub_rn(b.w, b.y) > -1.f;
}
__device__ __forceinline__ bool pair_intersects(const float4 &a, const float4 &b) {
    // branch-free: the minimum of the four differences against -1 (`&&` made the compiler emit a divergent branch
    // region per pair: 135 BSSY/BSYNC pairs in the unrolled tile loop)
    const float dx = fminf(__fsub_rn(a.z, b.x), __fsub_rn(b.z, a.x)), dy = fminf(__fsub_rn(a.w, b.y), __fsub_rn(b.w, a.y));
    return fminf(dx, dy) > -1.f;
}

// One 64-thread CTA per 64 x 64 tile of the upper triangle (linear block id -> (row tile, col tile)); thread =
// row.  Two phases per thread: a cheap exact intersection test over the 64 columns (boxes broadcast from shared
// memory), then the full IoU (IEEE division + double compare) only for the columns that intersect -- with an
// empty intersection the reference computes ovr = +-0 or NaN, which is >= thresh for no positive threshold,
// so those pairs can never be suppressed.  (A lanes-own-columns / ballot variant was measured 1.5x slower: it
// runs the division path for every row, because some lane of 32 random columns always intersects.)
constexpr int MASK_THREADS = 64;
constexpr int SCAN_THREADS = 1024;
constexpr int SUPER = 16;                      // column tiles per super-tile
constexpr int SUPER_UPDATERS = 48;             // CTAs of a launch that push the previous super-tile's kept rows

// id_base: the launch covers the blocks [id_base, id_base + gridDim.x) of the row-major triangle -- azn_nms launches
// the mask one super-row (16 row tiles = 1024 boxes) at a time so that the greedy pass of super-tile s can start as soon
// as ITS rows exist and runs concurrently with the mask of the rows behind it (see azn_nms).  A diagonal block also
// writes its transpose: diag_t[t][j] bit i <=> row i of tile t suppresses row j of tile t (i < j); the greedy pass
// resolves a tile from these columns (which kept rows suppress me?) in a few ballot rounds.
__global__ void __launch_bounds__(MASK_THREADS)
nms_mask_kernel(const float4 *__restrict__ boxes, const float *__restrict__ areas, int n, double thresh,
                u64 *__restrict__ mask, int col_tiles, long id_base, u64 *__restrict__ diag_t, int *__restrict__ row_done,
                u64 *__restrict__ sup_t) {
    // block id -> (rt, ct), rt <= ct, COLUMN-major over the triangle (column ct starts at id0(ct) = ct (ct + 1) / 2): the
    // cheap columns come first, so the greedy pass -- which consumes the mask column by column -- starts at once and
    // is nearly done when the last, most expensive column super-block lands
    const long id = id_base + blockIdx.x;
    int ct = (int)((sqrt(8.0 * (double)id + 1.0) - 1.0) * 0.5);
    ct = max(0, min(ct, col_tiles - 1));
    while (ct > 0 && (long)ct * (ct + 1) / 2 > id) --ct;
    while ((long)(ct + 1) * (ct + 2) / 2 <= id) ++ct;
    const int rt = (int)(id - (long)ct * (ct + 1) / 2);
    __shared__ float4 cb[64];
    __shared__ float ca[64];
    __shared__ u64 s_d[64];
    const int cn = min(64, n - ct * 64);
    if ((int)threadIdx.x < cn) {
        float4 b = boxes[ct * 64 + threadIdx.x];
        // a box whose own extent is empty under the reference's arithmetic ((x2 - x1) + 1 <= 0) intersects nothing:
        // poison it so that the pair test below fails without a per-pair check (only when the test is used at all)
        if (thresh > 0.0 && !box_valid(b)) b.x = INFINITY;
        cb[threadIdx.x] = b;
        ca[threadIdx.x] = areas[ct * 64 + threadIdx.x];
    }
    __syncthreads();
    const float thresh_f = thresh_as_float(thresh);
    u64 *dst = mask + ((size_t)rt * col_tiles + ct) * 64 + threadIdx.x;      // blocked: [row tile][col tile][64 rows]
    const int row = rt * 64 + threadIdx.x;
    u64 bits = 0;                                                            // rows past the end: no suppression bits
    if (row < n) {
        const float4 rb = boxes[row];
        const float ra = areas[row];
        const int start = (rt == ct) ? threadIdx.x + 1 : 0;
        unsigned cand_lo = 0, cand_hi = 0;
        if (thresh > 0.0) {
            if (!box_valid(rb)) {
                // nothing intersects an empty row box
            } else if (rt != ct && cn == 64) {               // the bulk of the triangle: full off-diagonal tiles, no guards
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    cand_lo |= pair_intersects(rb, cb[k]) ? (1u << k) : 0u;
                    cand_hi |= pair_intersects(rb, cb[k + 32]) ? (1u << k) : 0u;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    cand_lo |= ((k >= start) & (k < cn) & pair_intersects(rb, cb[min(k, 63)])) ? (1u << k) : 0u;
                    cand_hi |= ((k + 32 >= start) & (k + 32 < cn) & pair_intersects(rb, cb[k + 32])) ? (1u << k) : 0u;
                }
            }
        } else {
            const u64 all = (cn == 64 ? ~0ull : ((1ull << cn) - 1ull)) & (start >= 64 ? 0ull : (~0ull << start));
            cand_lo = (unsigned)all;
            cand_hi = (unsigned)(all >> 32);
        }
        // Area-ratio filter between the two phases: the clamped intersection is never larger than either box
        // (w <= w_a, h <= h_a and every operation is monotone), so inter <= m = min(area_a, area_b) exactly, the union
        // fl(fl(a + b) - inter) >= max(area_a, area_b) (1 - 2^-22), and ovr <= (m / M)(1 + 2^-21).  A pair with
        // m < thresh_f * M * (1 - 2^-19) therefore cannot reach the threshold: it is dropped here for two multiplies,
        // and the division loop below runs over a mask that is several times sparser (boxes of unlike size rarely
        // suppress each other) -- which matters because that loop runs at the pace of the slowest lane of the warp.
        if (thresh > 0.0) {
            const float tq = __fmul_rn(thresh_f, 0.999998f);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                unsigned c = half ? cand_hi : cand_lo, keep = c;
                while (c) {
                    const int k = __ffs((int)c) - 1;
                    c &= c - 1;
                    const float ak = ca[half * 32 + k];
                    if (fminf(ra, ak) < __fmul_rn(tq, fmaxf(ra, ak))) keep &= ~(1u << k);
                }
                if (half) cand_hi = keep; else cand_lo = keep;
            }
        }
        u64 cand = (u64)cand_lo | ((u64)cand_hi << 32);
        while (cand) {
            const int k = __ffsll((long long)cand) - 1;
            cand &= cand - 1;
            if (suppresses(rb, ra, cb[k], ca[k], thresh_f)) bits |= 1ull << k;
        }
    }
    *dst = bits;
    if (rt / SUPER == ct / SUPER) {                                          // block-uniform: a tile of a diagonal super-block
        // its transpose, for the greedy pass: sup_t[ct][rt % 16][j] bit i <=> row i of tile rt suppresses row j of tile ct
        // (the block-wise pass keeps, per row, the 16 words that say which rows of its super-tile suppress it)
        s_d[threadIdx.x] = bits;
        __syncthreads();
        u64 col = 0;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) col |= ((s_d[i] >> threadIdx.x) & 1ull) << i;
        if (rt == ct) diag_t[(size_t)rt * 64 + threadIdx.x] = col;
        sup_t[((size_t)ct * SUPER + rt % SUPER) * 64 + threadIdx.x] = col;
    }
    // publish: one more tile of super-column ct / 16 is in memory (the greedy pass polls these counters)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(row_done + ct / 16, 1);
    }
}


__device__ __forceinline__ u64 warp_or(u64 v) {
    // redux.sync.or.b32: one instruction per half instead of a 5-step shuffle butterfly (ten dependent SHFLs for
    // 64 bits) -- this sits on the serial path of the greedy pass once per tile
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v), hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
    return ((u64)hi << 32) | lo;
}
// OR of the words of one 64-row block whose row bit is set in `kb`; lane l holds rows 2l and 2l+1 (one 16-byte load)
__device__ __forceinline__ u64 select2(const ulonglong2 &w, u64 kb, int lane) {
    return (((kb >> (2 * lane)) & 1ull) ? w.x : 0ull) | (((kb >> (2 * lane + 1)) & 1ull) ? w.y : 0ull);
}

// ---- mask kernel, second version (thresh > 0): conservative packed-half screen + warp-balanced exact evaluation ----------
// ncu of nms_mask_kernel at N = 20 000 (profiles/r2_ncu_nms_mask_v1.csv): 1757 warp instructions per warp and tile, of which 650
// the float32 intersection screen of the 64 columns (10 per pair), 390 the per-lane area-ratio loops (12 trips at 8.6 live
// lanes), 420 the per-lane division loop (9 trips at 7.7 live lanes): most of the time goes to loops that run at the pace
// of the lane with the most candidates.  This version
//   (1) screens two columns per instruction in packed half precision with DIRECTED rounding, so that the screen can only
//       err towards "candidate": the exact test fl(a.x2 - b.x1) > -1 implies a.x2 + 1 > b.x1 in the reals, hence
//       RU(a.x2 + 1) > RD(b.x1) for any rounding up / down to half (+-inf included); likewise the other three
//       differences.  The area-ratio condition min >= t' max that the exact filter applies becomes |log2 A - log2 B| <= L
//       with L = -log2 t' plus 0.05 of slack for the half rounding of the logs (their error is below 0.02).  Five packed
//       compares and three logic operations per TWO pairs; ~3 % of the pairs come out as candidates;
//   (2) gathers the candidates of the warp's 32 rows into one list (prefix sum of the per-lane counts) and evaluates them
//       32 at a time, one per lane: the exact intersection test, the exact area filter and the IEEE division of the
//       first version, in the same order -- so the bits are the same -- but in ~T/32 trips with all lanes busy.
// Column p and column p + 32 share a half2: accumulator bit p % 16 of the low / high half.
constexpr int MASK_LIST = 256;                 // candidate list entries per warp and pass

__device__ __forceinline__ float bump_up(float v) { return nextafterf(v, INFINITY); }
__device__ __forceinline__ unsigned pack_h2(__half lo, __half hi) {
    return (unsigned)__half_as_ushort(lo) | ((unsigned)__half_as_ushort(hi) << 16);
}
__device__ __forceinline__ __half2 as_h2(unsigned u) { return *reinterpret_cast<const __half2 *>(&u); }

__global__ void __launch_bounds__(MASK_THREADS)
nms_mask2_kernel(const float4 *__restrict__ boxes, const float *__restrict__ areas, int n, double thresh,
                 u64 *__restrict__ mask, int col_tiles, u64 *__restrict__ diag_t, int *__restrict__ row_done, u64 *__restrict__ sup_t) {
    const long id = blockIdx.x;                                               // column-major over the triangle, as in version 1
    int ct = (int)((sqrt(8.0 * (double)id + 1.0) - 1.0) * 0.5);
    ct = max(0, min(ct, col_tiles - 1));
    while (ct > 0 && (long)ct * (ct + 1) / 2 > id) --ct;
    while ((long)(ct + 1) * (ct + 2) / 2 <= id) ++ct;
    const int rt = (int)(id - (long)ct * (ct + 1) / 2);
    __shared__ float4 cb[64], rbx[64];
    __shared__ float ca[64], rba[64];
    __shared__ __align__(16) unsigned hb[32][4];                              // pair p: x1 | y1 | RU(x2 + 1) | RU(y2 + 1), halves = columns p, p + 32
    __shared__ unsigned hl[32];                                               // pair p: log2(area)
    __shared__ __align__(8) unsigned s_bits[64][2];
    __shared__ unsigned short s_list[MASK_THREADS / 32][MASK_LIST];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cn = min(64, n - ct * 64);
    const int row = rt * 64 + tid;
    const float thresh_f = thresh_as_float(thresh);
    {
        // column tid: exact copy + its packed-half screen values (an absent or empty column can match nothing: x1 = +inf)
        float4 b = make_float4(INFINITY, 0.f, 0.f, 0.f);
        float ar = 0.f;
        if (tid < cn) {
            b = boxes[ct * 64 + tid];
            ar = areas[ct * 64 + tid];
            if (!box_valid(b)) b.x = INFINITY;
        }
        cb[tid] = b;
        ca[tid] = ar;
        unsigned short *h16 = reinterpret_cast<unsigned short *>(&hb[tid & 31][0]) + (tid >> 5);
        h16[0] = __half_as_ushort(__float2half_rd(b.x));
        h16[2] = __half_as_ushort(__float2half_rd(b.y));
        h16[4] = __half_as_ushort(__float2half_ru(bump_up(__fadd_rn(b.z, 1.f))));
        h16[6] = __half_as_ushort(__float2half_ru(bump_up(__fadd_rn(b.w, 1.f))));
        reinterpret_cast<unsigned short *>(&hl[tid & 31])[tid >> 5] = __half_as_ushort(__float2half_rn(__log2f(ar)));
        s_bits[tid][0] = 0u;
        s_bits[tid][1] = 0u;
    }
    float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
    float ra = 0.f;
    bool row_ok = false;
    if (row < n) {
        rb = boxes[row];
        ra = areas[row];
        row_ok = box_valid(rb);
    }
    rbx[tid] = rb;
    rba[tid] = ra;
    __syncthreads();
    // (1) the screen
    u64 cand = 0ull;
    if (row_ok) {
        const __half a1x = __float2half_rd(rb.x), a1y = __float2half_rd(rb.y);
        const __half a2x = __float2half_ru(bump_up(__fadd_rn(rb.z, 1.f))), a2y = __float2half_ru(bump_up(__fadd_rn(rb.w, 1.f)));
        const __half2 A1x = __halves2half2(a1x, a1x), A1y = __halves2half2(a1y, a1y), A2x = __halves2half2(a2x, a2x), A2y = __halves2half2(a2y, a2y);
        const __half la = __float2half_rn(__log2f(ra));
        const __half2 LA = __halves2half2(la, la);
        const float Lf = -__log2f(__fmul_rn(thresh_f, 0.999998f)) + 0.05f;
        const __half lh = __float2half_ru(fmaxf(Lf, 0.05f));
        const __half2 LH = __halves2half2(lh, lh);
        unsigned acc0 = 0u, acc1 = 0u;
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const uint4 c = *reinterpret_cast<const uint4 *>(&hb[p][0]);
            const unsigned m1 = __hgt2_mask(A2x, as_h2(c.x)), m2 = __hgt2_mask(as_h2(c.z), A1x);
            const unsigned m3 = __hgt2_mask(A2y, as_h2(c.y)), m4 = __hgt2_mask(as_h2(c.w), A1y);
            const unsigned m5 = __hle2_mask(__habs2(__hsub2(LA, as_h2(hl[p]))), LH);
            const unsigned hit = (m1 & m2 & m3) & (m4 & m5) & ((1u << (p & 15)) * 0x00010001u);
            if (p < 16) acc0 |= hit; else acc1 |= hit;
        }
        const unsigned lo = (acc0 & 0xffffu) | (acc1 << 16), hi = (acc0 >> 16) | (acc1 & 0xffff0000u);
        cand = (u64)lo | ((u64)hi << 32);
        if (rt == ct) cand &= tid >= 63 ? 0ull : (~0ull << (tid + 1));          // the diagonal tile: columns behind the row only
    }
    // (2) the warp's candidates, 32 at a time
    const int cnt = __popcll(cand);
    const int incl = warp_incl_scan(cnt, lane);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const float tq = __fmul_rn(thresh_f, 0.999998f);
    unsigned short *list = s_list[warp];
    for (int base = 0; base < total; base += MASK_LIST) {
        if (total <= MASK_LIST) {                                            // warp-uniform, the usual case: everything fits, no window checks
            unsigned short *dst = list + (incl - cnt);
            const unsigned tag = (unsigned)lane << 6;
            for (unsigned c = (unsigned)cand; c; c &= c - 1) *dst++ = (unsigned short)(tag | (unsigned)(__ffs((int)c) - 1));
            for (unsigned c = (unsigned)(cand >> 32); c; c &= c - 1) *dst++ = (unsigned short)(tag | 32u | (unsigned)(__ffs((int)c) - 1));
        } else {
            u64 c = cand;
            int e = incl - cnt - base;                                       // position of this lane's first candidate in the window
            while (c) {
                const int k = __ffsll((long long)c) - 1;
                c &= c - 1;
                if (e >= 0 && e < MASK_LIST) list[e] = (unsigned short)((lane << 6) | k);
                ++e;
            }
        }
        __syncwarp();
        const int m = min(MASK_LIST, total - base);
        for (int e = lane; e < m; e += 32) {
            const int ent = list[e], r = warp * 32 + (ent >> 6), k = ent & 63;
            const float4 a = rbx[r], b = cb[k];
            if (!pair_intersects(a, b)) continue;
            const float aa = rba[r], ab = ca[k];
            if (fminf(aa, ab) < __fmul_rn(tq, fmaxf(aa, ab))) continue;
            if (suppresses(a, aa, b, ab, thresh_f)) atomicOr(&s_bits[r][k >> 5], 1u << (k & 31));
        }
        __syncwarp();
    }
    __syncthreads();
    const u64 bits = (u64)s_bits[tid][0] | ((u64)s_bits[tid][1] << 32);
    mask[((size_t)rt * col_tiles + ct) * 64 + tid] = bits;                   // blocked: [row tile][col tile][64 rows]
    if (rt / SUPER == ct / SUPER) {                                          // block-uniform: see version 1
        u64 *s_d = reinterpret_cast<u64 *>(&s_bits[0][0]);
        u64 col = 0;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) col |= ((s_d[i] >> tid) & 1ull) << i;
        if (rt == ct) diag_t[(size_t)rt * 64 + tid] = col;
        sup_t[((size_t)ct * SUPER + rt % SUPER) * 64 + tid] = col;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(row_done + ct / 16, 1);
    }
}

// Bulk update (left-looking), the CTAs 1 .. gridDim.x - 1 of a greedy launch: kept rows of ALL super-tiles before s ->
// removed[] of the columns of super-tile s+1 (the rows of super-tile s itself reach them through CTA 0's push-ahead).
__device__ __forceinline__ void nms_bulk_update(const u64 *__restrict__ mask, int col_tiles, int s_idx, u64 *__restrict__ removed,
                                                const u64 *__restrict__ kept_bits, const int *__restrict__ row_done) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T0 = s_idx * SUPER;
    pdl_enter();
    if (s_idx == 0) return;
    const int pt0 = 0, c0 = T0 + SUPER;
    const int ncols = min(SUPER, col_tiles - c0);
    if (ncols <= 0) return;
    if (tid == 0) {                                          // the mask of super-column s+1 must have landed
        int need = 0;
        for (int c = c0; c < c0 + ncols; ++c) need += c + 1;
        const volatile int *flag = row_done + s_idx + 1;
        while (*flag < need) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
    const long nblk = (long)T0 * ncols;
    const long wid = (long)(blockIdx.x - 1) * (SCAN_THREADS / 32) + warp, nw = (long)(gridDim.x - 1) * (SCAN_THREADS / 32);
    for (long k0 = wid; k0 < nblk; k0 += 4 * nw) {
        ulonglong2 w[4];
        u64 kb[4];
        int col[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                       // four independent 512-byte blocks in flight per warp
            const long k = k0 + u * nw;
            col[u] = -1;
            kb[u] = 0ull;
            w[u] = make_ulonglong2(0ull, 0ull);
            if (k < nblk) {
                const int tp = (int)(k / ncols), c = c0 + (int)(k - (long)tp * ncols);
                kb[u] = kept_bits[pt0 + tp];
                col[u] = c;
                if (kb[u]) w[u] = __ldcg(reinterpret_cast<const ulonglong2 *>(mask + ((size_t)(pt0 + tp) * col_tiles + c) * 64) + lane);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (col[u] < 0 || kb[u] == 0ull) continue;      // warp-uniform
            const u64 v = warp_or(select2(w[u], kb[u], lane));
            if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(removed + col[u]), (unsigned long long)v);
        }
    }
}

// Launch s of the greedy pass (see the header).  mask is blocked [row tile][col tile][64]; `removed[c]` collects
// the suppression bits of column tile c from all kept rows of super-tiles that are at least two behind c's (written
// by the bulk updaters of the launch just before c's predecessor -- left-looking: complete exactly when c's turn comes);
// `kept_bits[t]` is the kept mask of row tile t; `nkept` the running number of kept boxes.
__global__ void __launch_bounds__(SCAN_THREADS)
nms_super_kernel(const u64 *__restrict__ mask, const u64 *__restrict__ diag_t, const int *__restrict__ order, int n, int col_tiles, int s_idx,
                 u64 *__restrict__ removed, u64 *__restrict__ kept_bits, int *__restrict__ nkept_ptr,
                 int64_t *__restrict__ keep, int32_t *__restrict__ keep_count, int last, const int *__restrict__ row_done) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T0 = s_idx * SUPER;
    if (blockIdx.x > 0) {
        nms_bulk_update(mask, col_tiles, s_idx, removed, kept_bits, row_done);
        return;
    }
    // ---------------- CTA 0: resolve super-tile s
#ifdef AZN_NMS_TRACE
    const long long tt0 = clock64();
#endif
    extern __shared__ u64 s_blk[];                               // [SUPER][SUPER][64] diagonal super-block
    __shared__ u64 s_col[SUPER], s_kept[SUPER];
    __shared__ int s_order[SUPER * 64];
    const int nt = min(SUPER, col_tiles - T0);
    // The mask kernel runs concurrently (other stream / SM partition) and publishes, per super-COLUMN, how many of its
    // tiles are in memory; CTA 0 needs super-column s (its diagonal blocks) and s+1 (push-ahead), the bulk updaters
    // super-column s+1.  The sort order was written before the first launch.  So the
    // launch may stage while the previous one is still resolving (PDL), and only then waits for that one's results.
    const int c0n = T0 + SUPER, ncols_next = max(0, min(SUPER, col_tiles - c0n));       // the next super-tile's columns
    if (tid == 0) {
        int need = 0;
        for (int c = T0; c < T0 + nt; ++c) need += c + 1;
        const volatile int *flag = row_done + s_idx;
        while (*flag < need) __nanosleep(64);
        if (ncols_next > 0) {                                // push-ahead (below) reads this super-tile's rows of super-column s+1
            need = 0;
            for (int c = c0n; c < c0n + ncols_next; ++c) need += c + 1;
            flag = row_done + s_idx + 1;
            while (*flag < need) __nanosleep(64);
        }
        __threadfence();
    }
    __syncthreads();
    if (T0 * 64 + tid < n) s_order[tid] = order[T0 * 64 + tid];
    // A: stage the upper triangle of the diagonal super-block (16-byte cp.async chunks, 32 per 64-row block)
    for (int ch = tid; ch < SUPER * SUPER * 32; ch += SCAN_THREADS) {
        const int blk = ch >> 5, tp = blk / SUPER, b = blk - tp * SUPER;
        if (tp <= b && b < nt) {                                 // slot (b, b) holds the TRANSPOSED diagonal block of tile b
            const u64 *src = (tp == b ? diag_t + (size_t)(T0 + b) * 64 : mask + ((size_t)(T0 + tp) * col_tiles + (T0 + b)) * 64) + (ch & 31) * 2;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s_blk + (size_t)blk * 64 + (ch & 31) * 2)), "l"(src) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pdl_enter();
    if (tid < SUPER) {
        // complete: the bulk updaters of the previous launch pushed the kept rows of super-tiles < s-1, its CTA 0 those of
        // super-tile s-1 (push-ahead, below) -- there is no update left to do before the first tile can be resolved
        s_col[tid] = tid < nt ? removed[T0 + tid] : 0ull;
        s_kept[tid] = 0ull;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#ifdef AZN_NMS_TRACE
    const long long tt1 = clock64();
#endif
    // The 16 tiles in order.  Iteration b: warp 0 resolves tile b; warps 1..31 gather, for column b+1, the words of
    // the rows kept in tiles 0..b-1 of this super-tile; warp 0 adds tile b's own rows once it knows them.
    int nkept = *nkept_ptr;
    const int push_col = (warp >= 16 && warp - 16 < ncols_next) ? c0n + (warp - 16) : -1;       // warp-uniform
    ulonglong2 pend_w = make_ulonglong2(0ull, 0ull);
    u64 pend_kb = 0ull;
    for (int b = 0; b < nt; ++b) {
        if (warp == 0) {
            const int row0 = (T0 + b) * 64;
            // Lane l owns rows l and l+32 and holds their columns of the diagonal block: t_lo / t_hi bit i <=> row i
            // (i < own row) suppresses it.  Rounds over the undecided set U: a row with a KEPT suppressor is removed,
            // a row with no undecided suppressor left is kept; the lowest undecided row always decides, and
            // independent rows decide together (typically 2-4 rounds per tile instead of one step per kept row).
            const u64 *diag = s_blk + (size_t)(b * SUPER + b) * 64;
            const u64 t_lo = diag[lane], t_hi = diag[lane + 32];
            const int rows = min(64, n - row0);
            u64 und = ~s_col[b] & (rows == 64 ? ~0ull : ((1ull << rows) - 1ull));       // warp-uniform
            u64 kept = 0;
            while (und) {
                const bool u0 = (und >> lane) & 1ull, u1 = (und >> (lane + 32)) & 1ull;
                const bool rem0 = u0 && (t_lo & kept) != 0ull, rem1 = u1 && (t_hi & kept) != 0ull;
                const bool ok0 = u0 && !rem0 && (t_lo & und) == 0ull, ok1 = u1 && !rem1 && (t_hi & und) == 0ull;
                const u64 new_kept = (u64)__ballot_sync(0xffffffffu, ok0) | ((u64)__ballot_sync(0xffffffffu, ok1) << 32);
                const u64 new_rem = (u64)__ballot_sync(0xffffffffu, rem0) | ((u64)__ballot_sync(0xffffffffu, rem1) << 32);
                kept |= new_kept;
                und &= ~(new_kept | new_rem);
            }
            if (b + 1 < nt) {                                 // tile b's rows against column b+1
                const u64 *nb = s_blk + (size_t)(b * SUPER + b + 1) * 64;
                const u64 v = warp_or((((kept >> lane) & 1ull) ? nb[lane] : 0ull) | (((kept >> (lane + 32)) & 1ull) ? nb[lane + 32] : 0ull));
                if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(&s_col[b + 1]), (unsigned long long)v);
            }
            if ((kept >> lane) & 1ull) keep[nkept + __popcll(kept & ((1ull << lane) - 1ull))] = s_order[b * 64 + lane];
            if ((kept >> (lane + 32)) & 1ull) keep[nkept + __popcll(kept & ((1ull << (lane + 32)) - 1ull))] = s_order[b * 64 + 32 + lane];
            if (lane == 0) { s_kept[b] = kept; kept_bits[T0 + b] = kept; }
            nkept += __popcll(kept);
        } else {
            // Push-ahead: the kept rows of tile b-1 go into the `removed` words of the NEXT super-tile's columns while
            // warp 0 resolves tile b (warps 16..31, one column each; the 512-byte block is loaded in one iteration and
            // folded in the next, so its L2 latency never sits on the per-tile barrier).  The next launch then finds its
            // columns complete and starts resolving at once: no "urgent" update on the serial path.
            if (push_col >= 0) {
                if (pend_kb) {
                    const u64 v = warp_or(select2(pend_w, pend_kb, lane));
                    if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(removed + push_col), (unsigned long long)v);
                    pend_kb = 0ull;
                }
                if (b >= 1) {
                    const u64 kb = s_kept[b - 1];
                    if (kb) {
                        pend_w = __ldcg(reinterpret_cast<const ulonglong2 *>(mask + ((size_t)(T0 + b - 1) * col_tiles + push_col) * 64) + lane);
                        pend_kb = kb;
                    }
                }
            }
            if (b + 1 < nt && b >= 1) {
                u64 acc = 0;
                for (int p = tid - 32; p < b * 64; p += SCAN_THREADS - 32) {
                    const int tp = p >> 6, i = p & 63;
                    if ((s_kept[tp] >> i) & 1ull) acc |= s_blk[(size_t)(tp * SUPER + b + 1) * 64 + i];
                }
                acc = warp_or(acc);
                if (lane == 0 && acc) atomicOr(reinterpret_cast<unsigned long long *>(&s_col[b + 1]), (unsigned long long)acc);
            }
        }
        __syncthreads();
    }
    if (push_col >= 0) {                                     // the block still in flight, then the last tile's rows
        if (pend_kb) {
            const u64 v = warp_or(select2(pend_w, pend_kb, lane));
            if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(removed + push_col), (unsigned long long)v);
        }
        const u64 kb = s_kept[nt - 1];
        if (kb) {
            const ulonglong2 w = __ldcg(reinterpret_cast<const ulonglong2 *>(mask + ((size_t)(T0 + nt - 1) * col_tiles + push_col) * 64) + lane);
            const u64 v = warp_or(select2(w, kb, lane));
            if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(removed + push_col), (unsigned long long)v);
        }
    }
    if (tid == 0) {
        *nkept_ptr = nkept;
        if (last) *keep_count = nkept;
#ifdef AZN_NMS_TRACE
        printf("nms super %d: stage %lld cycles, 16 tiles %lld cycles\n", s_idx, tt1 - tt0, clock64() - tt1);
#endif
    }
}

// Block-wise greedy launch (default): CTA 0 resolves ALL 1024 rows of super-tile s together instead of its 16 tiles one
// after the other.  Thread j owns row j of the super-tile and keeps, in registers, the (up to) 16 words that say which rows
// of the super-tile suppress it (sup_t, written transposed by the mask kernel); the kept set K and the undecided set U live
// in shared memory, 16 words each.  A round: a row with a KEPT suppressor is removed, a row with no undecided suppressor
// left is kept -- the same rule as the per-tile ballot rounds of nms_super_kernel, applied to the whole super-tile, so
// rows in different tiles decide in the same round and the per-tile barrier / column hand-over (16 x ~0.7 us) disappears:
// a super-tile takes as many rounds as its longest chain of "suppressed only by a row that is itself still undecided"
// (a handful for detection boxes; the worst case, 1024 rows that each depend on their predecessor, is 1024 rounds --
// correct, merely as slow as the serial scan).  The lowest undecided row always decides, so the loop terminates.  Then the
// kept rows are compacted into the keep list (score order) and pushed into the `removed` words of the next super-tile's
// columns; CTAs 1.. are the same bulk updaters as before.
__global__ void __launch_bounds__(SCAN_THREADS)
nms_block_kernel(const u64 *__restrict__ mask, const u64 *__restrict__ sup_t, const int *__restrict__ order, int n, int col_tiles, int s_idx,
                 u64 *__restrict__ removed, u64 *__restrict__ kept_bits, int *__restrict__ nkept_ptr,
                 int64_t *__restrict__ keep, int32_t *__restrict__ keep_count, int last, const int *__restrict__ row_done) {
    if (blockIdx.x > 0) {
        nms_bulk_update(mask, col_tiles, s_idx, removed, kept_bits, row_done);
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T0 = s_idx * SUPER;
    const int nt = min(SUPER, col_tiles - T0);
    const int c0n = T0 + SUPER, ncols_next = max(0, min(SUPER, col_tiles - c0n));       // the next super-tile's columns
    __shared__ u64 s_K[SUPER], s_U[SUPER], s_push[SUPER];
    __shared__ int s_pre[SUPER + 1];
    if (tid == 0) {                                          // the mask of super-columns s (this block) and s+1 (push) must have landed
        int need = 0;
        for (int c = T0; c < T0 + nt; ++c) need += c + 1;
        const volatile int *flag = row_done + s_idx;
        while (*flag < need) __nanosleep(64);
        if (ncols_next > 0) {
            need = 0;
            for (int c = c0n; c < c0n + ncols_next; ++c) need += c + 1;
            flag = row_done + s_idx + 1;
            while (*flag < need) __nanosleep(64);
        }
        __threadfence();
    }
    __syncthreads();
    // the row's suppressor words (rows of tiles 0 .. c of this super-tile; tile c's own word only has bits below the row) and
    // its words of the next super-tile's columns: all independent loads, one L2 round trip, issued before the previous
    // launch's results are awaited (they depend on the mask alone)
    const int c = tid >> 6, jj = tid & 63;
    const int row = T0 * 64 + tid;
    const bool live = c < nt && row < n;
    u64 col[SUPER];
#pragma unroll
    for (int r = 0; r < SUPER; ++r) col[r] = (live && r <= c) ? __ldcg(sup_t + ((size_t)(T0 + c) * SUPER + r) * 64 + jj) : 0ull;
    // the rows' words of the next super-tile's columns go to shared memory ([row tile][column][64 rows], 16-byte cp.async
    // chunks): 32 more registers per thread would not fit next to col[] at 1024 threads per CTA
    extern __shared__ u64 s_nxt[];
    for (int ch = tid; ch < nt * ncols_next * 32; ch += SCAN_THREADS) {
        const int blk = ch >> 5, tp = blk / ncols_next, k = blk - tp * ncols_next;
        const u64 *src = mask + ((size_t)(T0 + tp) * col_tiles + c0n + k) * 64 + (ch & 31) * 2;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s_nxt + ((size_t)tp * SUPER + k) * 64 + (ch & 31) * 2)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int my_order = live ? order[row] : 0;
    pdl_enter();                                             // the previous launch: removed[] of our columns, nkept, kept_bits
    if (tid < SUPER) {
        const int rows = min(64, n - (T0 + tid) * 64);
        const u64 valid = tid < nt ? (rows >= 64 ? ~0ull : ((1ull << max(rows, 0)) - 1ull)) : 0ull;
        s_U[tid] = tid < nt ? (~__ldcg(removed + T0 + tid) & valid) : 0ull;
        s_K[tid] = 0ull;
        s_push[tid] = 0ull;
    }
    __syncthreads();
    bool und = live && ((s_U[c] >> jj) & 1ull);
    bool kept = false;
    for (;;) {
        bool rem = false, ok = false;
        if (und) {
            u64 hit = 0ull, pend = 0ull;
#pragma unroll
            for (int r = 0; r < SUPER; ++r) { hit |= col[r] & s_K[r]; pend |= col[r] & s_U[r]; }
            rem = hit != 0ull;
            ok = !rem && pend == 0ull;
        }
        const unsigned bk = __ballot_sync(0xffffffffu, ok), br = __ballot_sync(0xffffffffu, rem);
        __syncthreads();                                     // every thread has read K / U of this round
        if (lane == 0 && (bk | br)) {                        // a warp owns one 32-bit half of its tile's words
            unsigned *K32 = reinterpret_cast<unsigned *>(s_K) + (warp >> 1) * 2 + (warp & 1);
            unsigned *U32 = reinterpret_cast<unsigned *>(s_U) + (warp >> 1) * 2 + (warp & 1);
            *K32 |= bk;
            *U32 &= ~(bk | br);
        }
        kept |= ok;
        und = und && !rem && !ok;
        if (!__syncthreads_or(und ? 1 : 0)) break;
    }
    // keep list, score order: kept rows before this one = the kept rows of earlier tiles + the lower bits of its own tile
    if (tid == 0) {
        int run = *nkept_ptr;
        for (int t = 0; t < SUPER; ++t) { s_pre[t] = run; run += __popcll(s_K[t]); }
        s_pre[SUPER] = run;
    }
    if (tid < nt) kept_bits[T0 + tid] = s_K[tid];
    // push-ahead: kept rows of this super-tile -> removed[] of the next super-tile's columns
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int k = 0; k < ncols_next; ++k) {
        const u64 v = warp_or(kept ? s_nxt[((size_t)c * SUPER + k) * 64 + jj] : 0ull);
        if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(&s_push[k]), (unsigned long long)v);
    }
    __syncthreads();
    if (kept) keep[s_pre[c] + __popcll(s_K[c] & ((1ull << jj) - 1ull))] = my_order;
    if (tid < ncols_next && s_push[tid]) atomicOr(reinterpret_cast<unsigned long long *>(removed + c0n + tid), (unsigned long long)s_push[tid]);
    if (tid == 0) {
        *nkept_ptr = s_pre[SUPER];
        if (last) *keep_count = s_pre[SUPER];
    }
}

// Persistent greedy pass (A/B variant, azn_nms_tune(64); NOT the default: at N = 20 000 it measured 0.233-0.244 ms against
// 0.208-0.219 ms for one PDL launch per super-tile -- ~10 us per super-tile instead of ~6: the flag hand-overs through L2
// (resolver -> updaters -> resolver, each a poll with a sleep) cost more than the hardware hand-over between programmatic
// dependent launches, whose successor has its operands fetched before the predecessor exits).
// ONE launch for the whole chain instead of one per super-tile.  CTA 0 runs the
// block-wise rounds of nms_block_kernel for the super-tiles in a loop, the bulk updaters loop over their target
// super-columns.  What the kernel boundaries did is done by flags in global memory (ctl, zeroed per call):
//   done[s]  = 2           a resolver has published kept_bits / nkept of super-tile s and pushed its kept rows into removed[]
//                          of super-column s + 1      (the other resolver waits for it; so do the updaters of target s + 2)
//   upd[s]  += 1 per CTA   an updater has pushed the kept rows of super-tiles <= s - 2 into removed[] of super-column s
//                                                                                        (the resolver of s waits for all)
// CTAs 0 and 1 alternate as resolvers (even / odd super-tiles), so that the mask words of super-tile s + 1 are fetched while
// s is being resolved; CTAs 2.. are the updaters.  The CTAs spin on each other: the grid (<= 8 CTAs of 1024 threads) is always
// co-resident eventually -- nothing it waits for waits for it -- and every spin has a time-out that raises ctl.abort
// (keep_count = -1) instead of hanging the GPU.
struct ChainCtl {
    int *done, *upd, *abort, *nkept;
};

__device__ __forceinline__ bool chain_wait(const volatile int *flag, int need, volatile int *abort_flag) {
    unsigned it = 0;
    while (*flag < need) {
        if (*abort_flag || ++it > (1u << 23)) { *abort_flag = 1; return false; }      // ~2 s: a bug, not a wait
        __nanosleep(it < 32 ? 20 : 200);
    }
    __threadfence();
    return true;
}
__device__ __forceinline__ int super_col_tiles(int s, int col_tiles) {      // mask tiles of super-column s (what row_done[s] counts up to)
    int need = 0;
    for (int c = s * SUPER; c < min((s + 1) * SUPER, col_tiles); ++c) need += c + 1;
    return need;
}

__global__ void __launch_bounds__(SCAN_THREADS)
nms_chain_kernel(const u64 *__restrict__ mask, const u64 *__restrict__ sup_t, const int *__restrict__ order, int n, int col_tiles, int n_super,
                 u64 *__restrict__ removed, u64 *__restrict__ kept_bits, int64_t *__restrict__ keep, int32_t *__restrict__ keep_count,
                 const int *__restrict__ row_done, ChainCtl ctl, int n_res) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_ok, s_ok2;                                  // two flags: thread 0 may reach the second wait before a slow warp has read the first
    if ((int)blockIdx.x >= n_res) {
        // ---------------- bulk updaters: target super-column t gets the kept rows of super-tiles 0 .. t - 2
        const int n_upd = gridDim.x - n_res;
        for (int t = 2; t < n_super; ++t) {
            const int c0 = t * SUPER, ncols = min(SUPER, col_tiles - c0);
            if (tid == 0) s_ok = chain_wait(ctl.done + (t - 2), 2, ctl.abort) && chain_wait(row_done + t, super_col_tiles(t, col_tiles), ctl.abort);
            __syncthreads();
            if (!s_ok) return;
            const long nblk = (long)(t - 1) * SUPER * ncols;
            const long wid = (long)(blockIdx.x - n_res) * (SCAN_THREADS / 32) + warp, nw = (long)n_upd * (SCAN_THREADS / 32);
            for (long k0 = wid; k0 < nblk; k0 += 4 * nw) {
                ulonglong2 w[4];
                u64 kb[4];
                int col[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {                   // four independent 512-byte blocks in flight per warp
                    const long k = k0 + u * nw;
                    col[u] = -1;
                    kb[u] = 0ull;
                    w[u] = make_ulonglong2(0ull, 0ull);
                    if (k < nblk) {
                        const int tp = (int)(k / ncols), c = c0 + (int)(k - (long)tp * ncols);
                        kb[u] = __ldcg(kept_bits + tp);
                        col[u] = c;
                        if (kb[u]) w[u] = __ldcg(reinterpret_cast<const ulonglong2 *>(mask + ((size_t)tp * col_tiles + c) * 64) + lane);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (col[u] < 0 || kb[u] == 0ull) continue;  // warp-uniform
                    const u64 v = warp_or(select2(w[u], kb[u], lane));
                    if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(removed + col[u]), (unsigned long long)v);
                }
            }
            __syncthreads();                                     // every warp's atomics are issued ...
            if (tid == 0) {
                __threadfence();                                 // ... and ordered before the count
                atomicAdd(ctl.upd + t, 1);
            }
        }
        return;
    }
    // ---------------- CTAs 0 and 1: the super-tiles in order, alternating -- while one resolves super-tile s the other has
    // already fetched what super-tile s + 1 needs from the mask (256 KB per super-tile: on one SM that is ~2.5 us, which a
    // single resolver would pay between every two super-tiles)
    extern __shared__ u64 s_nxt[];                               // [row tile][next column][64 rows]
    __shared__ u64 s_K[SUPER], s_U[SUPER], s_push[SUPER];
    __shared__ int s_pre[SUPER + 1];
    const int n_upd = gridDim.x - n_res;
    const int c = tid >> 6, jj = tid & 63;
    for (int s = blockIdx.x; s < n_super; s += n_res) {
        const int T0 = s * SUPER;
        const int nt = min(SUPER, col_tiles - T0);
        const int c0n = T0 + SUPER, ncols_next = max(0, min(SUPER, col_tiles - c0n));
        if (tid == 0)
            s_ok = chain_wait(row_done + s, super_col_tiles(s, col_tiles), ctl.abort) &&
                   (ncols_next == 0 || chain_wait(row_done + s + 1, super_col_tiles(s + 1, col_tiles), ctl.abort));
        __syncthreads();
        if (!s_ok) break;
        const int row = T0 * 64 + tid;
        const bool live = c < nt && row < n;
        u64 col[SUPER];
#pragma unroll
        for (int r = 0; r < SUPER; ++r) col[r] = (live && r <= c) ? __ldcg(sup_t + ((size_t)(T0 + c) * SUPER + r) * 64 + jj) : 0ull;
        for (int ch = tid; ch < nt * ncols_next * 32; ch += SCAN_THREADS) {
            const int blk = ch >> 5, tp = blk / ncols_next, k = blk - tp * ncols_next;
            const u64 *src = mask + ((size_t)(T0 + tp) * col_tiles + c0n + k) * 64 + (ch & 31) * 2;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s_nxt + ((size_t)tp * SUPER + k) * 64 + (ch & 31) * 2)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int my_order = live ? order[row] : 0;
        // the previous super-tile (the other resolver: its kept rows are in removed[] of our columns, nkept is current) and
        // the updaters' share
        if (tid == 0)
            s_ok2 = (s == 0 || chain_wait(ctl.done + (s - 1), 2, ctl.abort)) && (s < 2 || n_upd <= 0 || chain_wait(ctl.upd + s, n_upd, ctl.abort));
        __syncthreads();
        if (!s_ok2) break;
        if (tid < SUPER) {
            const int rows = min(64, n - (T0 + tid) * 64);
            const u64 valid = tid < nt ? (rows >= 64 ? ~0ull : ((1ull << max(rows, 0)) - 1ull)) : 0ull;
            s_U[tid] = tid < nt ? (~__ldcg(removed + T0 + tid) & valid) : 0ull;
            s_K[tid] = 0ull;
            s_push[tid] = 0ull;
        }
        const int nkept0 = tid == 0 ? __ldcg(ctl.nkept) : 0;
        __syncthreads();
        bool und = live && ((s_U[c] >> jj) & 1ull);
        bool kept = false;
        for (;;) {
            bool rem = false, ok = false;
            if (und) {
                u64 hit = 0ull, pend = 0ull;
#pragma unroll
                for (int r = 0; r < SUPER; ++r) { hit |= col[r] & s_K[r]; pend |= col[r] & s_U[r]; }
                rem = hit != 0ull;
                ok = !rem && pend == 0ull;
            }
            const unsigned bk = __ballot_sync(0xffffffffu, ok), br = __ballot_sync(0xffffffffu, rem);
            __syncthreads();                                     // every thread has read K / U of this round
            if (lane == 0 && (bk | br)) {                        // a warp owns one 32-bit half of its tile's words
                reinterpret_cast<unsigned *>(s_K)[warp] |= bk;
                reinterpret_cast<unsigned *>(s_U)[warp] &= ~(bk | br);
            }
            kept |= ok;
            und = und && !rem && !ok;
            if (!__syncthreads_or(und ? 1 : 0)) break;
        }
        if (tid == 0) {
            int run = nkept0;
            for (int t = 0; t < SUPER; ++t) { s_pre[t] = run; run += __popcll(s_K[t]); }
            s_pre[SUPER] = run;
            *ctl.nkept = run;
        }
        if (tid < nt) kept_bits[T0 + tid] = s_K[tid];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                         // s_nxt and s_pre are complete
        for (int k = 0; k < ncols_next; ++k) {
            const u64 v = warp_or(kept ? s_nxt[((size_t)c * SUPER + k) * 64 + jj] : 0ull);
            if (lane == 0 && v) atomicOr(reinterpret_cast<unsigned long long *>(&s_push[k]), (unsigned long long)v);
        }
        __syncthreads();
        if (tid < ncols_next && s_push[tid]) atomicOr(reinterpret_cast<unsigned long long *>(removed + c0n + tid), (unsigned long long)s_push[tid]);
        if (kept) keep[s_pre[c] + __popcll(s_K[c] & ((1ull << jj) - 1ull))] = my_order;
        __syncthreads();                                         // the pushes, kept_bits and nkept are issued ...
        if (tid == 0) {
            __threadfence();                                     // ... and ordered before the flag; done[s] = 2: resolved and pushed
            *reinterpret_cast<volatile int *>(ctl.done + s) = 2;
            if (s == n_super - 1) *keep_count = s_pre[SUPER];
        }
    }
    if (tid == 0 && *reinterpret_cast<volatile int *>(ctl.abort)) *keep_count = -1;
}

// One warp per segment.  Shared memory per warp: 6 * max_seg floats/ints.
constexpr int BATCH_WARPS = 4;

__global__ void __launch_bounds__(BATCH_WARPS * 32)
nms_batched_kernel(const float *__restrict__ dets, const int32_t *__restrict__ seg_off, const int32_t *__restrict__ seg_len,
                   int n_seg, double thresh, int max_seg, int64_t *__restrict__ keep, int32_t *__restrict__ keep_count) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float thresh_f = thresh_as_float(thresh);
    float *sx1 = smem + (size_t)warp * 6 * max_seg;
    float *sy1 = sx1 + max_seg, *sx2 = sy1 + max_seg, *sy2 = sx2 + max_seg, *sar = sy2 + max_seg;
    int *sid = (int *)(sar + max_seg);
    for (int seg = blockIdx.x * BATCH_WARPS + warp; seg < n_seg; seg += gridDim.x * BATCH_WARPS) {
        const int o = seg_off[seg], n = seg_len ? seg_len[seg] : seg_off[seg + 1] - o;
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
                if (suppresses(bi, ai, bj, sar[j], thresh_f)) dead |= 1u << (j >> 5);
            }
        }
        if (lane == 0) keep_count[seg] = nkept;
        __syncwarp();
    }
}

struct NmsWorkspace {
    float4 *boxes;
    float *areas;
    int *order, *rank, *nkept, *row_done, *ctl;
    u64 *removed, *kept_bits, *diag_t, *sup_t, *mask;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline NmsWorkspace carve(void *ws, int64_t n, int col_tiles) {
    char *p = (char *)ws;
    NmsWorkspace w;
    w.boxes = (float4 *)p;  p += align_up(sizeof(float4) * n, 256);
    w.areas = (float *)p;   p += align_up(sizeof(float) * n, 256);
    w.order = (int *)p;     p += align_up(sizeof(int) * n, 256);
    w.rank = (int *)p;      p += align_up(sizeof(int) * n, 256);
    // zeroed per call in one memset: rank | removed | kept_bits | nkept | row_done | ctl
    w.removed = (u64 *)p;   p += align_up(sizeof(u64) * col_tiles, 256);
    w.kept_bits = (u64 *)p; p += align_up(sizeof(u64) * col_tiles, 256);
    w.nkept = (int *)p;     p += 256;
    w.row_done = (int *)p;  p += align_up(sizeof(int) * ((col_tiles + SUPER - 1) / SUPER), 256);
    w.ctl = (int *)p;       p += align_up(sizeof(int) * (2 * ((col_tiles + SUPER - 1) / SUPER) + 4), 256);     // done[n_super] | upd[n_super] | abort
    w.diag_t = (u64 *)p;    p += align_up(sizeof(u64) * col_tiles * 64, 256);
    w.sup_t = (u64 *)p;     p += align_up(sizeof(u64) * col_tiles * SUPER * 64, 256);     // [col tile][row tile in its super-tile][64]
    w.mask = (u64 *)p;
    return w;
}

// Internal streams + events of azn_nms, one set per device, created on first use (never destroyed: process lifetime).
// The greedy chain is a latency-bound single CTA; sharing its SM with the mask kernel's CTAs stretches it from ~17 to
// ~27 us per super-tile (issue-slot contention), and that serial term bounds the whole call.  So the two run on
// DISJOINT SM partitions when the driver offers green contexts (CUDA 12.4+): 16 SMs for the chain (CTA 0 + its bulk
// updaters), the rest for the mask.  Without them: one extra high-priority stream on the whole device.
struct NmsStreams {
    cudaStream_t chain = nullptr;       // greedy pass
    cudaStream_t mask = nullptr;        // mask launches (green-context stream), or null: the caller's stream
    cudaEvent_t done = nullptr, fork = nullptr, mask_done = nullptr;
    std::vector<cudaEvent_t> ev;
    bool tried = false;
    int chain_sms = 8;                  // SMs of the chain's partition
};

template <typename Fn> Fn driver_fn(const char *name) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (Fn)p;
}

// 16 SMs for the chain, the remaining SMs for the mask.  Returns false (and leaves ns untouched) when anything is missing.
bool make_partitions(int dev, NmsStreams &ns) {
    typedef CUresult (*GetRes)(CUdevice, CUdevResource *, CUdevResourceType);
    typedef CUresult (*Split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
    typedef CUresult (*GenDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
    typedef CUresult (*CtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*StreamCreate)(CUstream *, CUgreenCtx, unsigned int, int);
    typedef CUresult (*DevGet)(CUdevice *, int);
    const GetRes get_res = driver_fn<GetRes>("cuDeviceGetDevResource");
    const Split split = driver_fn<Split>("cuDevSmResourceSplitByCount");
    const GenDesc gen = driver_fn<GenDesc>("cuDevResourceGenerateDesc");
    const CtxCreate ctx_create = driver_fn<CtxCreate>("cuGreenCtxCreate");
    const StreamCreate stream_create = driver_fn<StreamCreate>("cuGreenCtxStreamCreate");
    const DevGet dev_get = driver_fn<DevGet>("cuDeviceGet");
    if (!get_res || !split || !gen || !ctx_create || !stream_create || !dev_get) return false;
    if (getenv("AZN_NMS_NO_PARTITION")) return false;
    CUdevice cu_dev;
    if (dev_get(&cu_dev, dev) != CUDA_SUCCESS) return false;
    CUdevResource all, part, rest;
    if (get_res(cu_dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned int groups = 1;
    // 16 SMs: CTA 0 + 15 bulk updaters.  The updaters read the whole upper triangle of the mask once (25 MB at N = 20 000) and
    // 7 of them could not keep up with the block-wise resolver: measured 0.205-0.215 ms with 8 SMs, 0.170-0.179 with 12-24
    // (the mask kernel loses the SMs the chain gets: 0.167-0.189 with 32)
    int want = 16;
    if (const char *e = getenv("AZN_NMS_CHAIN_SMS")) want = std::max(8, std::min(64, atoi(e)));
    if (split(&part, &groups, &all, &rest, 0, (unsigned)want) != CUDA_SUCCESS || groups != 1) return false;
    if (part.sm.smCount < 8 || rest.sm.smCount < 32) return false;
    ns.chain_sms = (int)part.sm.smCount;
    CUdevResourceDesc d_chain, d_mask;
    if (gen(&d_chain, &part, 1) != CUDA_SUCCESS || gen(&d_mask, &rest, 1) != CUDA_SUCCESS) return false;
    CUgreenCtx g_chain, g_mask;
    if (ctx_create(&g_chain, d_chain, cu_dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (ctx_create(&g_mask, d_mask, cu_dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream s_chain, s_mask;
    if (stream_create(&s_chain, g_chain, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    if (stream_create(&s_mask, g_mask, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    ns.chain = (cudaStream_t)s_chain;
    ns.mask = (cudaStream_t)s_mask;
    return true;
}

NmsStreams *nms_streams(int n_events) {
    static std::mutex mu;
    static NmsStreams per_dev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    NmsStreams &ns = per_dev[dev];
    if (!ns.tried) {
        ns.tried = true;
        if (!make_partitions(dev, ns)) {
            ns.mask = nullptr;
            int least = 0, greatest = 0;
            cudaDeviceGetStreamPriorityRange(&least, &greatest);
            if (cudaStreamCreateWithPriority(&ns.chain, cudaStreamNonBlocking, greatest) != cudaSuccess) ns.chain = nullptr;
        }
        if (cudaEventCreateWithFlags(&ns.done, cudaEventDisableTiming) != cudaSuccess) ns.chain = nullptr;
        if (cudaEventCreateWithFlags(&ns.fork, cudaEventDisableTiming) != cudaSuccess) ns.chain = nullptr;
        if (cudaEventCreateWithFlags(&ns.mask_done, cudaEventDisableTiming) != cudaSuccess) ns.chain = nullptr;
    }
    if (!ns.chain) return nullptr;
    while ((int)ns.ev.size() < n_events) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ns.ev.push_back(e);
    }
    return &ns;
}

}  // namespace

// Diagnostic hook (like azn_fc_tune): 0 = default schedule; 1 = everything on the caller's stream, mask then chain;
// 2 = stop after the mask; 3 = stop after the sort (rank + scatter).  Modes 2 / 3 leave keep_count untouched.
// + 8: the all-pairs rank sort for every n (A/B of the bucket sort that large n take by default); + 16: the tile-by-tile
// greedy pass (nms_super_kernel) instead of the block-wise one; + 32: the float32 mask kernel (version 1) for every threshold;
// + 64: the persistent chain (nms_chain_kernel, one launch for all super-tiles) instead of one launch per super-tile.
static int g_nms_mode = 0;
extern "C" void azn_nms_tune(int mode) { g_nms_mode = mode; }

extern "C" size_t azn_nms_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    const size_t ct = (size_t)((n + 63) / 64);
    return align_up(sizeof(float4) * n, 256) + align_up(sizeof(float) * n, 256) + 2 * align_up(sizeof(int) * n, 256) +
           2 * align_up(sizeof(u64) * ct, 256) + 256 + align_up(sizeof(int) * ((ct + 15) / 16), 256) + align_up(sizeof(int) * (2 * ((ct + 15) / 16) + 4), 256) +
           align_up(sizeof(u64) * ct * 64, 256) +
           align_up(sizeof(u64) * ct * SUPER * 64, 256) + align_up(sizeof(u64) * ct * ct * 64, 256);
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
    // zeroed per call: removed | kept_bits | nkept | row_done | ctl -- by the first sort kernel; plus `rank` (a memset) when the
    // all-pairs rank sort runs over several score tiles and accumulates its counts there
    uint4 *zero = (uint4 *)w.removed;
    const int n_zero = (int)(((char *)w.diag_t - (char *)w.removed) / 16);
    if (n >= BKT_MIN_N && !(g_nms_mode & 8)) {
        // (key, index) pairs in bucket order and the bucket offsets: at the head of the mask
        // area (64 x 64 tiles x 512 B >= 2 MB at this n; the mask kernel, which runs after the sort, overwrites them)
        uint2 *bpair = (uint2 *)w.mask;
        int *boff = (int *)(bpair + n);
        const size_t stage = (size_t)n * sizeof(uint2);
        const int staged = stage <= 200 * 1024 ? 1 : 0;
#define AZN_BKT_LAUNCH(IT)                                                                                                          \
    do {                                                                                                                            \
        static bool attr_set = false;                                                                                               \
        if (!attr_set) {                                                                                                            \
            AZN_CUDA(cudaFuncSetAttribute(nms_bucket_kernel<IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));         \
            attr_set = true;                                                                                                        \
        }                                                                                                                           \
        nms_bucket_kernel<IT><<<1, BKT_THREADS, staged ? stage : 0, s>>>(dets, (int)n, bpair, boff, staged, zero, n_zero);                        \
    } while (0)
        if (n <= 8 * BKT_THREADS) AZN_BKT_LAUNCH(8);
        else if (n <= 20 * BKT_THREADS) AZN_BKT_LAUNCH(20);
        else if (n <= 32 * BKT_THREADS) AZN_BKT_LAUNCH(32);
        else AZN_BKT_LAUNCH(0);
#undef AZN_BKT_LAUNCH
        AZN_LAUNCH_CHECK();
        AZN_CUDA(azn_launch_pdl(nms_bucket_rank_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, dets, (int)n, (const uint2 *)bpair,
                                (const int *)boff, w.boxes, w.areas, w.order));
    } else if (n <= RANK_TILE && !(g_nms_mode & 8)) {
        nms_rank1_kernel<<<(unsigned)((n + RANK1_DETS - 1) / RANK1_DETS), RANK_THREADS, 0, s>>>(dets, (int)n, w.boxes, w.areas, w.order, zero, n_zero);
        AZN_LAUNCH_CHECK();
    } else {
        if (n > RANK_TILE) AZN_CUDA(cudaMemsetAsync(w.rank, 0, sizeof(int) * (size_t)n, s));
        nms_rank_kernel<<<dim3((unsigned)((n + RANK_THREADS - 1) / RANK_THREADS), (unsigned)((n + RANK_TILE - 1) / RANK_TILE)),
                          RANK_THREADS, 0, s>>>(dets, (int)n, w.rank, w.boxes, w.areas, w.order, zero, n_zero);
        AZN_LAUNCH_CHECK();
        if (n > RANK_TILE) {
            nms_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dets, (int)n, w.rank, w.boxes, w.areas, w.order);
            AZN_LAUNCH_CHECK();
        }
    }
    {
        const size_t smem = (size_t)SUPER * SUPER * 64 * sizeof(u64);      // 128 KB diagonal super-block
        static bool attr_set = false;
        if (!attr_set) {
            AZN_CUDA(cudaFuncSetAttribute(nms_super_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            AZN_CUDA(cudaFuncSetAttribute(nms_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int n_super = (col_tiles + SUPER - 1) / SUPER;
        // Two streams: the mask is ONE launch (column-major over the triangle, cheapest columns first, so super-columns
        // complete roughly in order) that publishes per-super-column tile counters; the greedy chain runs concurrently
        // -- on its own SM partition when the driver offers green contexts, else on a high-priority stream -- and launch
        // si polls the counters of super-columns si (CTA 0) and si + 1 (bulk updaters).  The serial chain thus runs
        // alongside the mask instead of after it; the caller's stream joins at the end.
        // Inside a stream capture (a caller building a CUDA graph) everything stays on the caller's stream: the graph
        // would serialise the fork anyway and the lazily created stream / events must not be born inside a capture.
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        AZN_CUDA(cudaStreamIsCapturing(s, &cap));
        if ((g_nms_mode & 7) == 3) return AZN_OK;
        const bool two_streams = cap == cudaStreamCaptureStatusNone && n_super > 1 && (g_nms_mode & 7) == 0;
        NmsStreams *ns = two_streams ? nms_streams(0) : nullptr;
        AZN_REQUIRE(!two_streams || ns != nullptr, "azn_nms: could not create the internal stream / events");
        cudaStream_t chain = two_streams ? ns->chain : s;
        cudaStream_t ms = (two_streams && ns->mask) ? ns->mask : s;      // partitioned: the mask has its own SMs and stream
        // The internal streams and their events are shared by every azn_nms call on this device.  The enqueue of one
        // call -- fork, launches, join -- is therefore one critical section: another host thread must not re-record
        // `fork` / `done` between this call's record and its wait (an already enqueued wait keeps the record it saw).
        // And whatever happens after the fork, the caller's stream is joined with the internal ones before returning,
        // error or not: the caller may free the workspace as soon as its own stream is idle.
        static std::mutex enqueue_mu;
        std::unique_lock<std::mutex> enqueue_lock(enqueue_mu, std::defer_lock);
        if (two_streams) enqueue_lock.lock();
        struct Join {
            NmsStreams *ns; cudaStream_t s, chain, ms; bool armed;
            ~Join() {
                if (!armed) return;
                if (cudaEventRecord(ns->done, chain) == cudaSuccess) cudaStreamWaitEvent(s, ns->done, 0);
                if (ms != s && cudaEventRecord(ns->mask_done, ms) == cudaSuccess) cudaStreamWaitEvent(s, ns->mask_done, 0);
            }
        } join{ns, s, chain, ms, false};
        if (two_streams) {
            AZN_CUDA(cudaEventRecord(ns->fork, s));                       // sorted boxes / zeroed state are ready
            join.armed = true;
            if (ms != s) AZN_CUDA(cudaStreamWaitEvent(ms, ns->fork, 0));
            AZN_CUDA(cudaStreamWaitEvent(chain, ns->fork, 0));
        }
        if (thresh > 0.0 && !(g_nms_mode & 32))
            nms_mask2_kernel<<<(unsigned)((long)col_tiles * (col_tiles + 1) / 2), MASK_THREADS, 0, ms>>>(w.boxes, w.areas, (int)n, thresh, w.mask, col_tiles,
                                                                                                      w.diag_t, w.row_done, w.sup_t);
        else
            nms_mask_kernel<<<(unsigned)((long)col_tiles * (col_tiles + 1) / 2), MASK_THREADS, 0, ms>>>(w.boxes, w.areas, (int)n, thresh, w.mask, col_tiles, 0,
                                                                                                     w.diag_t, w.row_done, w.sup_t);
        AZN_LAUNCH_CHECK();
        // The persistent chain, A/B only: measured slower than one launch per super-tile for long chains, and no faster for
        // the one or two super-tiles of n <= 2048 (one resolver CTA, no hand-over at all: 0.045-0.051 vs 0.044-0.048 ms)
        if ((g_nms_mode & 64) && !(g_nms_mode & 16) && (g_nms_mode & 7) != 2) {
            // the persistent chain: CTA 0 + up to 7 updaters (the chain's SM partition has 8 SMs; fewer when the columns are few)
            const long blocks = n_super > 2 ? ((long)(n_super - 2) * SUPER * SUPER + 127) / 128 : 0;
            const int upd = (int)std::min<long>(blocks, ms != s ? ns->chain_sms - 2 : SUPER_UPDATERS - 2);      // partitioned: what the chain's SMs hold
            ChainCtl ctl;
            ctl.done = w.ctl; ctl.upd = w.ctl + n_super; ctl.abort = w.ctl + 2 * n_super; ctl.nkept = w.nkept;
            static bool chain_attr = false;
            if (!chain_attr) {
                AZN_CUDA(cudaFuncSetAttribute(nms_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                chain_attr = true;
            }
            const int n_res = n_super > 2 ? 2 : 1;
            nms_chain_kernel<<<n_res + upd, SCAN_THREADS, smem, chain>>>((const u64 *)w.mask, (const u64 *)w.sup_t, (const int *)w.order, (int)n, col_tiles, n_super,
                                                                    w.removed, w.kept_bits, keep, keep_count, (const int *)w.row_done, ctl, n_res);
            AZN_LAUNCH_CHECK();
        } else
        for (int si = 0; si < n_super && (g_nms_mode & 7) != 2; ++si) {
            // the updaters of launch si push the kept rows of super-tiles < si into the columns of super-tile si + 1
            const int upd_cols = min(SUPER, col_tiles - (si + 1) * SUPER);
            long updaters = (si == 0 || upd_cols <= 0) ? 0 : ((long)si * SUPER * upd_cols + 127) / 128;      // ~4 blocks per warp
            const int upd_cap = ms != s ? ns->chain_sms - 1 : SUPER_UPDATERS;      // partitioned: CTA 0 keeps one SM of the chain's partition to itself
            if (updaters > upd_cap) updaters = upd_cap;
            if (g_nms_mode & 16)                               // A/B: the tile-by-tile pass
                AZN_CUDA(azn_launch_pdl(nms_super_kernel, dim3(1 + (unsigned)updaters), dim3(SCAN_THREADS), smem, chain, (const u64 *)w.mask, (const u64 *)w.diag_t,
                                        (const int *)w.order, (int)n, col_tiles, si, w.removed, w.kept_bits, w.nkept, keep, keep_count,
                                        si == n_super - 1 ? 1 : 0, (const int *)w.row_done));
            else
                AZN_CUDA(azn_launch_pdl(nms_block_kernel, dim3(1 + (unsigned)updaters), dim3(SCAN_THREADS), smem, chain, (const u64 *)w.mask, (const u64 *)w.sup_t,
                                        (const int *)w.order, (int)n, col_tiles, si, w.removed, w.kept_bits, w.nkept, keep, keep_count,
                                        si == n_super - 1 ? 1 : 0, (const int *)w.row_done));
        }
        // `join` (above) makes the caller's stream wait for the chain and the mask on every path out of this scope
    }
    return AZN_OK;
}

static int launch_batched(const float *dets, const int32_t *seg_off, const int32_t *seg_len, int n_seg, int max_seg,
                          double thresh, int64_t *keep, int32_t *keep_count, cudaStream_t stream) {
    const size_t smem = (size_t)BATCH_WARPS * 6 * max_seg * sizeof(float);   // 96 KB at AZN_NMS_SEG_MAX
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        AZN_CUDA(cudaFuncSetAttribute(nms_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    int blocks = (n_seg + BATCH_WARPS - 1) / BATCH_WARPS;
    size_t per_sm = (200 * 1024) / (smem > 0 ? smem : 1);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int cap = azn_num_sms() * (int)per_sm;
    if (blocks > cap) blocks = cap;
    nms_batched_kernel<<<blocks, BATCH_WARPS * 32, smem, stream>>>(dets, seg_off, seg_len, n_seg, thresh, max_seg, keep,
                                                                keep_count);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_nms_batched(const float *dets, const int32_t *seg_off, int n_seg, double thresh,
                               int64_t *keep, int32_t *keep_count, azn_stream_t stream) {
    AZN_REQUIRE(n_seg >= 0, "azn_nms_batched: n_seg < 0");
    if (n_seg == 0) return AZN_OK;
    AZN_REQUIRE(dets && seg_off && keep && keep_count, "azn_nms_batched: null pointer");
    return launch_batched(dets, seg_off, nullptr, n_seg, AZN_NMS_SEG_MAX, thresh, keep, keep_count, (cudaStream_t)stream);
}

extern "C" int azn_nms_segments(const float *dets, const int32_t *seg_off, const int32_t *seg_len, int n_seg, int max_len,
                                double thresh, int64_t *keep, int32_t *keep_count, azn_stream_t stream) {
    AZN_REQUIRE(n_seg >= 0, "azn_nms_segments: n_seg < 0");
    if (n_seg == 0) return AZN_OK;
    AZN_REQUIRE(dets && seg_off && seg_len && keep && keep_count, "azn_nms_segments: null pointer");
    AZN_REQUIRE(max_len > 0 && max_len <= AZN_NMS_SEG_MAX, "azn_nms_segments: max_len must be in [1, %d]", AZN_NMS_SEG_MAX);
    return launch_batched(dets, seg_off, seg_len, n_seg, (max_len + 31) / 32 * 32, thresh, keep, keep_count,
                          (cudaStream_t)stream);
}
