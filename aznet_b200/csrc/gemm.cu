// Fully connected layers of the AZ-Net / Fast R-CNN heads on the 5th-generation tensor cores.
//
//   out[M, N] = act(A[M, K] . W[N, K]^T + bias[N])        A, W bf16 (K-major), fp32 accumulate
//
// replaces InnerProductLayer::Forward (caffe-fast-rcnn/src/caffe/layers/inner_product_layer.cpp:80-93:
// one cblas/cublas sgemm for the product, a second rank-1 sgemm for the bias) and the ReLU / Sigmoid
// layers that follow it in models/Pascal/VGG16/az-net/test_fc.prototxt:26-232.
//
// Kernel anatomy (one persistent CTA per SM, 64 + 128*MH threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of MH 128x64 A boxes and a BLOCK_Nx64 W box
//               (128-byte swizzle) into a STAGES-deep shared-memory ring, completion on mbarriers.
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma.cta_group::1
//               .kind::f16 (UMMA 128 x BLOCK_N x 16, bf16 in, fp32 accumulate) straight from the
//               swizzled shared tiles through UMMA smem descriptors; tcgen05.commit releases the
//               ring slot / publishes the accumulator.
//   warps 2..   epilogue, 4 warps per 128-row half: tcgen05.ld the accumulator (one TMEM lane = one
//               output row per thread), bias + ReLU / sigmoid, 16-byte stores.
// Tile = (128*MH) x BLOCK_N.  The wide layers (int6 / fc6 / fc7) run MH = 2, BLOCK_N = 256: two
// 128 x 256 accumulators fill the 512 TMEM columns and every W box is used for two UMMAs, which
// halves the L2 -> SM bytes per FLOP of a 128 x 256 tile (measured: the 128-row kernel was bound by
// L2 bandwidth at 72 % tensor-pipe activity).  A half whose rows are all past the live count is
// neither loaded nor multiplied.  Narrow tiles (MH = 1) keep two accumulator buffers in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1.
// Scheduling is decided ON THE DEVICE from the live row count (the search keeps its region counts
// in HBM): tiles = ceil(m_live/TILE_M) x ceil(N/BLOCK_N).  Whole waves of tiles go one per CTA
// (data-parallel); the K loop of the remaining tiles is cut into P equal parts, P chosen on the
// device to minimise waves(P)/P plus a fix-up charge, and the (tile, part) units are dealt out
// part-major so that CTAs running side by side sweep the SAME k-range and share A / W panels
// through L2 (a k-staggered stream-K schedule was measured 1.3x slower: every CTA streamed its own
// panels from HBM).  A CTA that computes a non-final part dumps its raw fp32 accumulator into a
// workspace slot (coalesced layout) and raises the slot's flag (release); the CTA that computes the
// final part waits for the flags (acquire), adds the partials in part order -- deterministic, no
// atomics on data -- and runs the epilogue.  With few tiles (the shallow search levels:
// M = 64 .. 2048) this is an even split of the K loop across the whole chip: every SM streams its
// share of the 205 MB int6 weight matrix.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

namespace {

constexpr int HALF_M = 128;            // rows of one UMMA / one accumulator
constexpr int BLOCK_K = 64;            // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int SMEM_RING_BYTES = 192 * 1024;
constexpr size_t WS_DATA_BYTES = (size_t)128 << 20;    // partial-accumulator area of the workspace
constexpr int WS_MAX_SLOTS = 1024;                     // flag words (one per slot)


// RU = 1 (3x3 convolution, "row re-use"): a stage holds ONE A box of TILE_M + 8 rows -- the tile's pixels of filter row dy
// from one pixel to the left to one to the right (+ 5 rows of slack for the box granularity) -- and the W boxes of the
// three taps (dy, -1), (dy, 0), (dy, +1): the three taps multiply the same shared-memory rows, shifted by one row (128 B)
// each, instead of three separately loaded boxes.
template <int BLOCK_N, int MH, int RU = 0> struct Cfg {
    static constexpr int TILE_M = HALF_M * MH;
    static constexpr int A_HALF_BYTES = HALF_M * BLOCK_K * 2;
    static constexpr int A_TAIL_BYTES = RU ? 8 * BLOCK_K * 2 : 0;
    static constexpr int A_BYTES = MH * A_HALF_BYTES + A_TAIL_BYTES;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int B_STAGE_BYTES = (RU ? 3 : 1) * B_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
    static constexpr int RING_BYTES = RU ? 216 * 1024 : SMEM_RING_BYTES;
    static constexpr int STAGES = RING_BYTES / STAGE_BYTES < 8 ? RING_BYTES / STAGE_BYTES : 8;
    static constexpr int ACC_COLS = MH * BLOCK_N;                       // TMEM columns of one accumulator set
    static constexpr int NUM_ACC = 2 * ACC_COLS <= 512 ? 2 : 1;        // double-buffered when it fits
    static constexpr int TMEM_COLS = NUM_ACC * ACC_COLS < 32 ? 32 : NUM_ACC * ACC_COLS;
    static constexpr int THREADS = 64 + 128 * MH;
    static constexpr int EPI_THREADS = 128 * MH;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr size_t SLOT_BYTES = (size_t)TILE_M * BLOCK_N * sizeof(float);
    static constexpr int WS_SLOTS = (int)(WS_DATA_BYTES / SLOT_BYTES) < WS_MAX_SLOTS ? (int)(WS_DATA_BYTES / SLOT_BYTES) : WS_MAX_SLOTS;
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation must be a power of two <= 512");
    static_assert(STAGES >= (RU ? 2 : 3), "pipeline too shallow");
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major tile, 128-byte swizzle, rows of 128 bytes, 8-row groups
// 1024 bytes apart (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64)).
// The start address need not be 1024-byte aligned: the tensor core applies the 128-byte swizzle to the ABSOLUTE shared
// address (bits [4,7) ^= bits [7,10)), exactly like the TMA write that filled the box, so a descriptor that starts r rows
// (r * 128 B) into a box addresses the same data rows r .. r + 127 -- with the "matrix base offset" field [49,52) left 0.
// MEASURED (round 2, the RU convolution kernels): base offset 0 is bit-correct against the fp32 convolution on every test
// shape; setting it to (start >> 7) & 7 -- what the field's description suggests -- gives wrong sums.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// UMMA instruction descriptor, kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int MIN_KB_PER_PART = 4;      // do not cut the K loop finer than this many 64-wide k-blocks
constexpr int MAX_PARTS = 16;
// measured fix-up charges, in SM cycles: an owner CTA adds one peer's 256 x 256 partial in ~9 us (latency-bound
// L2 reads by its epilogue threads); the grid-wide finish phase costs a barrier plus one pass over all partials
constexpr int PEER_CYCLES = 18000;      // per 256 x 256 fp32 partial added by an owner CTA (measured ~9 us)
constexpr int FINISH_CYCLES = 4000;     // grid barrier + start of the finish phase

// Work decomposition, identical on every CTA and every warp role (pure function of m_live and the grid).
struct Plan {
    int m_live, m_tiles, n_tiles, tiles, kblocks;
    int dp_tiles, rem_tiles, parts;        // data-parallel tiles (first), split tiles (last), parts per split tile
    int finish;                            // 1: every part dumps a partial and finish_phase() reduces them after a grid barrier
    int n_major;                           // raster order of tile ids
};
struct Work {
    int tile, kb0, kb1, kind, part, rem;   // rem = index of the tile among the split tiles
};
enum { WORK_FULL = 0, WORK_PARTIAL = 1, WORK_OWNER = 2 };

__device__ __forceinline__ Plan make_plan(const int32_t *m_live_ptr, int M_cap, int N, int K, int block_n, int tile_m,
                                          int ws_slots, int grid, int force_parts = 0, int force_finish = -1) {
    Plan p;
    p.m_live = m_live_ptr ? min(max(*m_live_ptr, 0), M_cap) : M_cap;
    p.m_tiles = (p.m_live + tile_m - 1) / tile_m;
    p.n_tiles = (N + block_n - 1) / block_n;
    p.tiles = p.m_tiles * p.n_tiles;
    p.kblocks = K / BLOCK_K;
    p.rem_tiles = p.tiles % grid;
    p.dp_tiles = p.tiles - p.rem_tiles;
    p.parts = 1;
    p.finish = 0;
    if (p.rem_tiles > 0) {
        const int mh = tile_m / HALF_M;
        // per k-block: 4*mh UMMAs of 128 x block_n x 16 (2*block_n*mh cycles), but never faster than the SM can
        // take the operands in (measured ~50 B per cycle per SM through TMA)
        const int kb_bytes = (mh * HALF_M + block_n) * BLOCK_K * 2;
        const int kb_cycles = max(2 * block_n * mh, kb_bytes / 50);
        const long slot_bytes = (long)tile_m * block_n * 4;
        const int peer = (int)(PEER_CYCLES * slot_bytes / 262144);            // in-kernel: the owner adds one peer's partial
        int best = p.kblocks * kb_cycles;                  // P = 1: one wave of whole tiles
        for (int P = 2; P <= MAX_PARTS; ++P) {
            if (p.kblocks / P < MIN_KB_PER_PART || p.rem_tiles * P > ws_slots) break;
            const int waves = (p.rem_tiles * P + grid - 1) / grid;
            const int mma = waves * ((p.kblocks + P - 1) / P) * kb_cycles;
            // finish phase: grid barrier + all partials read once by the whole grid (measured ~0.8 KB per cycle)
            const int fin = FINISH_CYCLES + (int)((long)P * p.rem_tiles * slot_bytes / 800);
            const int c_in = mma + peer * (P - 1), c_fin = mma + fin;
            if (c_in < best) { best = c_in; p.parts = P; p.finish = 0; }
            if (c_fin < best) { best = c_fin; p.parts = P; p.finish = 1; }
        }
    }
    if (force_parts > 0 && p.rem_tiles > 0) {
        p.parts = force_parts;
        while (p.parts > 1 && (p.kblocks / p.parts < 1 || p.rem_tiles * p.parts > ws_slots)) --p.parts;
        p.finish = force_finish > 0 ? 1 : 0;
    } else if (force_finish >= 0 && p.parts > 1) {
        p.finish = force_finish;
    }
    // operands are re-read from HBM once per wave of tiles that does not share them: keep the bigger one
    // (W: N*K, A: m_live*K) shared inside a wave
    p.n_major = p.m_live < N ? 1 : 0;
    return p;
}

// idx-th piece of work of CTA `cta`: data-parallel tiles first, then (tile, part) units, part-major.
__device__ __forceinline__ bool get_work(const Plan &p, int cta, int grid, int idx, Work &w) {
    const int n_dp = p.dp_tiles / grid;
    if (idx < n_dp) {
        w.tile = cta + idx * grid; w.kb0 = 0; w.kb1 = p.kblocks; w.kind = WORK_FULL; w.part = 0; w.rem = 0;
        return true;
    }
    const int u = cta + (idx - n_dp) * grid;
    if (u >= p.rem_tiles * p.parts) return false;
    w.part = u / p.rem_tiles;
    w.rem = u - w.part * p.rem_tiles;
    w.tile = p.dp_tiles + w.rem;
    w.kb0 = (int)((long)p.kblocks * w.part / p.parts);
    w.kb1 = (int)((long)p.kblocks * (w.part + 1) / p.parts);
    w.kind = p.parts == 1 ? WORK_FULL : ((w.part == p.parts - 1 && !p.finish) ? WORK_OWNER : WORK_PARTIAL);
    return true;
}
__device__ __forceinline__ void tile_coords(const Plan &p, int tile, int &m_tile, int &n_tile) {
    if (p.n_major) { n_tile = tile / p.m_tiles; m_tile = tile - n_tile * p.m_tiles; }
    else           { m_tile = tile / p.n_tiles; n_tile = tile - m_tile * p.n_tiles; }
}
// Partial accumulators are stored so that the 32 lanes of a warp (32 consecutive tile rows) touch 32
// consecutive 16-byte words: float4 index ((chunk*8 + j4) * tile_m + row).
__device__ __forceinline__ size_t partial_f4(int chunk, int j4, int row, int tile_m) { return ((size_t)(chunk * 8 + j4) * tile_m + row); }

__device__ __forceinline__ void flag_release(int *flag) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
__device__ __forceinline__ int flag_acquire(const int *flag) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
template <int N_THREADS> __device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_THREADS) : "memory"); }
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// sigmoid_layer.cpp:11-13 evaluates `1. / (1. + exp(-x))` with a float exp and a double division; here the
// division is IEEE float (<= 1 ulp from the reference formula, which the reference's own test pins to 4 ulp,
// test_neuron_layer.cpp:202-217; the bf16 operands of the product move the argument by ~2^-9 relative anyway).
// A double division per score made the 128-thread epilogue of the heads layer latency-bound (measured 9 us).
__device__ __forceinline__ float sigmoid_caffe(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// Bias + activation of 32 consecutive columns held in registers.  The activation switch sits OUTSIDE the
// unrolled column loops: with the switch inside, every column carried all activation bodies and the epilogue
// thrashed the instruction cache (measured 12 us per 128 x 128 tile instead of ~1).
// AZN_ACT_AZ_HEAD: columns are [adj_score (nsub) | adj_bbox (4*nsub) | zoom_score (1)]; the two score
// groups go through the Sigmoid layers adj_prob / zoom_prob (test_fc.prototxt:221-232).
// Caffe's ReLU is std::max(x, 0) (relu_layer.cpp:16-19), which hands a NaN through (the comparison is false); one
// FMNMX.NAN instead of the NaN-dropping `v > 0 ? v : 0` -- the skip-layer head's GRN produces NaNs by design for
// all-zero pooled positions (grn_layer.cpp has no epsilon).
__device__ __forceinline__ float relu_caffe(float v) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void bias_act32(float (&v)[32], const float *__restrict__ bias, int col0, int N, int act, int nsub) {
    if (col0 + 32 <= N && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + col0 + j));     // col0 % 32 == 0
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += col0 + j < N ? __ldg(bias + col0 + j) : 0.f;
    }
    if (act == AZN_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = relu_caffe(v[j]);
    } else if (act == AZN_ACT_AZ_HEAD) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;                             // the same column in every lane: a uniform branch
            if (col < nsub || col == 5 * nsub) v[j] = sigmoid_caffe(v[j]);
        }
    }
}
__device__ __forceinline__ float apply_act(float v, int act, int col, int nsub) {
    if (act == AZN_ACT_RELU) return relu_caffe(v);
    if (act == AZN_ACT_AZ_HEAD) return (col < nsub || col == 5 * nsub) ? sigmoid_caffe(v) : v;
    return v;
}

struct EpiParams {
    const float *bias;
    void *out;
    int out_dtype, ldo, N, act, act_aux;
    int force_parts, force_finish;   // tuning hook (azn_fc_tune): 0 / -1 = automatic
    float *ws;           // split partials: [slot][BLOCK_N/4][TILE_M] float4, slot = part * rem_tiles + rem
    int *flags;          // [WS_MAX_SLOTS] "the partial of this slot is complete" (zero between launches)
    unsigned long long *sync;   // grid-barrier ticket counter of the finish phase (monotonic, never reset)
    long long *trace;    // azn_fc_trace: per CTA 8 globaltimer stamps, or NULL
    // CONV kernels only -- 3x3 / pad 1 / stride 1 convolution as an implicit GEMM over a zero-bordered NHWC map
    // (see azn_conv3x3_forward): k-blocks per filter tap (Cin / 64), padded width / height, rows per image,
    // and whether the output map is written without the border (the last layer of the backbone)
    int cv_kbt, cv_wp, cv_hp, cv_plane, cv_unpad, cv_taps;   // cv_taps: 9, or 1 = the taps are already gathered into K (patches)
    int cv_reuse;        // 1: launched as an RU kernel (one A box per filter row, three taps per stage)
};

// Row of the padded pixel grid -> is it a border pixel, and which output row does it go to.
__device__ __forceinline__ bool conv_row(const EpiParams &ep, int row, size_t &orow) {
    const int n = row / ep.cv_plane, pp = row - n * ep.cv_plane;
    const int y = pp / ep.cv_wp, x = pp - y * ep.cv_wp;
    const bool border = y == 0 || y == ep.cv_hp - 1 || x == 0 || x == ep.cv_wp - 1;
    orow = ep.cv_unpad ? ((size_t)n * (ep.cv_hp - 2) + (y - 1)) * (ep.cv_wp - 2) + (x - 1) : (size_t)row;
    return border;
}
enum { TR_START = 0, TR_SETUP = 1, TR_FIRST_FULL = 2, TR_MMA_DONE = 3, TR_ACC_READY = 4, TR_EPI_DONE = 5, TR_END = 6, TR_UNITS = 7 };

// Finish phase of the split schedule with many parts: every part has dumped a raw partial; after a grid-wide
// barrier (all CTAs are co-resident: one persistent CTA per SM) ALL threads of the grid sum the P partials of
// every split tile in part order and apply bias + activation.  One thread per (row, 4 columns).
template <bool CONV>
__device__ __forceinline__ void finish_phase(const Plan &pl, int N, int block_n, int tile_m, const EpiParams &ep) {
    const int q4 = block_n / 4;                              // float4 columns per tile row
    const size_t slot_f4 = (size_t)tile_m * q4;
    const long total = (long)pl.rem_tiles * tile_m * q4;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int trow = (int)(e % tile_m);                  // consecutive threads -> consecutive rows: coalesced partial reads
        const int c4 = (int)((e / tile_m) % q4);
        const int rem = (int)(e / ((long)tile_m * q4));
        int m_tile, n_tile;
        tile_coords(pl, pl.dp_tiles + rem, m_tile, n_tile);
        const int row = m_tile * tile_m + trow, col0 = n_tile * block_n + c4 * 4;
        if (row >= pl.m_live || col0 >= N) continue;
        size_t orow = (size_t)row;
        bool border = false;
        if (CONV) {
            border = conv_row(ep, row, orow);
            if (border && ep.cv_unpad) continue;
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *src = (const float4 *)ep.ws + (size_t)rem * slot_f4 + partial_f4(c4 >> 3, c4 & 7, trow, tile_m);
        const size_t part_stride = (size_t)pl.rem_tiles * slot_f4;
        int part = 0;
        for (; part + 4 <= pl.parts; part += 4) {            // 4 independent L2 reads in flight, summed in part order
            const float4 t0 = __ldcg(src + (size_t)part * part_stride), t1 = __ldcg(src + (size_t)(part + 1) * part_stride);
            const float4 t2 = __ldcg(src + (size_t)(part + 2) * part_stride), t3 = __ldcg(src + (size_t)(part + 3) * part_stride);
            acc.x += t0.x; acc.y += t0.y; acc.z += t0.z; acc.w += t0.w;
            acc.x += t1.x; acc.y += t1.y; acc.z += t1.z; acc.w += t1.w;
            acc.x += t2.x; acc.y += t2.y; acc.z += t2.z; acc.w += t2.w;
            acc.x += t3.x; acc.y += t3.y; acc.z += t3.z; acc.w += t3.w;
        }
        for (; part < pl.parts; ++part) {
            const float4 t = __ldcg(src + (size_t)part * part_stride);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const float v[4] = {acc.x, acc.y, acc.z, acc.w};
        if (ep.out_dtype == AZN_DTYPE_BF16 && col0 + 4 <= N && (ep.ldo & 3) == 0) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = border ? 0.f : apply_act(v[j] + ep.bias[col0 + j], ep.act, col0 + j, ep.act_aux);
            __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
            *reinterpret_cast<uint2 *>((__nv_bfloat16 *)ep.out + orow * ep.ldo + col0) = make_uint2(*(uint32_t *)&lo, *(uint32_t *)&hi);
            continue;
        }
        for (int j = 0; j < 4; ++j) {
            const int col = col0 + j;
            if (col >= N) break;
            const float o = border ? 0.f : apply_act(v[j] + ep.bias[col], ep.act, col, ep.act_aux);
            if (ep.out_dtype == AZN_DTYPE_BF16) ((__nv_bfloat16 *)ep.out)[orow * ep.ldo + col] = __float2bfloat16_rn(o);
            else ((float *)ep.out)[orow * ep.ldo + col] = o;
        }
    }
}

// Grid-wide barrier on a monotonic ticket counter: CTA k takes ticket t and waits until the counter reaches
// the end of t's generation (every launch that comes here adds exactly gridDim.x tickets).
__device__ __forceinline__ void grid_barrier(unsigned long long *counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd(counter, 1ull);
        const unsigned long long target = (t / gridDim.x + 1ull) * gridDim.x;
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

template <int BLOCK_N, int MH, bool CONV, int RU = 0>
__global__ void __launch_bounds__(Cfg<BLOCK_N, MH, RU>::THREADS, 1)
fc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const int32_t *__restrict__ m_live_ptr, int M_cap, int N, int K, EpiParams ep,
               const __grid_constant__ CUtensorMap tmap_a8) {
    // (tmap_a8: RU only -- the same tensor as tmap_a with boxes of 8 rows, for the tail of the A box)
    using C = Cfg<BLOCK_N, MH, RU>;
    constexpr int TILE_M = C::TILE_M;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + C::STAGES * C::A_BYTES;
    uint64_t *bars = (uint64_t *)(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + C::STAGES, *tmem_full = bars + 2 * C::STAGES, *tmem_empty = bars + 2 * C::STAGES + 2;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x, grid = gridDim.x;
    long long *tr = ep.trace ? ep.trace + (size_t)cta * 8 : nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4 * MH); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above is on-chip setup: it overlaps the tail of the preceding kernel (PDL); global memory from here on
    pdl_enter();
    if (tr && threadIdx.x == 0) tr[TR_START] = tr[TR_SETUP] = global_ns();
    const Plan pl = make_plan(m_live_ptr, M_cap, N, K, BLOCK_N, TILE_M, C::WS_SLOTS, gridDim.x, ep.force_parts, ep.force_finish);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            Work w;
            for (int idx = 0; get_work(pl, cta, grid, idx, w); ++idx) {
                int m_tile, n_tile;
                tile_coords(pl, w.tile, m_tile, n_tile);
                // 128-row halves of this tile that hold live rows: the others are neither loaded nor multiplied
                const int halves = min(MH, (pl.m_live - m_tile * TILE_M + HALF_M - 1) / HALF_M);
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (RU) {
                        // k-block -> (filter row dy, channel block): ONE A box of the tile's pixels of that row, from
                        // one pixel to the left of the tile (row -1) on, and the W boxes of the row's three taps
                        const int dy = kb / ep.cv_kbt, cb = kb - dy * ep.cv_kbt;
                        const int a_col = cb * BLOCK_K, a_row = m_tile * TILE_M + (dy - 1) * ep.cv_wp - 1;
                        mbar_expect_tx(&full[s], halves * C::A_HALF_BYTES + C::A_TAIL_BYTES + C::B_STAGE_BYTES);
                        for (int h = 0; h < halves; ++h)
                            tma_load_2d(smem_a + s * C::A_BYTES + h * C::A_HALF_BYTES, &tmap_a, &full[s], a_col, a_row + h * HALF_M);
                        tma_load_2d(smem_a + s * C::A_BYTES + halves * C::A_HALF_BYTES, &tmap_a8, &full[s], a_col, a_row + halves * HALF_M);
                        for (int dxi = 0; dxi < 3; ++dxi)
                            tma_load_2d(smem_b + s * C::B_STAGE_BYTES + dxi * C::B_BYTES, &tmap_w, &full[s],
                                        ((dy * 3 + dxi) * ep.cv_kbt + cb) * BLOCK_K, n_tile * BLOCK_N);
                        continue;
                    }
                    mbar_expect_tx(&full[s], halves * C::A_HALF_BYTES + C::B_BYTES);
                    int a_col = kb * BLOCK_K, a_row = m_tile * TILE_M;
                    if (CONV && ep.cv_taps != 1) {
                        // k-block -> (filter tap, channel block): the A box of tap (dy, dx) is the same 128 pixels
                        // shifted by dy rows and dx columns of the padded grid; rows outside the tensor read as zero
                        const int tap = kb / ep.cv_kbt, dy = tap / 3;
                        a_col = (kb - tap * ep.cv_kbt) * BLOCK_K;
                        a_row += (dy - 1) * ep.cv_wp + (tap - dy * 3 - 1);
                    }
                    for (int h = 0; h < halves; ++h)
                        tma_load_2d(smem_a + s * C::A_BYTES + h * C::A_HALF_BYTES, &tmap_a, &full[s], a_col, a_row + h * HALF_M);
                    tma_load_2d(smem_b + s * C::B_BYTES, &tmap_w, &full[s], kb * BLOCK_K, n_tile * BLOCK_N);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(HALF_M, BLOCK_N);
            uint32_t it = 0;
            Work w;
            for (int idx = 0; get_work(pl, cta, grid, idx, w); ++idx) {
                int m_tile, n_tile;
                tile_coords(pl, w.tile, m_tile, n_tile);
                const int halves = min(MH, (pl.m_live - m_tile * TILE_M + HALF_M - 1) / HALF_M);
                const int a = C::NUM_ACC == 2 ? (idx & 1) : 0;
                const uint32_t aph = C::NUM_ACC == 2 ? (((uint32_t)idx >> 1) & 1) : ((uint32_t)idx & 1);
                mbar_wait(&tmem_empty[a], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(a * C::ACC_COLS);
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (tr && it == 0) tr[TR_FIRST_FULL] = global_ns();
                    if (RU) {
                        // tap (dy, dx): the same rows, dx + 1 rows (128 B each) further (see umma_smem_desc: the swizzle
                        // follows the absolute address, so the shifted descriptor reads the shifted rows)
#pragma unroll
                        for (int dxi = 0; dxi < 3; ++dxi) {
                            const uint32_t a_addr = smem_u32(smem_a + s * C::A_BYTES) + (uint32_t)dxi * 128u;
                            const uint64_t adesc = umma_smem_desc(a_addr);
                            const uint64_t bdesc = umma_smem_desc(smem_u32(smem_b + s * C::B_STAGE_BYTES + dxi * C::B_BYTES));
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                const uint32_t acc = (kb > w.kb0 || k > 0 || dxi > 0) ? 1u : 0u;
                                umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                                if (MH == 2 && halves == 2)
                                    umma_f16(tmem_d + BLOCK_N, adesc + (uint64_t)(C::A_HALF_BYTES >> 4) + (uint64_t)(2 * k),
                                             bdesc + (uint64_t)(2 * k), idesc, acc);
                            }
                        }
                        umma_commit(&empty[s]);
                        continue;
                    }
                    const uint64_t adesc = umma_smem_desc(smem_u32(smem_a + s * C::A_BYTES));
                    const uint64_t bdesc = umma_smem_desc(smem_u32(smem_b + s * C::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field;
                        // the second half's A box sits A_HALF_BYTES further and accumulates BLOCK_N columns further
                        const uint32_t acc = (kb > w.kb0 || k > 0) ? 1u : 0u;
                        umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                        if (MH == 2 && halves == 2)
                            umma_f16(tmem_d + BLOCK_N, adesc + (uint64_t)(C::A_HALF_BYTES >> 4) + (uint64_t)(2 * k),
                                     bdesc + (uint64_t)(2 * k), idesc, acc);
                    }
                    umma_commit(&empty[s]);          // frees the smem slot when these MMAs retire
                }
                umma_commit(&tmem_full[a]);          // accumulator (or partial accumulator) complete
            }
            if (tr) { tr[TR_MMA_DONE] = global_ns(); }
        }
    } else {
        // ===================== epilogue (4 warps per 128-row half) =====================
        const int q = warp & 3;                      // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;            // which accumulator half
        const int trow = half * HALF_M + q * 32 + lane;   // row of the tile owned by this thread
        const bool leader = threadIdx.x == 64;
        Work w;
        int units = 0;
        for (int idx = 0; get_work(pl, cta, grid, idx, w); ++idx, ++units) {
            int m_tile, n_tile;
            tile_coords(pl, w.tile, m_tile, n_tile);
            const int a = C::NUM_ACC == 2 ? (idx & 1) : 0;
            const uint32_t aph = C::NUM_ACC == 2 ? (((uint32_t)idx >> 1) & 1) : ((uint32_t)idx & 1);
            mbar_wait(&tmem_full[a], aph);
            tc_fence_after();
            if (tr && leader && idx == 0) tr[TR_ACC_READY] = global_ns();
            const int row = m_tile * TILE_M + trow;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * C::ACC_COLS + half * BLOCK_N);
            bool row_ok = row < pl.m_live;
            size_t orow = (size_t)row;
            bool border = false;
            if (CONV && row_ok) {
                border = conv_row(ep, row, orow);
                if (border && ep.cv_unpad) row_ok = false;
            }
            const bool half_live = m_tile * TILE_M + half * HALF_M < pl.m_live;    // warp-uniform
            // the final part of a split tile waits for the earlier parts (always scheduled on lower unit ids)
            const int peers = w.kind == WORK_OWNER ? pl.parts - 1 : 0;
            if (w.kind == WORK_OWNER) {
                if (leader) {
                    for (int j = 0; j < peers; ++j)
                        while (flag_acquire(ep.flags + j * pl.rem_tiles + w.rem) == 0) { }
                }
                epi_bar_sync<C::EPI_THREADS>();
            }
            constexpr size_t SLOT_F4 = (size_t)TILE_M * BLOCK_N / 4;
            if (half_live) {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    tmem_ld_wait();
                    const int col0 = n_tile * BLOCK_N + c * 32;
                    if (w.kind == WORK_PARTIAL) {
                        uint4 *dst = (uint4 *)ep.ws + (size_t)(w.part * pl.rem_tiles + w.rem) * SLOT_F4;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            __stcg(dst + partial_f4(c, j >> 2, trow, TILE_M), make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
                        continue;
                    }
                    if (!(row_ok && col0 < ep.N)) continue;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    for (int pj = 0; pj < peers; pj += 2) {         // fixed part order => deterministic sum
                        const float4 *s0 = (const float4 *)ep.ws + (size_t)(pj * pl.rem_tiles + w.rem) * SLOT_F4;
                        const bool two = pj + 1 < peers;
                        const float4 *s1 = two ? s0 + (size_t)pl.rem_tiles * SLOT_F4 : s0;
                        float4 t0[8], t1[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {               // 16 independent L2 reads in flight per thread
                            t0[j] = __ldcg(s0 + partial_f4(c, j, trow, TILE_M));
                            t1[j] = __ldcg(s1 + partial_f4(c, j, trow, TILE_M));
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[4 * j] += t0[j].x; v[4 * j + 1] += t0[j].y; v[4 * j + 2] += t0[j].z; v[4 * j + 3] += t0[j].w;
                        }
                        if (two) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                v[4 * j] += t1[j].x; v[4 * j + 1] += t1[j].y; v[4 * j + 2] += t1[j].z; v[4 * j + 3] += t1[j].w;
                            }
                        }
                    }
                    bias_act32(v, ep.bias, col0, ep.N, ep.act, ep.act_aux);
                    if (CONV && border) {                    // the zero border of the next layer's input
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
                    if (ep.out_dtype == AZN_DTYPE_BF16) {
                        __nv_bfloat16 *dst = (__nv_bfloat16 *)ep.out + orow * ep.ldo + col0;
                        if (col0 + 32 <= ep.N && ((uintptr_t)dst & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint32_t pk[4];
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                                    pk[t] = *(uint32_t *)&h2;
                                }
                                *(uint4 *)(dst + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < ep.N) dst[j] = __float2bfloat16_rn(v[j]);
                        }
                    } else {
                        float *dst = (float *)ep.out + orow * ep.ldo + col0;
                        if (col0 + 32 <= ep.N && ((uintptr_t)dst & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *(float4 *)(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < ep.N) dst[j] = v[j];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[a]);
            if (w.kind == WORK_PARTIAL && !pl.finish) {
                __threadfence();                      // partial visible device-wide before the flag
                epi_bar_sync<C::EPI_THREADS>();
                if (leader) flag_release(ep.flags + w.part * pl.rem_tiles + w.rem);
            } else if (w.kind == WORK_OWNER) {
                epi_bar_sync<C::EPI_THREADS>();       // every epilogue thread is done reading the peers' slots
                if (leader)
                    for (int j = 0; j < peers; ++j) ep.flags[j * pl.rem_tiles + w.rem] = 0;   // clean for the next launch
            }
        }
        if (tr && leader) { tr[TR_EPI_DONE] = global_ns(); tr[TR_UNITS] = units; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (pl.finish && pl.parts > 1) {                 // uniform over the grid: the plan is a pure function of m_live
        __threadfence();                             // this CTA's partial dumps, before the barrier publishes them
        grid_barrier(ep.sync);
        finish_phase<CONV>(pl, N, BLOCK_N, TILE_M, ep);
    }
    if (tr && threadIdx.x == 0) tr[TR_END] = global_ns();
}

// Row softmax over the first `ncls` columns of a f32 [M, ld] matrix, in place
// (softmax_layer.cpp:28-60: max-subtract, exp, sum, divide).  One warp per row.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float *__restrict__ x, const int32_t *__restrict__ m_live_ptr, int M_cap, int ld, int ncls) {
    const int m = m_live_ptr ? min(max(*m_live_ptr, 0), M_cap) : M_cap;
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < m; row += gridDim.x * (blockDim.x >> 5)) {
        float *r = x + (size_t)row * ld;
        float mx = -INFINITY;
        for (int j = lane; j < ncls; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        float s = 0.f;
        for (int j = lane; j < ncls; j += 32) { float e = expf(r[j] - mx); r[j] = e; s += e; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        for (int j = lane; j < ncls; j += 32) r[j] = __fdiv_rn(r[j], s);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct MapKey {
    const void *ptr;
    int rows, cols, box_rows;
    bool operator==(const MapKey &o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows; }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        return std::hash<const void *>()(k.ptr) ^ ((size_t)k.rows * 1315423911u) ^ ((size_t)k.cols << 20) ^ (size_t)k.box_rows;
    }
};

// bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle, zero OOB fill.
int make_tmap(const void *ptr, int rows, int cols, int box_rows, CUtensorMap *out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, rows, cols, box_rows};
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return AZN_OK; }
    EncodeTiledFn enc = get_encode();
    if (!enc) { azn_set_error("cuTensorMapEncodeTiled entry point not found"); return AZN_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { azn_set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d box_rows=%d", (int)r, rows, cols, box_rows); return AZN_ERR_CUDA; }
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return AZN_OK;
}

// Tile choice: 256 x 256 (two accumulators, 128 FLOP per L2 byte) for the wide layers that carry the FLOPs
// (int6 / fc6 / fc7); narrower single-accumulator tiles for the small layers, where more tiles beat splitting
// the K loop.
int g_force_parts = 0, g_force_finish = -1, g_force_bn = 0, g_force_mh = 0;
int g_conv_reuse = 1, g_conv_bn128_cin = 1 << 30;          // azn_conv_tune; AZN_CONV_REUSE=0 in the environment starts without
long long *g_trace = nullptr;
void pick_tile(int N, int &bn, int &mh) {
    bn = N >= 2048 ? 256 : (N > 64 ? 128 : 64);
    mh = N >= 2048 ? 2 : 1;
    // (measured, round 2: for the mid-width int7 layer, N = 1280, the 128 x 128 tile beats 256 x 128, 128 x 256 and
    // 256 x 256 in CUDA-graph replay -- 0.789 vs 0.814 / 0.809 / 0.830 ms per step with four batches in flight)
    if (g_force_bn == 64 || g_force_bn == 128 || g_force_bn == 256) bn = g_force_bn;
    if (g_force_mh == 1 || g_force_mh == 2) mh = g_force_mh;
    if (bn == 64) mh = 1;
}

template <int BLOCK_N, int MH, bool CONV = false, int RU = 0>
int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tw, const int32_t *m_live, int M_cap, int N, int K,
                const EpiParams &ep, int grid, cudaStream_t s, const CUtensorMap *ta8 = nullptr) {
    using C = Cfg<BLOCK_N, MH, RU>;
    static bool attr = false;
    if (!attr) {
        AZN_CUDA(cudaFuncSetAttribute(fc_gemm_kernel<BLOCK_N, MH, CONV, RU>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    // every CTA of the persistent grid must be co-resident (flag spin-waits, grid barrier): cooperative launch
    static int max_grid = 0;
    if (!max_grid) {
        int per_sm = 0;
        AZN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fc_gemm_kernel<BLOCK_N, MH, CONV, RU>, C::THREADS, C::SMEM_BYTES));
        max_grid = per_sm * azn_num_sms();
    }
    if (grid > max_grid) {
        azn_set_error("azn_fc_forward: a grid of %d persistent CTAs cannot be co-resident on this device (max %d)", grid, max_grid);
        return AZN_ERR_CUDA;
    }
    AZN_CUDA(azn_launch_coop(fc_gemm_kernel<BLOCK_N, MH, CONV, RU>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, s, ta, tw, m_live, M_cap, N, K, ep,
                             ta8 ? *ta8 : ta));
    return AZN_OK;
}

}  // namespace

extern "C" void azn_fc_tune(int parts, int finish_mode, int block_n) {
    g_force_parts = parts;
    g_force_finish = finish_mode;
    g_force_bn = block_n % 1000;             // block_n + 1000 * rows-halves: 1256 = 128 x 256 tiles, 2256 = 256 x 256
    g_force_mh = block_n / 1000;
}

extern "C" void azn_fc_trace(long long *device_buffer) { g_trace = device_buffer; }

// Tuning hook of azn_conv3x3_forward: reuse 1 (default) = the RU kernels (one A box per filter row), 0 = one A box per tap
// (round 1); bn128_max_cin: wide layers (Cout >= 256) with at most this many input channels run RU on 256 x 128 tiles
// (default: all of them), the others keep 256 x 256 tiles with one A box per tap.
extern "C" void azn_conv_tune(int reuse, int bn128_max_cin) {
    g_conv_reuse = reuse ? 1 : 0;
    g_conv_bn128_cin = bn128_max_cin < 0 ? (1 << 30) : bn128_max_cin;
}

extern "C" size_t azn_fc_workspace_bytes(int M_cap, int N, int K) {
    (void)M_cap; (void)K; (void)N;
    // partial-accumulator slots (128 MB: 512 slots of 256 x 256 fp32, or 1024 narrower ones) + one flag word per slot
    return WS_DATA_BYTES + 8192;
}

extern "C" int azn_fc_forward(const void *A, const void *W, const float *bias, void *out, int out_dtype, int ldo,
                              int M_cap, const int32_t *m_live, int N, int K, int act, int act_aux, void *workspace,
                              size_t workspace_bytes, azn_stream_t stream) {
    AZN_REQUIRE(A && W && bias && out, "azn_fc_forward: null pointer");
    AZN_REQUIRE(M_cap > 0 && N > 0 && K > 0, "azn_fc_forward: bad shape M=%d N=%d K=%d", M_cap, N, K);
    AZN_REQUIRE(K % BLOCK_K == 0, "azn_fc_forward: K=%d must be a multiple of %d", K, BLOCK_K);
    AZN_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "azn_fc_forward: A and W must be 16-byte aligned");
    AZN_REQUIRE(out_dtype == AZN_DTYPE_F32 || out_dtype == AZN_DTYPE_BF16, "azn_fc_forward: bad out dtype");
    AZN_REQUIRE(ldo >= N, "azn_fc_forward: ldo=%d < N=%d", ldo, N);
    AZN_REQUIRE(act >= AZN_ACT_NONE && act <= AZN_ACT_SOFTMAX_BBOX, "azn_fc_forward: bad activation %d", act);
    AZN_REQUIRE(act != AZN_ACT_AZ_HEAD || (act_aux > 0 && 5 * act_aux + 1 <= N), "azn_fc_forward: AZ head needs act_aux = nsub with 5*nsub+1 <= N");
    AZN_REQUIRE(act != AZN_ACT_SOFTMAX_BBOX || (out_dtype == AZN_DTYPE_F32 && act_aux > 0 && act_aux <= N),
                "azn_fc_forward: softmax needs f32 output and 0 < classes <= N");
    cudaStream_t s = (cudaStream_t)stream;
    int bn, mh;
    pick_tile(N, bn, mh);
    const int grid = azn_num_sms();
    const size_t need = WS_DATA_BYTES + 8192;        // partials + flag block + barrier counter
    static_assert(WS_MAX_SLOTS * sizeof(int) <= 4096, "flag block");
    if (!workspace || workspace_bytes < need) {
        azn_set_error("azn_fc_forward: workspace %zu < %zu bytes", workspace_bytes, need);
        return AZN_ERR_CAPACITY;
    }
    CUtensorMap ta, tw;
    int rc = make_tmap(A, M_cap, K, HALF_M, &ta);
    if (rc) return rc;
    rc = make_tmap(W, N, K, bn, &tw);
    if (rc) return rc;
    EpiParams ep;
    ep.bias = bias; ep.out = out; ep.out_dtype = out_dtype; ep.ldo = ldo; ep.N = N;
    ep.act = act == AZN_ACT_SOFTMAX_BBOX ? AZN_ACT_NONE : act;
    ep.act_aux = act_aux;
    ep.force_parts = g_force_parts;
    ep.force_finish = g_force_finish;
    ep.ws = (float *)workspace;
    ep.flags = (int *)((char *)workspace + WS_DATA_BYTES);
    ep.sync = (unsigned long long *)((char *)workspace + WS_DATA_BYTES + WS_MAX_SLOTS * sizeof(int));
    ep.trace = g_trace;
    ep.cv_kbt = ep.cv_wp = ep.cv_hp = ep.cv_plane = ep.cv_unpad = ep.cv_taps = ep.cv_reuse = 0;
    if (bn == 256 && mh == 2) rc = launch_gemm<256, 2>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    else if (bn == 256) rc = launch_gemm<256, 1>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    else if (bn == 128 && mh == 2) rc = launch_gemm<128, 2>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    else if (bn == 128) rc = launch_gemm<128, 1>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    else rc = launch_gemm<64, 1>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    if (rc) return rc;
    if (act == AZN_ACT_SOFTMAX_BBOX) {
        softmax_rows_kernel<<<grid, 256, 0, s>>>((float *)out, m_live, M_cap, ldo, act_aux);
        AZN_LAUNCH_CHECK();
    }
    return AZN_OK;
}

// 3x3 / pad 1 / stride 1 convolution + bias + ReLU on the tensor cores, as an implicit GEMM.
// replaces: ConvolutionLayer::Forward (caffe-fast-rcnn/src/caffe/layers/conv_layer.cpp: im2col + sgemm per image,
// base_conv_layer.cpp forward_cpu_gemm / forward_cpu_bias) followed by the in-place ReLU layers of
// models/Pascal/VGG16/az-net/test.prototxt:16-384.
//
// Layout trick: every map lives in HBM as a ZERO-BORDERED channels-last grid [n_img, H+2, W+2, C] bf16.  Seen as a
// row-major matrix [P = n_img*(H+2)*(W+2), C], the input of filter tap (dy, dx) for output pixel p is simply row
// p + dy*(W+2) + dx -- the zero border supplies the padding and keeps a shifted row from wrapping into the
// neighbouring image row.  So the convolution is the fc GEMM  Y[P, Cout] = sum_tap X[P + shift(tap), Cin] . W_tap^T
// with a K loop over 9 taps x Cin/64 channel blocks, fed by the same TMA boxes (the A box just starts `shift` rows
// further; rows before 0 / past P are zero-filled by TMA) into the same tcgen05 pipeline.  Border rows of the
// output are computed like any other and then stored as zeros (they are the next layer's padding): (H+2)(W+2)/(HW)
// of the minimum work, 1.03x at 240x400, 1.11x at 30x50.
static int conv_grid_forward(const void *X, const void *Wt, const float *bias, void *Y, int n_img, int H, int W,
                             int Cin, int Cout, int relu, int out_unpadded, void *workspace, size_t workspace_bytes,
                             azn_stream_t stream, int taps) {
    AZN_REQUIRE(X && Wt && bias && Y, "azn_conv3x3_forward: null pointer");
    AZN_REQUIRE(n_img > 0 && H > 0 && W > 0, "azn_conv3x3_forward: bad shape n=%d H=%d W=%d", n_img, H, W);
    AZN_REQUIRE(Cin > 0 && Cin % BLOCK_K == 0, "azn_conv3x3_forward: Cin=%d must be a multiple of %d (pad the channels)", Cin, BLOCK_K);
    AZN_REQUIRE(Cout > 0 && Cout % 8 == 0, "azn_conv3x3_forward: Cout=%d must be a multiple of 8", Cout);
    AZN_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Wt % 16 == 0) && ((uintptr_t)Y % 16 == 0), "azn_conv3x3_forward: pointers must be 16-byte aligned");
    const long long P = (long long)n_img * (H + 2) * (W + 2);
    AZN_REQUIRE(P < (1ll << 31) - 4096, "azn_conv3x3_forward: %lld padded pixels exceed the 32-bit row index", P);
    const size_t need = WS_DATA_BYTES + 8192;
    if (!workspace || workspace_bytes < need) {
        azn_set_error("azn_conv3x3_forward: workspace %zu < %zu bytes", workspace_bytes, need);
        return AZN_ERR_CAPACITY;
    }
    const int K = taps * Cin;
    // Wide layers (Cout >= 256) of a 3x3 convolution also run the RU kernel, on 256 x 128 tiles: 155 FLOP per delivered byte
    // against 131 for 256 x 256 tiles with one A box per tap.  MEASURED per 16 images: conv3_1 0.249 -> 0.201 ms, conv3_2 / 3_3
    // 0.387 -> 0.34-0.355, conv4_1 0.213 -> 0.182-0.193, conv4_2 / 4_3 0.349 -> 0.345, conv5_x 0.114 -> 0.103; backbone 3985 ->
    // 4130-4220 images/s.  azn_conv_tune(1, c) keeps 256 x 256 tiles for the layers with more than c input channels (A/B).
    const int reuse_env = g_conv_reuse, bn128_cin = g_conv_bn128_cin;
    const int bn = (Cout >= 256 && !(taps == 9 && reuse_env != 0 && Cin <= bn128_cin)) ? 256 : (Cout > 64 ? 128 : 64);
    CUtensorMap ta, tw;
    int rc = make_tmap(X, (int)P, Cin, HALF_M, &ta);
    if (rc) return rc;
    rc = make_tmap(Wt, Cout, K, bn, &tw);
    if (rc) return rc;
    EpiParams ep;
    ep.bias = bias; ep.out = Y; ep.out_dtype = AZN_DTYPE_BF16; ep.ldo = Cout; ep.N = Cout;
    ep.act = relu ? AZN_ACT_RELU : AZN_ACT_NONE;
    ep.act_aux = 0;
    ep.force_parts = g_force_parts;
    ep.force_finish = g_force_finish;
    ep.ws = (float *)workspace;
    ep.flags = (int *)((char *)workspace + WS_DATA_BYTES);
    ep.sync = (unsigned long long *)((char *)workspace + WS_DATA_BYTES + WS_MAX_SLOTS * sizeof(int));
    ep.trace = nullptr;
    ep.cv_kbt = Cin / BLOCK_K; ep.cv_wp = W + 2; ep.cv_hp = H + 2; ep.cv_plane = (H + 2) * (W + 2); ep.cv_unpad = out_unpadded ? 1 : 0; ep.cv_taps = taps;
    ep.cv_reuse = 0;
    const int grid = azn_num_sms();
    cudaStream_t s = (cudaStream_t)stream;
    // Row re-use (RU kernels) for the layers whose tiles are bound by operand delivery, not by the tensor pipe: Cout <= 128
    // (conv1_2, conv2_x: 52 / 85 FLOP per delivered byte with one A box per tap; 110 / 150 with one per filter row).
    // MEASURED per 16 images of 480x800: conv1_2 0.893 -> 0.676 ms, conv2_1 0.315 -> 0.252, conv2_2 0.496 -> 0.373; backbone
    // 3483 -> 3973 images/s on the same box.  azn_conv_tune(0, 0) goes without (A/B).
    if (taps == 9 && bn <= 128 && reuse_env != 0) {
        CUtensorMap ta8;
        rc = make_tmap(X, (int)P, Cin, 8, &ta8);
        if (rc) return rc;
        ep.cv_reuse = 1;
        if (bn == 128) return launch_gemm<128, 2, true, 1>(ta, tw, nullptr, (int)P, Cout, 3 * Cin, ep, grid, s, &ta8);
        return launch_gemm<64, 2, true, 1>(ta, tw, nullptr, (int)P, Cout, 3 * Cin, ep, grid, s, &ta8);
    }
    if (bn == 256) rc = launch_gemm<256, 2, true>(ta, tw, nullptr, (int)P, Cout, K, ep, grid, s);
    else if (bn == 128) rc = launch_gemm<128, 2, true>(ta, tw, nullptr, (int)P, Cout, K, ep, grid, s);
    else rc = launch_gemm<64, 2, true>(ta, tw, nullptr, (int)P, Cout, K, ep, grid, s);
    return rc;
}

extern "C" int azn_conv3x3_forward(const void *X, const void *Wt, const float *bias, void *Y, int n_img, int H, int W,
                                   int Cin, int Cout, int relu, int out_unpadded, void *workspace, size_t workspace_bytes,
                                   azn_stream_t stream) {
    return conv_grid_forward(X, Wt, bias, Y, n_img, H, W, Cin, Cout, relu, out_unpadded, workspace, workspace_bytes, stream, 9);
}

// The same convolution over rows whose 3x3 neighbourhood is already gathered into K (azn_patches3x3): one tap, no
// row shifts -- a pointwise GEMM over the padded grid with the border-zeroing epilogue.
extern "C" int azn_conv_patches_forward(const void *Xp, const void *Wt, const float *bias, void *Y, int n_img, int H, int W,
                                        int Kp, int Cout, int relu, int out_unpadded, void *workspace, size_t workspace_bytes,
                                        azn_stream_t stream) {
    return conv_grid_forward(Xp, Wt, bias, Y, n_img, H, W, Kp, Cout, relu, out_unpadded, workspace, workspace_bytes, stream, 1);
}
