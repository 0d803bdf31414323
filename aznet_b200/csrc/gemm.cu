// Fully connected layers of the AZ-Net / Fast R-CNN heads on the 5th-generation tensor cores.
//
//   out[M, N] = act(A[M, K] . W[N, K]^T + bias[N])        A, W bf16 (K-major), fp32 accumulate
//
// replaces InnerProductLayer::Forward (caffe-fast-rcnn/src/caffe/layers/inner_product_layer.cpp:80-93:
// one cblas/cublas sgemm for the product, a second rank-1 sgemm for the bias) and the ReLU / Sigmoid
// layers that follow it in models/Pascal/VGG16/az-net/test_fc.prototxt:26-232.
//
// Kernel anatomy (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of a 128x64 A box and a BLOCK_Nx64 W box
//               (128-byte swizzle) into a STAGES-deep shared-memory ring, completion on mbarriers.
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma.cta_group::1
//               .kind::f16 (UMMA 128 x BLOCK_N x 16, bf16 in, fp32 accumulate) straight from the
//               swizzled shared tiles through UMMA smem descriptors; tcgen05.commit releases the
//               ring slot / publishes the accumulator.
//   warps 2-5   epilogue: tcgen05.ld the accumulator (one TMEM lane = one output row per thread),
//               bias + ReLU / sigmoid, 16-byte stores.  Two accumulator buffers in TMEM
//               (2 x BLOCK_N columns) let the epilogue of tile i overlap the main loop of tile i+1.
// Scheduling is decided ON THE DEVICE from the live row count (the search keeps its region counts
// in HBM): tiles = ceil(m_live/128) x ceil(N/BLOCK_N).  Whole waves of tiles go one per CTA
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

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;


template <int BLOCK_N> struct Cfg {
    static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
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
constexpr int WS_SLOTS = 1024;         // partial-accumulator slots in the workspace (one flag each)
// measured fix-up charges, in SM cycles: an owner CTA adds one peer's partial in ~4 us (latency-bound L2
// reads by 128 threads); the stand-alone finish kernel costs ~12 us but reduces all parts in parallel
constexpr int PEER_CYCLES = 8000;
constexpr int FINISH_CYCLES = 24000;

// Work decomposition, identical on every CTA and every warp role (pure function of m_live and the grid).
struct Plan {
    int m_live, m_tiles, n_tiles, tiles, kblocks;
    int dp_tiles, rem_tiles, parts;        // data-parallel tiles (first), split tiles (last), parts per split tile
    int finish;                            // 1: every part dumps a partial and fc_finish_kernel reduces them
    int n_major;                           // raster order of tile ids
};
struct Work {
    int tile, kb0, kb1, kind, part, rem;   // rem = index of the tile among the split tiles
};
enum { WORK_FULL = 0, WORK_PARTIAL = 1, WORK_OWNER = 2 };

__device__ __forceinline__ Plan make_plan(const int32_t *m_live_ptr, int M_cap, int N, int K, int block_n, int grid,
                                          int force_parts = 0, int force_finish = -1) {
    Plan p;
    p.m_live = m_live_ptr ? min(max(*m_live_ptr, 0), M_cap) : M_cap;
    p.m_tiles = (p.m_live + BLOCK_M - 1) / BLOCK_M;
    p.n_tiles = (N + block_n - 1) / block_n;
    p.tiles = p.m_tiles * p.n_tiles;
    p.kblocks = K / BLOCK_K;
    p.rem_tiles = p.tiles % grid;
    p.dp_tiles = p.tiles - p.rem_tiles;
    p.parts = 1;
    p.finish = 0;
    if (p.rem_tiles > 0) {
        const int kb_cycles = 2 * block_n;                 // 4 UMMAs of 128 x block_n x 16
        int best = p.kblocks * kb_cycles;                  // P = 1: one wave of whole tiles
        for (int P = 2; P <= MAX_PARTS; ++P) {
            if (p.kblocks / P < MIN_KB_PER_PART || p.rem_tiles * P > WS_SLOTS) break;
            const int waves = (p.rem_tiles * P + grid - 1) / grid;
            const int mma = waves * ((p.kblocks + P - 1) / P) * kb_cycles;
            const int c_in = mma + PEER_CYCLES * (P - 1), c_fin = mma + FINISH_CYCLES;
            if (c_in < best) { best = c_in; p.parts = P; p.finish = 0; }
            if (c_fin < best) { best = c_fin; p.parts = P; p.finish = 1; }
        }
    }
    if (force_parts > 0 && p.rem_tiles > 0) {
        p.parts = force_parts;
        while (p.parts > 1 && (p.kblocks / p.parts < 1 || p.rem_tiles * p.parts > WS_SLOTS)) --p.parts;
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
// consecutive 16-byte words: float4 index ((chunk*8 + j4) * 128 + row).
__device__ __forceinline__ size_t partial_f4(int chunk, int j4, int row) { return ((size_t)(chunk * 8 + j4) * BLOCK_M + row); }

__device__ __forceinline__ void flag_release(int *flag) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
__device__ __forceinline__ int flag_acquire(const int *flag) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float sigmoid_caffe(float x) {
    // sigmoid_layer.cpp:11-13: exp in float, `1. / (1. + e)` in double, rounded to float
    return (float)(1.0 / (1.0 + (double)expf(-x)));
}
// AZN_ACT_AZ_HEAD: columns are [adj_score (nsub) | adj_bbox (4*nsub) | zoom_score (1)]; the two score
// groups go through the Sigmoid layers adj_prob / zoom_prob (test_fc.prototxt:221-232).
__device__ __forceinline__ float apply_act(float v, int act, int col, int nsub) {
    if (act == AZN_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == AZN_ACT_AZ_HEAD) return (col < nsub || col == 5 * nsub) ? sigmoid_caffe(v) : v;
    return v;
}

struct EpiParams {
    const float *bias;
    void *out;
    int out_dtype, ldo, N, act, act_aux;
    int force_parts, force_finish;   // tuning hook (azn_fc_tune): 0 / -1 = automatic
    float *ws;           // split partials: [slot][BLOCK_N/4][BLOCK_M] float4, slot = part * rem_tiles + rem
    int *flags;          // [WS_SLOTS] "the partial of this slot is complete" (zero between launches)
};

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
fc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const int32_t *__restrict__ m_live_ptr, int M_cap, int N, int K, EpiParams ep) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + C::STAGES * C::A_BYTES;
    uint64_t *bars = (uint64_t *)(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + C::STAGES, *tmem_full = bars + 2 * C::STAGES, *tmem_empty = bars + 2 * C::STAGES + 2;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Plan pl = make_plan(m_live_ptr, M_cap, N, K, BLOCK_N, gridDim.x, ep.force_parts, ep.force_finish);
    const int cta = blockIdx.x, grid = gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            Work w;
            for (int idx = 0; get_work(pl, cta, grid, idx, w); ++idx) {
                int m_tile, n_tile;
                tile_coords(pl, w.tile, m_tile, n_tile);
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], C::STAGE_BYTES);
                    tma_load_2d(smem_a + s * C::A_BYTES, &tmap_a, &full[s], kb * BLOCK_K, m_tile * BLOCK_M);
                    tma_load_2d(smem_b + s * C::B_BYTES, &tmap_w, &full[s], kb * BLOCK_K, n_tile * BLOCK_N);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(BLOCK_M, BLOCK_N);
            uint32_t it = 0;
            Work w;
            for (int idx = 0; get_work(pl, cta, grid, idx, w); ++idx) {
                const int a = idx & 1;
                const uint32_t aph = ((uint32_t)idx >> 1) & 1;
                mbar_wait(&tmem_empty[a], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(a * BLOCK_N);
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t adesc = umma_smem_desc(smem_u32(smem_a + s * C::A_BYTES));
                    const uint64_t bdesc = umma_smem_desc(smem_u32(smem_b + s * C::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field
                        umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);          // frees the smem slot when these MMAs retire
                }
                umma_commit(&tmem_full[a]);          // accumulator (or partial accumulator) complete
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                      // TMEM lane quadrant this warp may access
        const int trow = q * 32 + lane;              // row of the tile owned by this thread
        Work w;
        for (int idx = 0; get_work(pl, cta, grid, idx, w); ++idx) {
            int m_tile, n_tile;
            tile_coords(pl, w.tile, m_tile, n_tile);
            const int a = idx & 1;
            const uint32_t aph = ((uint32_t)idx >> 1) & 1;
            mbar_wait(&tmem_full[a], aph);
            tc_fence_after();
            const int row = m_tile * BLOCK_M + trow;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BLOCK_N);
            const bool row_ok = row < pl.m_live;
            // the final part of a split tile waits for the earlier parts (always scheduled on lower unit ids)
            const int peers = w.kind == WORK_OWNER ? pl.parts - 1 : 0;
            if (w.kind == WORK_OWNER) {
                if (threadIdx.x == 64) {
                    for (int j = 0; j < peers; ++j)
                        while (flag_acquire(ep.flags + j * pl.rem_tiles + w.rem) == 0) { }
                }
                epi_bar_sync();
            }
            constexpr size_t SLOT_F4 = (size_t)BLOCK_M * BLOCK_N / 4;
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
                        __stcg(dst + partial_f4(c, j >> 2, trow), make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
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
                        t0[j] = __ldcg(s0 + partial_f4(c, j, trow));
                        t1[j] = __ldcg(s1 + partial_f4(c, j, trow));
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
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = col0 + j;
                    const float b = col < ep.N ? __ldg(ep.bias + col) : 0.f;
                    v[j] = apply_act(v[j] + b, ep.act, col, ep.act_aux);
                }
                if (ep.out_dtype == AZN_DTYPE_BF16) {
                    __nv_bfloat16 *dst = (__nv_bfloat16 *)ep.out + (size_t)row * ep.ldo + col0;
                    if (col0 + 32 <= ep.N && ((uintptr_t)dst & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint32_t pk[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                                pk[t] = *(uint32_t *)&h;
                            }
                            *(uint4 *)(dst + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < ep.N) dst[j] = __float2bfloat16_rn(v[j]);
                    }
                } else {
                    float *dst = (float *)ep.out + (size_t)row * ep.ldo + col0;
                    if (col0 + 32 <= ep.N && ((uintptr_t)dst & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *(float4 *)(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < ep.N) dst[j] = v[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[a]);
            if (w.kind == WORK_PARTIAL && !pl.finish) {
                __threadfence();                      // partial visible device-wide before the flag
                epi_bar_sync();
                if (threadIdx.x == 64) flag_release(ep.flags + w.part * pl.rem_tiles + w.rem);
            } else if (w.kind == WORK_OWNER) {
                epi_bar_sync();                       // every epilogue thread is done reading the peers' slots
                if (threadIdx.x == 64)
                    for (int j = 0; j < peers; ++j) ep.flags[j * pl.rem_tiles + w.rem] = 0;   // clean for the next launch
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// Finish-kernel mode of the split schedule (many parts): sums the P partials of every split tile in part
// order and applies bias + activation.  One thread per (row, 4 columns); exits at once in the other modes.
__global__ void __launch_bounds__(256)
fc_finish_kernel(const int32_t *__restrict__ m_live_ptr, int M_cap, int N, int K, int block_n, int grid_gemm, EpiParams ep) {
    const Plan pl = make_plan(m_live_ptr, M_cap, N, K, block_n, grid_gemm, ep.force_parts, ep.force_finish);
    if (!pl.finish || pl.parts <= 1) return;
    const int q4 = block_n / 4;                              // float4 columns per tile row
    const size_t slot_f4 = (size_t)BLOCK_M * q4;
    const long total = (long)pl.rem_tiles * BLOCK_M * q4;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int trow = (int)(e % BLOCK_M);                 // consecutive threads -> consecutive rows: coalesced partial reads
        const int c4 = (int)((e / BLOCK_M) % q4);
        const int rem = (int)(e / ((long)BLOCK_M * q4));
        int m_tile, n_tile;
        tile_coords(pl, pl.dp_tiles + rem, m_tile, n_tile);
        const int row = m_tile * BLOCK_M + trow, col0 = n_tile * block_n + c4 * 4;
        if (row >= pl.m_live || col0 >= N) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int part = 0; part < pl.parts; ++part) {
            const float4 t = __ldcg((const float4 *)ep.ws + (size_t)(part * pl.rem_tiles + rem) * slot_f4 +
                                    partial_f4(c4 >> 3, c4 & 7, trow));
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const float v[4] = {acc.x, acc.y, acc.z, acc.w};
        for (int j = 0; j < 4; ++j) {
            const int col = col0 + j;
            if (col >= N) break;
            const float o = apply_act(v[j] + ep.bias[col], ep.act, col, ep.act_aux);
            if (ep.out_dtype == AZN_DTYPE_BF16) ((__nv_bfloat16 *)ep.out)[(size_t)row * ep.ldo + col] = __float2bfloat16_rn(o);
            else ((float *)ep.out)[(size_t)row * ep.ldo + col] = o;
        }
    }
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

// 128 x 256 tiles (87 FLOP per L2 byte) for the wide layers that carry the FLOPs (int6 / fc6 / fc7);
// narrower tiles for the small layers, where more tiles beat splitting the K loop.
int g_force_parts = 0, g_force_finish = -1, g_force_bn = 0;
int pick_block_n(int N) {
    if (g_force_bn == 64 || g_force_bn == 128 || g_force_bn == 256) return g_force_bn;
    return N >= 2048 ? 256 : (N > 64 ? 128 : 64);
}

template <int BLOCK_N>
int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tw, const int32_t *m_live, int M_cap, int N, int K,
                const EpiParams &ep, int grid, cudaStream_t s) {
    using C = Cfg<BLOCK_N>;
    static bool attr = false;
    if (!attr) {
        AZN_CUDA(cudaFuncSetAttribute(fc_gemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    fc_gemm_kernel<BLOCK_N><<<grid, GEMM_THREADS, C::SMEM_BYTES, s>>>(ta, tw, m_live, M_cap, N, K, ep);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

}  // namespace

extern "C" void azn_fc_tune(int parts, int finish_mode, int block_n) {
    g_force_parts = parts;
    g_force_finish = finish_mode;
    g_force_bn = block_n;
}

extern "C" size_t azn_fc_workspace_bytes(int M_cap, int N, int K) {
    (void)M_cap; (void)K;
    // WS_SLOTS partial accumulators of 128 x BLOCK_N fp32 + one flag word per slot
    (void)N;
    return (size_t)WS_SLOTS * BLOCK_M * 256 * sizeof(float) + 4096;    // sized for the widest tile
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
    const int bn = pick_block_n(N);
    const int grid = azn_num_sms();
    const size_t slots = (size_t)WS_SLOTS * BLOCK_M * 256 * sizeof(float);
    const size_t need = slots + 4096;
    static_assert(WS_SLOTS * sizeof(int) <= 4096, "flag block");
    if (!workspace || workspace_bytes < need) {
        azn_set_error("azn_fc_forward: workspace %zu < %zu bytes", workspace_bytes, need);
        return AZN_ERR_CAPACITY;
    }
    CUtensorMap ta, tw;
    int rc = make_tmap(A, M_cap, K, BLOCK_M, &ta);
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
    ep.flags = (int *)((char *)workspace + slots);
    if (bn == 256) rc = launch_gemm<256>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    else if (bn == 128) rc = launch_gemm<128>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    else rc = launch_gemm<64>(ta, tw, m_live, M_cap, N, K, ep, grid, s);
    if (rc) return rc;
    // the split mode is decided on the device; the finish kernel returns at once unless it is needed, and is
    // not even launched when the capacity rules a split out (tile count of M_cap a multiple of the grid is
    // not knowable for a live count, so only the static case is skipped)
    {
        const long tiles_cap = (long)((M_cap + BLOCK_M - 1) / BLOCK_M) * ((N + bn - 1) / bn);
        if (m_live != nullptr || tiles_cap % grid != 0) {
            fc_finish_kernel<<<grid * 4, 256, 0, s>>>(m_live, M_cap, N, K, bn, grid, ep);
            AZN_LAUNCH_CHECK();
        }
    }
    if (act == AZN_ACT_SOFTMAX_BBOX) {
        softmax_rows_kernel<<<grid, 256, 0, s>>>((float *)out, m_live, M_cap, ldo, act_aux);
        AZN_LAUNCH_CHECK();
    }
    return AZN_OK;
}
