// Shared helpers of libaznet_b200.so: error plumbing and small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/aznet_b200.h"

void azn_set_error(const char *fmt, ...);

#define AZN_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            azn_set_error(__VA_ARGS__);             \
            return AZN_ERR_INVALID;                 \
        }                                           \
    } while (0)

#define AZN_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            azn_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return AZN_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define AZN_LAUNCH_CHECK() AZN_CUDA(cudaGetLastError())

static inline int azn_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The level loop is a chain of ~26 short kernels per step.  Launched with the programmatic-stream-
// serialization attribute, kernel K+1 may be scheduled while kernel K drains: its prologue (barrier
// init, TMEM allocation, descriptor prefetch, smem carve-out) and the launch latency overlap K's tail.
// Contract of every kernel launched through azn_launch_pdl: it touches NO global memory (reads or
// writes) before pdl_grid_wait(), which returns when the preceding grid has completed and flushed;
// right after it the kernel lets its own successor start (pdl_launch_dependents), so at most one
// successor is ever waiting.  azn_set_pdl(0) launches the same kernels fully serialised.
__device__ __forceinline__ void pdl_grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_grid_wait(); pdl_launch_dependents(); }

extern int g_azn_pdl;
extern int g_azn_coop;

template <typename... KArgs, typename... Args>
static inline cudaError_t azn_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_azn_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- cooperative launch ---------------------------------------------------------------------------
// The persistent GEMM spins on device-side flags and on a grid-wide barrier: every CTA of its grid (one per SM) must be
// resident at the same time.  A plain launch gives no such guarantee -- if another kernel holds some SMs (a second
// persistent GEMM on another stream, the NMS chain's SM partition, an NCCL kernel) part of the grid waits for SMs that
// the running part never frees: deadlock.  Launched with cudaLaunchAttributeCooperative the driver either places the
// whole grid together (waiting for the SMs it needs) or fails the launch with cudaErrorCooperativeLaunchTooLarge,
// which surfaces as AZN_ERR_CUDA instead of a hang.  Cooperative and programmatic-dependent launch are not combined:
// griddepcontrol.* in the kernel are no-ops without the PDL attribute (measured neutral under CUDA-graph replay,
// DESIGN.md).  azn_set_coop(0) restores the plain / PDL launch (A/B).
template <typename... KArgs, typename... Args>
static inline cudaError_t azn_launch_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                          Args &&...args) {
    if (!g_azn_coop) return azn_launch_pdl(kernel, grid, block, smem, s, static_cast<Args &&>(args)...);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
