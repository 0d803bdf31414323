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
