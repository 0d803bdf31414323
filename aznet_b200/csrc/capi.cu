// Version / error plumbing of the C ABI (include/aznet_b200.h).
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void azn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int g_azn_pdl = 1;
extern "C" void azn_set_pdl(int on) { g_azn_pdl = on ? 1 : 0; }
int g_azn_coop = 1;
extern "C" void azn_set_coop(int on) { g_azn_coop = on ? 1 : 0; }

extern "C" const char *azn_version(void) { return "aznet_b200 0.1 (sm_100a)"; }
extern "C" const char *azn_last_error(void) { return g_err; }

extern "C" int azn_check_device(void) {
    int dev = 0;
    AZN_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    AZN_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) {
        azn_set_error("device %d is sm_%d%d; libaznet_b200 is built for sm_100a only", dev, p.major, p.minor);
        return AZN_ERR_CUDA;
    }
    return AZN_OK;
}

// Write-only HBM probe: every thread streams 16-byte stores (L1 no-allocate), grid-stride, fully coalesced.
__global__ void __launch_bounds__(512) hbm_write_probe_kernel(uint4 *__restrict__ dst, size_t n_vec) {
    const uint4 v = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

extern "C" int azn_hbm_write_probe(void *dst, size_t bytes, azn_stream_t stream) {
    AZN_REQUIRE(dst && ((uintptr_t)dst % 16 == 0) && bytes >= 16, "azn_hbm_write_probe: bad buffer");
    hbm_write_probe_kernel<<<azn_num_sms() * 8, 512, 0, (cudaStream_t)stream>>>((uint4 *)dst, bytes / 16);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
