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
