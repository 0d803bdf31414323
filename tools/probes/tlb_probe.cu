// Store-side page-locality probe: every warp instruction writes 8 x 64-byte pieces (4 lanes x 16 B each) -- the store
// shape of the staged ROI pool.  mode 0: the 8 pieces of an instruction go to 8 different 2 MB pages drawn from the
// CTA's working set of P pages; mode 1: all 8 pieces go to the same page (stride 7 KB), pages drawn per instruction.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 1) probe(uint4 *buf, size_t cta_bytes, int P, int iters, int mode) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, j = lane & 3;
    char *base = (char *)buf + (size_t)blockIdx.x * cta_bytes;
    unsigned s = (blockIdx.x * 32 + warp) * 2654435761u + 12345u;
    const uint4 v = make_uint4(lane, warp, blockIdx.x, 7);
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        unsigned pg, off;
        if (mode == 0) {
            const unsigned h = (s >> 8) + g * 0x9E3779B1u;
            pg = (h >> 11) % (unsigned)P;
            off = (h & 2047u) * 1024u;                  // 1 KB-aligned row inside the 2 MB page
        } else {
            pg = (s >> 19) % (unsigned)P;
            off = (((s >> 8) & 1023u) * 1024u + g * 7168u) & (2097152u - 1024u);
        }
        uint4 *p = (uint4 *)(base + (size_t)pg * 2097152u + off + (size_t)((s >> 4) & 15u) * 64u) + j;
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}
int main() {
    const int ctas = 148, iters = 4096;
    for (int mode = 0; mode < 2; ++mode)
        for (int P = 1; P <= 256; P *= 2) {
            const size_t cta_bytes = (size_t)P * 2097152u;
            uint4 *buf;
            if (cudaMalloc(&buf, cta_bytes * ctas) != cudaSuccess) { printf("alloc failed at P=%d\n", P); break; }
            cudaMemset(buf, 0, cta_bytes * ctas);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            probe<<<ctas, 1024>>>(buf, cta_bytes, P, 64, mode);
            cudaEventRecord(e0);
            probe<<<ctas, 1024>>>(buf, cta_bytes, P, iters, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)ctas * 32 * iters * 512;
            printf("mode %d pages/CTA %3d: %.3f ms  %.0f GB/s\n", mode, P, ms, bytes / ms / 1e6);
            cudaFree(buf);
        }
    return 0;
}
