// Peer windows: a device buffer of one rank mapped into the address space of the other ranks of the box
// (one process per GPU), so that the last kernel of a search step writes its proposal lists straight into
// the collecting rank's memory over NVLink -- no NCCL kernel, no SMs taken from the persistent GEMMs.
//
// Reference semantics: test_proposals appends every image's list to ONE all_boxes and writes ONE
// proposals.pkl after the loop (lib/detect/test.py:508-539); with the images sharded over the ranks the
// collecting rank is rank 0 and the "append" of the other ranks is a store through the window.
//
// The window is plain cudaMalloc memory exported with cudaIpcGetMemHandle (64 opaque bytes the host side
// hands to the other processes however it likes -- torch.distributed object broadcast in aznet_b200/dist.py)
// and opened with cudaIpcOpenMemHandle + cudaIpcMemLazyEnablePeerAccess.  Everything here is host code; the
// stores themselves are the ordinary stores of collect_kernel (search.cu) through the mapped pointer.
#include <string.h>
#include "common.cuh"

extern "C" int azn_peer_alloc(size_t bytes, void **ptr, unsigned char *handle64) {
    AZN_REQUIRE(ptr && handle64 && bytes > 0, "azn_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    AZN_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        azn_set_error("azn_peer_alloc: %s", cudaGetErrorString(e));
        return AZN_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return AZN_OK;
}

extern "C" int azn_peer_open(const unsigned char *handle64, void **ptr) {
    AZN_REQUIRE(ptr && handle64, "azn_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        azn_set_error("azn_peer_open: %s", cudaGetErrorString(e));
        return AZN_ERR_CUDA;
    }
    *ptr = p;
    return AZN_OK;
}

extern "C" int azn_peer_close(void *ptr) {
    AZN_REQUIRE(ptr, "azn_peer_close: null pointer");
    AZN_CUDA(cudaIpcCloseMemHandle(ptr));
    return AZN_OK;
}

extern "C" int azn_peer_free(void *ptr) {
    AZN_REQUIRE(ptr, "azn_peer_free: null pointer");
    AZN_CUDA(cudaFree(ptr));
    return AZN_OK;
}
