// Peer-memory plumbing for the fused tile exchange of a multi-GPU render.
//
// Instead of rendering a tile and then calling an all-gather, the compositing kernel - the last kernel of every chunk -
// stores each finished packed pixel row straight into the image buffer of EVERY rank: its own with a local store, the
// others' through NVLink peer memory (cudaIpc mappings of buffers that each process allocates here).  The exchange then
// overlaps the render chunk by chunk and costs 48 B x ranks of posted stores per ray; what remains at the end of a frame
// is one 4-byte all-reduce that orders "every rank's kernels are complete" (ucnerf_b200/peer.py).
//
// One process per GPU (torchrun): device memory cannot be shared through torch's caching allocator safely, so the image
// buffers are plain cudaMalloc allocations owned by this library; their IPC handles travel through torch.distributed.
#include "../../include/ucnerf_b200.h"
#include <cstring>

#include "common.cuh"

using namespace ucnerf;

extern "C" int ucnerf_peer_alloc(uint64_t bytes, void** dptr_out, uint8_t* handle64_out) {
    UC_REQUIRE(dptr_out && handle64_out && bytes > 0, "peer_alloc: null argument / zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* p = nullptr;
    UC_CUDA_OK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
        return 2;
    }
    UC_CUDA_OK(cudaMemset(p, 0, bytes));
    std::memcpy(handle64_out, &h, 64);
    *dptr_out = p;
    return 0;
}

extern "C" int ucnerf_peer_open(const uint8_t* handle64, void** dptr_out) {
    UC_REQUIRE(handle64 && dptr_out, "peer_open: null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* p = nullptr;
    UC_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr_out = p;
    return 0;
}

extern "C" int ucnerf_peer_close(void* dptr) {
    if (dptr) UC_CUDA_OK(cudaIpcCloseMemHandle(dptr));
    return 0;
}

extern "C" int ucnerf_peer_free(void* dptr) {
    if (dptr) UC_CUDA_OK(cudaFree(dptr));
    return 0;
}
