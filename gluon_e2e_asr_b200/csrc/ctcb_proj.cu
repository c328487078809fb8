// ctcb_proj.cu -- translation unit of the tcgen05 projection kernel (ctcb_proj.cuh) and its launcher.
#define CTCB_PROJ_IMPL
#define CTCB_NO_PLAIN_KERNELS
#include "ctcb_proj.cuh"

namespace ctcb {

cudaError_t launch_proj_emit(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const ProjArgs& a, dim3 grid, size_t smem,
                             cudaStream_t stream) {
    // the opt-in shared-memory size is a per-device function attribute: raised when a launch needs more than any before
    static thread_local int done_dev = -1;
    static thread_local size_t done_bytes = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (done_dev != dev || done_bytes < smem) {
        const cudaError_t rc = cudaFuncSetAttribute(k_proj_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        done_dev = dev; done_bytes = smem;
    }
    k_proj_emit<<<grid, kPThreads, smem, stream>>>(tmA, tmB, tmC, a);
    return cudaGetLastError();
}

}  // namespace ctcb
