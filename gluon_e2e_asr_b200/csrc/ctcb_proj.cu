// ctcb_proj.cu -- translation unit of the tcgen05 projection kernel (ctcb_proj.cuh) and its launcher.
#define CTCB_PROJ_IMPL
#define CTCB_NO_PLAIN_KERNELS
#include "ctcb_proj.cuh"

namespace ctcb {

cudaError_t launch_proj_emit(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const ProjArgs& a, dim3 grid, size_t smem,
                             cudaStream_t stream) {
    // the opt-in shared-memory size is a per-device function attribute: raised when a launch needs more than any before
    static thread_local int done_dev = -1;
    static thread_local size_t done_bytes[2] = {0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (done_dev != dev) { done_dev = dev; done_bytes[0] = done_bytes[1] = 0; }
    const int pair = a.ctas == 2 ? 1 : 0;
    const void* fn = pair ? reinterpret_cast<const void*>(k_proj_emit<2>) : reinterpret_cast<const void*>(k_proj_emit<1>);
    if (done_bytes[pair] < smem) {
        const cudaError_t rc = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        done_bytes[pair] = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(kPThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = pair ? 1 : 0;
    const cudaError_t rc = pair ? cudaLaunchKernelEx(&cfg, k_proj_emit<2>, tmA, tmB, tmC, a) : cudaLaunchKernelEx(&cfg, k_proj_emit<1>, tmA, tmB, tmC, a);
    if (rc != cudaSuccess) return rc;
    return cudaGetLastError();
}

}  // namespace ctcb
