// Instantiates the forward/backward recurrence kernels for one tile height (compile with -DLSTHM_MT=n).
#include "mab_kernels.cuh"

#ifndef LSTHM_MT
#error "compile with -DLSTHM_MT=<1..8>"
#endif
#define LSTHM_CAT2(a, b) a##b
#define LSTHM_CAT(a, b) LSTHM_CAT2(a, b)

namespace lsthm {

int LSTHM_CAT(launch_fwd_, LSTHM_MT)(const FwdArgs &a, int grid, size_t smem_bytes, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(mab_fwd_kernel<LSTHM_MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
    if (e != cudaSuccess) return set_error("lsthm_mab_fwd shared-memory opt-in", e);
    mab_fwd_kernel<LSTHM_MT><<<grid, a.L.nt, smem_bytes, st>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_mab_fwd launch", e);
}

int LSTHM_CAT(launch_bwd_, LSTHM_MT)(const BwdArgs &a, int grid, size_t smem_bytes, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(mab_bwd_kernel<LSTHM_MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
    if (e != cudaSuccess) return set_error("lsthm_mab_bwd shared-memory opt-in", e);
    mab_bwd_kernel<LSTHM_MT><<<grid, a.L.nt, smem_bytes, st>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_mab_bwd launch", e);
}

}  // namespace lsthm
