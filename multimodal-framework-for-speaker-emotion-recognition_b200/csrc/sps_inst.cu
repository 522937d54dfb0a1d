// Instantiates the speaker-state cell kernels for one tile height (compile with -DLSTHM_MT=n).
#include "sps_kernels.cuh"

#ifndef LSTHM_MT
#error "compile with -DLSTHM_MT=<1..8>"
#endif
#define LSTHM_CAT2(a, b) a##b
#define LSTHM_CAT(a, b) LSTHM_CAT2(a, b)

namespace lsthm {

int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);

template <typename K, typename A>
static int coop_launch(K kernel, const A &args, int grid, size_t smem_bytes, cudaStream_t st, const char *what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return set_error(what, e);
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSpsThreads, smem_bytes);
    if (e != cudaSuccess) return set_error(what, e);
    if (grid > per_sm * sms)
        return fail_msg("lsthm_sps: the shard needs more co-resident CTAs than the device has SMs (8 dialogues per CTA: N <= 8 x #SMs = "
                        "1184 on a B200).  The reference couples the dialogues of a shard through its packed speaker rows "
                        "(model/lsthm_sps.py:238-259), so a shard cannot be split inside one call: shard the batch across GPUs or steps.");
    void *params[] = {const_cast<A *>(&args)};
    e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(kSpsThreads), params, smem_bytes, st);
    return e == cudaSuccess ? 0 : set_error(what, e);
}

// MODE 1 (GRU speaker state): dialogues are independent, so an ordinary launch of any grid size
template <typename K, typename A>
static int plain_launch(K kernel, const A &args, int grid, size_t smem_bytes, cudaStream_t st, const char *what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return set_error(what, e);
    kernel<<<grid, kSpsThreads, smem_bytes, st>>>(args);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error(what, e);
}

int LSTHM_CAT(launch_sps_fwd_, LSTHM_MT)(const SpsFwdArgs &a, int grid, size_t smem_bytes, cudaStream_t st) {
    return coop_launch(sps_fwd_kernel<LSTHM_MT, 0>, a, grid, smem_bytes, st, "lsthm_sps_fwd launch");
}
int LSTHM_CAT(launch_sps_bwd_, LSTHM_MT)(const SpsBwdArgs &a, int grid, size_t smem_bytes, cudaStream_t st) {
    return coop_launch(sps_bwd_kernel<LSTHM_MT, 0>, a, grid, smem_bytes, st, "lsthm_sps_bwd launch");
}
int LSTHM_CAT(launch_gsp_fwd_, LSTHM_MT)(const SpsFwdArgs &a, int grid, size_t smem_bytes, cudaStream_t st) {
    return plain_launch(sps_fwd_kernel<LSTHM_MT, 1>, a, grid, smem_bytes, st, "lsthm_gsp_fwd launch");
}
int LSTHM_CAT(launch_gsp_bwd_, LSTHM_MT)(const SpsBwdArgs &a, int grid, size_t smem_bytes, cudaStream_t st) {
    return plain_launch(sps_bwd_kernel<LSTHM_MT, 1>, a, grid, smem_bytes, st, "lsthm_gsp_bwd launch");
}

}  // namespace lsthm
