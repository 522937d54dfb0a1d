// MaskedLoss.forward of the reference (loss.py:13-21, weight = None) as one pass each way:
//     loss = sum_r L(pred[r] * mask[r], target[r]) / sum(mask)
// with L = cross entropy (log-softmax + NLL, the train.py default) or NLL.  A padded row (mask 0) becomes an all-zero logit
// row and contributes the constant log C to the cross entropy, with no gradient — reproduced, not "fixed" (SURVEY.md a-10).
// The library route is a mask multiply, a log-softmax and a single-block nll reduction over all L*B rows (0.19 ms per ATV step
// for the two nll kernels alone); here every row is read once and the row sums are reduced in a fixed order (deterministic).
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>

#include "../../include/lsthm_b200.h"

namespace lsthm {
int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);

constexpr int kLossThreads = 256, kLossMaxBlocks = 296, kLossMaxC = 32;

__device__ __forceinline__ float block_sum(float v, float *sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < kLossThreads / 32; ++w) s += sh[w];      // fixed order
    __syncthreads();
    return s;
}

// partial[b] = sum of row losses of block b, partial[nb + b] = sum of its mask values
__global__ void __launch_bounds__(kLossThreads) masked_loss_rows_kernel(const float *__restrict__ pred, const long long *__restrict__ target,
                                                                        const float *__restrict__ mask, long long R, int C, int kind,
                                                                        float *__restrict__ partial) {
    __shared__ float sh[kLossThreads / 32];
    float acc = 0.f, macc = 0.f;
    for (long long r = (long long)blockIdx.x * kLossThreads + threadIdx.x; r < R; r += (long long)gridDim.x * kLossThreads) {
        const float m = __ldg(mask + r);
        const float *p = pred + r * C;
        const int t = (int)__ldg(target + r);
        float pt = 0.f, mx = -INFINITY;
        float v[kLossMaxC];
#pragma unroll 8
        for (int c = 0; c < C; ++c) { v[c] = __ldg(p + c) * m; mx = fmaxf(mx, v[c]); if (c == t) pt = v[c]; }
        float row = -pt;
        if (kind == 0) {
            float s = 0.f;
#pragma unroll 8
            for (int c = 0; c < C; ++c) s += expf(v[c] - mx);
            row += mx + logf(s);
        }
        acc += row;
        macc += m;
    }
    const float s = block_sum(acc, sh), ms = block_sum(macc, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s; partial[gridDim.x + blockIdx.x] = ms; }
}
// out[0] = loss, out[1] = sum(mask)
__global__ void masked_loss_final_kernel(const float *__restrict__ partial, int nb, float *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f, ms = 0.f;
        for (int b = 0; b < nb; ++b) { s += partial[b]; ms += partial[nb + b]; }
        out[0] = s / ms;
        out[1] = ms;
    }
}
// dpred[r][c] = g / sum(mask) * mask[r] * (softmax(pred[r] * mask[r])[c] - [c == target[r]])      (NLL: without the softmax term)
__global__ void __launch_bounds__(kLossThreads) masked_loss_bwd_kernel(const float *__restrict__ pred, const long long *__restrict__ target,
                                                                       const float *__restrict__ mask, const float *__restrict__ stats,
                                                                       const float *__restrict__ gout, long long R, int C, int kind,
                                                                       float *__restrict__ dpred) {
    const float scale = __ldg(gout) / __ldg(stats + 1);
    for (long long r = (long long)blockIdx.x * kLossThreads + threadIdx.x; r < R; r += (long long)gridDim.x * kLossThreads) {
        const float m = __ldg(mask + r);
        const float *p = pred + r * C;
        float *d = dpred + r * C;
        const int t = (int)__ldg(target + r);
        const float gs = scale * m;
        if (kind == 0) {
            float v[kLossMaxC], mx = -INFINITY, s = 0.f;
#pragma unroll 8
            for (int c = 0; c < C; ++c) { v[c] = __ldg(p + c) * m; mx = fmaxf(mx, v[c]); }
#pragma unroll 8
            for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - mx); s += v[c]; }
            const float inv = 1.f / s;
#pragma unroll 8
            for (int c = 0; c < C; ++c) d[c] = gs * (v[c] * inv - (c == t ? 1.f : 0.f));
        } else {
#pragma unroll 8
            for (int c = 0; c < C; ++c) d[c] = c == t ? -gs : 0.f;
        }
    }
}
}  // namespace lsthm
using namespace lsthm;

static int loss_blocks(long long R) { return (int)std::min<long long>((R + kLossThreads - 1) / kLossThreads, kLossMaxBlocks); }

extern "C" {

size_t lsthm_masked_loss_workspace_floats(int64_t R) { return R < 1 ? 0 : 2 * (size_t)loss_blocks(R); }

int lsthm_masked_loss_fwd(int64_t R, int32_t C, int32_t kind, const float *pred, const int64_t *target, const float *mask,
                          float *workspace, float *out2, void *stream) {
    if (R < 1 || C < 1 || C > kLossMaxC) return fail_msg("lsthm_masked_loss: need R >= 1 and 1 <= C <= 32");
    if (kind != 0 && kind != 1) return fail_msg("lsthm_masked_loss: kind must be 0 (cross entropy) or 1 (NLL)");
    if (!pred || !target || !mask || !workspace || !out2) return fail_msg("lsthm_masked_loss_fwd: null pointer");
    const int nb = loss_blocks(R);
    cudaStream_t st = (cudaStream_t)stream;
    masked_loss_rows_kernel<<<nb, kLossThreads, 0, st>>>(pred, reinterpret_cast<const long long *>(target), mask, R, C, kind, workspace);
    masked_loss_final_kernel<<<1, 32, 0, st>>>(workspace, nb, out2);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_masked_loss_fwd launch", e);
}

int lsthm_masked_loss_bwd(int64_t R, int32_t C, int32_t kind, const float *pred, const int64_t *target, const float *mask,
                          const float *out2, const float *gout, float *dpred, void *stream) {
    if (R < 1 || C < 1 || C > kLossMaxC) return fail_msg("lsthm_masked_loss: need R >= 1 and 1 <= C <= 32");
    if (kind != 0 && kind != 1) return fail_msg("lsthm_masked_loss: kind must be 0 (cross entropy) or 1 (NLL)");
    if (!pred || !target || !mask || !out2 || !gout || !dpred) return fail_msg("lsthm_masked_loss_bwd: null pointer");
    masked_loss_bwd_kernel<<<loss_blocks(R), kLossThreads, 0, (cudaStream_t)stream>>>(pred, reinterpret_cast<const long long *>(target), mask, out2,
                                                                                      gout, R, C, kind, dpred);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_masked_loss_bwd launch", e);
}

}  // extern "C"
