// Fused Adam step over a flat fp32 buffer (one launch per gradient bucket) — the optimizer the reference's
// trainer uses: torch.optim.Adam(lr, weight_decay) with L2-style decay (model_trainer.py:82), applied to the
// flat parameter/gradient buckets of ddp.GradAllReducer right after the allreduce.
#include <cmath>

#include "../../include/lsthm_b200.h"
#include "common.cuh"

namespace lsthm {
int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, size_t n, float lr_over_bc1, float inv_sqrt_bc2, float beta1,
                                                   float beta2, float eps, float wd) {
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gg = reinterpret_cast<const float4 *>(g)[i];
        float *pa = &pp.x, *ma = &mm.x, *va = &vv.x;
        const float *ga = &gg.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = ga[k] + wd * pa[k];                       // L2 decay folded into the gradient (torch Adam)
            ma[k] = ma[k] + (gr - ma[k]) * (1.f - beta1);              // exp_avg.lerp_(grad, 1 - beta1)
            va[k] = va[k] * beta2 + (1.f - beta2) * gr * gr;
            pa[k] -= lr_over_bc1 * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
        }
        reinterpret_cast<float4 *>(p)[i] = pp; reinterpret_cast<float4 *>(m)[i] = mm; reinterpret_cast<float4 *>(v)[i] = vv;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gr = g[i] + wd * p[i];
        const float mi = m[i] + (gr - m[i]) * (1.f - beta1), vi = v[i] * beta2 + (1.f - beta2) * gr * gr;
        m[i] = mi; v[i] = vi;
        p[i] -= lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}
}  // namespace lsthm
using namespace lsthm;

extern "C" int lsthm_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, size_t n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int32_t step, void *stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq) return fail_msg("lsthm_adam_step: null pointer");
    if (step < 1) return fail_msg("lsthm_adam_step: step counts from 1");
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return fail_msg("lsthm_adam_step: buffers must be 16-byte aligned");
    if (n == 0) return 0;
    const double bc1 = 1.0 - std::pow((double)beta1, (double)step), bc2 = 1.0 - std::pow((double)beta2, (double)step);
    const int blocks = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, 148 * 8);
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1),
                                                          (float)(1.0 / std::sqrt(bc2)), beta1, beta2, eps, weight_decay);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_adam_step launch", e);
}
