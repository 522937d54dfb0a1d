// Host side of lsthm_attn_fwd / lsthm_attn_bwd (include/lsthm_b200.h).
#include "../../include/lsthm_b200.h"
#include "attn_kernels.cuh"

namespace lsthm {
int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);
}  // namespace lsthm
using namespace lsthm;

static int attn_check(const lsthm_attn_desc *d) {
    if (!d) return fail_msg("null attention descriptor");
    if (d->B < 1 || d->H < 1 || d->L < 1 || d->L > kAttLP) return fail_msg("lsthm_attn: need 1 <= L <= 128");
    if (d->d_head != kAttD) return fail_msg("lsthm_attn: d_k = d_v = 40 only (encoder.py: d_k = d_v = 40)");
    if ((d->ldq | d->ldk | d->ldv | d->ldo) & 3) return fail_msg("lsthm_attn: row strides must be multiples of 4 floats");
    if (d->p_drop < 0.f || d->p_drop >= 1.f) return fail_msg("lsthm_attn: p_drop must be in [0,1)");
    if (d->precision != 0 && d->precision != 1) return fail_msg("lsthm_attn: precision must be 0 (fp32-accurate split) or 1 (bf16 operands)");
    if (d->row_stride_b < 0 || d->row_stride_i < 0) return fail_msg("lsthm_attn: row strides must be >= 0");
    return 0;
}
static void attn_fill(const lsthm_attn_desc *d, AttnArgs &a) {
    a.B = d->B; a.L = d->L; a.H = d->H; a.ldq = d->ldq; a.ldk = d->ldk; a.ldv = d->ldv; a.ldo = d->ldo;
    a.scale = d->scale; a.p_drop = d->p_drop; a.seed = d->seed;
    // row index of (dialogue b, position i) = b*row_stride_b + i*row_stride_i; both 0 = batch-major [B][L]
    const bool dflt = d->row_stride_b == 0 && d->row_stride_i == 0;
    a.sb = dflt ? d->L : d->row_stride_b;
    a.si = dflt ? 1 : d->row_stride_i;
}

extern "C" {

int lsthm_attn_fwd(const lsthm_attn_desc *d, const float *q, const float *k, const float *v, float *out, float *lse,
                   void *stream) {
    if (attn_check(d)) return 1;
    if (!q || !k || !v || !out) return fail_msg("lsthm_attn_fwd: null pointer");
    AttnArgs a{};
    attn_fill(d, a);
    a.q = q; a.k = k; a.v = v; a.out = out; a.lse = lse;
    const size_t smem = 2 * kSqTile + 2 * kRowTile + 1024;
    auto kern = d->precision == 1 ? attn_fwd_kernel<true> : attn_fwd_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lsthm_attn_fwd shared-memory opt-in", e);
    kern<<<d->B * d->H, 256, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_attn_fwd launch", e);
}

int lsthm_attn_bwd(const lsthm_attn_desc *d, const float *q, const float *k, const float *v, const float *out, const float *lse,
                   const float *dout, float *dq, float *dk, float *dv, void *stream) {
    if (attn_check(d)) return 1;
    if (!q || !k || !v || !out || !lse || !dout || !dq || !dk || !dv) return fail_msg("lsthm_attn_bwd: null pointer");
    AttnArgs a{};
    attn_fill(d, a);
    a.q = q; a.k = k; a.v = v; a.o = out; a.lse = const_cast<float *>(lse); a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv;
    const size_t smem = 8 * kRowTile + 4 * kSqTile;
    auto kern = d->precision == 1 ? attn_bwd_kernel<true> : attn_bwd_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lsthm_attn_bwd shared-memory opt-in", e);
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
            sm_count = 148;
    }
    const int items = d->B * d->H;
    kern<<<items < sm_count ? items : sm_count, kAttBwdThreads, smem, (cudaStream_t)stream>>>(a);   // persistent: one CTA per SM
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_attn_bwd launch", e);
}

}  // extern "C"
