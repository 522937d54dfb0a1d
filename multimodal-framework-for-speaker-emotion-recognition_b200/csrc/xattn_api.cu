// Host side of lsthm_xattn_fwd / lsthm_xattn_bwd (include/lsthm_b200.h).
#include "../../include/lsthm_b200.h"
#include "xattn_kernels.cuh"

namespace lsthm {
int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);
}  // namespace lsthm
using namespace lsthm;

static int xattn_check(const lsthm_xattn_desc *d) {
    if (!d) return fail_msg("null attention descriptor");
    if (d->B < 1 || d->L < 1 || d->L > 128) return fail_msg("lsthm_xattn: need B >= 1 and 1 <= L <= 128");
    if (d->D < 4 || d->D > kXD || (d->D & 3)) return fail_msg("lsthm_xattn: width D must be a multiple of 4 in [4, 128]");
    if ((d->ldq | d->ldk | d->ldv | d->ldo) & 3) return fail_msg("lsthm_xattn: row strides must be multiples of 4 floats");
    if (d->ldq < d->D || d->ldk < d->D || d->ldv < d->D || d->ldo < d->D) return fail_msg("lsthm_xattn: row strides must be >= D");
    if (d->p_drop < 0.f || d->p_drop >= 1.f) return fail_msg("lsthm_xattn: p_drop must be in [0,1)");
    if (d->row_stride_b < 0 || d->row_stride_i < 0) return fail_msg("lsthm_xattn: row strides must be >= 0");
    return 0;
}
static void xattn_fill(const lsthm_xattn_desc *d, XAttnArgs &a) {
    a.B = d->B; a.L = d->L; a.D = d->D; a.ldq = d->ldq; a.ldk = d->ldk; a.ldv = d->ldv; a.ldo = d->ldo;
    a.lddq = d->ldq; a.lddk = d->ldk; a.lddv = d->ldv;
    a.scale = d->scale; a.p_drop = d->p_drop; a.seed = d->seed;
    const bool dflt = d->row_stride_b == 0 && d->row_stride_i == 0;
    a.sb = dflt ? d->L : d->row_stride_b;
    a.si = dflt ? 1 : d->row_stride_i;
}
static bool misaligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

extern "C" {

int lsthm_xattn_fwd(const lsthm_xattn_desc *d, const float *q, const float *k, const float *v, float *out, float *lse,
                    void *stream) {
    if (xattn_check(d)) return 1;
    if (!q || !k || !v || !out) return fail_msg("lsthm_xattn_fwd: null pointer");
    if (misaligned(q) || misaligned(k) || misaligned(v) || misaligned(out)) return fail_msg("lsthm_xattn_fwd: operands must be 16-byte aligned");
    XAttnArgs a{};
    xattn_fill(d, a);
    a.q = q; a.k = k; a.v = v; a.out = out; a.lse = lse;
    const size_t smem = 3 * kXSlot;
    cudaError_t e = cudaFuncSetAttribute(xattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lsthm_xattn_fwd shared-memory opt-in", e);
    xattn_fwd_kernel<<<d->B, 256, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_xattn_fwd launch", e);
}

int lsthm_xattn_bwd(const lsthm_xattn_desc *d, const float *q, const float *k, const float *v, const float *dout, float *dq,
                    float *dk, float *dv, void *stream) {
    if (xattn_check(d)) return 1;
    if (!q || !k || !v || !dout || !dq || !dk || !dv) return fail_msg("lsthm_xattn_bwd: null pointer");
    if (misaligned(q) || misaligned(k) || misaligned(v) || misaligned(dout) || misaligned(dq) || misaligned(dk) || misaligned(dv))
        return fail_msg("lsthm_xattn_bwd: operands must be 16-byte aligned");
    XAttnArgs a{};
    xattn_fill(d, a);
    a.q = q; a.k = k; a.v = v; a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv;
    const size_t smem = 3 * kXSlot;
    cudaError_t e = cudaFuncSetAttribute(xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lsthm_xattn_bwd shared-memory opt-in", e);
    xattn_bwd_kernel<<<d->B, 512, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_xattn_bwd launch", e);
}

}  // extern "C"
