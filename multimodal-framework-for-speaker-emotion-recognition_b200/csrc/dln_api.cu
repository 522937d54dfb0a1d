// Host side of lsthm_dln_fwd / lsthm_dln_bwd (include/lsthm_b200.h): fused dropout + residual + LayerNorm.
#include <algorithm>

#include "../../include/lsthm_b200.h"
#include "dln_kernels.cuh"

namespace lsthm {
int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);
}  // namespace lsthm
using namespace lsthm;

static constexpr int kDlnMaxGrid = 148 * 4;   // 4 resident CTAs of 8 warps per SM

static int dln_grid(long long R) {
    const long long need = (R + kDlnWarps - 1) / kDlnWarps;
    return (int)(need < kDlnMaxGrid ? (need < 1 ? 1 : need) : kDlnMaxGrid);
}
static int dln_check(const lsthm_dln_desc *d) {
    if (!d) return fail_msg("null dln descriptor");
    if (d->R < 0) return fail_msg("lsthm_dln: R < 0");
    if (d->d < 4 || d->d > 128 * kDlnMaxC || (d->d & 3)) return fail_msg("lsthm_dln: need 4 <= d <= 512, d % 4 == 0");
    if (d->p_drop < 0.f || d->p_drop >= 1.f) return fail_msg("lsthm_dln: p_drop must be in [0,1)");
    return 0;
}
static bool bad_ld(int ld, int d) { return ld < d || (ld & 3); }

extern "C" {

size_t lsthm_dln_workspace_floats(int32_t d) { return (size_t)kDlnMaxGrid * 3 * (size_t)d; }

int lsthm_dln_fwd(const lsthm_dln_desc *d, const float *y, int32_t ldy, const float *bias, const float *res, int32_t ldres, const float *gamma,
                  const float *beta, float *v, int32_t ldv, float *out, int32_t ldo, void *stream) {
    if (dln_check(d)) return 1;
    if (!y || !res || !gamma || !beta || !out) return fail_msg("lsthm_dln_fwd: null pointer");
    if (bad_ld(ldy, d->d) || bad_ld(ldres, d->d) || bad_ld(ldo, d->d) || (v && bad_ld(ldv, d->d)))
        return fail_msg("lsthm_dln_fwd: row strides must be >= d and multiples of 4 floats");
    if (d->R == 0) return 0;
    DlnArgs a{};
    a.R = d->R; a.d = d->d; a.eps = d->eps; a.p_drop = d->p_drop; a.seed = d->seed;
    a.y = y; a.ldy = ldy; a.bias = bias; a.res = res; a.ldres = ldres; a.gamma = gamma; a.beta = beta; a.v = v; a.ldv = ldv; a.out = out; a.ldo = ldo;
    const int grid = dln_grid(d->R);
    if (d->d <= 128) dln_fwd_kernel<1><<<grid, 32 * kDlnWarps, 0, (cudaStream_t)stream>>>(a);
    else if (d->d <= 256) dln_fwd_kernel<2><<<grid, 32 * kDlnWarps, 0, (cudaStream_t)stream>>>(a);
    else dln_fwd_kernel<4><<<grid, 32 * kDlnWarps, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_dln_fwd launch", e);
}

int lsthm_dln_bwd(const lsthm_dln_desc *d, const float *dout, int32_t lddo, const float *v, int32_t ldv, const float *gamma,
                  float *dy, int32_t lddy, float *dres, int32_t lddres, float *dgamma, float *dbeta, float *dbias,
                  float *workspace, size_t workspace_floats, void *stream) {
    if (dln_check(d)) return 1;
    if (!dout || !v || !gamma || !dres || !workspace) return fail_msg("lsthm_dln_bwd: null pointer");
    if (d->p_drop > 0.f && !dy) return fail_msg("lsthm_dln_bwd: dy is required when p_drop > 0 (it equals dres otherwise)");
    if (bad_ld(lddo, d->d) || bad_ld(ldv, d->d) || bad_ld(lddres, d->d) || (dy && bad_ld(lddy, d->d)))
        return fail_msg("lsthm_dln_bwd: row strides must be >= d and multiples of 4 floats");
    if (workspace_floats < lsthm_dln_workspace_floats(d->d)) return fail_msg("lsthm_dln_bwd: workspace too small");
    DlnArgs a{};
    a.R = d->R; a.d = d->d; a.eps = d->eps; a.p_drop = d->p_drop; a.seed = d->seed;
    a.dout = dout; a.lddo = lddo; a.vin = v; a.ldv = ldv; a.gamma = gamma; a.dy = dy; a.lddy = lddy; a.dres = dres; a.lddres = lddres;
    a.partial = workspace;
    const int grid = dln_grid(d->R);
    if (d->d <= 128) dln_bwd_kernel<1><<<grid, 32 * kDlnWarps, 0, (cudaStream_t)stream>>>(a);
    else if (d->d <= 256) dln_bwd_kernel<2><<<grid, 32 * kDlnWarps, 0, (cudaStream_t)stream>>>(a);
    else dln_bwd_kernel<4><<<grid, 32 * kDlnWarps, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("lsthm_dln_bwd launch", e);
    dln_reduce_kernel<<<(3 * d->d + 31) / 32, 256, 0, (cudaStream_t)stream>>>(workspace, grid, d->d, dgamma, dbeta, dbias);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_dln_bwd reduce launch", e);
}

static int colsum_chunks(long long R, int C) {
    const int coltiles = (C + 127) / 128;
    long long n = kDlnMaxGrid / coltiles;
    if (n < 1) n = 1;
    const long long maxc = (R + 63) / 64;                         // at least 64 rows per chunk
    return (int)(n < maxc ? n : (maxc < 1 ? 1 : maxc));
}

size_t lsthm_colsum_workspace_floats(int64_t R, int32_t C) { return (size_t)colsum_chunks(R, C) * (size_t)C; }

int lsthm_colsum(int64_t R, int32_t C, const float *A, int32_t ld, float *out, float *workspace, size_t workspace_floats, void *stream) {
    if (R < 1 || C < 4 || (C & 3) || !A || !out || !workspace) return fail_msg("lsthm_colsum: need R >= 1, C % 4 == 0 and non-null pointers");
    if (ld < C || (ld & 3) || (reinterpret_cast<uintptr_t>(A) & 15)) return fail_msg("lsthm_colsum: rows must be 16-byte aligned (ld % 4 == 0)");
    const int nchunk = colsum_chunks(R, C);
    if (workspace_floats < (size_t)nchunk * C) return fail_msg("lsthm_colsum: workspace too small");
    const int rpc = (int)((R + nchunk - 1) / nchunk);
    colsum_partial_kernel<<<dim3((C + 127) / 128, nchunk), 256, 0, (cudaStream_t)stream>>>(A, R, C, ld, rpc, workspace);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("lsthm_colsum launch", e);
    colsum_final_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(workspace, nchunk, C, out);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_colsum final launch", e);
}

int lsthm_assemble_input(int64_t R, int32_t d_text, int32_t d_audio, const float *r1, const float *r2, const float *r3, const float *r4,
                         const float *acouf, float *x, void *stream) {
    if (R < 0 || d_text < 4 || d_audio < 4 || (d_text & 3) || (d_audio & 3)) return fail_msg("lsthm_assemble_input: widths must be positive multiples of 4");
    if (!r1 || !r2 || !r3 || !r4 || !acouf || !x) return fail_msg("lsthm_assemble_input: null pointer");
    if ((reinterpret_cast<uintptr_t>(r1) | reinterpret_cast<uintptr_t>(r2) | reinterpret_cast<uintptr_t>(r3) | reinterpret_cast<uintptr_t>(r4) |
         reinterpret_cast<uintptr_t>(acouf) | reinterpret_cast<uintptr_t>(x)) & 15)
        return fail_msg("lsthm_assemble_input: operands must be 16-byte aligned");
    if (R == 0) return 0;
    const long long total = R * ((d_text + d_audio) / 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    assemble_input_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(r1), reinterpret_cast<const float4 *>(r2), reinterpret_cast<const float4 *>(r3),
        reinterpret_cast<const float4 *>(r4), reinterpret_cast<const float4 *>(acouf), reinterpret_cast<float4 *>(x), R, d_text / 4, d_audio / 4);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_assemble_input launch", e);
}


int lsthm_reverse_seq(int32_t L, int32_t B, int32_t w, const float *X, const int32_t *len, float *out, void *stream) {
    if (L < 1 || B < 1 || w < 2 || (w & 1)) return fail_msg("lsthm_reverse_seq: need L, B >= 1 and an even row width");
    if (!X || !len || !out) return fail_msg("lsthm_reverse_seq: null pointer");
    if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(out)) & 7) return fail_msg("lsthm_reverse_seq: operands must be 8-byte aligned");
    const long long total = (long long)L * B * (w / 2);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    reverse_seq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(X), len, reinterpret_cast<float2 *>(out), L, B, w / 2);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_reverse_seq launch", e);
}

}  // extern "C"
