// Fused  out = LayerNorm(dropout(y + bias) + residual)  of the utterance encoder, forward and backward
// (model/encoder.py:54-58 — `q = self.dropout(self.fc(q)); q += residual; q = self.layer_norm(q)` — and
// :106-112, the same tail after the position-wise feed-forward).  Row-wise and HBM-bound: one warp owns a
// row (d <= 512 floats in registers as float4 chunks), rows are walked grid-stride, every tensor is read or
// written exactly once.  The reference (and stock PyTorch) spends a dropout kernel + mask tensor, an add and a
// LayerNorm kernel on this in the forward and four kernels in the backward.
//
// Dropout is regenerated from (seed, row, column) in the backward — no mask tensor.  The backward also
// emits the column sums every caller needs next: dgamma, dbeta and the bias gradient of the Linear that
// produced y (= column sums of dy), as per-block partials reduced in a fixed order (deterministic).
#pragma once
#include "common.cuh"

namespace lsthm {

constexpr int kDlnMaxC = 4;          // float4 chunks per lane -> d <= 512
constexpr int kDlnWarps = 8;

struct DlnArgs {
    long long R;
    int d;
    const float *y, *bias, *res, *gamma, *beta, *dout, *vin;   // bias (may be NULL): y + bias before the dropout
    float *v, *out, *dy, *dres, *partial;     // partial: [gridDim.x][3][d]  (dgamma, dbeta, dbias)
    int ldy, ldres, ldv, ldo, lddo, lddy, lddres;
    float eps, p_drop;
    unsigned long long seed;
};

// 16 random bits per element, one 32-bit hash per pair of elements (same construction as the attention dropout)
struct RowDrop {
    uint32_t key, thr;
    float scale;
    bool on;
    __device__ __forceinline__ RowDrop(unsigned long long seed, float p) {
        on = p > 0.f;
        thr = (uint32_t)(p * 65536.0f + 0.5f);
        scale = 65536.0f / (65536.0f - (float)thr);
        uint32_t x = (uint32_t)seed * 0x9E3779B9u + 0x7F4A7C15u;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        key = x ^ (uint32_t)(seed >> 32);
    }
    // scales of the 4 elements of float4 chunk number `chunk` (global index over the whole matrix)
    __device__ __forceinline__ float4 quad(unsigned long long chunk) const {
        const uint32_t base = key + (uint32_t)(chunk >> 31) * 0x85EBCA6Bu;
        uint32_t x = base + (uint32_t)(2 * chunk) * 0xC2B2AE35u, z = base + (uint32_t)(2 * chunk + 1) * 0xC2B2AE35u;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        z ^= z >> 16; z *= 0x7feb352du; z ^= z >> 15; z *= 0x846ca68bu; z ^= z >> 16;
        return make_float4((x & 0xffffu) >= thr ? scale : 0.f, (x >> 16) >= thr ? scale : 0.f,
                           (z & 0xffffu) >= thr ? scale : 0.f, (z >> 16) >= thr ? scale : 0.f);
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NC>   // float4 chunks per lane: d <= 128 * NC
__global__ void __launch_bounds__(32 * kDlnWarps, NC == 4 ? 2 : NC == 2 ? 3 : 4) dln_fwd_kernel(const __grid_constant__ DlnArgs a) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nc = a.d >> 2;
    const float inv_d = 1.0f / (float)a.d;
    const RowDrop drop(a.seed, a.p_drop);
    float4 gm[NC], bt[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = lane + 32 * i;
        gm[i] = bt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nc) {
            gm[i] = __ldg(reinterpret_cast<const float4 *>(a.gamma) + c);
            bt[i] = __ldg(reinterpret_cast<const float4 *>(a.beta) + c);
        }
    }
    for (long long row = (long long)blockIdx.x * kDlnWarps + wib; row < a.R; row += (long long)gridDim.x * kDlnWarps) {
        float4 x[NC];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nc) {
                float4 y = __ldcs(reinterpret_cast<const float4 *>(a.y + row * a.ldy) + c);
                const float4 r = __ldg(reinterpret_cast<const float4 *>(a.res + row * a.ldres) + c);
                if (a.bias != nullptr) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias) + c);
                    y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
                }
                if (drop.on) {
                    const float4 m = drop.quad((unsigned long long)row * nc + c);
                    y.x *= m.x; y.y *= m.y; y.z *= m.z; y.w *= m.w;
                }
                x[i] = make_float4(y.x + r.x, y.y + r.y, y.z + r.z, y.w + r.w);
                s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
            }
        }
        const float mean = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            if (lane + 32 * i < nc) {
                const float dx = x[i].x - mean, dy = x[i].y - mean, dz = x[i].z - mean, dw = x[i].w - mean;
                q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * inv_d + a.eps);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            if (c < nc) {
                if (a.v != nullptr) reinterpret_cast<float4 *>(a.v + row * a.ldv)[c] = x[i];
                float4 o;
                o.x = (x[i].x - mean) * rstd * gm[i].x + bt[i].x;
                o.y = (x[i].y - mean) * rstd * gm[i].y + bt[i].y;
                o.z = (x[i].z - mean) * rstd * gm[i].z + bt[i].z;
                o.w = (x[i].w - mean) * rstd * gm[i].w + bt[i].w;
                reinterpret_cast<float4 *>(a.out + row * a.ldo)[c] = o;
            }
        }
    }
}

// dv = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dout * gamma,  xhat = (v - mean) * rstd
// dres = dv,  dy = dv * dropout scale;  partial sums of (dout * xhat, dout, dy) per column
template <int NC>
__global__ void __launch_bounds__(32 * kDlnWarps, NC == 4 ? 2 : NC == 2 ? 3 : 4) dln_bwd_kernel(const __grid_constant__ DlnArgs a) {
    __shared__ float4 red[kDlnWarps][3][32 * NC];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nc = a.d >> 2;
    const float inv_d = 1.0f / (float)a.d;
    const RowDrop drop(a.seed, a.p_drop);
    float4 gm[NC], ag[NC], ab[NC], ay[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = lane + 32 * i;
        gm[i] = ag[i] = ab[i] = ay[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nc) gm[i] = __ldg(reinterpret_cast<const float4 *>(a.gamma) + c);
    }
    for (long long row = (long long)blockIdx.x * kDlnWarps + wib; row < a.R; row += (long long)gridDim.x * kDlnWarps) {
        float4 x[NC], g[NC];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            x[i] = g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nc) {
                x[i] = __ldcs(reinterpret_cast<const float4 *>(a.vin + row * a.ldv) + c);
                g[i] = __ldcs(reinterpret_cast<const float4 *>(a.dout + row * a.lddo) + c);
                s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
            }
        }
        const float mean = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            if (lane + 32 * i < nc) {
                x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
                q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * inv_d + a.eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            if (lane + 32 * i < nc) {
                x[i].x *= rstd; x[i].y *= rstd; x[i].z *= rstd; x[i].w *= rstd;           // xhat
                ag[i].x += g[i].x * x[i].x; ag[i].y += g[i].y * x[i].y; ag[i].z += g[i].z * x[i].z; ag[i].w += g[i].w * x[i].w;
                ab[i].x += g[i].x; ab[i].y += g[i].y; ab[i].z += g[i].z; ab[i].w += g[i].w;
                g[i].x *= gm[i].x; g[i].y *= gm[i].y; g[i].z *= gm[i].z; g[i].w *= gm[i].w;
                sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                sgx += (g[i].x * x[i].x + g[i].y * x[i].y) + (g[i].z * x[i].z + g[i].w * x[i].w);
            }
        }
        const float mg = warp_sum(sg) * inv_d, mgx = warp_sum(sgx) * inv_d;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            if (c < nc) {
                float4 dv;
                dv.x = rstd * (g[i].x - mg - x[i].x * mgx);
                dv.y = rstd * (g[i].y - mg - x[i].y * mgx);
                dv.z = rstd * (g[i].z - mg - x[i].z * mgx);
                dv.w = rstd * (g[i].w - mg - x[i].w * mgx);
                reinterpret_cast<float4 *>(a.dres + row * a.lddres)[c] = dv;
                if (drop.on) {
                    const float4 m = drop.quad((unsigned long long)row * nc + c);
                    dv.x *= m.x; dv.y *= m.y; dv.z *= m.z; dv.w *= m.w;
                    reinterpret_cast<float4 *>(a.dy + row * a.lddy)[c] = dv;
                }
                ay[i].x += dv.x; ay[i].y += dv.y; ay[i].z += dv.z; ay[i].w += dv.w;
            }
        }
    }
    // block partials, fixed warp order
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        red[wib][0][lane + 32 * i] = ag[i];
        red[wib][1][lane + 32 * i] = ab[i];
        red[wib][2][lane + 32 * i] = ay[i];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * nc; idx += blockDim.x) {
        const int which = idx / nc, c = idx - which * nc;
        float4 t = red[0][which][c];
        for (int w = 1; w < kDlnWarps; ++w) {
            const float4 u = red[w][which][c];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        reinterpret_cast<float4 *>(a.partial + ((size_t)blockIdx.x * 3 + which) * a.d)[c] = t;
    }
}

// out[which][col] = sum over blocks of partial[blk][which][col]  (fixed order: 8 interleaved row groups, then the groups)
static __global__ void __launch_bounds__(256) dln_reduce_kernel(const float *__restrict__ partial, int nblk, int d, float *dgamma,
                                                                float *dbeta, float *dbias) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cx;                      // column of the [3*d] vector
    float s = 0.f;
    if (i < 3 * d)
        for (int b = ry; b < nblk; b += 8) s += partial[(size_t)b * 3 * d + i];
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && i < 3 * d) {
        float t = red[0][cx];
#pragma unroll
        for (int r = 1; r < 8; ++r) t += red[r][cx];
        const int which = i / d, col = i - which * d;
        float *dst = which == 0 ? dgamma : which == 1 ? dbeta : dbias;
        if (dst != nullptr) dst[col] = t;
    }
}

}  // namespace lsthm

// ---------------------------------------------------------------------------------------------
// Column sums of a row-major matrix A[R][C] (bias gradients: db = sum over all T*N rows of dy), two stages, fixed
// summation order.  Stage 1: a CTA owns 128 columns (32 float4 lanes) x one chunk of rows, 8 row lanes.
// ---------------------------------------------------------------------------------------------
namespace lsthm {
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float *__restrict__ A, long long R, int C, int ld, int rows_per_chunk,
                                                             float *__restrict__ partial) {
    __shared__ float4 red[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c4 = blockIdx.x * 32 + cx;                         // float4 column index
    const long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * c4 < C) {
        const float4 *p = reinterpret_cast<const float4 *>(A) + c4;
        const size_t ld4 = (size_t)(ld >> 2);
        long long r = r0 + ry;
        for (; r + 24 < r1; r += 32) {                            // 4 independent loads in flight per thread
            const float4 a = __ldcs(p + (size_t)r * ld4), b = __ldcs(p + (size_t)(r + 8) * ld4);
            const float4 c = __ldcs(p + (size_t)(r + 16) * ld4), d = __ldcs(p + (size_t)(r + 24) * ld4);
            s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y);
            s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
        }
        for (; r < r1; r += 8) {
            const float4 a = __ldcs(p + (size_t)r * ld4);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && 4 * c4 < C) {
        float4 t = red[0][cx];
#pragma unroll
        for (int k = 1; k < 8; ++k) { t.x += red[k][cx].x; t.y += red[k][cx].y; t.z += red[k][cx].z; t.w += red[k][cx].w; }
        reinterpret_cast<float4 *>(partial + (size_t)blockIdx.y * C)[c4] = t;
    }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float *__restrict__ partial, int nchunk, int C, float *__restrict__ out) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (c < C)
        for (int b = ry; b < nchunk; b += 8) s += partial[(size_t)b * C + c];
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < C) {
        float t = red[0][cx];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += red[k][cx];
        out[c] = t;
    }
}
}  // namespace lsthm

// ---------------------------------------------------------------------------------------------
// Batch assembly of the trainer (model_trainer.py:104-105):  x = cat(((r1 + r2) + r3 + r4) / 4, acouf)  per utterance,
// one pass: 4 x d_text + d_audio floats read, d_text + d_audio written (the reference runs 3 adds, a divide and a cat:
// seven passes over the 1024-wide RoBERTa layers).  Same operation order as the reference -> bit-identical.
// ---------------------------------------------------------------------------------------------
namespace lsthm {
__global__ void __launch_bounds__(256) assemble_input_kernel(const float4 *__restrict__ r1, const float4 *__restrict__ r2,
                                                             const float4 *__restrict__ r3, const float4 *__restrict__ r4,
                                                             const float4 *__restrict__ ac, float4 *__restrict__ x, long long R, int dt4, int da4) {
    const int w4 = dt4 + da4;
    const long long total = R * w4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / w4;
        const int c = (int)(i - row * w4);
        float4 o;
        if (c < dt4) {
            const long long j = row * dt4 + c;
            const float4 a = __ldcs(r1 + j), b = __ldcs(r2 + j), cc = __ldcs(r3 + j), d = __ldcs(r4 + j);
            o.x = (((a.x + b.x) + cc.x) + d.x) / 4.0f; o.y = (((a.y + b.y) + cc.y) + d.y) / 4.0f;
            o.z = (((a.z + b.z) + cc.z) + d.z) / 4.0f; o.w = (((a.w + b.w) + cc.w) + d.w) / 4.0f;
        } else {
            o = __ldcs(ac + row * da4 + (c - dt4));
        }
        x[i] = o;
    }
}

// MARN1_sps._reverse_seq (model/lsthm_sps.py:396-410; same code in every member of the family): each dialogue flipped over
// its own length, zero padded:  out[t][b][:] = t < len[b] ? X[len[b] - 1 - t][b][:] : 0.   One pass (the reference loops
// over the batch in Python; a gather + mask product + their autograd scatter is what a tensor-op version costs).  The map is
// its own adjoint, so the backward is the same kernel on the gradient.  Rows of w floats (w % 2 == 0: float2 lanes so that
// qmask, w = 2, qualifies).
__global__ void __launch_bounds__(256) reverse_seq_kernel(const float2 *__restrict__ X, const int *__restrict__ len, float2 *__restrict__ out,
                                                          int L, int B, int w2) {
    const long long total = (long long)L * B * w2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / w2;
        const int c = (int)(i - row * w2), t = (int)(row / B), b = (int)(row - (long long)t * B);
        const int n = __ldg(len + b);
        out[i] = t < n ? __ldg(X + ((long long)(n - 1 - t) * B + b) * w2 + c) : make_float2(0.f, 0.f);
    }
}
}  // namespace lsthm
