// Sequence-level cross-modal attention core of the speaker-state family (model/lsthm_sps.py:88-101, 116-129 —
// CrossAttention2 / CrossAttention3.forward after their three projections; model/lsthm_nsps.py:90-106):
//     out = dropout(softmax(Q K^T / sqrt(d_k))) V        per dialogue, UNMASKED over its L <= 128 utterances,
// single head, width D <= 128 (128 in lsthm_sps / lsthm_onlysp, 100 in lsthm_nsps), forward and backward, on the
// tcgen05 tensor cores with the fp32-accurate three-term bf16 split.  The L x L scores, probabilities and dropout
// masks never leave the SM (the reference materialises all three in HBM: [B,L,L] x 4 attentions, both passes).
//
// Same machinery as attn_kernels.cuh (operand tiles stored once, K-major, in the canonical no-swizzle UMMA layout,
// re-read through the MN-major view for the transposed products; a thread owns a score row out of TMEM), but a
// 128-wide operand tile is 64 KB (hi + lo), so the operands cannot all be resident: the forward keeps Q, K, V
// (192 KB, P aliases Q), the backward walks three 64 KB slots through the chain
//     S = Q K^T | dPd = dO V^T | Pd -> dV = Pd^T dO | dS -> dQ = dS K, dK = dS^T Q
// re-staging an operand when its slot has been recycled.  One CTA per dialogue.
#pragma once
#include "attn_kernels.cuh"

namespace lsthm {

constexpr int kXD = 128;                    // padded operand width
constexpr int kXSlot = 2 * kSqTile;         // hi + lo image of a [128 x 128] tile: 64 KB

struct XAttnArgs {
    const float *q, *k, *v, *o, *dout;      // row i of dialogue b: base + (b*sb + i*si)*ld
    float *out, *dq, *dk, *dv;
    float *lse;                             // [B][L] row log-sum-exp of the scaled scores, log2 units
    int B, L, D;                            // D = operand width (multiple of 4, <= 128)
    int ldq, ldk, ldv, ldo, lddq, lddk, lddv;
    long long sb, si;
    float scale, p_drop;
    unsigned long long seed;
};

// Stage columns [32*g4, 32*g4 + 32) of one row (NCH = 4 chunks of 8 floats) — or [64*g, 64*g + 64) with NCH = 8 —
// into a K-major [128 x 128] tile (hi at `hi`, lo at hi + kSqTile).  Columns >= D and rows with src == nullptr are zeros.
template <int NCH>
struct XRow {
    float4 v[2 * NCH];
};
template <int NCH>
__device__ __forceinline__ void xrow_load(const float *src, int D, int c0, XRow<NCH> &r) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int col = 8 * (c0 + i);
        if (src != nullptr && col + 8 <= D) {
            r.v[2 * i] = __ldg(reinterpret_cast<const float4 *>(src + col));
            r.v[2 * i + 1] = __ldg(reinterpret_cast<const float4 *>(src + col) + 1);
        } else if (src != nullptr && col + 4 <= D) {           // D % 8 == 4: the last chunk is half full
            r.v[2 * i] = __ldg(reinterpret_cast<const float4 *>(src + col));
            r.v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            r.v[2 * i] = r.v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
template <int NCH>
__device__ __forceinline__ void xrow_store(const XRow<NCH> &r, int row, float scale, uint8_t *hi, int c0) {
    const int roff = (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const float4 a = r.v[2 * i], b = r.v[2 * i + 1];
        const float x[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
        split_store8(x, hi + (c0 + i) * 2048 + roff, hi + kSqTile + (c0 + i) * 2048 + roff);
    }
}

// ---------------------------------------------------------------------------------------------
// forward: 256 threads, thread = (query row r = tid % 128, column half g = tid / 128); 192 KB -> one CTA per SM
// TMEM: S [0,128) | O [128,256)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) xattn_fwd_kernel(const __grid_constant__ XAttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    __shared__ float ex[512];
    uint8_t *sQ = smem, *sK = smem + kXSlot, *sV = smem + 2 * kXSlot, *sP = smem;   // P aliases Q once S is done
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, g = tid >> 7;
    const int b = blockIdx.x, L = a.L, D = a.D;
    const size_t grow = (size_t)((long long)b * a.sb + (long long)r * a.si);
    const bool rv = r < L;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        XRow<8> rq, rk;
        xrow_load<8>(rv ? a.q + grow * a.ldq : nullptr, D, 8 * g, rq);
        xrow_load<8>(rv ? a.k + grow * a.ldk : nullptr, D, 8 * g, rk);
        xrow_store<8>(rq, r, a.scale * kLog2e, sQ, 8 * g);
        xrow_store<8>(rk, r, 1.f, sK, 8 * g);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {   // S = Qs K^T : M = query, N = key, K = d (128)
        umma3<false>(tmem, smem_u32(sQ), kSqTile, 4096u, 2048u, 128u, smem_u32(sK), kSqTile, 4096u, 2048u, 128u, att_idesc(128, 0, 0), 8);
        umma_commit(&bar[0]);
    }
    {   // V is staged while the score product runs
        XRow<8> rw;
        xrow_load<8>(rv ? a.v + grow * a.ldv : nullptr, D, 8 * g, rw);
        xrow_store<8>(rw, r, 1.f, sV, 8 * g);
    }
    mbar_wait(&bar[0], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int cb = 64 * g;
    float mx = -INFINITY, sum = 0.f;
    {
        float v[16];
        for (int c0 = cb; c0 < cb + 64; c0 += 16) {
            tmem_ld16(trow + c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c0 + j < L) mx = fmaxf(mx, v[j]);
        }
        ex[tid] = mx;
        __syncthreads();
        mx = fmaxf(mx, ex[tid ^ 128]);
        const AttDrop drop(a.seed, blockIdx.x, r, a.p_drop);
        const int roff = (r >> 3) * 128 + (r & 7) * 16;
        float p8[8];
        for (int c0 = cb; c0 < cb + 64; c0 += 16) {
            tmem_ld16(trow + c0, v);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const int col = c0 + half * 8 + j;
                    float e0 = (rv && col < L) ? fast_exp2(v[half * 8 + j] - mx) : 0.f;
                    float e1 = (rv && col + 1 < L) ? fast_exp2(v[half * 8 + j + 1] - mx) : 0.f;
                    sum += e0 + e1;
                    if (drop.on) {
                        float s0, s1;
                        drop.pair(col >> 1, s0, s1);
                        e0 *= s0; e1 *= s1;
                    }
                    p8[j] = e0; p8[j + 1] = e1;
                }
                const int chunk = (c0 >> 3) + half;
                split_store8(p8, sP + chunk * 2048 + roff, sP + kSqTile + chunk * 2048 + roff);
            }
        }
        ex[256 + tid] = sum;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {   // O = P V : M = query, N = d (MN-major view of V), K = key (128)
        umma3<false>(tmem + 128, smem_u32(sP), kSqTile, 4096u, 2048u, 128u, smem_u32(sV), kSqTile, 256u, 128u, 2048u, att_idesc(128, 0, 1), 8);
        umma_commit(&bar[1]);
    }
    sum += ex[256 + (tid ^ 128)];
    const float inv = rv ? 1.0f / sum : 0.f;
    if (g == 0 && rv && a.lse != nullptr) a.lse[(size_t)b * L + r] = mx + log2f(sum);
    mbar_wait(&bar[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        float v[16];
        float *orow = a.out + grow * a.ldo;
        for (int c0 = cb; c0 < cb + 64; c0 += 16) {
            if (c0 >= D) break;
            tmem_ld16(trow + 128 + c0, v);
            if (rv) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    if (c0 + 4 * q4 < D)
                        reinterpret_cast<float4 *>(orow + c0)[q4] =
                            make_float4(v[4 * q4] * inv, v[4 * q4 + 1] * inv, v[4 * q4 + 2] * inv, v[4 * q4 + 3] * inv);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
    }
}

// ---------------------------------------------------------------------------------------------
// backward: 512 threads, thread = (row r = tid % 128, column quarter g = tid / 128); three 64 KB slots X, Y, Z
// TMEM: S [0,128) -> later dK | dPd [128,256) | dV [256,384) | dQ [384,512)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) xattn_bwd_kernel(const __grid_constant__ XAttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base;
    __shared__ float ex[512];
    uint8_t *sX = smem, *sY = smem + kXSlot, *sZ = smem + 2 * kXSlot;
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, g = tid >> 7;
    const int b = blockIdx.x, L = a.L, D = a.D;
    const size_t grow = (size_t)((long long)b * a.sb + (long long)r * a.si);
    const bool rv = r < L;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const float qscale = a.scale * kLog2e;
    const uint32_t uX = smem_u32(sX), uY = smem_u32(sY), uZ = smem_u32(sZ);
    const int roff = (r >> 3) * 128 + (r & 7) * 16, cb = 32 * g;
    // ---- X <- Qs, Y <- K
    {
        XRow<4> rq, rk;
        xrow_load<4>(rv ? a.q + grow * a.ldq : nullptr, D, 4 * g, rq);
        xrow_load<4>(rv ? a.k + grow * a.ldk : nullptr, D, 4 * g, rk);
        xrow_store<4>(rq, r, qscale, sX, 4 * g);
        xrow_store<4>(rk, r, 1.f, sY, 4 * g);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    if (tid == 0) {   // S = Qs K^T
        umma3<false>(tmem, uX, kSqTile, 4096u, 2048u, 128u, uY, kSqTile, 4096u, 2048u, 128u, att_idesc(128, 0, 0), 8);
        umma_commit(&bar[0]);
    }
    // ---- Z <- dO (while S runs); delta = rowsum(dO * O)
    float delta = 0.f;
    {
        XRow<4> rd, ro;
        xrow_load<4>(rv ? a.dout + grow * a.ldo : nullptr, D, 4 * g, rd);
        xrow_load<4>(rv ? a.o + grow * a.ldo : nullptr, D, 4 * g, ro);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            delta += rd.v[i].x * ro.v[i].x + rd.v[i].y * ro.v[i].y + rd.v[i].z * ro.v[i].z + rd.v[i].w * ro.v[i].w;
        xrow_store<4>(rd, r, 1.f, sZ, 4 * g);
        ex[tid] = delta;
    }
    const float lse = rv ? __ldg(a.lse + (size_t)b * L + r) : 0.f;
    mbar_wait(&bar[0], 0);                                   // S done: X is free
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- X <- V
    {
        XRow<4> rw;
        xrow_load<4>(rv ? a.v + grow * a.ldv : nullptr, D, 4 * g, rw);
        xrow_store<4>(rw, r, 1.f, sX, 4 * g);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {   // dPd = dO V^T
        umma3<false>(tmem + 128, uZ, kSqTile, 4096u, 2048u, 128u, uX, kSqTile, 4096u, 2048u, 128u, att_idesc(128, 0, 0), 8);
        umma_commit(&bar[1]);
    }
    delta = (ex[r] + ex[r + 128]) + (ex[r + 256] + ex[r + 384]);
    mbar_wait(&bar[1], 0);                                   // dPd done: X (V) is free
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const AttDrop drop(a.seed, blockIdx.x, r, a.p_drop);
    // ---- pass 1: Pd -> X
    {
        float s[16], p8[8];
#pragma unroll
        for (int c0 = cb; c0 < cb + 32; c0 += 16) {
            tmem_ld16(trow + c0, s);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const int col = c0 + half * 8 + j, i0 = half * 8 + j;
                    const float p0 = (rv && col < L) ? fast_exp2(s[i0] - lse) : 0.f;
                    const float p1 = (rv && col + 1 < L) ? fast_exp2(s[i0 + 1] - lse) : 0.f;
                    float s0 = 1.f, s1 = 1.f;
                    if (drop.on) drop.pair(col >> 1, s0, s1);
                    p8[j] = p0 * s0; p8[j + 1] = p1 * s1;
                }
                const int chunk = (c0 >> 3) + half;
                split_store8(p8, sX + chunk * 2048 + roff, sX + kSqTile + chunk * 2048 + roff);
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {   // dV = Pd^T dO : M = key (MN view of Pd), N = d (MN view of dO), K = query
        umma3<false>(tmem + 256, uX, kSqTile, 256u, 128u, 2048u, uZ, kSqTile, 256u, 128u, 2048u, att_idesc(128, 1, 1), 8);
        umma_commit(&bar[2]);
    }
    // the Q rows for the last product are fetched while dV runs
    XRow<4> rq2;
    xrow_load<4>(rv ? a.q + grow * a.ldq : nullptr, D, 4 * g, rq2);
    mbar_wait(&bar[2], 0);                                   // dV done: X (Pd) and Z (dO) are free
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    xrow_store<4>(rq2, r, qscale, sZ, 4 * g);                // Z <- Qs
    // ---- pass 2: dS -> X
    {
        float s[16], gg[16], d8[8];
#pragma unroll
        for (int c0 = cb; c0 < cb + 32; c0 += 16) {
            tmem_ld16x2(trow + c0, trow + 128 + c0, s, gg);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const int col = c0 + half * 8 + j, i0 = half * 8 + j;
                    const float p0 = (rv && col < L) ? fast_exp2(s[i0] - lse) : 0.f;
                    const float p1 = (rv && col + 1 < L) ? fast_exp2(s[i0 + 1] - lse) : 0.f;
                    float s0 = 1.f, s1 = 1.f;
                    if (drop.on) drop.pair(col >> 1, s0, s1);
                    d8[j] = p0 * (s0 * gg[i0] - delta); d8[j + 1] = p1 * (s1 * gg[i0 + 1] - delta);
                }
                const int chunk = (c0 >> 3) + half;
                split_store8(d8, sX + chunk * 2048 + roff, sX + kSqTile + chunk * 2048 + roff);
            }
        }
    }
    // drain dV while the other threads finish their pass (it does not touch the slots)
    {
        float v[16];
        float *dst = a.dv + grow * a.lddv;
        for (int c0 = cb; c0 < cb + 32; c0 += 16) {
            if (c0 >= D) break;
            tmem_ld16(trow + 256 + c0, v);
            if (rv) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    if (c0 + 4 * q4 < D)
                        reinterpret_cast<float4 *>(dst + c0)[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                         // dS and Qs tiles complete; every thread has read S and dPd
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        // dQs = dS K : M = query, N = d (MN view of K), K = key ;  dK = dS^T Qs : M = key (MN view of dS), N = d (MN view of Qs), K = query
        umma3<false>(tmem + 384, uX, kSqTile, 4096u, 2048u, 128u, uY, kSqTile, 256u, 128u, 2048u, att_idesc(128, 0, 1), 8);
        umma3<false>(tmem, uX, kSqTile, 256u, 128u, 2048u, uZ, kSqTile, 256u, 128u, 2048u, att_idesc(128, 1, 1), 8);
        umma_commit(&bar[3]);
    }
    mbar_wait(&bar[3], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        float v[16];
        for (int which = 0; which < 2; ++which) {
            // d/dq = scale * (dS K);  the staged q carries an extra log2(e): d/dk = (dS^T Qs) / log2(e)
            float *dst = which == 0 ? a.dq + grow * a.lddq : a.dk + grow * a.lddk;
            const float mul = which == 0 ? a.scale : kLn2;
            const uint32_t tcol = which == 0 ? 384u : 0u;
            for (int c0 = cb; c0 < cb + 32; c0 += 16) {
                if (c0 >= D) break;
                tmem_ld16(trow + tcol + c0, v);
                if (rv) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        if (c0 + 4 * q4 < D)
                            reinterpret_cast<float4 *>(dst + c0)[q4] =
                                make_float4(v[4 * q4] * mul, v[4 * q4 + 1] * mul, v[4 * q4 + 2] * mul, v[4 * q4 + 3] * mul);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

}  // namespace lsthm
