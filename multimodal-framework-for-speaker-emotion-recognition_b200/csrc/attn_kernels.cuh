// Fused multi-head self-attention of the utterance encoder (model/encoder.py:27-86: per head
// softmax(q k^T / sqrt(d_k)) -> dropout -> . v, UNMASKED over the L <= 128 utterances of a dialogue,
// d_k = d_v = 40) on the tcgen05 tensor cores, forward and backward.  One CTA per (dialogue, head):
// the whole L x L score tile lives in tensor memory, a thread owns one query row for the softmax
// (tcgen05.ld gives a thread its row), the probabilities go back to shared memory as the next UMMA's
// operand; nothing of size L x L ever touches HBM (the reference materialises scores, softmax and
// dropout mask there: [B,8,L,L] x 3 encoders, forward and backward).
//
// fp32 accuracy: as in gemm3, every operand is split into bf16 hi + lo and each product is three UMMAs.
//
// Operand tiles are stored ONCE, row-major-K ("K-major") in the canonical no-swizzle layout
//   offset(row, col) = (col/8)*2048 + (row/8)*128 + (row%8)*16 + (col%8)*2      [128 rows]
// and the same bytes double as the MN-major operand of the transposed product (a core matrix is its
// own transpose under the major flag): view(m = col, k = row) has LBO = 128, SBO = 2048.
#pragma once
#include "gemm3_kernels.cuh"

namespace lsthm {

constexpr int kAttD = 40, kAttDP = 48, kAttLP = 128;
constexpr int kRowTile = 6 * 2048;          // [128 x 48] bf16
constexpr int kSqTile = 16 * 2048;          // [128 x 128] bf16
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnArgs {
    const float *q, *k, *v, *o, *dout;      // row i of dialogue b, head h: base + (b*sb + i*si)*ld + h*40
    float *out, *dq, *dk, *dv;
    float *lse;                             // [B*H][L] row log-sum-exp of the scaled scores, in log2 units (fwd writes, bwd reads)
    int B, L, H;
    int ldq, ldk, ldv, ldo;                 // row strides (floats) of q/k/v (and dq/dk/dv) and of o/out/dout
    long long sb, si;                       // row index of (dialogue b, position i) = b*sb + i*si  (batch-major: L,1; time-major: 1,B)
    float scale, p_drop;
    unsigned long long seed;
};

__device__ __forceinline__ uint32_t att_idesc(int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[128 x N] (+)= A . B over `ksteps` k16 steps, three split terms.  A: hi at a, lo at a + a_lo; same for B.
template <bool BF16 = false>
__device__ __forceinline__ void umma3(uint32_t tmem_d, uint32_t a, uint32_t a_lo, uint32_t a_step, uint32_t a_lbo, uint32_t a_sbo,
                                      uint32_t b, uint32_t b_lo, uint32_t b_step, uint32_t b_lbo, uint32_t b_sbo,
                                      uint32_t idesc, int ksteps) {
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t ah = umma_desc(a + ks * a_step, a_lbo, a_sbo), al = umma_desc(a + a_lo + ks * a_step, a_lbo, a_sbo);
        const uint64_t bh = umma_desc(b + ks * b_step, b_lbo, b_sbo), bl = umma_desc(b + b_lo + ks * b_step, b_lbo, b_sbo);
        umma_f16(tmem_d, ah, bh, idesc, ks > 0 ? 1u : 0u);
        if (!BF16) {                                   // bf16 mode: operands rounded to bf16, one UMMA per k-step
            umma_f16(tmem_d, ah, bl, idesc, 1u);
            umma_f16(tmem_d, al, bh, idesc, 1u);
        }
    }
}
// operand triples used below:  K-major row tile: (k16 step, LBO, SBO) = (4096, 2048, 128);
//                              MN-major view of the same bytes:         = ( 256,  128, 2048)

#define LSTHM_TMEM_LD16_REGS(r, o) \
    "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), \
    "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 "tcgen05.wait::ld.sync.aligned;"
                 : LSTHM_TMEM_LD16_REGS(r, 0)
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// two 16-column loads (different TMEM addresses) behind one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float (&va)[16], float (&vb)[16]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n"
                 "tcgen05.wait::ld.sync.aligned;"
                 : LSTHM_TMEM_LD16_REGS(r, 0), LSTHM_TMEM_LD16_REGS(r, 16)
                 : "r"(ta), "r"(tb) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { va[i] = __uint_as_float(r[i]); vb[i] = __uint_as_float(r[16 + i]); }
}

template <bool BF16>
__device__ __forceinline__ void put8(const float (&x)[8], uint8_t *hi, uint8_t *lo) {
    if (BF16) store8_hi(x, hi);
    else split_store8(x, hi, lo);
}

// Staging of one 40-float row into a K-major [128 x 48] row tile (hi, lo), split in two phases so that a thread's
// global loads for ALL operand tiles are in flight before the first conversion: a thread owns three of the six
// 8-float chunks of its row (chunk 5 is the zero padding 40 -> 48); src == nullptr (row >= L) stages zeros.
struct RowRegs {
    float4 v[6];
};
__device__ __forceinline__ void row_load(const float *src, int c0, RowRegs &r) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int c = c0 + i;
        if (c < 5 && src != nullptr) {
            r.v[2 * i] = __ldg(reinterpret_cast<const float4 *>(src) + 2 * c);
            r.v[2 * i + 1] = __ldg(reinterpret_cast<const float4 *>(src) + 2 * c + 1);
        } else {
            r.v[2 * i] = r.v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
template <bool BF16>
__device__ __forceinline__ void row_store(const RowRegs &r, int row, float scale, uint8_t *hi, int c0) {
    uint8_t *lo = hi + kRowTile;
    const int roff = (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float4 a = r.v[2 * i], b = r.v[2 * i + 1];
        const float x[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
        put8<BF16>(x, hi + (c0 + i) * 2048 + roff, lo + (c0 + i) * 2048 + roff);
    }
}

// In-kernel attention dropout: one 32-bit hash per PAIR of adjacent key columns, 16 random bits per element
// (keep iff bits >= thr, thr = round(p * 65536); the scale is 65536 / (65536 - thr) so the mask is exactly unbiased).
// Forward and backward evaluate the same function, so no mask is stored.
struct AttDrop {
    uint32_t key, thr;
    float scale;
    bool on;
    __device__ __forceinline__ AttDrop(unsigned long long seed, int bh, int row, float p) {
        on = p > 0.f;
        thr = (uint32_t)(p * 65536.0f + 0.5f);
        scale = 65536.0f / (65536.0f - (float)thr);
        uint32_t x = (uint32_t)seed ^ ((uint32_t)bh * 0x9E3779B9u);
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        key = (x ^ (uint32_t)(seed >> 32)) + (uint32_t)row * 0x85EBCA6Bu;
    }
    // scales for columns (2*pair, 2*pair + 1)
    __device__ __forceinline__ void pair(int pair_idx, float &s0, float &s1) const {
        uint32_t x = key + (uint32_t)pair_idx * 0xC2B2AE35u;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        s0 = (x & 0xffffu) >= thr ? scale : 0.f;
        s1 = (x >> 16) >= thr ? scale : 0.f;
    }
};

// ---------------------------------------------------------------------------------------------
// forward:  out = dropout(softmax(q k^T * scale)) v      256 threads, SMEM 89 KB -> 2 CTAs / SM
// thread = (query row r = tid%128, column half g = tid/128).  Scores are produced in log2 units (log2(e) is folded
// into the staged q) so the softmax is one FADD + EX2 per element; the normalisation is applied to the 40 output
// columns instead of the 128 probabilities.
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256, 2) attn_fwd_kernel(const __grid_constant__ AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    __shared__ float ex[512];
    uint8_t *sQ = smem, *sK = smem + 2 * kRowTile, *sV = smem + 2 * kSqTile;   // P (64 KB) aliases Q,K after S is done
    uint8_t *sP = smem;
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, g = tid >> 7;
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H, L = a.L;
    const size_t grow = (size_t)((long long)b * a.sb + (long long)r * a.si);
    const bool rv = r < L;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        RowRegs rq, rk, rw;
        row_load(rv ? a.q + grow * a.ldq + h * kAttD : nullptr, 3 * g, rq);
        row_load(rv ? a.k + grow * a.ldk + h * kAttD : nullptr, 3 * g, rk);
        row_load(rv ? a.v + grow * a.ldv + h * kAttD : nullptr, 3 * g, rw);
        row_store<BF16>(rq, r, a.scale * kLog2e, sQ, 3 * g);
        row_store<BF16>(rk, r, 1.f, sK, 3 * g);
        row_store<BF16>(rw, r, 1.f, sV, 3 * g);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {   // S = Qs K^T : M = query, N = key, K = d (48)
        umma3<BF16>(tmem, smem_u32(sQ), kRowTile, 4096u, 2048u, 128u, smem_u32(sK), kRowTile, 4096u, 2048u, 128u, att_idesc(128, 0, 0), 3);
        umma_commit(&bar[0]);
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
                put8<BF16>(p8, sP + chunk * 2048 + roff, sP + kSqTile + chunk * 2048 + roff);
            }
        }
        ex[256 + tid] = sum;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {   // O = P V : M = query, N = d (48, MN-major view of V), K = key (128); overwrites S columns [0,48)
        umma3<BF16>(tmem, smem_u32(sP), kSqTile, 4096u, 2048u, 128u, smem_u32(sV), kRowTile, 256u, 128u, 2048u, att_idesc(kAttDP, 0, 1), 8);
        umma_commit(&bar[1]);
    }
    sum += ex[256 + (tid ^ 128)];
    const float inv = rv ? 1.0f / sum : 0.f;
    if (g == 0 && rv && a.lse != nullptr) a.lse[(size_t)blockIdx.x * L + r] = mx + log2f(sum);
    mbar_wait(&bar[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        float v[16];
        float *orow = a.out + grow * a.ldo + h * kAttD;
        for (int c0 = 16 * g; c0 < kAttDP; c0 += 32) {
            tmem_ld16(trow + c0, v);
            if (rv) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    if (c0 + 4 * q4 < kAttD)
                        reinterpret_cast<float4 *>(orow + c0)[q4] =
                            make_float4(v[4 * q4] * inv, v[4 * q4 + 1] * inv, v[4 * q4 + 2] * inv, v[4 * q4 + 3] * inv);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  TMEM columns: S [0,128) | dPd [128,256) | dV [256,304) | dQ [304,352) | dK [352,400)
// One pass over the score tile: with the forward's row log-sum-exp the probabilities are p = 2^(s - lse), so the
// dropped probabilities Pd (operand of dV = Pd^T dO) and dS = p (sc dPd - delta) (operand of dQ, dK) are produced
// together into two shared tiles and the three remaining products are issued back to back.
// ---------------------------------------------------------------------------------------------
// 512 threads: four threads share a query row (tid, tid+128, tid+256, tid+384 address the same TMEM lanes), each
// owns 32 of the 128 key columns -> 4 warps per scheduler hide the TMEM / MUFU / shared-memory latencies of the pass.
constexpr int kAttBwdThreads = 512;

// the one or two 8-float chunks of a 40 (+8 zero) float row a thread of group g stages: chunks g and g + 4 (< 6)
struct RowRegs2 {
    float4 v[4];
};
__device__ __forceinline__ void row_load2(const float *src, int g, RowRegs2 &r) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = g + 4 * i;
        if (c < 5 && src != nullptr) {
            r.v[2 * i] = reinterpret_cast<const float4 *>(src)[2 * c];
            r.v[2 * i + 1] = reinterpret_cast<const float4 *>(src)[2 * c + 1];
        } else {
            r.v[2 * i] = r.v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
template <bool BF16>
__device__ __forceinline__ void row_store2(const RowRegs2 &r, int row, float scale, uint8_t *hi, int g) {
    uint8_t *lo = hi + kRowTile;
    const int roff = (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = g + 4 * i;
        if (c >= 6) continue;
        const float4 a = r.v[2 * i], b = r.v[2 * i + 1];
        const float x[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
        put8<BF16>(x, hi + c * 2048 + roff, lo + c * 2048 + roff);
    }
}

// Persistent: one CTA per SM walks the (dialogue, head) items blockIdx.x, blockIdx.x + gridDim.x, ...  With 224 KB of
// shared memory only one CTA fits an SM, so nothing else can hide an item's memory phases; instead the rows of item
// i+1 (q, k, v, dO, O: 5 x L slices of 160 bytes) are fetched by bulk async copies (TMA engine, mbarrier
// complete_tx) into a raw fp32 patch that aliases the Pd / dS tiles, issued as soon as item i's last products have
// retired, so they land while item i's gradients are drained from TMEM and stored.
template <bool BF16>
__global__ void __launch_bounds__(kAttBwdThreads, 1) attn_bwd_kernel(const __grid_constant__ AttnArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];   // no-swizzle operands need 16-byte alignment only
    __shared__ __align__(8) uint64_t bar[3];           // 0: S,dPd done   1: dV,dQ,dK done   2: raw rows of the next item landed
    __shared__ uint32_t tmem_base;
    uint8_t *sQ = smem, *sK = sQ + 2 * kRowTile, *sV = sK + 2 * kRowTile, *sdO = sV + 2 * kRowTile, *sP = sdO + 2 * kRowTile;
    uint8_t *sD = sP + 2 * kSqTile;
    constexpr int kRawLd = 44;                                  // floats per raw row: 176 B stride -> conflict-free 16-byte reads by row
    float *raw = reinterpret_cast<float *>(sP);                 // [5][128][kRawLd] = 112.6 KB <= 128 KB (sP + sD)
    constexpr int kPatchRows = (4 * kSqTile - 5 * 128 * kRawLd * 4) / (kAttD * 4);   // rows of the epilogue patch behind it: 115
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, g = tid >> 7;
    const int L = a.L, nitems = a.B * a.H;
    const bool rv = r < L;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV), udO = smem_u32(sdO), uP = smem_u32(sP), uD = smem_u32(sD);
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int roff = (r >> 3) * 128 + (r & 7) * 16, cb = 32 * g;

    // bulk copies of one item's rows: 5 arrays x L rows x 160 B, spread over the CTA's threads; thread 0 posts the byte count
    auto fetch = [&](int item) {
        const int b = item / a.H, h = item % a.H;
        if (tid == 0) mbar_expect_tx(&bar[2], (uint32_t)(5 * L * kAttD * sizeof(float)));
        for (int idx = tid; idx < 5 * L; idx += kAttBwdThreads) {
            const int arr = idx / L, row = idx - arr * L;
            const size_t gr = (size_t)((long long)b * a.sb + (long long)row * a.si);
            const float *src = (arr == 0 ? a.q + gr * a.ldq : arr == 1 ? a.k + gr * a.ldk : arr == 2 ? a.v + gr * a.ldv
                                : arr == 3 ? a.dout + gr * a.ldo : a.o + gr * a.ldo) + h * kAttD;
            bulk_g2s(raw + (arr * 128 + row) * kRawLd, src, kAttD * sizeof(float), &bar[2]);
        }
    };
    int item = blockIdx.x;
    float lse_next = 0.f;
    if (item < nitems) {
        fetch(item);
        if (rv) lse_next = __ldg(a.lse + (size_t)item * L + r);
    }
    for (int it = 0; item < nitems; item += gridDim.x, ++it) {
        const int b = item / a.H, h = item % a.H;
        const uint32_t ph = it & 1;
        const size_t grow = (size_t)((long long)b * a.sb + (long long)r * a.si);
        const float lse = lse_next;
        float delta = 0.f;
        mbar_wait(&bar[2], ph);                                   // this item's raw rows have landed
        {
            RowRegs2 rq, rk, rw, rd;
            auto raw_row = [&](int arr) { return rv ? raw + (arr * 128 + r) * kRawLd : nullptr; };
            row_load2(raw_row(0), g, rq);
            row_load2(raw_row(1), g, rk);
            row_load2(raw_row(2), g, rw);
            row_load2(raw_row(3), g, rd);
            if (rv) {   // delta = rowsum(dO * O) = sum_j Pd_ij dPd_ij (each of the row's four threads computes it from the raw patch)
                const float4 *d4 = reinterpret_cast<const float4 *>(raw + (3 * 128 + r) * kRawLd);
                const float4 *o4 = reinterpret_cast<const float4 *>(raw + (4 * 128 + r) * kRawLd);
#pragma unroll
                for (int d = 0; d < 10; ++d) {
                    const float4 x = d4[d], y = o4[d];
                    delta += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
                }
            }
            row_store2<BF16>(rq, r, a.scale * kLog2e, sQ, g);
            row_store2<BF16>(rk, r, 1.f, sK, g);
            row_store2<BF16>(rw, r, 1.f, sV, g);
            row_store2<BF16>(rd, r, 1.f, sdO, g);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                          // tiles complete; raw patch fully consumed; previous item's TMEM reads done
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            // S = Qs K^T ; dPd = dO V^T   (both M = query, N = key, K = d)
            umma3<BF16>(tmem, uQ, kRowTile, 4096u, 2048u, 128u, uK, kRowTile, 4096u, 2048u, 128u, att_idesc(128, 0, 0), 3);
            umma3<BF16>(tmem + 128, udO, kRowTile, 4096u, 2048u, 128u, uV, kRowTile, 4096u, 2048u, 128u, att_idesc(128, 0, 0), 3);
            umma_commit(&bar[0]);
        }
        mbar_wait(&bar[0], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const AttDrop drop(a.seed, item, r, a.p_drop);
            float s[16], gg[16], p8[8], d8[8];
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
                        p8[j] = p0 * s0; p8[j + 1] = p1 * s1;
                        d8[j] = p0 * (s0 * gg[i0] - delta); d8[j + 1] = p1 * (s1 * gg[i0 + 1] - delta);
                    }
                    const int chunk = (c0 >> 3) + half;
                    put8<BF16>(p8, sP + chunk * 2048 + roff, sP + kSqTile + chunk * 2048 + roff);
                    put8<BF16>(d8, sD + chunk * 2048 + roff, sD + kSqTile + chunk * 2048 + roff);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            // dV = Pd^T dO : M = key (MN-major view of Pd), N = d (MN-major view of dO), K = query
            umma3<BF16>(tmem + 256, uP, kSqTile, 256u, 128u, 2048u, udO, kRowTile, 256u, 128u, 2048u, att_idesc(kAttDP, 1, 1), 8);
            // dQs = dS K : M = query, N = d (MN view of K), K = key ;  dK = dS^T Qs : M = key (MN view of dS), N = d (MN view of Qs), K = query
            umma3<BF16>(tmem + 304, uD, kSqTile, 4096u, 2048u, 128u, uK, kRowTile, 256u, 128u, 2048u, att_idesc(kAttDP, 0, 1), 8);
            umma3<BF16>(tmem + 352, uD, kSqTile, 256u, 128u, 2048u, uQ, kRowTile, 256u, 128u, 2048u, att_idesc(kAttDP, 1, 1), 8);
            umma_commit(&bar[1]);
        }
        mbar_wait(&bar[1], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // every product of this item has retired: the Pd / dS tiles (= the raw patch) are free -> start the next item's rows
        const int next = item + gridDim.x;
        if (next < nitems) {
            fetch(next);
            if (rv) lse_next = __ldg(a.lse + (size_t)next * L + r);
        }
        if (L <= kPatchRows) {
            // Coalesced epilogue: a thread owns an accumulator ROW, so storing from registers would put 16 bytes into
            // each of 32 rows per instruction.  Each gradient matrix goes through a [L][40] fp32 patch in the 18 KB of
            // the Pd/dS region that the next item's raw rows do not use, and leaves as whole 160-byte row slices.
            float *patch = reinterpret_cast<float *>(sP) + 5 * 128 * kRawLd;
            float v[16];
            for (int which = 0; which < 3; ++which) {
                // d/dq = scale * (dS K);  the staged q carries an extra log2(e): d/dk = (dS^T Qs) / log2(e)
                const float mul = which == 1 ? a.scale : which == 2 ? kLn2 : 1.f;
                if (g < 3) {                                  // warp-uniform: column block 16*g of the 40 (+8) columns
                    tmem_ld16(trow + 256 + 48 * which + 16 * g, v);
                    if (rv) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            if (16 * g + 4 * q4 < kAttD)
                                *reinterpret_cast<float4 *>(patch + r * kAttD + 16 * g + 4 * q4) =
                                    make_float4(v[4 * q4] * mul, v[4 * q4 + 1] * mul, v[4 * q4 + 2] * mul, v[4 * q4 + 3] * mul);
                    }
                }
                __syncthreads();
                float *base = which == 0 ? a.dv : which == 1 ? a.dq : a.dk;
                const int ld = which == 0 ? a.ldv : which == 1 ? a.ldq : a.ldk;
                for (int idx = tid; idx < L * (kAttD / 4); idx += kAttBwdThreads) {
                    const int row = idx / (kAttD / 4), f4 = idx - row * (kAttD / 4);
                    const size_t gr = (size_t)((long long)b * a.sb + (long long)row * a.si);
                    reinterpret_cast<float4 *>(base + gr * ld + h * kAttD)[f4] = reinterpret_cast<const float4 *>(patch + row * kAttD)[f4];
                }
                __syncthreads();
            }
        } else {
            float v[16];
            // 9 (matrix, 16-column block) items over the four threads of a row: g takes the items with index % 4 == g
            for (int blk = g; blk < 9; blk += 4) {
                const int which = blk / 3, c0 = 16 * (blk % 3);
                float *dst = (which == 0 ? a.dv : which == 1 ? a.dq : a.dk) + grow * (which == 0 ? a.ldv : which == 1 ? a.ldq : a.ldk) + h * kAttD;
                const float mul = which == 1 ? a.scale : which == 2 ? kLn2 : 1.f;
                tmem_ld16(trow + 256 + 48 * which + c0, v);
                if (rv) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        if (c0 + 4 * q4 < kAttD)
                            reinterpret_cast<float4 *>(dst + c0)[q4] =
                                make_float4(v[4 * q4] * mul, v[4 * q4 + 1] * mul, v[4 * q4 + 2] * mul, v[4 * q4 + 3] * mul);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

}  // namespace lsthm
