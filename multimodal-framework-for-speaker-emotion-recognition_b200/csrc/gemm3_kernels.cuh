// fp32-accurate dense GEMM on the 5th-generation tensor cores (tcgen05 + TMEM) for the TIME-PARALLEL
// products of the path: input projections W.x over all T*N utterances, encoder projections, heads and
// the hoisted weight-gradient products  dW = adj^T . act  (K = T*N = 112 640).
//
// fp32 parity rules out a single bf16/TF32 pass, so each fp32 operand is split while it
// is staged into shared memory:  x = hi + lo,  hi = bf16(x), lo = bf16(x - hi)  (16 mantissa bits), and
// each k-step issues three UMMAs   D += Ahi.Bhi + Ahi.Blo + Alo.Bhi   accumulating in fp32 in TMEM.
//
//   C[M,N] = op(A) . op(B)  (+ bias[N]),  A/B/C fp32 row-major
//     AMN = 0: A is [M][K] (K contiguous -> K-major operand)      AMN = 1: A is [K][M] (MN-major operand)
//     BMN = 0: B is [N][K] (K-major, nn.Linear weight layout)     BMN = 1: B is [K][N] (MN-major)
//   NT (0,0): y = x W^T         NN (0,1): dx = dy W         TN (1,1): dW = dy^T x  (split-K over gridDim.z)
//
// CTA = one 128x128 output tile (x one K split): 8 producer warps stage + split the operands
// (global fp32 -> registers -> bf16 hi/lo -> canonical no-swizzle UMMA layouts, padded strides so the
// 16-byte stores are bank-conflict free), one thread of warp 8 issues tcgen05.mma, tcgen05.commit
// releases the stage back to the producers (3-stage mbarrier ring), warps 0-3 drain TMEM with
// tcgen05.ld and store C.  99 KB shared memory -> 2 CTAs per SM, so one CTA's epilogue overlaps the
// other's main loop.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace lsthm {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 32, kGemmStages = 3;
constexpr int kGemmThreads = 288;                      // 8 producer warps + 1 MMA warp
constexpr int kTileBytes = 8448;                       // one bf16 operand tile (hi or lo), either major
constexpr int kStageBytes = 4 * kTileBytes;            // A_hi, A_lo, B_hi, B_lo
constexpr int kKLbo = 2080, kKSbo = 128;               // K-major: 16-byte chunk c of row r at c*2080 + (r/8)*128 + (r%8)*16
constexpr int kMnLbo = 128, kMnSbo = 528;              // MN-major: 8 rows mc at mc*528 + (k/8)*128 + (k%8)*16
constexpr size_t kGemmSmemBytes = (size_t)kGemmStages * kStageBytes + 1024;

struct GemmArgs {
    const float *A, *B, *bias;
    float *C;          // output, or split-K workspace [splits][M][N] when splits > 1
    int M, N, K, lda, ldb, ldc;
    int k_per_split;   // multiple of kGemmBK
    int splits;
    int relu;          // NT epilogue: C = max(0, A.B^T + bias)  (splits == 1 only)
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);     // version 1, no swizzle
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
// split 8 consecutive fp32 values into bf16 hi / lo and store each as one 16-byte chunk.
// Two values per packed conversion (cvt.rn.bf16x2.f32 -> one F2FP): hi = rn(x), lo = rn(x - hi).
__device__ __forceinline__ void split_store8(const float (&x)[8], uint8_t *hi, uint8_t *lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = pack_bf16(x[2 * i] - h0, x[2 * i + 1] - h1);
    }
    *reinterpret_cast<uint4 *>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// three-way split x = hi + mid + lo (24 mantissa bits: exact for fp32 up to the last rounding) for the six-term product
//   D += Ah.Bh + Ah.Bm + Am.Bh + Ah.Bl + Al.Bh + Am.Bm      (dropped terms are <= 2^-24 |a||b|)
// used where a three-term product is not enough: sums that cancel structurally, e.g. a LayerNorm output times the
// all-ones projections of the reference's sequence-level cross attention (model/lsthm_sps.py:82-84).
__device__ __forceinline__ void split_store8_3(const float (&x)[8], uint8_t *hi, uint8_t *mid, uint8_t *lo) {
    uint32_t h[4], m[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
        const float r0 = x[2 * i] - __uint_as_float(h[i] << 16), r1 = x[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
        m[i] = pack_bf16(r0, r1);
        l[i] = pack_bf16(r0 - __uint_as_float(m[i] << 16), r1 - __uint_as_float(m[i] & 0xffff0000u));
    }
    *reinterpret_cast<uint4 *>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(mid) = make_uint4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<uint4 *>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// bf16 mode (one UMMA per k-step, operands rounded to bf16): only the hi image is produced
__device__ __forceinline__ void store8_hi(const float (&x)[8], uint8_t *hi) {
    *reinterpret_cast<uint4 *>(hi) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
}

// Load 8 consecutive floats p[0..7] where only the first `valid` (<= 8) are in range; the rest are zero.
__device__ __forceinline__ void load8(const float *p, int valid, float (&x)[8]) {
    if (valid >= 8) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = i < valid ? __ldg(p + i) : 0.f;
    }
}

// One operand tile (128 rows of the M/N axis x 32 of K) is 512 items of 8 consecutive floats; each of the 256
// producer threads owns two.  Loading (global -> registers) and storing (split -> shared) are separate so
// that the loads of k-block kb+1 are in flight while block kb is converted and stored.
//   MN = 0: src[row][k], k contiguous.   item i: row = i*64 + ptid/4, k-chunk = ptid%4 (4 lanes read a 128 B line)
//   MN = 1: src[k][row], row contiguous. item i: k = i*16 + ptid/16, row-chunk = ptid%16 (16 lanes read 512 B)
struct TileRegs {
    float x[2][8];
};

template <int MN>
__device__ __forceinline__ void load_tile(const float *__restrict__ src, int ld, int row0, int nrows, int k0, int kend,
                                          int ptid, TileRegs &t) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int valid;
        const float *p;
        if (MN == 0) {
            const int r = i * 64 + (ptid >> 2), c = ptid & 3, gr = row0 + r, gk = k0 + 8 * c;
            valid = gr < nrows ? max(0, min(8, kend - gk)) : 0;
            p = src + (size_t)gr * ld + gk;
        } else {
            const int k = i * 16 + (ptid >> 4), mc = ptid & 15, gk = k0 + k, gr = row0 + 8 * mc;
            valid = gk < kend ? max(0, min(8, nrows - gr)) : 0;
            p = src + (size_t)gk * ld + gr;
        }
        if (valid > 0) load8(p, valid, t.x[i]);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) t.x[i][j] = 0.f;
        }
    }
}

// PREC: 0 = three-term split (fp32 parity mode), 1 = bf16 operands (one term), 2 = six-term split (third image at lo + (lo - hi))
template <int MN, int PREC = 0>
__device__ __forceinline__ void store_tile(const TileRegs &t, uint8_t *hi, uint8_t *lo, int ptid) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int off;
        if (MN == 0) {
            const int r = i * 64 + (ptid >> 2), c = ptid & 3;
            off = c * kKLbo + (r >> 3) * kKSbo + (r & 7) * 16;
        } else {
            const int k = i * 16 + (ptid >> 4), mc = ptid & 15;
            off = mc * kMnSbo + (k >> 3) * kMnLbo + (k & 7) * 16;
        }
        if (PREC == 1) store8_hi(t.x[i], hi + off);
        else if (PREC == 2) split_store8_3(t.x[i], hi + off, lo + off, lo + (lo - hi) + off);
        else split_store8(t.x[i], hi + off, lo + off);
    }
}

template <int PREC> struct GemmCfg {
    static constexpr int kParts = PREC == 2 ? 3 : 2;                    // images per operand
    static constexpr int kStages = PREC == 2 ? 2 : kGemmStages;        // 2 x 6 tiles = 101 KB: still two CTAs per SM
    static constexpr int kStage = 2 * kParts * kTileBytes;
    static constexpr size_t kSmem = (size_t)kStages * kStage + 1024;
};

template <int AMN, int BMN, int PREC = 0>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm3_kernel(const __grid_constant__ GemmArgs g) {
    constexpr int kGemmStages = GemmCfg<PREC>::kStages, kStageBytes = GemmCfg<PREC>::kStage, kParts = GemmCfg<PREC>::kParts;
    constexpr bool BF16 = PREC == 1;
    // six-term mode: the hi.hi products accumulate in columns [0,128), the five correction terms (2^-9 and smaller) in
    // [128,256); the epilogue adds the two in fp32.  The tensor core's accumulation is not round-to-nearest — measured: one
    // accumulator taking all 42 UMMAs of a K = 100 product is 3-5x less accurate than an fp32 FMA chain — and its error
    // scales with the number of UMMAs into the LARGE accumulator; this way that number is 7 instead of 42.
    constexpr int kTmemCols = PREC == 2 ? 2 * kGemmBN : kGemmBN;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    __shared__ __align__(8) uint64_t full_bar[kGemmStages], empty_bar[kGemmStages], done_bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
    const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
    const int nkb = (kend - kbeg + kGemmBK - 1) / kGemmBK;

    if (tid == 0) {
        for (int s = 0; s < kGemmStages; ++s) { mbar_init(&full_bar[s], 8); mbar_init(&empty_bar[s], 1); }
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (warp < 8) {
        // ------------------------------ producers ------------------------------
        TileRegs ta, tb;
        if (nkb > 0) {
            load_tile<AMN>(g.A, g.lda, m0, g.M, kbeg, kend, tid, ta);
            load_tile<BMN>(g.B, g.ldb, n0, g.N, kbeg, kend, tid, tb);
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kGemmStages, round = kb / kGemmStages;
            if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);      // the MMAs that read this stage have retired
            uint8_t *st = smem + (size_t)s * kStageBytes;
            store_tile<AMN, PREC>(ta, st, st + kTileBytes, tid);
            store_tile<BMN, PREC>(tb, st + kParts * kTileBytes, st + (kParts + 1) * kTileBytes, tid);
            if (kb + 1 < nkb) {                                             // next block's loads fly during the fence/arrive/wait
                const int k1 = kbeg + (kb + 1) * kGemmBK;
                load_tile<AMN>(g.A, g.lda, m0, g.M, k1, kend, tid, ta);
                load_tile<BMN>(g.B, g.ldb, n0, g.N, k1, kend, tid, tb);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the UMMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        }
    } else if (lane == 0) {
        // ------------------------------ MMA issuer (one thread) ------------------------------
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)AMN << 15) | ((uint32_t)BMN << 16) |
                               ((uint32_t)(kGemmBN >> 3) << 17) | ((uint32_t)(kGemmBM >> 4) << 24);
        constexpr uint32_t a_lbo = AMN ? kMnLbo : kKLbo, a_sbo = AMN ? kMnSbo : kKSbo, a_step = AMN ? 2 * kMnLbo : 2 * kKLbo;
        constexpr uint32_t b_lbo = BMN ? kMnLbo : kKLbo, b_sbo = BMN ? kMnSbo : kKSbo, b_step = BMN ? 2 * kMnLbo : 2 * kKLbo;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kGemmStages, round = kb / kGemmStages;
            mbar_wait(&full_bar[s], round & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(smem + (size_t)s * kStageBytes);
#pragma unroll
            for (int ks = 0; ks < kGemmBK / 16; ++ks) {
                const uint64_t ah = umma_desc(base + ks * a_step, a_lbo, a_sbo);
                const uint64_t al = umma_desc(base + kTileBytes + ks * a_step, a_lbo, a_sbo);
                const uint64_t bh = umma_desc(base + kParts * kTileBytes + ks * b_step, b_lbo, b_sbo);
                const uint64_t bl = umma_desc(base + (kParts + 1) * kTileBytes + ks * b_step, b_lbo, b_sbo);
                const uint32_t acc0 = (kb > 0 || ks > 0) ? 1u : 0u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(ah), "l"(bh), "r"(idesc), "r"(acc0) : "memory");
                if (!BF16) {
                    const uint32_t tc = PREC == 2 ? tmem + kGemmBN : tmem;          // correction accumulator (six-term mode)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tc), "l"(ah), "l"(bl), "r"(idesc), "r"(PREC == 2 ? acc0 : 1u) : "memory");
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tc), "l"(al), "l"(bh), "r"(idesc), "r"(1u) : "memory");
                }
                if (PREC == 2) {     // here al / bl are the MID images; the LO images follow them
                    const uint64_t a3 = umma_desc(base + 2 * kTileBytes + ks * a_step, a_lbo, a_sbo);
                    const uint64_t b3 = umma_desc(base + (kParts + 2) * kTileBytes + ks * b_step, b_lbo, b_sbo);
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tmem + kGemmBN), "l"(ah), "l"(b3), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tmem + kGemmBN), "l"(a3), "l"(bh), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tmem + kGemmBN), "l"(al), "l"(bl), "r"(idesc), "r"(1u) : "memory");
                }
            }
            // commit: arrives on the barrier when all MMAs issued so far have completed (implies before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    }

    // ------------------------------ epilogue: warps 0-3 drain TMEM ------------------------------
    if (warp < 4) {
        mbar_wait(&done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int gm = m0 + warp * 32 + lane;
        float *crow = g.C + (g.splits > 1 ? (size_t)blockIdx.z * g.M * g.ldc : 0) + (size_t)gm * g.ldc + n0;
        const bool add_bias = g.bias != nullptr && g.splits == 1;
#pragma unroll 1
        for (int c0 = 0; c0 < kGemmBN; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (PREC == 2) {
                uint32_t w[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                               "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                             : "r"(taddr + (uint32_t)kGemmBN));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
            }
            if (gm < g.M) {
                if (n0 + c0 + 16 <= g.N && (g.ldc & 3) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                               __uint_as_float(v[4 * q + 3]));
                        if (add_bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(g.bias + n0 + c0) + q);
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        reinterpret_cast<float4 *>(crow + c0)[q] = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c0 + j < g.N) {
                            const float o = __uint_as_float(v[j]) + (add_bias ? __ldg(g.bias + n0 + c0 + j) : 0.f);
                            crow[c0 + j] = g.relu ? fmaxf(o, 0.f) : o;
                        }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
    }
}

// C[i] = bias[i % N] + sum_s ws[s][i]   (fixed order -> deterministic)
static __global__ void gemm3_reduce_kernel(const float *__restrict__ ws, const float *__restrict__ bias, float *__restrict__ C,
                                    int M, int N, int ldc, int splits) {
    const size_t total = (size_t)M * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
        float s = bias ? __ldg(bias + n) : 0.f;
        for (int k = 0; k < splits; ++k) s += ws[(size_t)k * M * N + i];
        C[(size_t)m * ldc + n] = s;
    }
}

// The same for MANY partials of a SMALL output (K = T*N weight-gradient products of narrow layers: up to 294 splits of a few
// thousand outputs): with one thread per output the 294 dependent-address loads of a thread ran at L2 latency (38 us per
// launch).  Here a block owns 32 consecutive outputs, warp w sums the splits w, w + 8, ... (coalesced 128-byte rows, four
// independent accumulators), and the eight partial sums are combined in fixed order: deterministic, ~6 us.
static __global__ void __launch_bounds__(256) gemm3_reduce_wide_kernel(const float *__restrict__ ws, const float *__restrict__ bias,
                                                                      float *__restrict__ C, int M, int N, int ldc, int splits) {
    __shared__ float part[8][32];
    const size_t total = (size_t)M * N, i = (size_t)blockIdx.x * 32 + (threadIdx.x & 31);
    const int w = threadIdx.x >> 5;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (i < total) {
        int k = w;
        for (; k + 24 < splits; k += 32) {
            s0 += __ldg(ws + (size_t)k * total + i);
            s1 += __ldg(ws + (size_t)(k + 8) * total + i);
            s2 += __ldg(ws + (size_t)(k + 16) * total + i);
            s3 += __ldg(ws + (size_t)(k + 24) * total + i);
        }
        for (; k < splits; k += 8) s0 += __ldg(ws + (size_t)k * total + i);
    }
    part[w][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (w == 0 && i < total) {
        float s = bias ? __ldg(bias + (int)(i % N)) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += part[j][threadIdx.x];
        C[(i / N) * (size_t)ldc + (i % N)] = s;
    }
}

// =============================================================================================
// gemm3w: the same split-bf16 product for the case "B is a WEIGHT" (y = x W^T: NT; dx = dy W: NN), where M = T*N rows
// is huge and N, K are a layer's widths.  Differences to gemm3_kernel:
//   * the weight is split ONCE per call into bf16 hi/lo tile images already in the canonical UMMA layout
//     (gemm3w_pack_kernel, a few hundred KB), so the main kernel's producers only convert the activation operand and
//     the weight tiles arrive by one 32 KB bulk async copy (TMA engine) per k-block;
//   * a CTA owns 128 rows x up to 256 output columns (one UMMA of N = 256 per term): half the redundant activation
//     conversions of the 128-wide tile, 256 TMEM columns -> still 2 CTAs per SM;
//   * all 8 producer warps drain the accumulator (2 warps per TMEM lane quarter).
// =============================================================================================
constexpr int kWNC = 256, kWStages = 2;
constexpr int kWImgBytes = kWNC * kGemmBK * 2;                 // one bf16 image (hi or lo) of a [256 x 32] weight block
constexpr int kWStageBytes = 2 * kTileBytes + 2 * kWImgBytes;   // A_hi, A_lo (padded gemm3 layout) + B_hi, B_lo images
constexpr size_t kWSmemBytes = (size_t)kWStages * kWStageBytes + 1024;

// image index: ((n_chunk * nkb + kb) * 2 + part) ; within an image
//   BMN = 0 (W[n][k], K-major):  (k/8)*4096 + (n/8)*128 + (n%8)*16 + (k%8)*2        LBO 4096, SBO 128, k16 step 8192
//   BMN = 1 (W[k][n], MN-major): (n/8)*512 + (k/8)*128 + (k%8)*16 + (n%8)*2         LBO 128,  SBO 512, k16 step 256
template <int BMN>
__global__ void __launch_bounds__(256) gemm3w_pack_kernel(const float *__restrict__ W, int ldw, int N, int K, int nkb, uint8_t *__restrict__ img,
                                                          size_t total_chunks) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_chunks; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 1023);                         // 16-byte chunk inside an image
        const size_t blk = i >> 10;                            // n_chunk * nkb + kb
        const int kb = (int)(blk % nkb), nch = (int)(blk / nkb);
        float x[8];
        if (BMN == 0) {
            const int kc = c >> 8, n = c & 255;                // chunk = 8 consecutive k of row n
            const int gn = nch * kWNC + n, gk = kb * kGemmBK + 8 * kc;
            const int valid = gn < N ? max(0, min(8, K - gk)) : 0;
            if (valid > 0) load8(W + (size_t)gn * ldw + gk, valid, x);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = 0.f;
            }
        } else {
            const int mc = c >> 5, k = c & 31;                 // chunk = 8 consecutive n at row k: c = mc*32 + (k/8)*8 + k%8
            const int gk = kb * kGemmBK + k, gn = nch * kWNC + 8 * mc;
            const int valid = gk < K ? max(0, min(8, N - gn)) : 0;
            if (valid > 0) load8(W + (size_t)gk * ldw + gn, valid, x);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = 0.f;
            }
        }
        uint8_t *hi = img + blk * (2 * (size_t)kWImgBytes) + (size_t)c * 16;
        split_store8(x, hi, hi + kWImgBytes);
    }
}

struct GemmWArgs {
    const float *A, *bias;
    const uint8_t *img;   // packed weight images
    float *C;
    int M, N, K, lda, ldc, nkb, relu;
};

template <int BMN, bool BF16 = false>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm3w_kernel(const __grid_constant__ GemmWArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    __shared__ __align__(8) uint64_t full_bar[kWStages], empty_bar[kWStages], done_bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * kGemmBM, nch = blockIdx.x, n0 = nch * kWNC;
    const int ncols = min(kWNC, g.N - n0);                     // valid output columns of this CTA
    const int nkb = g.nkb;

    if (tid == 0) {
        for (int s = 0; s < kWStages; ++s) { mbar_init(&full_bar[s], 9); mbar_init(&empty_bar[s], 1); }
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(kWNC));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (warp < 8) {
        // ------------------------------ producers: convert A, fetch the packed weight block ------------------------------
        TileRegs ta;
        load_tile<0>(g.A, g.lda, m0, g.M, 0, g.K, tid, ta);
        const uint8_t *wsrc = g.img + (size_t)nch * nkb * (2 * (size_t)kWImgBytes);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kWStages, round = kb / kWStages;
            if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);      // the MMAs that read this stage have retired
            uint8_t *st = smem + (size_t)s * kWStageBytes;
            if (tid == 0) {
                constexpr uint32_t wbytes = BF16 ? kWImgBytes : 2 * kWImgBytes;           // bf16 mode: the hi image only
                mbar_expect_tx(&full_bar[s], wbytes);
                bulk_g2s(st + 2 * kTileBytes, wsrc + (size_t)kb * (2 * (size_t)kWImgBytes), wbytes, &full_bar[s]);
            }
            store_tile<0, BF16 ? 1 : 0>(ta, st, st + kTileBytes, tid);
            if (kb + 1 < nkb) load_tile<0>(g.A, g.lda, m0, g.M, (kb + 1) * kGemmBK, g.K, tid, ta);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        }
    } else if (lane == 0) {
        // ------------------------------ MMA issuer ------------------------------
        const int n_mma = (ncols + 15) & ~15;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)BMN << 16) | ((uint32_t)(n_mma >> 3) << 17) |
                               ((uint32_t)(kGemmBM >> 4) << 24);
        constexpr uint32_t b_lbo = BMN ? 128 : 4096, b_sbo = BMN ? 512 : 128, b_step = BMN ? 256 : 8192;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kWStages, round = kb / kWStages;
            mbar_wait(&full_bar[s], round & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(smem + (size_t)s * kWStageBytes);
#pragma unroll
            for (int ks = 0; ks < kGemmBK / 16; ++ks) {
                const uint64_t ah = umma_desc(base + ks * 2 * kKLbo, kKLbo, kKSbo);
                const uint64_t al = umma_desc(base + kTileBytes + ks * 2 * kKLbo, kKLbo, kKSbo);
                const uint64_t bh = umma_desc(base + 2 * kTileBytes + ks * b_step, b_lbo, b_sbo);
                const uint64_t bl = umma_desc(base + 2 * kTileBytes + kWImgBytes + ks * b_step, b_lbo, b_sbo);
                const uint32_t acc0 = (kb > 0 || ks > 0) ? 1u : 0u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(ah), "l"(bh), "r"(idesc), "r"(acc0) : "memory");
                if (!BF16) {
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tmem), "l"(ah), "l"(bl), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tmem), "l"(al), "l"(bh), "r"(idesc), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    }

    // ------------------------------ epilogue: warps 0-7, two per TMEM lane quarter ------------------------------
    // A thread owns an accumulator ROW (TMEM lane); storing it directly would scatter 16-byte pieces over 32 rows per
    // instruction.  Each warp therefore transposes 64 columns at a time through its own shared-memory patch (the
    // operand stages are dead by now) and writes whole 256-byte row segments: 4 cache lines per store instruction.
    if (warp < 8) {
        mbar_wait(&done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, chalf = warp >> 2;
        const int gm0 = m0 + q * 32;
        const bool vec = (g.ldc & 3) == 0 && (g.N & 3) == 0;
        float *patch = reinterpret_cast<float *>(smem) + warp * (32 * 68);
#pragma unroll 1
        for (int cb = chalf * 128; cb < chalf * 128 + 128 && cb < ncols; cb += 64) {
#pragma unroll 1
            for (int c0 = cb; c0 < cb + 64 && c0 < ncols; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             "tcgen05.wait::ld.sync.aligned;"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                if (vec) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        float4 o = make_float4(__uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1]), __uint_as_float(v[4 * q4 + 2]),
                                               __uint_as_float(v[4 * q4 + 3]));
                        if (g.bias != nullptr && c0 + 4 * q4 < ncols) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(g.bias + n0 + c0) + q4);
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4 *>(patch + lane * 68 + (c0 - cb) + 4 * q4) = o;
                    }
                } else if (gm0 + lane < g.M) {
                    float *crow = g.C + (size_t)(gm0 + lane) * g.ldc + n0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < ncols) {
                            const float o = __uint_as_float(v[j]) + (g.bias != nullptr ? __ldg(g.bias + n0 + c0 + j) : 0.f);
                            crow[c0 + j] = g.relu ? fmaxf(o, 0.f) : o;
                        }
                }
            }
            if (vec) {
                __syncwarp();
                const int col = (lane & 15) * 4, rsub = lane >> 4;
                if (cb + col < ncols) {
#pragma unroll 4
                    for (int i = 0; i < 16; ++i) {
                        const int row = 2 * i + rsub;
                        if (gm0 + row < g.M)
                            *reinterpret_cast<float4 *>(g.C + (size_t)(gm0 + row) * g.ldc + n0 + cb + col) =
                                *reinterpret_cast<const float4 *>(patch + row * 68 + col);
                    }
                }
                __syncwarp();
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kWNC));
    }
}

}  // namespace lsthm
