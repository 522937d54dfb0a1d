// Shared device helpers for the LSTHM recurrence kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lsthm {

constexpr int kMaxMod = 3;
constexpr int kHeads = 4;

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) — used to stage each timestep's
// contiguous [rows][width] tile of per-utterance features into shared memory one step ahead.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// 16-byte cp.async (LDGSTS) for tiles that need a padded shared layout.
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Pointwise math.  fp32 mode must stay well inside 1e-4 (logits) / 1e-3 (grads) of the fp32
// reference, so no tanh.approx / ex2.approx-only shortcuts with 1e-3 relative error here.
// ---------------------------------------------------------------------------------------------
// The divisions are MUFU.RCP based (__fdividef: <= 2 ulp, denominators here are in [1, 2^126]) instead of the IEEE
// division sequence: a third of the instructions of the gate nonlinearities, same 1e-7 absolute accuracy.
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
    // 1 - 2/(e^{2x}+1): abs error ~1e-7, saturates cleanly for |x| large (e = inf -> 2/inf = 0).
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// ---------------------------------------------------------------------------------------------
// Register-tiled "tall-skinny" product: one thread owns 4 adjacent output columns for all MT
// dialogue rows of the CTA's tile.  Weights are streamed from L2 as float4 (k-major image, row k
// = 4 adjacent columns), activations are broadcast from shared memory in k-major [k][MTP].
//   acc[c][m] += W[k][col4 + c] * act[k][m]      k = 0 .. n-1
// Rows are held as pairs so the inner product runs on the packed fp32x2 FMA of sm_100
// (SASS FFMA2 with a scalar-broadcast weight operand): half the issue slots of scalar FFMA.
// The weight loads of the next 4 k-rows are issued before the current 4 are consumed
// (explicit double buffering) so one L2 round trip is always in flight per thread.
// ---------------------------------------------------------------------------------------------
template <int MT>
struct Acc {
    static constexpr int NP = (MT + 1) / 2;
    float2 v[4][NP];
    __device__ __forceinline__ float get(int c, int m) const { return (m & 1) ? v[c][m >> 1].y : v[c][m >> 1].x; }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < NP; ++p) v[c][p] = make_float2(0.f, 0.f);
    }
};

template <int MT, int MTP>
__device__ __forceinline__ void fma_row(Acc<MT> &acc, const float4 w, const float *__restrict__ a) {
    constexpr int NP = Acc<MT>::NP;
    float2 av[MTP / 2];
#pragma unroll
    for (int i = 0; i < (NP + 1) / 2; ++i) {
        const float4 t = reinterpret_cast<const float4 *>(a)[i];
        av[2 * i] = make_float2(t.x, t.y);
        av[2 * i + 1] = make_float2(t.z, t.w);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        acc.v[0][p] = __ffma2_rn(make_float2(w.x, w.x), av[p], acc.v[0][p]);
        acc.v[1][p] = __ffma2_rn(make_float2(w.y, w.y), av[p], acc.v[1][p]);
        acc.v[2][p] = __ffma2_rn(make_float2(w.z, w.z), av[p], acc.v[2][p]);
        acc.v[3][p] = __ffma2_rn(make_float2(w.w, w.w), av[p], acc.v[3][p]);
    }
}

#ifndef LSTHM_MAC_VARIANT
#define LSTHM_MAC_VARIANT 8
#endif

#if LSTHM_MAC_VARIANT == 1
// double-buffered blocks of 4 k-rows (8 weight quads live): deepest prefetch, highest register use
template <int MT, int MTP, int KB_UNUSED = 0>
__device__ __forceinline__ void mac(Acc<MT> &acc, const float4 *__restrict__ w, const int ldw4,
                                    const float *__restrict__ act, const int n) {
    const int n4 = n & ~3;
    const size_t ld = (size_t)ldw4;
    const float4 *p = w;
    float4 a0, a1, a2, a3, b0, b1, b2, b3;
    if (n4 > 0) {
        a0 = __ldg(p); a1 = __ldg(p + ld); a2 = __ldg(p + 2 * ld); a3 = __ldg(p + 3 * ld);
    }
    int k = 0;
#pragma unroll 1
    for (; k + 8 <= n4; k += 8) {
        const float4 *pb = p + 4 * ld;
        b0 = __ldg(pb); b1 = __ldg(pb + ld); b2 = __ldg(pb + 2 * ld); b3 = __ldg(pb + 3 * ld);
        fma_row<MT, MTP>(acc, a0, act);
        fma_row<MT, MTP>(acc, a1, act + MTP);
        fma_row<MT, MTP>(acc, a2, act + 2 * MTP);
        fma_row<MT, MTP>(acc, a3, act + 3 * MTP);
        // next block of 4 (clamped to a valid address on the last trip: the reload is never used)
        const float4 *pa = (k + 8 < n4) ? p + 8 * ld : p;
        a0 = __ldg(pa); a1 = __ldg(pa + ld); a2 = __ldg(pa + 2 * ld); a3 = __ldg(pa + 3 * ld);
        fma_row<MT, MTP>(acc, b0, act + 4 * MTP);
        fma_row<MT, MTP>(acc, b1, act + 5 * MTP);
        fma_row<MT, MTP>(acc, b2, act + 6 * MTP);
        fma_row<MT, MTP>(acc, b3, act + 7 * MTP);
        p += 8 * ld;
        act += 8 * MTP;
    }
    if (k < n4) {   // one remaining block of 4, already in a0..a3
        fma_row<MT, MTP>(acc, a0, act);
        fma_row<MT, MTP>(acc, a1, act + MTP);
        fma_row<MT, MTP>(acc, a2, act + 2 * MTP);
        fma_row<MT, MTP>(acc, a3, act + 3 * MTP);
        k += 4;
        p += 4 * ld;
        act += 4 * MTP;
    }
    for (; k < n; ++k) {
        fma_row<MT, MTP>(acc, __ldg(p), act);
        p += ld;
        act += MTP;
    }
}
#elif LSTHM_MAC_VARIANT == 2
// double-buffered blocks of 2 k-rows (4 weight quads live)
template <int MT, int MTP, int KB_UNUSED = 0>
__device__ __forceinline__ void mac(Acc<MT> &acc, const float4 *__restrict__ w, const int ldw4,
                                    const float *__restrict__ act, const int n) {
    const int n2 = n & ~1;
    const size_t ld = (size_t)ldw4;
    const float4 *p = w;
    float4 a0, a1, b0, b1;
    if (n2 > 0) {
        a0 = __ldg(p); a1 = __ldg(p + ld);
    }
    int k = 0;
#pragma unroll 1
    for (; k + 4 <= n2; k += 4) {
        const float4 *pb = p + 2 * ld;
        b0 = __ldg(pb); b1 = __ldg(pb + ld);
        fma_row<MT, MTP>(acc, a0, act);
        fma_row<MT, MTP>(acc, a1, act + MTP);
        const float4 *pa = (k + 4 < n2) ? p + 4 * ld : p;
        a0 = __ldg(pa); a1 = __ldg(pa + ld);
        fma_row<MT, MTP>(acc, b0, act + 2 * MTP);
        fma_row<MT, MTP>(acc, b1, act + 3 * MTP);
        p += 4 * ld;
        act += 4 * MTP;
    }
    if (k < n2) {
        fma_row<MT, MTP>(acc, a0, act);
        fma_row<MT, MTP>(acc, a1, act + MTP);
        k += 2;
        p += 2 * ld;
        act += 2 * MTP;
    }
    if (k < n) fma_row<MT, MTP>(acc, __ldg(p), act);
}
#else
// single-buffered blocks of KB (4 or 8) k-rows: all loads of a block first, then the FMAs.  Measured on B200 (N=1024,
// T=110): mab_fwd is fastest with KB = 8 (5.00 vs 5.22 ms), mab_bwd with KB = 4 (5.40 vs 5.82 ms), both sps kernels
// with KB = 8 (sps_bwd 7.86 vs 8.50 ms).
#ifndef LSTHM_MAC_BWD
#define LSTHM_MAC_BWD 4
#endif
template <int MT, int MTP, int KB = LSTHM_MAC_VARIANT>
__device__ __forceinline__ void mac(Acc<MT> &acc, const float4 *__restrict__ w, const int ldw4,
                                    const float *__restrict__ act, const int n) {
    const size_t ld = (size_t)ldw4;
    const float4 *p = w;
    int k = 0;
#pragma unroll 1
    for (; k + KB <= n; k += KB) {
        float4 wv[KB];
#pragma unroll
        for (int u = 0; u < KB; ++u) wv[u] = __ldg(p + u * ld);
#pragma unroll
        for (int u = 0; u < KB; ++u) fma_row<MT, MTP>(acc, wv[u], act + u * MTP);
        p += KB * ld;
        act += KB * MTP;
    }
    for (; k < n; ++k) {
        fma_row<MT, MTP>(acc, __ldg(p), act);
        p += ld;
        act += MTP;
    }
}
#endif

// Split-K partial sums: part[(split*MTP + m)*J + col]; a thread stores its 4 columns as one float4.
template <int MT, int MTP>
__device__ __forceinline__ void store_partial(float *part, const int J, const int split, const int col4,
                                              const Acc<MT> &acc) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
        *reinterpret_cast<float4 *>(part + (size_t)(split * MTP + m) * J + col4) =
            make_float4(acc.get(0, m), acc.get(1, m), acc.get(2, m), acc.get(3, m));
}

// Counter-based in-kernel dropout for dense attention tiles: ONE one-round 32-bit integer finaliser ("lowbias32") per PAIR of adjacent columns, 16 random bits per
// element (keep iff bits >= round(p * 65536); scale 65536 / (65536 - thr), so the mask is exactly unbiased).  `a`
// names the (step, dialogue) tile, `row` / `col` the element; forward and backward evaluate the same function.
struct PairDrop {
    uint32_t base, thr;
    float scale;
    __device__ __forceinline__ PairDrop(unsigned long long seed, uint32_t a, float p) {
        thr = (uint32_t)(p * 65536.0f + 0.5f);
        scale = 65536.0f / (65536.0f - (float)thr);
        uint32_t x = (uint32_t)seed ^ (a * 0x9E3779B9u);
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        base = x ^ (uint32_t)(seed >> 32);
    }
    __device__ __forceinline__ void pair(uint32_t row, uint32_t pair_idx, float &s0, float &s1) const {
        uint32_t x = base + row * 0x85EBCA6Bu + pair_idx * 0xC2B2AE35u;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        s0 = (x & 0xffffu) >= thr ? scale : 0.f;
        s1 = (x >> 16) >= thr ? scale : 0.f;
    }
    __device__ __forceinline__ float one(uint32_t row, uint32_t col) const {
        float s0, s1;
        pair(row, col >> 1, s0, s1);
        return (col & 1) ? s1 : s0;
    }
};

// k-major [k][MTP] vector of one unit: load / store all rows at once (conflict-free 16B accesses).
template <int MTP>
__device__ __forceinline__ void load_rows(float (&v)[MTP], const float *p) {
#pragma unroll
    for (int i = 0; i < MTP / 4; ++i) {
        const float4 t = reinterpret_cast<const float4 *>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
template <int MTP>
__device__ __forceinline__ void store_rows(float *p, const float (&v)[MTP]) {
#pragma unroll
    for (int i = 0; i < MTP / 4; ++i)
        reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

}  // namespace lsthm
