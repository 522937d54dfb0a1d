// Weight-stationary tensor-core recurrence of HybridRNN_AT / HybridRNN_ATV (LSTHM cells + multi-attention block).
//
// What it replaces in the reference: the body of the time loop of MARN.forward (model/HybridRNN_ATV.py:117-143,
// AT: model/HybridRNN_AT.py:107-132) and its autograd BPTT — see include/lsthm_b200.h (lsthm_mab_*).
//
// Design (sm_100a):
//   * A GROUP of G co-resident CTAs (cooperative launch, G = 12 for ATV, 9 for AT) owns a block of up to 96 dialogues for
//     all T steps.  The chain weights (composite form: gates [U_m | W2], attention logits Watt, fused reduce+fc.0 W1) are
//     sharded over the ranks of the group and stay RESIDENT in shared memory for the whole launch as bf16 hi/lo images
//     in the canonical K-major UMMA layout (1.45 MB / 12 = 136 KB per CTA): nothing is streamed from L2 per step.
//   * Every product is a real dense GEMM tile on the 5th-generation tensor cores: M = the group's dialogues (one
//     tcgen05.mma M = 128), N = the rank's slice of output features, fp32 accumulation in TMEM, operands split
//     x = hi + lo (bf16 each) with three UMMAs per k-step (hi.hi + hi.lo + lo.hi): fp32-parity accurate (< 2e-5).
//   * rank r owns in stage 1 a slice of hidden units of ONE modality (their four gates, their cell state: K = dh_m + 64),
//     in stage 2/3 a (head, feature range) slice of the attention (logit rows, then the matching K-slice of W1).
//   * Three group exchanges per step through L2 (per-group monotonic counters, release/acquire, bulk-copy gathers):
//       A: c_t, h_t slices (already split into bf16 hi/lo operand images by their producer)  -> all-gather
//       B: per-rank partial  W1[:, slice] . (exp(e - m_r) * c)  + local softmax statistics (m_r, s_r)
//          -> reduce-scatter: rank i combines the dialogues [i*cd, (i+1)*cd) in fixed rank order (deferred softmax
//             normalisation), applies bias / ReLU / dropout mask: u_t
//       C: u_t operand image -> all-gather
//   * the epilogues run on 8 warps straight out of TMEM (tcgen05.ld): thread = dialogue row, so the LSTM cell update
//     and the softmax statistics are thread-local.
#pragma once
#include "gemm3_kernels.cuh"

namespace lsthm {

constexpr int kM2MaxRanks = 16;
constexpr int kM2EpiWarps = 8;
// 8 epilogue warps (two warpgroups) + one control warpgroup whose first warp issues MMAs, copies and barriers.  The control
// warpgroup hands most of its registers to the epilogue warps (setmaxnreg): with shared memory at ~215 KB the L1 is only a
// few KB, so a spilled register costs an L2 round trip — the epilogues must not spill.
constexpr int kM2Threads = (kM2EpiWarps + 4) * 32;
constexpr int kM2RegsEpi = 224, kM2RegsCtl = 56;
// the pool setmaxnreg draws from is the CTA's own allocation (launch registers x threads, 168 x 384 under these launch bounds),
// not the whole register file: an increase that the control warpgroup's release cannot cover blocks forever
static_assert(kM2EpiWarps * 32 * kM2RegsEpi + 4 * 32 * kM2RegsCtl <= kM2Threads * 168, "setmaxnreg budget exceeds the CTA's register pool");
constexpr int kM2MaxDG = 96;                            // dialogues per group (operand buffer budget)
constexpr int kM2MaxNJ = 80;                            // stage-2 feature range per rank (at most 5 chunks per epilogue warp)
constexpr int kM2MaxNU = 16;                            // stage-1 hidden units per rank: one 8-unit chunk per epilogue warp of a lane quarter
constexpr int kM2CPH = 1;                               // chunks per epilogue half (kM2MaxNU / 16)
constexpr long long kM2Timeout = 1LL << 31;             // cycles (~1 s): a stuck exchange traps instead of hanging the GPU

struct M2Rank {
    int m, u0, nu;        // stage 1: modality, first hidden unit (global index, multiple of 8), units (multiple of 8)
    int head, j0, nj;     // stage 2/3: head (-1 = none), feature range [j0, j0 + nj), nj multiple of 16
    int dhm, offm;        // cell size and first unit of the own modality
    int mr0, mr1;         // ranks of the own modality: [mr0, mr1)
    int pad0, pad1;       // 48 bytes: the kernels read their rank's record from global memory with three 16-byte loads
};

struct M2Plan {
    int T, N, nm, MH, D, G4;
    int dh[kMaxMod], off[kMaxMod];
    int G, nr;                       // ranks per group; ranges per head (stage-2 jobs are ranks 0 .. 4 nr - 1, head-major)
    M2Rank r[kM2MaxRanks];
    int DG, Mr, ngroups, nblocks, cd; // dialogues per block, rows rounded up to 8, co-resident groups, blocks, combine share
    int blob_f, blob_b;              // per-rank weight blob strides (bytes), forward / backward
    int act_f, act_b;                // operand buffer bytes
    // exchange workspace (bytes inside a group's area)
    int ws_xc, ws_xh, ws_xu, ws_xp, ws_xst, ws_xdc, ws_xdh, ws_xdu, ws_xdup, ws_group;
};

// ---- per-rank weight blob layouts (byte offsets; the blob is copied verbatim into shared memory) ----
struct M2FwdBlob { int wg, wa, w1, batt, b1, bv, total, kcg, ng; };
struct M2BwdBlob { int w1t, wat, wf, total, ng, nf; };

__host__ __device__ inline int m2_align(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline M2FwdBlob m2_fwd_blob(const M2Plan &P, const M2Rank &R) {
    M2FwdBlob b;
    b.ng = 4 * R.nu;
    b.kcg = (R.dhm + P.MH) / 8;
    int o = 0;
    b.wg = o; o += 2 * b.kcg * b.ng * 16;                            // [hi|lo][kc][n = ul*4+gate][8]  K = [h_m | u]
    b.wa = o; o += R.head >= 0 ? 2 * (P.D / 8) * R.nj * 16 : 0;      // [hi|lo][kc][n = j - j0][8]     K = c (D)
    b.w1 = o; o += R.head >= 0 ? 2 * (R.nj / 8) * P.MH * 16 : 0;     // [hi|lo][kc][n = q][8]          K = attended slice
    b.batt = o; o += kM2MaxNJ * 4;
    b.b1 = o; o += P.MH * 4;
    b.bv = o; o += 4 * kM2MaxNU * 4;
    b.total = m2_align(o, 128);
    return b;
}

__host__ __device__ inline M2BwdBlob m2_bwd_blob(const M2Plan &P, const M2Rank &R) {
    M2BwdBlob b;
    b.ng = 4 * R.nu;
    b.nf = P.MH + R.dhm;
    int o = 0;
    b.w1t = o; o += R.head >= 0 ? 2 * (P.MH / 8) * R.nj * 16 : 0;    // [N = nj][K = MH]          d attended = dup . W1
    b.wat = o; o += R.head >= 0 ? 2 * (R.nj / 8) * P.D * 16 : 0;     // [N = D][K = nj]           dc partial = de . Watt
    b.wf = o; o += 2 * (b.ng / 8) * b.nf * 16;                       // [N = MH + dh_m][K = ng]   [du | dh_m] partial = ds . [W2 | U_m]
    b.total = m2_align(o, 128);
    return b;
}

// ---------------------------------------------------------------------------------------------
// composite weights (fp64 accumulation, rounded once to fp32) into a small fp32 staging area:
//   W1 = Wf1 . blockdiag(Wr_m)  [MH][4D] (column k = head*D + j),   b1 = Wf1 br + bf1
//   W2 = Vcat . Wf2             [4D][MH] (rows in the native gate order),  bv = Vcat bf2
// ---------------------------------------------------------------------------------------------
struct M2CompArgs {
    M2Plan P;
    int R, rd[kMaxMod], roff[kMaxMod];
    const float *V[kMaxMod], *Wr[kMaxMod], *br[kMaxMod], *Wf1, *bf1, *Wf2, *bf2;
    float *W1, *W2, *b1, *bv;
};

__global__ void mab_compose_kernel(const __grid_constant__ M2CompArgs a) {
    const M2Plan &P = a.P;
    const int D = P.D, G4 = P.G4, MH = P.MH, R = a.R;
    const int n1 = MH * G4, n2 = G4 * MH, total = n1 + n2 + MH + G4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        if (idx < n1) {
            const int col = idx / G4, k = idx - col * G4, head = k / D, j = k - head * D;
            int m = 0;
            while (m + 1 < P.nm && j >= P.off[m + 1]) ++m;
            const int jl = j - P.off[m], dh = P.dh[m];
            double s = 0.0;
            for (int r = 0; r < a.rd[m]; ++r)
                s += (double)__ldg(a.Wf1 + (size_t)col * R + a.roff[m] + r) * (double)__ldg(a.Wr[m] + (size_t)r * 4 * dh + head * dh + jl);
            a.W1[idx] = (float)s;
        } else if (idx < n1 + n2) {
            const int i2 = idx - n1, g = i2 / MH, q = i2 - g * MH;
            int m = 0;
            while (m + 1 < P.nm && g >= 4 * P.off[m + 1]) ++m;
            const float *vrow = a.V[m] + (size_t)(g - 4 * P.off[m]) * D;
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += (double)__ldg(vrow + j) * (double)__ldg(a.Wf2 + (size_t)j * MH + q);
            a.W2[i2] = (float)s;
        } else if (idx < n1 + n2 + MH) {
            const int col = idx - n1 - n2;
            double s = (double)__ldg(a.bf1 + col);
            for (int m = 0; m < P.nm; ++m)
                for (int r = 0; r < a.rd[m]; ++r) s += (double)__ldg(a.Wf1 + (size_t)col * R + a.roff[m] + r) * (double)__ldg(a.br[m] + r);
            a.b1[col] = (float)s;
        } else {
            const int g = idx - n1 - n2 - MH;
            int m = 0;
            while (m + 1 < P.nm && g >= 4 * P.off[m + 1]) ++m;
            const float *vrow = a.V[m] + (size_t)(g - 4 * P.off[m]) * D;
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += (double)__ldg(vrow + j) * (double)__ldg(a.bf2 + j);
            a.bv[g] = (float)s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// operand images: one thread per 16-byte chunk (8 consecutive K values of one N row), writes the hi and the lo image
// ---------------------------------------------------------------------------------------------
struct M2ImgArgs {
    M2Plan P;
    const float *U[kMaxMod], *Watt, *batt, *W1, *W2, *b1, *bv;
    uint8_t *blob_f, *blob_b;
    M2Rank *ranktab;
};

// value (n, k) of image `img` of rank R;  forward: 0 WG, 1 WA, 2 W1;  backward: 3 W1T, 4 WAT, 5 WF = [W2T ; UT] stacked along N
__device__ __forceinline__ float m2_img_value(const M2ImgArgs &a, const M2Rank &R, int img, int n, int k) {
    const M2Plan &P = a.P;
    const int m = R.m, dh = P.dh[m], D = P.D, MH = P.MH, G4 = P.G4, u0l = R.u0 - P.off[m], goff = 4 * P.off[m];
    switch (img) {
    case 0: {
        const int ul = n >> 2, gate = n & 3, row = gate * dh + u0l + ul;
        return k < dh ? __ldg(a.U[m] + (size_t)row * dh + k) : __ldg(a.W2 + (size_t)(goff + row) * MH + (k - dh));
    }
    case 1: return __ldg(a.Watt + (size_t)(R.head * D + R.j0 + n) * D + k);
    case 2: return __ldg(a.W1 + (size_t)n * G4 + R.head * D + R.j0 + k);
    case 3: return __ldg(a.W1 + (size_t)k * G4 + R.head * D + R.j0 + n);
    case 4: return __ldg(a.Watt + (size_t)(R.head * D + R.j0 + k) * D + n);
    default: {
        const int ul = k >> 2, gate = k & 3, row = gate * dh + u0l + ul;
        return n < MH ? __ldg(a.W2 + (size_t)(goff + row) * MH + n) : __ldg(a.U[m] + (size_t)row * dh + (n - MH));
    }
    }
}

__global__ void __launch_bounds__(256) mab_image_kernel(const __grid_constant__ M2ImgArgs a) {
    const M2Plan &P = a.P;
    const int rank = blockIdx.y;
    const M2Rank R = P.r[rank];
    const M2FwdBlob F = m2_fwd_blob(P, R);
    const M2BwdBlob B = m2_bwd_blob(P, R);
    const bool s2 = R.head >= 0;
    const int dh = R.dhm;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ranktab[rank] = R;          // the kernels' per-rank record
    // images: N rows, K/8 chunks, byte offset of the hi image
    const int iN[6] = {F.ng, R.nj, P.MH, R.nj, P.D, B.nf};
    const int iKc[6] = {F.kcg, P.D / 8, R.nj / 8, P.MH / 8, R.nj / 8, B.ng / 8};
    const int iOff[6] = {F.wg, F.wa, F.w1, B.w1t, B.wat, B.wf};
    uint8_t *bf = a.blob_f + (size_t)rank * P.blob_f, *bb = a.blob_b + (size_t)rank * P.blob_b;
    for (int img = 0; img < 6; ++img) {
        const bool on = (img == 0 || img == 5) ? true : s2;
        const int chunks = on ? iN[img] * iKc[img] : 0;
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x) {
            const int kc = c / iN[img], n = c - kc * iN[img];
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = m2_img_value(a, R, img, n, 8 * kc + e);
            uint8_t *hi = (img < 3 ? bf : bb) + iOff[img] + (size_t)c * 16;
            split_store8(x, hi, hi + (size_t)chunks * 16);
        }
    }
    // fp32 vectors of the forward blob
    const int goff = 4 * R.offm, u0l = R.u0 - R.offm;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kM2MaxNJ + P.MH + 4 * kM2MaxNU; i += gridDim.x * blockDim.x) {
        if (i < kM2MaxNJ) {
            reinterpret_cast<float *>(bf + F.batt)[i] = (s2 && i < R.nj) ? __ldg(a.batt + R.head * P.D + R.j0 + i) : 0.f;
        } else if (i < kM2MaxNJ + P.MH) {
            const int q = i - kM2MaxNJ;
            reinterpret_cast<float *>(bf + F.b1)[q] = __ldg(a.b1 + q);
        } else {
            const int n = i - kM2MaxNJ - P.MH;
            float v = 0.f;
            if (n < F.ng) {
                const int ul = n >> 2, gate = n & 3;
                v = __ldg(a.bv + goff + gate * dh + u0l + ul);
            }
            reinterpret_cast<float *>(bf + F.bv)[n] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[128 x N] (+)= A . B^T over nk16 k-steps of 16; both operands K-major no-swizzle images [chunk][row][8 bf16]:
// chunk stride = lbo bytes, 8-row groups contiguous (SBO 128).  Three terms per k-step: hi.hi + hi.lo + lo.hi.
// Called by the WHOLE control warp with warp-uniform arguments (descriptors stay in uniform registers); only the elected
// lane issues.  A k-step advances a descriptor's 14-bit address field by 2 * lbo / 16.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void m2_issue3(bool leader, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t b_hi,
                                          uint32_t b_lo, uint32_t b_lbo, int nk16, uint32_t idesc, bool zero_first) {
    uint64_t ah = umma_desc(a_hi, a_lbo, 128), al = umma_desc(a_lo, a_lbo, 128);
    uint64_t bh = umma_desc(b_hi, b_lbo, 128), bl = umma_desc(b_lo, b_lbo, 128);
    const uint64_t da = (uint64_t)((2 * a_lbo) >> 4), db = (uint64_t)((2 * b_lbo) >> 4);
    for (int ks = 0; ks < nk16; ++ks) {
        if (leader) {
            umma_f16(tmem_d, ah, bh, idesc, (zero_first && ks == 0) ? 0u : 1u);
            umma_f16(tmem_d, ah, bl, idesc, 1u);
            umma_f16(tmem_d, al, bh, idesc, 1u);
        }
        ah += da; al += da; bh += db; bl += db;
    }
}
__device__ __forceinline__ uint32_t m2_idesc(int n) {      // kind::f16, bf16 x bf16 -> f32, K-major both, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// bounded waits: a protocol bug or a lost peer traps (launch failure) instead of hanging the device
__device__ __forceinline__ void m2_mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > kM2Timeout) __trap();
}
// Spin on the group counter with RELAXED gpu-scope loads (they are served by L2 and leave the SM's L1 alone) and issue ONE
// acquire fence after the counter has arrived: an acquire load per spin would invalidate the L1 on every iteration, which
// evicts the epilogue warps' cached lines (including their spill slots) for as long as the control warp waits.
__device__ __forceinline__ void m2_poll(const unsigned *ctr, unsigned target) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if ((int)(v - target) < 0) {
        const long long t0 = clock64();
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (clock64() - t0 > kM2Timeout) __trap();
        } while ((int)(v - target) < 0);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// publish: the CTA's writes were collected by an mbarrier wait (they happen-before this thread); the release is cumulative
__device__ __forceinline__ void m2_signal(unsigned *ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
// 256-bit global accesses (sm_100: LDG/STG.256): a thread's 8-float chunk of a row-major row is one instruction, i.e. half the
// L1 wavefronts of two float4 accesses when the 32 lanes of a warp touch 32 different rows.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg8(const float *p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg8(float *p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// one instruction pulls a contiguous range (multiple of 16 bytes) into L2
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Makes a pointer opaque to the optimiser at this point.  The epilogue loops are fully unrolled over ranks / chunks; without
// this the compiler hoists every per-rank address (dozens of 64-bit values) out of the time loop and then spills them.
template <typename T>
__device__ __forceinline__ void m2_launder(T *&p) { asm volatile("" : "+l"(p)); }

// 8 fp32 -> one bf16 hi chunk + one bf16 lo chunk (registers)
__device__ __forceinline__ void m2_split8(const float (&x)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = pack_bf16(x[2 * i] - h0, x[2 * i + 1] - h1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// bf16 hi chunk + lo chunk -> 8 fp32 (hi + lo: 16 mantissa bits of the original)
__device__ __forceinline__ void m2_join8(const uint4 hi, const uint4 lo, float (&x)[8]) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
        x[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
    }
}

// Development trace: when the host has set a buffer (lsthm_mab_set_trace), the control thread and lane 0 of epilogue
// warp 0 of CTA 0 record clock64() at their phase boundaries, [step][role][16] — read back by profiles/dev_mab_check.py.
__device__ long long *g_m2_trace = nullptr;
#ifdef LSTHM_M2_TRACE
#define M2_TRACE(role, slot)                                                                          \
    do {                                                                                              \
        if (trace != nullptr) trace[((size_t)tstep * 2 + (role)) * 16 + (slot)] = clock64();          \
    } while (0)
#else
#define M2_TRACE(role, slot) do { (void)tstep; } while (0)      // production build: no trace code in the kernels
#endif

// shared-memory control block: mbarriers
enum { M2B_W = 0, M2B_H, M2B_U, M2B_C, M2B_G, M2B_E, M2B_P, M2B_B, M2E_A, M2E_B, M2E_C, M2E_D, M2B_X0, M2B_X1, M2B_X2, M2B_X3, M2_NBAR };
constexpr int kM2CtrlBytes = 256;

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct M2FwdArgs {
    M2Plan P;
    const uint8_t *blob;          // [G][blob_f]
    const M2Rank *ranktab;        // [G]
    const float *gx, *mask;       // [T][N][4D], [T][N][MH] or null
    float *hz, *sU;               // [T][N][2D] (h half), [T][N][MH]
    float *sC;                    // [T][N][D] cell states, row-major (the host's weight-gradient products read it)
    // private stash of the kernel pair (all or none), PIECE-MAJOR inside a dialogue block so that a warp's access is one
    // contiguous run:  [t][block][column / 4][row][4]  with Mr padded rows per block
    float *sCp, *sG, *sE, *sMS, *sP;  // c [D], gates [4D], logits [4D], (max, 1/sum) [4][2], per-head W1 product [4*MH]
    uint8_t *ws;                  // exchange workspace [ngroups][ws_group]
    unsigned *bars;               // [ngroups][4][32] zero-initialised counters
};

__global__ void __launch_bounds__(kM2Threads, 1) mab_fwd_kernel(const __grid_constant__ M2FwdArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const M2Plan &P = a.P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = blockIdx.x % P.G, grp = blockIdx.x / P.G;
    M2Rank R;                     // from global memory: indexing the parameter struct by rank would put a copy of it on the stack
    {
        const int4 *rp = reinterpret_cast<const int4 *>(a.ranktab + rank);
        const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
        R.m = r0.x; R.u0 = r0.y; R.nu = r0.z; R.head = r0.w; R.j0 = r1.x; R.nj = r1.y; R.dhm = r1.z; R.offm = r1.w;
        R.mr0 = r2.x; R.mr1 = r2.y; R.pad0 = R.pad1 = 0;
    }
    const M2FwdBlob B = m2_fwd_blob(P, R);
    const int Mr = P.Mr, MH = P.MH, D = P.D, G4 = P.G4, N = P.N, T = P.T, G = P.G;
    long long *trace = blockIdx.x == 0 ? g_m2_trace : nullptr;
    const int dhm = R.dhm, u0l = R.u0 - R.offm, goff = 4 * R.offm;
    const bool s2 = R.head >= 0;
    const int nch1 = R.nu / 8, nch2 = s2 ? R.nj / 8 : 0;

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 192);
    uint8_t *blob = smem + kM2CtrlBytes;
    uint8_t *act = blob + P.blob_f;
    float *smax = reinterpret_cast<float *>(act + P.act_f), *ssum = smax + 256;
    const float *s_batt = reinterpret_cast<const float *>(blob + B.batt);
    const float *s_b1 = reinterpret_cast<const float *>(blob + B.b1);
    const float *s_bv = reinterpret_cast<const float *>(blob + B.bv);

    // operand buffer views (byte offsets inside `act`)
    const int imgC = D * Mr * 4;                           // c image: hi then lo, D/8 chunks of Mr rows each
    const int offH_lo = (dhm / 8) * Mr * 16, offU = 2 * offH_lo, offU_lo = offU + (MH / 8) * Mr * 16;
    const uint32_t rowb = (uint32_t)Mr * 16;               // chunk stride of every activation image

    uint8_t *wsg = a.ws + (size_t)grp * P.ws_group;
    uint8_t *xc = wsg + P.ws_xc, *xh = wsg + P.ws_xh, *xu = wsg + P.ws_xu;
    float *xp = reinterpret_cast<float *>(wsg + P.ws_xp), *xst = reinterpret_cast<float *>(wsg + P.ws_xst);
    unsigned *barA = a.bars + (size_t)grp * 128, *barB = barA + 32, *barC = barA + 64;

    if (tid == 0) {
        mbar_init(&bar[M2B_W], 1); mbar_init(&bar[M2B_H], 1); mbar_init(&bar[M2B_U], 1); mbar_init(&bar[M2B_C], 1);
        mbar_init(&bar[M2B_G], 1); mbar_init(&bar[M2B_E], 1); mbar_init(&bar[M2B_P], 1); mbar_init(&bar[M2B_B], 1);
        mbar_init(&bar[M2E_A], kM2EpiWarps); mbar_init(&bar[M2E_B], kM2EpiWarps); mbar_init(&bar[M2E_C], kM2EpiWarps);
        mbar_init(&bar[M2E_D], 4);
        mbar_fence_init();
    }
    if (warp == kM2EpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t accG = tmem, accE = tmem + 128, accP = tmem + 256;

    if (tid == 0) {           // resident weights: one pass of bulk copies, never touched again
        const uint8_t *src = a.blob + (size_t)rank * P.blob_f;
        mbar_expect_tx(&bar[M2B_W], (uint32_t)B.total);
        for (int o = 0; o < B.total; o += 32768) bulk_g2s(blob + o, src + o, (uint32_t)min(32768, B.total - o), &bar[M2B_W]);
    }

    // role split OUTSIDE the block loop: the control warpgroup shrinks its register budget once, the epilogue warpgroups grow theirs
    if (warp >= kM2EpiWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kM2RegsCtl));
    for (int blk = grp, wave = 0; blk < P.nblocks; blk += P.ngroups, ++wave) {
        const int n0 = blk * P.DG, rows = min(P.DG, N - n0);
        // zero operand buffer (h_{-1} = u_{-1} = 0)
        for (int i = tid; i < P.act_f / 16; i += kM2Threads) reinterpret_cast<uint4 *>(act)[i] = make_uint4(0, 0, 0, 0);
        proxy_fence_smem();
        __syncthreads();
        if (wave == 0) m2_mbar_wait(&bar[M2B_W], 0);

        if (warp == kM2EpiWarps) {
            // =============================== control warp (one elected lane issues) ===============================
            {
                const bool leader = elect_one();
                const uint32_t act_s = smem_u32(act), blob_s = smem_u32(blob);
                const uint32_t wg_hi = blob_s + B.wg, wg_lo = wg_hi + B.kcg * B.ng * 16, wg_lbo = B.ng * 16;
                const uint32_t wa_hi = blob_s + B.wa, wa_lo = wa_hi + (D / 8) * R.nj * 16, wa_lbo = R.nj * 16;
                const uint32_t w1_hi = blob_s + B.w1, w1_lo = w1_hi + nch2 * MH * 16, w1_lbo = MH * 16;
                const uint32_t idG = m2_idesc(B.ng), idE = m2_idesc(R.nj), idP = m2_idesc(MH);
                const unsigned base = (unsigned)wave * T;        // barrier epochs are monotonic over the whole launch
                if (lane != 0) trace = nullptr;
                for (int t = 0; t < T; ++t) {
                    const uint32_t ph = (uint32_t)((wave * T + t) & 1);
                    const uint32_t ph1 = (uint32_t)((wave * (T - 1) + (t - 1)) & 1);     // barriers used only for t >= 1
                    const int tstep = t;
                    M2_TRACE(0, 0);
                    // ---- gates: U_m h_{t-1} as soon as h has arrived, W2 u_{t-1} after exchange C ----
                    if (t > 0) { m2_mbar_wait(&bar[M2B_H], ph1); tc_fence_after(); }
                    M2_TRACE(0, 1);
                    m2_issue3(leader, accG, act_s, act_s + offH_lo, rowb, wg_hi, wg_lo, wg_lbo, dhm / 16, idG, true);
                    M2_TRACE(0, 2);
                    if (t > 0) {
                        m2_poll(barC, (base + t) * G);
                        M2_TRACE(0, 3);
                        if (leader) {
                            proxy_fence_all();
                            mbar_expect_tx(&bar[M2B_U], (uint32_t)(MH * Mr * 4));
                            bulk_g2s(act + offU, xu, (uint32_t)(MH * Mr * 4), &bar[M2B_U]);
                        }
                        m2_mbar_wait(&bar[M2B_U], ph1);
                        tc_fence_after();
                    }
                    M2_TRACE(0, 4);
                    m2_issue3(leader, accG, act_s + offU, act_s + offU_lo, rowb, wg_hi + (dhm / 8) * wg_lbo, wg_lo + (dhm / 8) * wg_lbo,
                              wg_lbo, MH / 16, idG, false);
                    if (leader) umma_commit(&bar[M2B_G]);
                    M2_TRACE(0, 5);
                    // ---- exchange A: c_t / h_t slices of all ranks ----
                    m2_mbar_wait(&bar[M2E_A], ph);
                    M2_TRACE(0, 6);
                    if (leader) m2_signal(barA);
                    M2_TRACE(0, 7);
                    m2_poll(barA, (base + t + 1) * G);
                    M2_TRACE(0, 8);
                    if (s2) {
                        if (leader) {
                            proxy_fence_all();
                            mbar_expect_tx(&bar[M2B_C], (uint32_t)imgC);
                            bulk_g2s(act, xc, (uint32_t)imgC, &bar[M2B_C]);
                        }
                        m2_mbar_wait(&bar[M2B_C], ph);
                        tc_fence_after();
                        M2_TRACE(0, 9);
                        m2_issue3(leader, accE, act_s, act_s + imgC / 2, rowb, wa_hi, wa_lo, wa_lbo, D / 16, idE, true);
                        if (leader) umma_commit(&bar[M2B_E]);
                        // ---- fused reduce + fc.0 over the own K slice ----
                        m2_mbar_wait(&bar[M2E_B], ph);
                        tc_fence_after();
                        M2_TRACE(0, 10);
                        m2_issue3(leader, accP, act_s + (R.j0 / 8) * rowb, act_s + imgC / 2 + (R.j0 / 8) * rowb, rowb, w1_hi, w1_lo, w1_lbo,
                                  R.nj / 16, idP, true);
                        if (leader) umma_commit(&bar[M2B_P]);
                        m2_mbar_wait(&bar[M2B_P], ph);
                    }
                    M2_TRACE(0, 11);
                    // the operand buffer is free: fetch h_t of the own modality for the next step's gates
                    if (t + 1 < T && leader) {
                        const uint8_t *hsrc = xh + (size_t)(t & 1) * imgC + (size_t)(R.offm / 8) * rowb;
                        proxy_fence_all();
                        mbar_expect_tx(&bar[M2B_H], 2u * (uint32_t)offH_lo);
                        bulk_g2s(act, hsrc, (uint32_t)offH_lo, &bar[M2B_H]);
                        bulk_g2s(act + offH_lo, hsrc + imgC / 2, (uint32_t)offH_lo, &bar[M2B_H]);
                    }
                    // ---- exchange B: partial products + softmax statistics ----
                    if (s2) m2_mbar_wait(&bar[M2E_C], ph);
                    M2_TRACE(0, 12);
                    if (leader) m2_signal(barB);
                    m2_poll(barB, (base + t + 1) * G);
                    M2_TRACE(0, 13);
                    if (leader) mbar_arrive(&bar[M2B_B]);
                    // ---- exchange C: u_t slices ----
                    m2_mbar_wait(&bar[M2E_D], ph);
                    M2_TRACE(0, 14);
                    if (leader) m2_signal(barC);
                    M2_TRACE(0, 15);
                }
            }
        }
        __syncthreads();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kM2RegsEpi));
    for (int blk = grp, wave = 0; blk < P.nblocks; blk += P.ngroups, ++wave) {
        const int n0 = blk * P.DG, rows = min(P.DG, N - n0);
        // zero operand buffer (h_{-1} = u_{-1} = 0)
        for (int i = tid; i < P.act_f / 16; i += kM2Threads) reinterpret_cast<uint4 *>(act)[i] = make_uint4(0, 0, 0, 0);
        proxy_fence_smem();
        __syncthreads();
        if (wave == 0) m2_mbar_wait(&bar[M2B_W], 0);

        {
            // =============================== epilogue warps ===============================
            const int q = warp & 3, hh = warp >> 2, row = 32 * q + lane;
            const bool rv = row < rows;
            const uint32_t lane_base = (uint32_t)(32 * q) << 16;
            float cprev[kM2CPH][8];
#pragma unroll
            for (int ci = 0; ci < kM2CPH; ++ci)
#pragma unroll
                for (int i = 0; i < 8; ++i) cprev[ci][i] = 0.f;
            const bool stash = a.sC != nullptr;
            // private stash addressing (floats): piece-major inside the block
            const size_t pvb = (size_t)P.nblocks * Mr;                  // padded rows per step
            auto priv = [&](int width, int tt, int col) { return ((size_t)tt * pvb * width) + ((size_t)blk * (width / 4) + col / 4) * Mr * 4; };
            // combine role (warps 0-3): dialogue dd of this rank's share (fastest over lanes), piece pc (4 of the MH outputs)
            const int cdd = tid & 7, cpc = (tid >> 3) & 15, cdia = rank * P.cd + cdd;
            const bool comb = tid < 128 && cdd < P.cd && cdia < rows;

            for (int t = 0; t < T; ++t) {
                const uint32_t ph = (uint32_t)((wave * T + t) & 1);
                const size_t tn = (size_t)t * N + n0 + row;
                const float bvon = t > 0 ? 1.f : 0.f;
                const int tstep = t;
                m2_launder(xp); m2_launder(xst); m2_launder(xc); m2_launder(xh); m2_launder(xu);
                long long *trace_ct = trace;
                if (tid != 0) trace = nullptr;
                M2_TRACE(1, 0);
                // ---- gate pre-activations of the hoisted W x (+ biases): the first chunk is fetched before the wait, and the
                //      lines of the next step are pulled into L2 now ----
                float gxr[32];
                auto load_gx = [&](int c) {
                    const float *g0 = a.gx + tn * G4 + goff + u0l + 8 * c;
#pragma unroll
                    for (int gate = 0; gate < 4; ++gate) {
                        float v8[8];
                        ldg8(g0 + gate * dhm, v8);
#pragma unroll
                        for (int i = 0; i < 8; ++i) gxr[gate * 8 + i] = v8[i];
                    }
                };
                if (rv) {
                    if (hh < nch1) load_gx(hh);
                    if (t + 1 < T)
                        for (int c = hh; c < nch1; c += 2)
#pragma unroll
                            for (int gate = 0; gate < 4; ++gate) prefetch_l2(a.gx + (tn + N) * G4 + goff + u0l + 8 * c + gate * dhm);
                }
                float4 mk0 = make_float4(1.f, 1.f, 1.f, 1.f);
                if (comb && a.mask != nullptr) mk0 = __ldg(reinterpret_cast<const float4 *>(a.mask + ((size_t)t * N + n0 + cdia) * MH + 4 * cpc));
                // ================= epilogue 1: LSTHM cell update of the own hidden units =================
                m2_mbar_wait(&bar[M2B_G], ph);
                tc_fence_after();
                M2_TRACE(1, 1);
#pragma unroll
                for (int ci = 0; ci < kM2CPH; ++ci) {
                    const int c = hh + 2 * ci;
                    if (c < nch1) {                                   // warp-uniform
                        uint32_t v[32];
                        tmem_ld32(accG + lane_base + 32 * c, v);
                        tmem_ld_wait();
                        if (rv) {
                            if (ci > 0) load_gx(c);
                            float hn[8], cn[8], gf[8], gi[8], go[8], gg[8];
#pragma unroll
                            for (int ul = 0; ul < 8; ++ul) {
                                const float *bvp = s_bv + 4 * (8 * c + ul);
                                const float f = sigmoidf_(__uint_as_float(v[4 * ul + 0]) + gxr[ul] + bvon * bvp[0]);
                                const float ig = sigmoidf_(__uint_as_float(v[4 * ul + 1]) + gxr[8 + ul] + bvon * bvp[1]);
                                const float og = sigmoidf_(__uint_as_float(v[4 * ul + 2]) + gxr[16 + ul] + bvon * bvp[2]);
                                const float g = tanhf_(__uint_as_float(v[4 * ul + 3]) + gxr[24 + ul] + bvon * bvp[3]);
                                const float cc = f * cprev[ci][ul] + ig * g;
                                cn[ul] = cc;
                                hn[ul] = tanhf_(cc) * og;
                                cprev[ci][ul] = cc;
                                gf[ul] = f; gi[ul] = ig; go[ul] = og; gg[ul] = g;
                            }
                            const int ug = R.u0 + 8 * c;               // global unit index of the chunk
                            stg8(a.hz + tn * 2 * D + ug, hn);
                            if (stash) {
                                stg8(a.sC + tn * D + ug, cn);
                                float *cq = a.sCp + priv(D, t, ug) + row * 4;
                                *reinterpret_cast<float4 *>(cq) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                                *reinterpret_cast<float4 *>(cq + Mr * 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                                float *gp = a.sG + priv(G4, t, goff + u0l + 8 * c) + row * 4;
                                const size_t gst = (size_t)(dhm / 4) * Mr * 4;       // one gate further = dhm columns
                                *reinterpret_cast<float4 *>(gp) = make_float4(gf[0], gf[1], gf[2], gf[3]);
                                *reinterpret_cast<float4 *>(gp + Mr * 4) = make_float4(gf[4], gf[5], gf[6], gf[7]);
                                *reinterpret_cast<float4 *>(gp + gst) = make_float4(gi[0], gi[1], gi[2], gi[3]);
                                *reinterpret_cast<float4 *>(gp + gst + Mr * 4) = make_float4(gi[4], gi[5], gi[6], gi[7]);
                                *reinterpret_cast<float4 *>(gp + 2 * gst) = make_float4(go[0], go[1], go[2], go[3]);
                                *reinterpret_cast<float4 *>(gp + 2 * gst + Mr * 4) = make_float4(go[4], go[5], go[6], go[7]);
                                *reinterpret_cast<float4 *>(gp + 3 * gst) = make_float4(gg[0], gg[1], gg[2], gg[3]);
                                *reinterpret_cast<float4 *>(gp + 3 * gst + Mr * 4) = make_float4(gg[4], gg[5], gg[6], gg[7]);
                            }
                            // exchange A: the producer splits once, every consumer bulk-copies the operand image
                            uint4 hi, lo;
                            const size_t xo = ((size_t)(ug / 8) * Mr + row) * 16;
                            m2_split8(cn, hi, lo);
                            *reinterpret_cast<uint4 *>(xc + xo) = hi;
                            *reinterpret_cast<uint4 *>(xc + imgC / 2 + xo) = lo;
                            m2_split8(hn, hi, lo);
                            uint8_t *xhb = xh + (size_t)(t & 1) * imgC;
                            *reinterpret_cast<uint4 *>(xhb + xo) = hi;
                            *reinterpret_cast<uint4 *>(xhb + imgC / 2 + xo) = lo;
                        }
                    }
                }
                M2_TRACE(1, 2);
                // (exchange images are read by the peers' bulk copies: their control thread orders its acquire against the
                // async proxy with fence.proxy.async before issuing the copy)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[M2E_A]);
                M2_TRACE(1, 3);

                if (s2) {
                    // ================= epilogue 2: logits -> local softmax statistics -> attended operand =================
                    m2_mbar_wait(&bar[M2B_C], ph);           // the gathered c image is visible to this thread
                    m2_mbar_wait(&bar[M2B_E], ph);
                    tc_fence_after();
                    M2_TRACE(1, 4);
                    // pass 1: the row's maximum over the own chunks.  Pass 2 re-reads the logits from TMEM (nothing is kept in
                    // registers across the exchange of the maxima) and writes the attended operand IN PLACE over the c chunks it
                    // was computed from: thread (row, chunk) is the only reader of that hi/lo pair, so no other thread's input is
                    // overwritten and the W1 product's A operand is simply the own-slice window of the c image.
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int c2 = hh; c2 < nch2; c2 += 2) {               // warp-uniform
                        uint32_t v[8];
                        tmem_ld8(accE + lane_base + 8 * c2, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) mx = fmaxf(mx, __uint_as_float(v[i]) + s_batt[8 * c2 + i]);
                    }
                    smax[hh * 128 + row] = rv ? mx : -INFINITY;
                    M2_TRACE(1, 5);
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    M2_TRACE(1, 6);
                    const float mfin = fmaxf(smax[row], smax[128 + row]);
                    float sum = 0.f;
#pragma unroll 1
                    for (int c2 = hh; c2 < nch2; c2 += 2) {               // warp-uniform
                        uint32_t v[8];
                        tmem_ld8(accE + lane_base + 8 * c2, v);
                        tmem_ld_wait();
                        if (rv) {
                            const size_t co = ((size_t)(R.j0 / 8 + c2) * Mr + row) * 16;
                            float cv[8], ev[8], at[8];
                            m2_join8(*reinterpret_cast<const uint4 *>(act + co), *reinterpret_cast<const uint4 *>(act + imgC / 2 + co), cv);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                ev[i] = __uint_as_float(v[i]) + s_batt[8 * c2 + i];
                                const float p = __expf(ev[i] - mfin);
                                sum += p;
                                at[i] = p * cv[i];
                            }
                            uint4 hi, lo;
                            m2_split8(at, hi, lo);
                            *reinterpret_cast<uint4 *>(act + co) = hi;
                            *reinterpret_cast<uint4 *>(act + imgC / 2 + co) = lo;
                            if (stash) {
                                float *ep = a.sE + priv(G4, t, R.head * D + R.j0 + 8 * c2) + row * 4;
                                *reinterpret_cast<float4 *>(ep) = make_float4(ev[0], ev[1], ev[2], ev[3]);
                                *reinterpret_cast<float4 *>(ep + Mr * 4) = make_float4(ev[4], ev[5], ev[6], ev[7]);
                            }
                        }
                    }
                    ssum[hh * 128 + row] = sum;
                    proxy_fence_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[M2E_B]);
                    M2_TRACE(1, 7);
                    // ================= epilogue 3: partial W1 product + statistics to the group =================
                    m2_mbar_wait(&bar[M2B_P], ph);
                    tc_fence_after();
                    M2_TRACE(1, 8);
                    {
                        uint32_t v[32];
                        tmem_ld32(accP + lane_base + 32 * hh, v);
                        tmem_ld_wait();
                        if (rv) {
                            // piece-major [rank][MH/4][Mr][4]: a warp's store of one piece is 512 contiguous bytes
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                *reinterpret_cast<float4 *>(xp + (((size_t)rank * (MH / 4) + 8 * hh + i) * Mr + row) * 4) =
                                    make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                                __uint_as_float(v[4 * i + 3]));
                            if (hh == 0)
                                *reinterpret_cast<float2 *>(xst + ((size_t)rank * Mr + row) * 2) =
                                    make_float2(fmaxf(smax[row], smax[128 + row]), ssum[row] + ssum[128 + row]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[M2E_C]);
                    M2_TRACE(1, 9);
                }
                // ================= combine (warps 0-3): u_t of this rank's share of the dialogues =================
                // thread = (dialogue, piece of 4 outputs); every load of the reduction is in flight before the first use
                if (warp < 4) {
                    m2_mbar_wait(&bar[M2B_B], ph);
                    M2_TRACE(1, 10);
                    float u4[4] = {0.f, 0.f, 0.f, 0.f};
                    if (comb) {
                        const int nr = P.nr;
                        const size_t tnc = (size_t)t * N + n0 + cdia;
                        float2 ms[kHeads][4];
                        float4 pp[kHeads][4];
#pragma unroll
                        for (int k = 0; k < kHeads; ++k)
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (i < nr) {
                                    const int r = k * nr + i;
                                    ms[k][i] = __ldcg(reinterpret_cast<const float2 *>(xst + ((size_t)r * Mr + cdia) * 2));
                                    pp[k][i] = ldcg4(xp + (((size_t)r * (MH / 4) + cpc) * Mr + cdia) * 4);
                                }
                        M2_TRACE(1, 12);
                        const float4 b1v = *reinterpret_cast<const float4 *>(s_b1 + 4 * cpc);
                        u4[0] = b1v.x; u4[1] = b1v.y; u4[2] = b1v.z; u4[3] = b1v.w;
#pragma unroll
                        for (int k = 0; k < kHeads; ++k) {
                            float Mk = -INFINITY;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (i < nr) Mk = fmaxf(Mk, ms[k][i].x);
                            float S = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (i < nr) {
                                    const float w = __expf(ms[k][i].x - Mk);
                                    S += ms[k][i].y * w;
                                    acc[0] += w * pp[k][i].x; acc[1] += w * pp[k][i].y; acc[2] += w * pp[k][i].z; acc[3] += w * pp[k][i].w;
                                }
                            const float inv = 1.0f / S;
#pragma unroll
                            for (int i = 0; i < 4; ++i) { acc[i] *= inv; u4[i] += acc[i]; }
                            if (stash) {
                                *reinterpret_cast<float4 *>(a.sP + priv(kHeads * MH, t, k * MH + 4 * cpc) + cdia * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                                if (cpc == 0) *reinterpret_cast<float2 *>(a.sMS + (((size_t)t * P.nblocks + blk) * kHeads + k) * Mr * 2 + cdia * 2) = make_float2(Mk, inv);
                            }
                        }
                        M2_TRACE(1, 13);
                        u4[0] = fmaxf(u4[0], 0.f) * mk0.x; u4[1] = fmaxf(u4[1], 0.f) * mk0.y;
                        u4[2] = fmaxf(u4[2], 0.f) * mk0.z; u4[3] = fmaxf(u4[3], 0.f) * mk0.w;
                        *reinterpret_cast<float4 *>(a.sU + tnc * MH + 4 * cpc) = make_float4(u4[0], u4[1], u4[2], u4[3]);
                    }
                    // operand image: the even / odd piece of a chunk sit 8 lanes apart; the even piece's lane stores the hi chunk,
                    // the odd one the lo chunk
                    {
                        const uint32_t h0 = pack_bf16(u4[0], u4[1]), h1 = pack_bf16(u4[2], u4[3]);
                        const uint32_t l0 = pack_bf16(u4[0] - __uint_as_float(h0 << 16), u4[1] - __uint_as_float(h0 & 0xffff0000u));
                        const uint32_t l1 = pack_bf16(u4[2] - __uint_as_float(h1 << 16), u4[3] - __uint_as_float(h1 & 0xffff0000u));
                        const uint32_t oh0 = __shfl_xor_sync(0xffffffffu, h0, 8), oh1 = __shfl_xor_sync(0xffffffffu, h1, 8);
                        const uint32_t ol0 = __shfl_xor_sync(0xffffffffu, l0, 8), ol1 = __shfl_xor_sync(0xffffffffu, l1, 8);
                        if (comb) {
                            const size_t uo = ((size_t)(cpc >> 1) * Mr + cdia) * 16;
                            if ((cpc & 1) == 0) *reinterpret_cast<uint4 *>(xu + uo) = make_uint4(h0, h1, oh0, oh1);
                            else *reinterpret_cast<uint4 *>(xu + (size_t)(MH / 8) * Mr * 16 + uo) = make_uint4(ol0, ol1, l0, l1);
                        }
                    }
                    M2_TRACE(1, 14);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[M2E_D]);
                    M2_TRACE(1, 11);
                } else {
                    m2_mbar_wait(&bar[M2E_D], ph);     // the next step's feature loads stay out of the combine's way
                }
                trace = trace_ct;
            }
        }
        __syncthreads();
    }
    }
    if (warp == kM2EpiWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}


// ---------------------------------------------------------------------------------------------
// backward (BPTT).  Same groups, same ownership; the adjoint chain of a step is
//   dup_t (replicated)  -> d attended slice = dup . W1[:, slice]  -> softmax backward (de slice, direct term)
//   -> dc partial = de_slice . Watt[slice, :]  (exchange X2: reduce over the ranks, consumed by the unit owners)
//   -> cell backward of the own units (ds)  -> [du | dh_m] partials = ds_own . [W2 | U_m][own rows, :]
//   (exchange X1a: dh_m reduced by the unit owners, du reduce-scattered over dialogues and turned into dup_{t-1};
//    exchange X1b: dup_{t-1} operand image + the softmax-backward dots  <dup, P_k>  all-gathered).
// The per-head dot  sum_j a_kj dv_kj c_j  equals  <dup, P_k>  with P_k = W1[:, head k] . attended_k stashed by the forward,
// so the softmax backward needs no cross-rank reduction of its own.
// ---------------------------------------------------------------------------------------------
struct M2BwdArgs {
    M2Plan P;
    const uint8_t *blob;          // [G][blob_b]
    const M2Rank *ranktab;        // [G]
    const float *dhz, *duz, *mask, *sCp, *sG, *sE, *sMS, *sP, *sU;
    float *dgx, *de, *dup, *att;
    uint8_t *ws;
    unsigned *bars;
};

enum { B2_W = 0, B2_DUP, B2_DV, B2_DC, B2_F, B2_X2, B2_X1A, E2_B, E2_D, E2_S, E2_F, E2_CMB, B2_NBAR };

__global__ void __launch_bounds__(kM2Threads, 1) mab_bwd_kernel(const __grid_constant__ M2BwdArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const M2Plan &P = a.P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = blockIdx.x % P.G, grp = blockIdx.x / P.G;
    M2Rank R;                     // from global memory: indexing the parameter struct by rank would put a copy of it on the stack
    {
        const int4 *rp = reinterpret_cast<const int4 *>(a.ranktab + rank);
        const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
        R.m = r0.x; R.u0 = r0.y; R.nu = r0.z; R.head = r0.w; R.j0 = r1.x; R.nj = r1.y; R.dhm = r1.z; R.offm = r1.w;
        R.mr0 = r2.x; R.mr1 = r2.y; R.pad0 = R.pad1 = 0;
    }
    const M2BwdBlob B = m2_bwd_blob(P, R);
    const int Mr = P.Mr, MH = P.MH, D = P.D, G4 = P.G4, N = P.N, T = P.T, G = P.G;
    long long *trace = blockIdx.x == 0 ? g_m2_trace : nullptr;
    const int dhm = R.dhm, u0l = R.u0 - R.offm, goff = 4 * R.offm;
    const bool s2 = R.head >= 0;
    const int nch1 = R.nu / 8, nch2 = s2 ? R.nj / 8 : 0, ng = B.ng, nF = MH + dhm;
    const int ns2 = 4 * P.nr;                                   // ranks holding a stage-2 slice (0 .. ns2-1)
    const int mr0 = R.mr0, mr1 = R.mr1;                         // ranks of the own modality (contiguous)

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 192);
    uint8_t *blob = smem + kM2CtrlBytes;
    uint8_t *act = blob + P.blob_b;
    const int offDE = MH * Mr * 4, offDS = offDE + R.nj * Mr * 4;
    const uint32_t rowb = (uint32_t)Mr * 16;
    const int imgU = MH * Mr * 4;

    uint8_t *wsg = a.ws + (size_t)grp * P.ws_group;
    float *xdc = reinterpret_cast<float *>(wsg + P.ws_xdc);       // [rank][D/8][Mr][8]  dc partials; then [4][D/8][Mr][8] direct terms
    // (the four heads' direct terms follow as pseudo-ranks G .. G+3 of the same piece-major layout)
    float *xdu = reinterpret_cast<float *>(wsg + P.ws_xdu);       // [rank][Mr][MH]
    float *xdh = reinterpret_cast<float *>(wsg + P.ws_xdh);       // [rank][16][Mr][8]
    uint8_t *xdup = wsg + P.ws_xdup;                              // dup image (hi|lo) then dots [Mr][4]
    float *xdot = reinterpret_cast<float *>(xdup + imgU);
    unsigned *barX2 = a.bars + (size_t)grp * 128, *barX1a = barX2 + 32, *barX1b = barX2 + 64;

    if (tid == 0) {
        mbar_init(&bar[B2_W], 1); mbar_init(&bar[B2_DUP], 1); mbar_init(&bar[B2_DV], 1); mbar_init(&bar[B2_DC], 1);
        mbar_init(&bar[B2_F], 1); mbar_init(&bar[B2_X2], 1); mbar_init(&bar[B2_X1A], 1);
        mbar_init(&bar[E2_B], kM2EpiWarps); mbar_init(&bar[E2_D], kM2EpiWarps); mbar_init(&bar[E2_S], kM2EpiWarps);
        mbar_init(&bar[E2_F], kM2EpiWarps); mbar_init(&bar[E2_CMB], 4);
        mbar_fence_init();
    }
    if (warp == kM2EpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t accDC = tmem, accDV = tmem + 256, accF = tmem + 256;     // DV is dead before F is produced

    if (tid == 0) {
        const uint8_t *src = a.blob + (size_t)rank * P.blob_b;
        mbar_expect_tx(&bar[B2_W], (uint32_t)B.total);
        for (int o = 0; o < B.total; o += 32768) bulk_g2s(blob + o, src + o, (uint32_t)min(32768, B.total - o), &bar[B2_W]);
    }

    // role split OUTSIDE the block loop: the control warpgroup shrinks its register budget once, the epilogue warpgroups grow theirs
    if (warp >= kM2EpiWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kM2RegsCtl));
    for (int blk = grp, wave = 0; blk < P.nblocks; blk += P.ngroups, ++wave) {
        const int n0 = blk * P.DG, rows = min(P.DG, N - n0);
        for (int i = tid; i < P.act_b / 16; i += kM2Threads) reinterpret_cast<uint4 *>(act)[i] = make_uint4(0, 0, 0, 0);
        proxy_fence_smem();
        __syncthreads();
        if (wave == 0) m2_mbar_wait(&bar[B2_W], 0);

        if (warp == kM2EpiWarps) {
            // =============================== control warp (one elected lane issues) ===============================
            {
                const bool leader = elect_one();
                const uint32_t act_s = smem_u32(act), blob_s = smem_u32(blob);
                const uint32_t w1t_hi = blob_s + B.w1t, w1t_lo = w1t_hi + (MH / 8) * R.nj * 16, w1t_lbo = R.nj * 16;
                const uint32_t wat_hi = blob_s + B.wat, wat_lo = wat_hi + nch2 * D * 16, wat_lbo = D * 16;
                const uint32_t wf_hi = blob_s + B.wf, wf_lo = wf_hi + (ng / 8) * nF * 16, wf_lbo = nF * 16;
                const uint32_t idV = m2_idesc(R.nj), idC = m2_idesc(D), idF = m2_idesc(nF);
                const unsigned baseS = (unsigned)wave * T, baseF = (unsigned)wave * (T - 1);
                if (lane != 0) trace = nullptr;
                // the pre-step combine publishes dup_{T-1}
                m2_mbar_wait(&bar[E2_CMB], (uint32_t)((wave * T) & 1));
                if (leader) m2_signal(barX1b);
                for (int t = T - 1, s = 0; t >= 0; --t, ++s) {         // s = steps done in this wave
                    const uint32_t ph = (uint32_t)((wave * T + s) & 1);
                    const uint32_t phF = (uint32_t)((wave * (T - 1) + s) & 1);      // barriers skipped at t == 0
                    const int tstep = s;
                    M2_TRACE(0, 0);
                    if (leader && t > 0) {
                        // the private stash of step t-1 is contiguous per (column range, block): pull this rank's slices into L2
                        // one step ahead of their use (they are first touches from HBM otherwise, on the critical path)
                        const size_t pvb = (size_t)P.nblocks * Mr, tb = (size_t)(t - 1);
                        const uint32_t piece = (uint32_t)Mr * 16;
                        const float *g0 = a.sG + tb * pvb * G4 + ((size_t)blk * (G4 / 4) + (goff + u0l) / 4) * Mr * 4;
#pragma unroll
                        for (int gate = 0; gate < 4; ++gate) bulk_prefetch_l2(g0 + (size_t)gate * (dhm / 4) * Mr * 4, (R.nu / 4) * piece);
                        const float *c0 = a.sCp + tb * pvb * D + ((size_t)blk * (D / 4) + R.u0 / 4) * Mr * 4;
                        bulk_prefetch_l2(c0, (R.nu / 4) * piece);
                        if (t > 1) bulk_prefetch_l2(c0 - pvb * D, (R.nu / 4) * piece);
                        if (s2) {
                            bulk_prefetch_l2(a.sE + tb * pvb * G4 + ((size_t)blk * (G4 / 4) + (R.head * D + R.j0) / 4) * Mr * 4, (R.nj / 4) * piece);
                            bulk_prefetch_l2(a.sCp + tb * pvb * D + ((size_t)blk * (D / 4) + R.j0 / 4) * Mr * 4, (R.nj / 4) * piece);
                            bulk_prefetch_l2(a.sMS + ((tb * P.nblocks + blk) * kHeads + R.head) * Mr * 2, (uint32_t)Mr * 8);
                        }
                        const int d0 = rank * P.cd, nd = min(P.cd, rows - d0);
                        if (nd > 0) {
                            const size_t tnc = tb * N + n0 + d0;
                            bulk_prefetch_l2(a.duz + tnc * MH, (uint32_t)nd * MH * 4);
                            bulk_prefetch_l2(a.sU + tnc * MH, (uint32_t)nd * MH * 4);
                            if (a.mask != nullptr) bulk_prefetch_l2(a.mask + tnc * MH, (uint32_t)nd * MH * 4);
                        }
                    }
                    if (s2) {
                        m2_poll(barX1b, (baseS + s + 1) * G);
                        M2_TRACE(0, 1);
                        if (leader) {
                            proxy_fence_all();
                            mbar_expect_tx(&bar[B2_DUP], (uint32_t)imgU);
                            bulk_g2s(act, xdup, (uint32_t)imgU, &bar[B2_DUP]);
                        }
                        m2_mbar_wait(&bar[B2_DUP], ph);
                        tc_fence_after();
                        M2_TRACE(0, 2);
                        m2_issue3(leader, accDV, act_s, act_s + imgU / 2, rowb, w1t_hi, w1t_lo, w1t_lbo, MH / 16, idV, true);
                        if (leader) umma_commit(&bar[B2_DV]);
                        M2_TRACE(0, 3);
                        m2_mbar_wait(&bar[E2_B], ph);
                        tc_fence_after();
                        M2_TRACE(0, 4);
                        m2_issue3(leader, accDC, act_s + offDE, act_s + offDE + nch2 * Mr * 16, rowb, wat_hi, wat_lo, wat_lbo, R.nj / 16, idC, true);
                        if (leader) umma_commit(&bar[B2_DC]);
                        M2_TRACE(0, 5);
                        m2_mbar_wait(&bar[E2_D], ph);
                    }
                    M2_TRACE(0, 6);
                    if (leader) m2_signal(barX2);
                    m2_poll(barX2, (baseS + s + 1) * G);
                    M2_TRACE(0, 7);
                    if (leader) mbar_arrive(&bar[B2_X2]);
                    m2_mbar_wait(&bar[E2_S], ph);
                    M2_TRACE(0, 8);
                    if (t > 0) {
                        tc_fence_after();
                        m2_issue3(leader, accF, act_s + offDS, act_s + offDS + (ng / 8) * Mr * 16, rowb, wf_hi, wf_lo, wf_lbo, ng / 16, idF, true);
                        if (leader) umma_commit(&bar[B2_F]);
                        M2_TRACE(0, 9);
                        m2_mbar_wait(&bar[E2_F], phF);
                        M2_TRACE(0, 10);
                        if (leader) m2_signal(barX1a);
                        m2_poll(barX1a, (baseF + s + 1) * G);
                        M2_TRACE(0, 11);
                        if (leader) mbar_arrive(&bar[B2_X1A]);
                        m2_mbar_wait(&bar[E2_CMB], (uint32_t)((wave * T + s + 1) & 1));
                        M2_TRACE(0, 12);
                        if (leader) m2_signal(barX1b);
                    }
                }
            }
        }
        __syncthreads();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kM2RegsEpi));
    for (int blk = grp, wave = 0; blk < P.nblocks; blk += P.ngroups, ++wave) {
        const int n0 = blk * P.DG, rows = min(P.DG, N - n0);
        for (int i = tid; i < P.act_b / 16; i += kM2Threads) reinterpret_cast<uint4 *>(act)[i] = make_uint4(0, 0, 0, 0);
        proxy_fence_smem();
        __syncthreads();
        if (wave == 0) m2_mbar_wait(&bar[B2_W], 0);

        {
            // =============================== epilogue warps ===============================
            const int q = warp & 3, hh = warp >> 2, row = 32 * q + lane;
            const bool rv = row < rows;
            const uint32_t lane_base = (uint32_t)(32 * q) << 16;
            float dhc[kM2CPH][8], dcc[kM2CPH][8];                   // carries of the own hidden units
#pragma unroll
            for (int ci = 0; ci < kM2CPH; ++ci)
#pragma unroll
                for (int i = 0; i < 8; ++i) dhc[ci][i] = dcc[ci][i] = 0.f;
            const size_t pvb = (size_t)P.nblocks * Mr;
            auto priv = [&](int width, int tt, int col) { return ((size_t)tt * pvb * width) + ((size_t)blk * (width / 4) + col / 4) * Mr * 4; };
            const int cdd = tid & 7, cpc = (tid >> 3) & 15, cdia = rank * P.cd + cdd;
            const bool comb = tid < 128 && cdd < P.cd && cdia < rows;
            float *s_dot = reinterpret_cast<float *>(act + P.act_b);    // [4 warps][8 dialogues][4 heads] partial dots

            // combine: du (sum of the ranks' partials, fixed order) -> dup_tt, its operand image and the dots <dup_tt, P_k>
            int tstep = 0;
            auto combine = [&](int tt, bool first) {
                // thread = (dialogue, piece of 4 outputs); all partial loads in flight before the first add (fixed rank order)
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
                float dots[kHeads] = {0.f, 0.f, 0.f, 0.f};
                if (comb) {
                    const size_t tnc = (size_t)tt * N + n0 + cdia;
                    M2_TRACE(1, 13);
                    if (!first) {
                        // two batches of eight partial pieces: a batch is in flight together; more live registers would make the
                        // compiler put a spill store behind every load, which serialises them
#pragma unroll
                        for (int b0 = 0; b0 < kM2MaxRanks; b0 += 8) {
                            float4 pr[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (b0 + i < G) pr[i] = ldcg4(xdu + (((size_t)(b0 + i) * (MH / 4) + cpc) * Mr + cdia) * 4);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (b0 + i < G) { s4[0] += pr[i].x; s4[1] += pr[i].y; s4[2] += pr[i].z; s4[3] += pr[i].w; }
                        }
                        M2_TRACE(1, 14);
                    }
                    const float4 z = __ldg(reinterpret_cast<const float4 *>(a.duz + tnc * MH + 4 * cpc));
                    const float4 u = __ldg(reinterpret_cast<const float4 *>(a.sU + tnc * MH + 4 * cpc));
                    float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (a.mask != nullptr) mk = __ldg(reinterpret_cast<const float4 *>(a.mask + tnc * MH + 4 * cpc));
                    float4 pk[kHeads];
#pragma unroll
                    for (int k = 0; k < kHeads; ++k) pk[k] = __ldg(reinterpret_cast<const float4 *>(a.sP + priv(kHeads * MH, tt, k * MH + 4 * cpc) + cdia * 4));
                    s4[0] = (u.x != 0.f ? s4[0] + z.x : 0.f) * mk.x;
                    s4[1] = (u.y != 0.f ? s4[1] + z.y : 0.f) * mk.y;
                    s4[2] = (u.z != 0.f ? s4[2] + z.z : 0.f) * mk.z;
                    s4[3] = (u.w != 0.f ? s4[3] + z.w : 0.f) * mk.w;
                    *reinterpret_cast<float4 *>(a.dup + tnc * MH + 4 * cpc) = make_float4(s4[0], s4[1], s4[2], s4[3]);
#pragma unroll
                    for (int k = 0; k < kHeads; ++k) dots[k] = s4[0] * pk[k].x + s4[1] * pk[k].y + s4[2] * pk[k].z + s4[3] * pk[k].w;
                }
                {
                    const uint32_t h0 = pack_bf16(s4[0], s4[1]), h1 = pack_bf16(s4[2], s4[3]);
                    const uint32_t l0 = pack_bf16(s4[0] - __uint_as_float(h0 << 16), s4[1] - __uint_as_float(h0 & 0xffff0000u));
                    const uint32_t l1 = pack_bf16(s4[2] - __uint_as_float(h1 << 16), s4[3] - __uint_as_float(h1 & 0xffff0000u));
                    const uint32_t oh0 = __shfl_xor_sync(0xffffffffu, h0, 8), oh1 = __shfl_xor_sync(0xffffffffu, h1, 8);
                    const uint32_t ol0 = __shfl_xor_sync(0xffffffffu, l0, 8), ol1 = __shfl_xor_sync(0xffffffffu, l1, 8);
                    if (comb) {
                        const size_t uo = ((size_t)(cpc >> 1) * Mr + cdia) * 16;
                        if ((cpc & 1) == 0) *reinterpret_cast<uint4 *>(xdup + uo) = make_uint4(h0, h1, oh0, oh1);
                        else *reinterpret_cast<uint4 *>(xdup + imgU / 2 + uo) = make_uint4(ol0, ol1, l0, l1);
                    }
                }
                // <dup, P_k>: a warp holds 4 of the 16 pieces of its 8 dialogues (8 and 16 lanes apart); the four warps' partial
                // sums meet in shared memory and are added in fixed warp order
#pragma unroll
                for (int k = 0; k < kHeads; ++k) {
                    dots[k] += __shfl_xor_sync(0xffffffffu, dots[k], 8);
                    dots[k] += __shfl_xor_sync(0xffffffffu, dots[k], 16);
                }
                M2_TRACE(1, 15);
                if (lane < 8) *reinterpret_cast<float4 *>(s_dot + (warp * 8 + lane) * 4) = make_float4(dots[0], dots[1], dots[2], dots[3]);
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (comb && cpc == 0) {
                    float4 d = *reinterpret_cast<const float4 *>(s_dot + cdd * 4);
#pragma unroll
                    for (int w = 1; w < 4; ++w) {
                        const float4 o = *reinterpret_cast<const float4 *>(s_dot + (w * 8 + cdd) * 4);
                        d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w;
                    }
                    *reinterpret_cast<float4 *>(xdot + (size_t)cdia * 4) = d;
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[E2_CMB]);
            };

            if (warp < 4) combine(T - 1, true);

            for (int t = T - 1, s = 0; t >= 0; --t, ++s) {
                const uint32_t ph = (uint32_t)((wave * T + s) & 1);
                const uint32_t phF = (uint32_t)((wave * (T - 1) + s) & 1);
                const size_t tn = (size_t)t * N + n0 + row;
                tstep = s;
                m2_launder(xdc); m2_launder(xdu); m2_launder(xdh); m2_launder(xdot); m2_launder(xdup);
                if (comb && t > 0) {                    // the combine at the end of this step reads P_k of step t-1: first touch from HBM
#pragma unroll
                    for (int k = 0; k < kHeads; ++k) prefetch_l2(a.sP + priv(kHeads * MH, t - 1, k * MH + 4 * cpc) + cdia * 4);
                }
                long long *trace_ct = trace;
                if (tid != 0) trace = nullptr;
                M2_TRACE(1, 0);
                if (s2) {
                    // ================= softmax backward of the own (head, range) slice =================
                    float2 ms = make_float2(0.f, 0.f);
                    if (rv) ms = __ldg(reinterpret_cast<const float2 *>(a.sMS + (((size_t)t * P.nblocks + blk) * kHeads + R.head) * Mr * 2 + row * 2));
                    m2_mbar_wait(&bar[B2_DV], ph);
                    tc_fence_after();
                    M2_TRACE(1, 1);
                    const float dot = rv ? __ldcg(xdot + (size_t)row * 4 + R.head) : 0.f;
#pragma unroll 1
                    for (int c2 = hh; c2 < nch2; c2 += 2) {               // warp-uniform
                        uint32_t v[8];
                        tmem_ld8(accDV + lane_base + 8 * c2, v);
                        tmem_ld_wait();
                        if (rv) {
                            const int j = R.j0 + 8 * c2;          // global feature index of the chunk
                            const float *ep = a.sE + priv(G4, t, R.head * D + j) + row * 4, *cq = a.sCp + priv(D, t, j) + row * 4;
                            const float4 e0 = __ldg(reinterpret_cast<const float4 *>(ep)), e1 = __ldg(reinterpret_cast<const float4 *>(ep + Mr * 4));
                            const float4 c0 = __ldg(reinterpret_cast<const float4 *>(cq)), c1 = __ldg(reinterpret_cast<const float4 *>(cq + Mr * 4));
                            const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w}, cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                            float dev[8], dir[8], atc[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float av = __expf(ev[i] - ms.x) * ms.y, dv = __uint_as_float(v[i]);
                                dev[i] = av * (dv * cv[i] - dot);
                                dir[i] = dv * av;
                                atc[i] = av * cv[i];
                            }
                            stg8(a.de + tn * G4 + R.head * D + j, dev);
                            if (a.att != nullptr) {
                                // modality of feature j (static indexing of the parameter arrays only)
                                int offj = 0, dhj = P.dh[0];
                                if (P.nm > 1 && j >= P.off[1]) { offj = P.off[1]; dhj = P.dh[1]; }
                                if (P.nm > 2 && j >= P.off[2]) { offj = P.off[2]; dhj = P.dh[2]; }
                                stg8(a.att + tn * G4 + 4 * offj + R.head * dhj + (j - offj), atc);
                            }
                            float *xd = xdc + (((size_t)(G + R.head) * (D / 4) + j / 4) * Mr + row) * 4;
                            *reinterpret_cast<float4 *>(xd) = make_float4(dir[0], dir[1], dir[2], dir[3]);
                            *reinterpret_cast<float4 *>(xd + (size_t)Mr * 4) = make_float4(dir[4], dir[5], dir[6], dir[7]);
                            uint4 hi, lo;
                            m2_split8(dev, hi, lo);
                            const size_t ao = ((size_t)c2 * Mr + row) * 16;
                            *reinterpret_cast<uint4 *>(act + offDE + ao) = hi;
                            *reinterpret_cast<uint4 *>(act + offDE + nch2 * Mr * 16 + ao) = lo;
                        }
                    }
                    proxy_fence_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[E2_B]);
                    M2_TRACE(1, 2);
                    // ================= dc partial of the own slice -> group =================
                    m2_mbar_wait(&bar[B2_DC], ph);
                    tc_fence_after();
                    M2_TRACE(1, 3);
                    for (int kc = hh; kc < D / 8; kc += 2) {
                        uint32_t v[8];
                        tmem_ld8(accDC + lane_base + 8 * kc, v);
                        tmem_ld_wait();
                        if (rv) {
                            *reinterpret_cast<float4 *>(xdc + (((size_t)rank * (D / 4) + 2 * kc) * Mr + row) * 4) =
                                make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
                            *reinterpret_cast<float4 *>(xdc + (((size_t)rank * (D / 4) + 2 * kc + 1) * Mr + row) * 4) =
                                make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[E2_D]);
                    M2_TRACE(1, 4);
                }
                // ================= cell backward of the own hidden units =================
                float gf[8], gi[8], go[8], gg[8], cc[8], cp[8], gh[8];
                auto ld8 = [](const float *p, size_t second, float (&x)[8]) {      // two pieces of 4 floats, `second` floats apart
                    const float4 v0 = __ldg(reinterpret_cast<const float4 *>(p)), v1 = __ldg(reinterpret_cast<const float4 *>(p + second));
                    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                };
                auto load_stash = [&](int c) {
                    const int ug = R.u0 + 8 * c;
                    const size_t pst = (size_t)Mr * 4, gst = (size_t)(dhm / 4) * Mr * 4;
                    const float *gp = a.sG + priv(G4, t, goff + u0l + 8 * c) + row * 4;
                    ld8(gp, pst, gf); ld8(gp + gst, pst, gi); ld8(gp + 2 * gst, pst, go); ld8(gp + 3 * gst, pst, gg);
                    ld8(a.sCp + priv(D, t, ug) + row * 4, pst, cc);
                    if (t > 0) ld8(a.sCp + priv(D, t - 1, ug) + row * 4, pst, cp);
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) cp[i] = 0.f;
                    }
                    ldg8(a.dhz + tn * 2 * D + ug, gh);
                };
                if (rv && t > 0)
                    for (int c = hh; c < nch1; c += 2) prefetch_l2(a.dhz + (tn - N) * 2 * D + R.u0 + 8 * c);
                M2_TRACE(1, 5);
                m2_mbar_wait(&bar[B2_X2], ph);
                M2_TRACE(1, 6);
#pragma unroll
                for (int ci = 0; ci < kM2CPH; ++ci) {
                    const int c = hh + 2 * ci;
                    if (c < nch1 && rv) {
                        const int ug = R.u0 + 8 * c;
                        float gc[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) gc[i] = dcc[ci][i];
                        // sources 0 .. ns2-1: the ranks' dc partials; then the four heads' direct terms (they follow xdc in memory
                        // as pseudo-ranks G .. G+3).  Batches of 8 sources: all loads of a batch in flight, adds in fixed order.
#pragma unroll
                        for (int b0 = 0; b0 < kM2MaxRanks + kHeads; b0 += 8) {
                            float4 p0[8], p1[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int sidx = b0 + i;
                                if (sidx < ns2 + kHeads) {
                                    const int src = sidx < ns2 ? sidx : G + (sidx - ns2);
                                    const float *pr = xdc + (((size_t)src * (D / 4) + ug / 4) * Mr + row) * 4;
                                    p0[i] = ldcg4(pr);
                                    p1[i] = ldcg4(pr + (size_t)Mr * 4);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (b0 + i < ns2 + kHeads) {
                                    gc[0] += p0[i].x; gc[1] += p0[i].y; gc[2] += p0[i].z; gc[3] += p0[i].w;
                                    gc[4] += p1[i].x; gc[5] += p1[i].y; gc[6] += p1[i].z; gc[7] += p1[i].w;
                                }
                        }
                        // the stash only now: the partial loads above must not compete with 56 live stash registers (a spill
                        // store behind every load would serialise them); the slices were bulk-prefetched into L2 a step ahead
                        load_stash(c);
                        float ds[4][8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float f = gf[i], ig = gi[i], og = go[i], g = gg[i];
                            const float ghv = gh[i] + dhc[ci][i];
                            const float tc = tanhf_(cc[i]);
                            const float gcj = gc[i] + ghv * og * (1.f - tc * tc);
                            ds[0][i] = gcj * cp[i] * f * (1.f - f);
                            ds[1][i] = gcj * g * ig * (1.f - ig);
                            ds[2][i] = ghv * tc * og * (1.f - og);
                            ds[3][i] = gcj * ig * (1.f - g * g);
                            dcc[ci][i] = gcj * f;
                        }
                        float *dg = a.dgx + tn * G4 + goff + u0l + 8 * c;
#pragma unroll
                        for (int gate = 0; gate < 4; ++gate) stg8(dg + gate * dhm, ds[gate]);
                        // ds operand image, K order = local unit * 4 + gate: K-chunk 4c + i holds units 2i, 2i+1 of the chunk
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float x8[8] = {ds[0][2 * i], ds[1][2 * i], ds[2][2 * i], ds[3][2 * i],
                                                 ds[0][2 * i + 1], ds[1][2 * i + 1], ds[2][2 * i + 1], ds[3][2 * i + 1]};
                            uint4 hi, lo;
                            m2_split8(x8, hi, lo);
                            const size_t so = ((size_t)(4 * c + i) * Mr + row) * 16;
                            *reinterpret_cast<uint4 *>(act + offDS + so) = hi;
                            *reinterpret_cast<uint4 *>(act + offDS + (ng / 8) * Mr * 16 + so) = lo;
                        }
                    }
                }
                proxy_fence_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[E2_S]);
                M2_TRACE(1, 7);
                if (t > 0) {
                    // ================= [du | dh_m] partials -> group =================
                    m2_mbar_wait(&bar[B2_F], phF);
                    tc_fence_after();
                    M2_TRACE(1, 8);
                    {
                        uint32_t v[32];
                        tmem_ld32(accF + lane_base + 32 * hh, v);
                        tmem_ld_wait();
                        if (rv) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                *reinterpret_cast<float4 *>(xdu + (((size_t)rank * (MH / 4) + 8 * hh + i) * Mr + row) * 4) =
                                    make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                                __uint_as_float(v[4 * i + 3]));
                        }
                    }
                    for (int kc = hh; kc < dhm / 8; kc += 2) {
                        uint32_t v[8];
                        tmem_ld8(accF + lane_base + MH + 8 * kc, v);
                        tmem_ld_wait();
                        if (rv) {
                            *reinterpret_cast<float4 *>(xdh + (((size_t)rank * 32 + 2 * kc) * Mr + row) * 4) =
                                make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
                            *reinterpret_cast<float4 *>(xdh + (((size_t)rank * 32 + 2 * kc + 1) * Mr + row) * 4) =
                                make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[E2_F]);
                    M2_TRACE(1, 9);
                    m2_mbar_wait(&bar[B2_X1A], phF);
                    M2_TRACE(1, 10);
                    if (warp < 4) combine(t - 1, false);            // first: the whole group waits for dup_{t-1}
                    else m2_mbar_wait(&bar[E2_CMB], (uint32_t)((wave * T + s + 1) & 1));   // keep the load queue free for the combine
                    M2_TRACE(1, 11);
                    // dh carry of the own units: sum over the ranks of the own modality (fixed order)
#pragma unroll
                    for (int ci = 0; ci < kM2CPH; ++ci) {
                        const int c = hh + 2 * ci;
                        if (c < nch1 && rv) {
                            const int kc = (u0l + 8 * c) / 8;
                            float s8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            // at most 8 ranks per modality (planner); two batches of four: off the critical path (the group is
                            // waiting for dup meanwhile), so a second round trip is cheaper than 64 live registers
#pragma unroll
                            for (int b0 = 0; b0 < 8; b0 += 4) {
                                float4 p0[4], p1[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (mr0 + b0 + i < mr1) {
                                        const float *pr = xdh + (((size_t)(mr0 + b0 + i) * 32 + 2 * kc) * Mr + row) * 4;
                                        p0[i] = ldcg4(pr);
                                        p1[i] = ldcg4(pr + (size_t)Mr * 4);
                                    }
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (mr0 + b0 + i < mr1) {
                                        s8[0] += p0[i].x; s8[1] += p0[i].y; s8[2] += p0[i].z; s8[3] += p0[i].w;
                                        s8[4] += p1[i].x; s8[5] += p1[i].y; s8[6] += p1[i].z; s8[7] += p1[i].w;
                                    }
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) dhc[ci][i] = s8[i];
                        }
                    }
                    M2_TRACE(1, 12);
                }
                trace = trace_ct;
            }
        }
        __syncthreads();
    }
    }
    if (warp == kM2EpiWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

}  // namespace lsthm
