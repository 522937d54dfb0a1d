// LSTHM + multi-attention-block (MAB) recurrence of HybridRNN_AT / HybridRNN_ATV — forward and BPTT.
//
// What the kernels replace in the reference: the body of the time loop of MARN.forward
// (model/HybridRNN_ATV.py:117-143; AT: model/HybridRNN_AT.py:107-132), see include/lsthm_b200.h.
//
// Design (sm_100a, fp32 FFMA path — DESIGN.md §3):
//   * one persistent CTA per tile of MT (<=8) dialogues walks all T steps; grid = ceil(N/MT) ~ one
//     CTA per SM at the headline batch (N=1024 -> MT=7 -> 147 CTAs on 148 SMs);
//   * every stage of a step is a tall-skinny product  out[MT][J] = act[MT][K] . W[K][J]  executed as
//     register-tiled FFMA: a thread owns 4 adjacent columns x MT rows, streams its weight column
//     quad from L2 as float4 (k-major packed image), and reads the activations as shared-memory
//     broadcasts; split-K across thread groups keeps all threads busy, partials are reduced in a
//     fixed order through shared memory (deterministic);
//   * recurrent state (c,h,z / the three adjoint carries) never leaves shared memory between steps;
//   * per-step inputs are staged one step ahead with a 1-D bulk async copy (TMA engine) on an
//     mbarrier (forward) or cp.async/L2 prefetch (backward);
//   * the forward stashes every post-nonlinearity activation so the backward does no recompute:
//     HBM is not the binding resource here, FFMA issue and L2 weight streaming are.
//   * weight gradients are NOT accumulated in the serial chain: the backward emits the per-step
//     adjoints and the host forms  dW = adj^T . act  as time-parallel products over all T*N rows.
#pragma once
#include "common.cuh"

namespace lsthm {


#ifndef LSTHM_MAXT
#define LSTHM_MAXT 448
#endif
constexpr int kMaxThreads = LSTHM_MAXT;  // 13-14 warps: leaves up to 144 registers per thread

struct MabLayout {
    int T, N, nm, MH, D, G, R;
    int dh[kMaxMod], off[kMaxMod], goff[kMaxMod], rd[kMaxMod], roff[kMaxMod];
    // packed weight image (float offsets)
    // packed image offsets (floats).  Composite weights (the chain has no nonlinearity between reduce_m and fc.0, nor
    // between fc.3 and the V term of the next step's gates):  W1 = Wf1 . blockdiag(Wr_m) [MH x 4D],  b1 = Wf1 br + bf1,
    // W2 = Vcat . Wf2 [4D x MH],  bv = Vcat bf2.
    //   wg[m]  [(dh_m + MH)][4 dh_m]  rows: U_m^T then W2_m^T, columns gate-interleaved        (forward gates)
    //   watt   [D][4D]                 Watt^T                                                   (forward logits)
    //   w1     [4D][MH]                W1^T, rows in the attended order k = head*D + j           (forward fused reduce+fc.0)
    //   w1n    [MH][4D]                W1                                                       (backward d attended)
    //   w2n    [4D][MH]                W2, rows in the native gate order                         (backward du carry)
    int wg[kMaxMod], watt, w1, w1n, w2n, batt, b1, bvz, total;
    int nt, nwarp, ldr, ldc, smchunk;
    // split-K plans (forward)
    int s34ns, s34chunk;
    // split-K plans (backward)
    int b4ns, b4chunk, b5uns, b5uchunk;
    int b5ns[kMaxMod], b5chunk[kMaxMod], b5items[kMaxMod], b5total;
};

struct FwdSmem {  // float offsets
    int h, c, km, row, u, red, fin, part, gx, mask, batt, total;
};
struct BwdSmem {
    int dh, du, dc, gh, dup, km, C, A, row, p2, red, fin, dhz, duz, uh, mk, total;
    int b5pb[kMaxMod];
};

struct FwdArgs {
    MabLayout L;
    FwdSmem S;
    const float *packed, *gx, *mask;
    float *hz, *sC, *sG, *sA, *sU;
};
struct BwdArgs {
    MabLayout L;
    BwdSmem S;
    const float *packed;
    const float *U[kMaxMod], *Watt;
    const float *dhz, *duz, *mask, *sC, *sG, *sA, *sU;
    float *dgx, *de, *dup;
    float *att;   // [T][N][G] attended = a * c regrouped per modality, head-major (HybridRNN_ATV.py:125-128): the operand of d reduce_m
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int rup(int a, int b) { return cdiv(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// weight packing: transposes into k-major images, gate-interleaves the LSTHM columns
// ---------------------------------------------------------------------------------------------
struct PackJob {
    const float *src;
    int dst, J, K, ld, row_off, gate_dh;
};
struct PackJobs {
    PackJob j[24];
    int n;
};
struct ComposeArgs {          // inputs of the composite-weight kernel (native nn.Linear layouts)
    MabLayout L;
    const float *V[kMaxMod], *Wr[kMaxMod], *br[kMaxMod], *Wf1, *bf1, *Wf2, *bf2;
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(kMaxThreads, 1) mab_fwd_kernel(const __grid_constant__ FwdArgs a) {
    constexpr int MTP = (MT + 3) & ~3;
    constexpr int JL = 32 / MTP;
    extern __shared__ __align__(16) float smem[];
    const MabLayout &L = a.L;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int D = L.D, G = L.G, MH = L.MH, N = L.N, T = L.T;
    const int n0 = blockIdx.x * MT;
    const int rows = min(MT, N - n0);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float *s_h = smem + a.S.h, *s_c = smem + a.S.c, *s_km = smem + a.S.km;
    float *s_row = smem + a.S.row, *s_u = smem + a.S.u, *s_red = smem + a.S.red;
    float *s_fin = smem + a.S.fin, *s_part = smem + a.S.part, *s_gx = smem + a.S.gx;
    float *s_mask = smem + a.S.mask, *s_batt = smem + a.S.batt;
    const float *__restrict__ packed = a.packed;
    const bool stash = a.sC != nullptr;
    const bool masked = a.mask != nullptr;

    for (int i = 8 + tid; i < a.S.total; i += nt) smem[i] = 0.f;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();
    for (int i = tid; i < G; i += nt) s_batt[i] = __ldg(packed + L.batt + i);
    // per-step tile = this CTA's rows of gx[t] (and of the dropout mask): contiguous in global memory
    const uint32_t gx_bytes = (uint32_t)rows * G * sizeof(float);
    const uint32_t mask_bytes = masked ? (uint32_t)rows * MH * sizeof(float) : 0u;
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, gx_bytes + mask_bytes);
        bulk_g2s(s_gx, a.gx + (size_t)n0 * G, gx_bytes, bar);
        if (masked) bulk_g2s(s_mask, a.mask + (size_t)n0 * MH, mask_bytes, bar);
    }

    // step-invariant role of this thread in the gate stage: (hidden unit j, K-half); K = [h_m (dh_m) ; u (MH)]
    const bool s1_on = tid < 2 * D;
    const int s1_j = tid % D, s1_half = tid / D;
    int m1 = 0;
    while (m1 + 1 < L.nm && s1_j >= L.off[m1 + 1]) ++m1;
    const int s1_dh = L.dh[m1], s1_jl = s1_j - L.off[m1], s1_goff = L.goff[m1];
    const int s1_kh = (s1_dh + MH) / 2, s1_k0 = s1_half * s1_kh, s1_k1 = s1_k0 + s1_kh;
    // bv = Vcat bf2 of this unit's four gates: the constant part of V z_{t-1} = W2 u_{t-1} + bv for t >= 1 (z_{-1} = 0)
    float bv4[4] = {0.f, 0.f, 0.f, 0.f};
    if (s1_on && s1_half == 0) {
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) bv4[g4] = __ldg(packed + L.bvz + s1_goff + g4 * s1_dh + s1_jl);
    }
    const int nq2 = G / 4;
    // softmax lane mapping: lane = (jj, mm); warp w owns features [jb, je)
    const int jj = lane / MTP, mm = lane % MTP;
    const bool mvalid = mm < MT;
    const int jb = warp * L.smchunk, je = min(D, jb + L.smchunk);

    for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        const size_t tn0 = (size_t)t * N + n0;
        const float bvon = t > 0 ? 1.f : 0.f;
        if (tid == 0 && t + 1 < T) {
            mbar_expect_tx(bar + (buf ^ 1), gx_bytes + mask_bytes);
            bulk_g2s(s_gx + (buf ^ 1) * MT * G, a.gx + (tn0 + N) * G, gx_bytes, bar + (buf ^ 1));
            if (masked) bulk_g2s(s_mask + (buf ^ 1) * MT * MH, a.mask + (tn0 + N) * MH, mask_bytes, bar + (buf ^ 1));
        }
        Acc<MT> acc;
        // ---- S1: gate pre-activations  U_m h_{t-1} + W2_m u_{t-1}  (+ gx + bv), then the LSTHM cell update
        if (s1_on) {
            acc.zero();
            const float4 *wp = reinterpret_cast<const float4 *>(packed + L.wg[m1]) + s1_jl;
            const int e1 = min(s1_k1, s1_dh);
            if (s1_k0 < e1)
                mac<MT, MTP>(acc, wp + (size_t)s1_k0 * s1_dh, s1_dh, s_h + (L.off[m1] + s1_k0) * MTP, e1 - s1_k0);
            const int b2 = max(s1_k0, s1_dh);
            if (b2 < s1_k1) mac<MT, MTP>(acc, wp + (size_t)b2 * s1_dh, s1_dh, s_u + (b2 - s1_dh) * MTP, s1_k1 - b2);
            if (s1_half == 1) store_partial<MT, MTP>(s_part, G, 0, s1_goff + 4 * s1_jl, acc);
        }
        __syncthreads();
        if (s1_on && s1_half == 0) {
            mbar_wait(bar + buf, (t >> 1) & 1);
            const float *gxs = s_gx + buf * MT * G + s1_goff + s1_jl;
            float cp[MTP], cn[MTP], hn[MTP];
            load_rows<MTP>(cp, s_c + s1_j * MTP);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float4 pp = *reinterpret_cast<const float4 *>(s_part + m * G + s1_goff + 4 * s1_jl);
                const float *gr = gxs + m * G;
                const float f = sigmoidf_(acc.get(0, m) + pp.x + gr[0] + bvon * bv4[0]);
                const float ig = sigmoidf_(acc.get(1, m) + pp.y + gr[s1_dh] + bvon * bv4[1]);
                const float og = sigmoidf_(acc.get(2, m) + pp.z + gr[2 * s1_dh] + bvon * bv4[2]);
                const float gg = tanhf_(acc.get(3, m) + pp.w + gr[3 * s1_dh] + bvon * bv4[3]);
                const float c = f * cp[m] + ig * gg;
                const float h = tanhf_(c) * og;
                cn[m] = c;
                hn[m] = h;
                if (m < rows) {
                    a.hz[(tn0 + m) * 2 * D + s1_j] = h;
                    if (stash) {
                        a.sC[(tn0 + m) * D + s1_j] = c;
                        float *go = a.sG + (tn0 + m) * G + s1_goff + s1_jl;
                        go[0] = f; go[s1_dh] = ig; go[2 * s1_dh] = og; go[3 * s1_dh] = gg;
                    }
                }
            }
#pragma unroll
            for (int m = MT; m < MTP; ++m) cn[m] = hn[m] = 0.f;
            store_rows<MTP>(s_c + s1_j * MTP, cn);
            store_rows<MTP>(s_h + s1_j * MTP, hn);
        }
        __syncthreads();
        // ---- S2: attention logits e = Watt c + b  (4 heads x D), K split in two halves; each half
        //      leaves its partial in a padded row layout [m][ldr] (half 0 -> s_row, half 1 -> s_part)
        if (tid < 2 * nq2) {
            const int quad = tid % nq2, half = tid / nq2, kh = D / 2;
            acc.zero();
            mac<MT, MTP>(acc, reinterpret_cast<const float4 *>(packed + L.watt) + (size_t)(half * kh) * nq2 + quad, nq2,
                         s_c + half * kh * MTP, kh);
            float *dst = (half ? s_part : s_row) + 4 * quad;
#pragma unroll
            for (int m = 0; m < MT; ++m)
                *reinterpret_cast<float4 *>(dst + m * L.ldr) =
                    make_float4(acc.get(0, m), acc.get(1, m), acc.get(2, m), acc.get(3, m));
        }
        __syncthreads();
        // ---- softmax over the D features per (head, dialogue): e = p0 + p1 + b, per-warp partial (max,sum) ...
#pragma unroll
        for (int k = 0; k < kHeads; ++k) {
            float mx = -INFINITY, sm = 0.f;
            if (mvalid) {
                float *e = s_row + mm * L.ldr + k * D;
                const float *e1 = s_part + mm * L.ldr + k * D, *bb = s_batt + k * D;
                for (int j = jb + jj; j < je; j += JL) {
                    const float x = e[j] + e1[j] + bb[j];
                    e[j] = x;
                    mx = fmaxf(mx, x);
                }
                for (int j = jb + jj; j < je; j += JL) sm += __expf(e[j] - mx);
            }
#pragma unroll
            for (int o = MTP; o < 32; o <<= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sm, o);
                const float M = fmaxf(mx, om);
                sm = (mx == -INFINITY ? 0.f : sm * __expf(mx - M)) + (om == -INFINITY ? 0.f : os * __expf(om - M));
                mx = M;
            }
            if (jj == 0) {
                s_red[((warp * kHeads + k) * MTP + mm) * 2] = mx;
                s_red[((warp * kHeads + k) * MTP + mm) * 2 + 1] = sm;
            }
        }
        __syncthreads();
        // ... combined in fixed warp order ...
        if (tid < kHeads * MTP) {
            float M = -INFINITY, S = 0.f;
            for (int w = 0; w < L.nwarp; ++w) M = fmaxf(M, s_red[(w * kHeads * MTP + tid) * 2]);
            for (int w = 0; w < L.nwarp; ++w) {
                const float mw = s_red[(w * kHeads * MTP + tid) * 2];
                if (mw != -INFINITY) S += s_red[(w * kHeads * MTP + tid) * 2 + 1] * __expf(mw - M);
            }
            s_fin[tid * 2] = M;
            s_fin[tid * 2 + 1] = 1.0f / S;
        }
        __syncthreads();
        // ... and applied: a = softmax, attended = a * c  (k-major, row index head*D + j: the K order of W1)
        if (mvalid) {
#pragma unroll
            for (int k = 0; k < kHeads; ++k) {
                const float M = s_fin[(k * MTP + mm) * 2], inv = s_fin[(k * MTP + mm) * 2 + 1];
                float *e = s_row + mm * L.ldr + k * D;
                for (int j = jb + jj; j < je; j += JL) {
                    const float av = __expf(e[j] - M) * inv;
                    e[j] = av;
                    s_km[(k * D + j) * MTP + mm] = av * s_c[j * MTP + mm];
                }
            }
        }
        __syncthreads();
        // ---- S34: fc hidden pre-activation straight from the attended features, v = W1 att + b1 with
        //      W1 = Wf1 . blockdiag(Wr_m) composed at pack time (reduce_m and fc.0 have no nonlinearity between them);
        //      + A tile copy-out
        if (stash) {
            for (int i = tid; i < rows * nq2; i += nt) {
                const int m = i / nq2, c4 = i - m * nq2;
                reinterpret_cast<float4 *>(a.sA + (tn0 + m) * G)[c4] =
                    *reinterpret_cast<const float4 *>(s_row + m * L.ldr + 4 * c4);
            }
        }
        {
            const int nq = MH / 4, items = nq * L.s34ns;
            for (int item = tid; item < items; item += nt) {
                const int quad = item % nq, sp = item / nq, k0 = sp * L.s34chunk, n = min(G, k0 + L.s34chunk) - k0;
                acc.zero();
                if (n > 0)
                    mac<MT, MTP>(acc, reinterpret_cast<const float4 *>(packed + L.w1) + (size_t)k0 * nq + quad, nq,
                                 s_km + k0 * MTP, n);
                store_partial<MT, MTP>(s_part, MH, sp, 4 * quad, acc);
            }
        }
        __syncthreads();
        // ---- ReLU (+ dropout mask): u_t, the state the next step's gates consume (z_t = fc.3(u_t) is formed by the
        //      host for all steps at once: nothing on the serial path needs it any more)
        // one (hidden unit, dialogue) pair per thread: 13 partials each instead of 7 x 13 on 64 threads
        for (int idx = tid; idx < MH * MTP; idx += nt) {
            const int j = idx % MH, q = idx / MH;
            float s = 0.f;
            if (q < rows) {
                s = __ldg(packed + L.b1 + j);
                for (int sp = 0; sp < L.s34ns; ++sp) s += s_part[(sp * MTP + q) * MH + j];
                s = fmaxf(s, 0.f);
                if (masked) s *= s_mask[buf * MT * MH + q * MH + j];
                a.sU[(tn0 + q) * MH + j] = s;      // always written: the host forms z_t = fc.3(u_t) from it
            }
            s_u[j * MTP + q] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// backward (BPTT).  Carries dh, du, dc live in shared memory across steps.  With the composite weights the adjoint
// chain of a step is:  du_t = duz_t + W2^T ds_{t+1}  ->  (ReLU, mask)  ->  d att = W1^T dup  ->  softmax backward
// ->  dc += Watt^T de  ->  cell backward (ds_t)  ->  carries  du = W2^T ds_t,  dh_m = U_m^T ds_{t,m}.
// duz_t = (dL/dz_t from the head) . Wf2 is formed by the host for all steps at once.
// ---------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(kMaxThreads, 1) mab_bwd_kernel(const __grid_constant__ BwdArgs a) {
    constexpr int MTP = (MT + 3) & ~3;
    constexpr int JL = 32 / MTP;
    extern __shared__ __align__(16) float smem[];
    const MabLayout &L = a.L;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int D = L.D, G = L.G, MH = L.MH, N = L.N, T = L.T;
    const int n0 = blockIdx.x * MT;
    const int rows = min(MT, N - n0);
    float *s_dh = smem + a.S.dh, *s_du = smem + a.S.du, *s_dc = smem + a.S.dc, *s_gh = smem + a.S.gh;
    float *s_dup = smem + a.S.dup, *s_km = smem + a.S.km;
    float *s_C = smem + a.S.C, *s_A = smem + a.S.A, *s_row = smem + a.S.row, *s_p2 = smem + a.S.p2;
    float *s_red = smem + a.S.red, *s_fin = smem + a.S.fin;
    float *s_pA = s_A;  // B4/B5 partials alias the (by then dead) A tile + dvec rows
    const int nq2 = G / 4, nqd = D / 4;
    const float *__restrict__ packed = a.packed;

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float *s_dhz = smem + a.S.dhz, *s_duz = smem + a.S.duz, *s_uh = smem + a.S.uh, *s_mk = smem + a.S.mk;
    const bool masked = a.mask != nullptr;
    for (int i = 8 + tid; i < a.S.total; i += nt) smem[i] = 0.f;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // per-step input tiles (dL/d[h|z], duz, fc hidden, dropout mask) are contiguous rows: bulk-copied one step ahead
    const uint32_t dhz_bytes = (uint32_t)rows * 2 * D * sizeof(float), uh_bytes = (uint32_t)rows * MH * sizeof(float);
    const uint32_t tile_bytes = dhz_bytes + 2 * uh_bytes + (masked ? uh_bytes : 0u);
    auto issue_tiles = [&](int t) {
        const int bf = t & 1;
        const size_t tn = (size_t)t * N + n0;
        mbar_expect_tx(bar + bf, tile_bytes);
        bulk_g2s(s_dhz + bf * MT * 2 * D, a.dhz + tn * 2 * D, dhz_bytes, bar + bf);
        bulk_g2s(s_duz + bf * MT * MH, a.duz + tn * MH, uh_bytes, bar + bf);
        bulk_g2s(s_uh + bf * MT * MH, a.sU + tn * MH, uh_bytes, bar + bf);
        if (masked) bulk_g2s(s_mk + bf * MT * MH, a.mask + tn * MH, uh_bytes, bar + bf);
    };
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_tiles(T - 1);
    }

    const int jj = lane / MTP, mm = lane % MTP;
    const bool mvalid = mm < MT;
    const int jb = warp * L.smchunk, je = min(D, jb + L.smchunk);
    // role in the cell stage: (unit j, group of 4 rows)
    const int c_j = tid % D, c_grp = tid / D;
    int mc = 0;
    while (mc + 1 < L.nm && c_j >= L.off[mc + 1]) ++mc;
    const int c_dh = L.dh[mc], c_jl = c_j - L.off[mc], c_goff = L.goff[mc];

    for (int t = T - 1; t >= 0; --t) {
        const size_t tn0 = (size_t)t * N + n0;
        // ---- stage this step's A and C tiles (padded row layout) and warm L2 for direct reads
        for (int i = tid; i < rows * nq2; i += nt) {
            const int m = i / nq2, c4 = i - m * nq2;
            cp_async16(s_A + m * L.ldr + 4 * c4, a.sA + (tn0 + m) * G + 4 * c4);
        }
        for (int i = tid; i < rows * nqd; i += nt) {
            const int m = i / nqd, c4 = i - m * nqd;
            cp_async16(s_C + m * L.ldc + 4 * c4, a.sC + (tn0 + m) * D + 4 * c4);
        }
        cp_async_commit();
        {
            const char *g = reinterpret_cast<const char *>(a.sG + tn0 * G);
            for (int i = tid * 128; i < rows * G * 4; i += nt * 128) prefetch_l2(g + i);
            if (t > 0) {
                const char *c = reinterpret_cast<const char *>(a.sC + (tn0 - N) * D);
                for (int i = tid * 128; i < rows * D * 4; i += nt * 128) prefetch_l2(c + i);
            }
        }
        const int buf = t & 1;
        if (tid == 0 && t > 0) issue_tiles(t - 1);      // the other slot was last read two barriers ago (step t+1)
        Acc<MT> acc;
        // ---- P0: gh = dL/dh_t + carry;  dup = (duz_t + carry_u) through ReLU and the dropout mask
        if (tid < D + MH) {
            float v[MTP];
#pragma unroll
            for (int q = 0; q < MTP; ++q) v[q] = 0.f;
            mbar_wait(bar + buf, ((T - 1 - t) >> 1) & 1);
            if (tid < D) {
                float carry[MTP];
                load_rows<MTP>(carry, s_dh + tid * MTP);
                const float *dh_s = s_dhz + buf * MT * 2 * D + tid;
#pragma unroll
                for (int q = 0; q < MT; ++q)
                    if (q < rows) v[q] = dh_s[q * 2 * D] + carry[q];
                store_rows<MTP>(s_gh + tid * MTP, v);
            } else {
                const int j = tid - D;
                float carry[MTP];
                load_rows<MTP>(carry, s_du + j * MTP);
#pragma unroll
                for (int q = 0; q < MT; ++q) {
                    if (q < rows) {
                        float s = s_duz[buf * MT * MH + q * MH + j] + carry[q];
                        const float uh = s_uh[buf * MT * MH + q * MH + j];
                        s = (uh != 0.f) ? s : 0.f;
                        if (masked) s *= s_mk[buf * MT * MH + q * MH + j];
                        a.dup[(tn0 + q) * MH + j] = s;
                        v[q] = s;
                    }
                }
                store_rows<MTP>(s_dup + j * MTP, v);
            }
        }
        __syncthreads();
        // ---- B23: d(attended) = W1^T dup  (K = MH in one go: no split, each thread writes its 4 columns of the dvec rows)
        for (int quad = tid; quad < nq2; quad += nt) {
            acc.zero();
            mac<MT, MTP, LSTHM_MAC_BWD>(acc, reinterpret_cast<const float4 *>(packed + L.w1n) + quad, nq2, s_dup, MH);
#pragma unroll
            for (int m = 0; m < MT; ++m)
                *reinterpret_cast<float4 *>(s_row + m * L.ldr + 4 * quad) =
                    make_float4(acc.get(0, m), acc.get(1, m), acc.get(2, m), acc.get(3, m));
        }
        cp_async_wait_all();
        __syncthreads();
        // ---- softmax backward: dot_k = sum_j a*dvec*c  (per head, per dialogue) ...
#pragma unroll
        for (int k = 0; k < kHeads; ++k) {
            float dot = 0.f;
            if (mvalid) {
                const float *av = s_A + mm * L.ldr + k * D, *dv = s_row + mm * L.ldr + k * D, *cv = s_C + mm * L.ldc;
                for (int j = jb + jj; j < je; j += JL) dot += av[j] * dv[j] * cv[j];
            }
#pragma unroll
            for (int o = MTP; o < 32; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (jj == 0) s_red[(warp * kHeads + k) * MTP + mm] = dot;
        }
        __syncthreads();
        if (tid < kHeads * MTP) {
            float s = 0.f;
            for (int w = 0; w < L.nwarp; ++w) s += s_red[w * kHeads * MTP + tid];
            s_fin[tid] = s;
        }
        __syncthreads();
        // ... de = a*(dvec*c - dot) (k-major for B4 + global), direct term dvec*a into the dc carry
        if (mvalid) {
            float dot[kHeads];
#pragma unroll
            for (int k = 0; k < kHeads; ++k) dot[k] = s_fin[k * MTP + mm];
            for (int j = jb + jj; j < je; j += JL) {
                const float cv = s_C[mm * L.ldc + j];
                float direct = 0.f;
                int mj = 0;
                while (mj + 1 < L.nm && j >= L.off[mj + 1]) ++mj;
                const int abase = 4 * L.off[mj] + (j - L.off[mj]), adh = L.dh[mj];   // column of (modality, head 0, feature) in `att`
#pragma unroll
                for (int k = 0; k < kHeads; ++k) {
                    const float av = s_A[mm * L.ldr + k * D + j], dv = s_row[mm * L.ldr + k * D + j];
                    direct += dv * av;
                    const float dev = av * (dv * cv - dot[k]);
                    s_km[(k * D + j) * MTP + mm] = dev;
                    if (mm < rows) {
                        a.de[(tn0 + mm) * G + k * D + j] = dev;
                        if (a.att != nullptr) a.att[(tn0 + mm) * G + abase + k * adh] = av * cv;
                    }
                }
                s_dc[j * MTP + mm] += direct;
            }
        }
        __syncthreads();
        // ---- B4: dc += Watt^T de
        {
            const int items = nqd * L.b4ns;
            for (int item = tid; item < items; item += nt) {
                const int quad = item % nqd, sp = item / nqd, k0 = sp * L.b4chunk, n = min(G, k0 + L.b4chunk) - k0;
                acc.zero();
                if (n > 0)
                    mac<MT, MTP, LSTHM_MAC_BWD>(acc, reinterpret_cast<const float4 *>(a.Watt) + (size_t)k0 * nqd + quad, nqd,
                                 s_km + k0 * MTP, n);
                store_partial<MT, MTP>(s_pA, D, sp, 4 * quad, acc);
            }
        }
        __syncthreads();
        // ---- cell backward: gates from the stash, writes ds (k-major, native gate order) + dgx
        if (c_grp < MTP / 4) {
            const int r0 = c_grp * 4;
            float gc[4], gh[4], dcn[4], ds[4][4];
            {
                const float4 t4 = *reinterpret_cast<const float4 *>(s_dc + c_j * MTP + r0);
                gc[0] = t4.x; gc[1] = t4.y; gc[2] = t4.z; gc[3] = t4.w;
                const float4 h4 = *reinterpret_cast<const float4 *>(s_gh + c_j * MTP + r0);
                gh[0] = h4.x; gh[1] = h4.y; gh[2] = h4.z; gh[3] = h4.w;
            }
            for (int sp = 0; sp < L.b4ns; ++sp)
#pragma unroll
                for (int r = 0; r < 4; ++r) gc[r] += s_pA[(sp * MTP + r0 + r) * D + c_j];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = r0 + r;
                dcn[r] = 0.f;
                ds[0][r] = ds[1][r] = ds[2][r] = ds[3][r] = 0.f;
                if (q < rows) {
                    const float *gp = a.sG + (tn0 + q) * G + c_goff + c_jl;
                    const float f = __ldg(gp), ig = __ldg(gp + c_dh), og = __ldg(gp + 2 * c_dh), gg = __ldg(gp + 3 * c_dh);
                    const float c = s_C[q * L.ldc + c_j];
                    const float cprev = t > 0 ? __ldg(a.sC + (tn0 - N + q) * D + c_j) : 0.f;
                    const float tc = tanhf_(c);
                    const float gcj = gc[r] + gh[r] * og * (1.f - tc * tc);
                    ds[0][r] = gcj * cprev * f * (1.f - f);
                    ds[1][r] = gcj * gg * ig * (1.f - ig);
                    ds[2][r] = gh[r] * tc * og * (1.f - og);
                    ds[3][r] = gcj * ig * (1.f - gg * gg);
                    dcn[r] = gcj * f;
                    float *dg = a.dgx + (tn0 + q) * G + c_goff + c_jl;
                    dg[0] = ds[0][r]; dg[c_dh] = ds[1][r]; dg[2 * c_dh] = ds[2][r]; dg[3 * c_dh] = ds[3][r];
                }
            }
            *reinterpret_cast<float4 *>(s_dc + c_j * MTP + r0) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                *reinterpret_cast<float4 *>(s_km + (c_goff + g * c_dh + c_jl) * MTP + r0) =
                    make_float4(ds[g][0], ds[g][1], ds[g][2], ds[g][3]);
        }
        __syncthreads();
        // ---- B5: carries into step t-1:  du = W2^T ds  (through fc.3 and V in one product),  dh_m = U_m^T ds_m
        {
            const int nq = MH / 4, items = nq * L.b5uns;
            for (int item = tid; item < items; item += nt) {
                const int quad = item % nq, sp = item / nq, k0 = sp * L.b5uchunk, n = min(G, k0 + L.b5uchunk) - k0;
                acc.zero();
                if (n > 0)
                    mac<MT, MTP, LSTHM_MAC_BWD>(acc, reinterpret_cast<const float4 *>(packed + L.w2n) + (size_t)k0 * nq + quad, nq,
                                 s_km + k0 * MTP, n);
                store_partial<MT, MTP>(s_pA, MH, sp, 4 * quad, acc);
            }
            for (int item = tid; item < L.b5total; item += nt) {
                int m = 0, local = item;
                while (local >= L.b5items[m]) { local -= L.b5items[m]; ++m; }
                const int nq5 = L.dh[m] / 4, quad = local % nq5, sp = local / nq5;
                const int k0 = sp * L.b5chunk[m], n = min(4 * L.dh[m], k0 + L.b5chunk[m]) - k0;
                acc.zero();
                if (n > 0)
                    mac<MT, MTP, LSTHM_MAC_BWD>(acc, reinterpret_cast<const float4 *>(a.U[m]) + (size_t)k0 * nq5 + quad, nq5,
                                 s_km + (L.goff[m] + k0) * MTP, n);
                store_partial<MT, MTP>(s_p2 + a.S.b5pb[m], L.dh[m], sp, 4 * quad, acc);
            }
        }
        __syncthreads();
        if (tid < D + MH) {
            float v[MTP];
#pragma unroll
            for (int q = 0; q < MTP; ++q) v[q] = 0.f;
            if (tid >= D) {
                const int j = tid - D;
#pragma unroll
                for (int q = 0; q < MT; ++q) {
                    float s = 0.f;
                    for (int sp = 0; sp < L.b5uns; ++sp) s += s_pA[(sp * MTP + q) * MH + j];
                    v[q] = s;
                }
                store_rows<MTP>(s_du + j * MTP, v);
            } else {
                const float *pb = s_p2 + a.S.b5pb[mc] + c_jl;
                const int ns = L.b5ns[mc];
#pragma unroll
                for (int q = 0; q < MT; ++q) {
                    float s = 0.f;
                    for (int sp = 0; sp < ns; ++sp) s += pb[(sp * MTP + q) * c_dh];
                    v[q] = s;
                }
                store_rows<MTP>(s_dh + c_j * MTP, v);
            }
        }
        __syncthreads();
    }
}


// One translation unit per MT instantiates these (mab_inst.cu, -DLSTHM_MT=n) so the build parallelises.
typedef int (*FwdLaunchFn)(const FwdArgs &, int grid, size_t smem_bytes, cudaStream_t);
typedef int (*BwdLaunchFn)(const BwdArgs &, int grid, size_t smem_bytes, cudaStream_t);
int set_error(const char *what, cudaError_t e);

}  // namespace lsthm
