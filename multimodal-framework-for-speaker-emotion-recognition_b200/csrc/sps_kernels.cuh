// Speaker-state LSTHM cell of lsthm_sps (MARN_cell.forward, model/lsthm_sps.py:156-221) — forward and
// BPTT — as persistent cooperative kernels (sm_100a, fp32).
//
// One step of the reference, in the index form validated against it:
//   pi_t  = dialogues sorted by (current speaker, dialogue id);  n0 = #speaker-0 dialogues
//   x0[r] = r<n0 ? Q[pi(r)][0] : 0          x1[r] = r<n1 ? Q[pi(n0+r)][1] : 0          (lines 177-178, 238-259)
//   h0[d] = Q[pi(d)][d<n0 ? 0 : 1]
//   if n0: (hq0,cq0) = LSTMCell_q0(x0,(hq0,cq0)); hq0 = drop(hq0)   (all N rows; lines 180-185)
//   if n1: (hq1,cq1) = LSTMCell_q1(x1,(hq1,cq1)); hq1 = drop(hq1)   (lines 186-189)
//   hq[d] = d<n0 ? hq0[d] : hq1[d-n0]                                 (lines 191-201)
//   Q[d][p] = h0[d](1-m) + hq[d] m,  m = qmask[t][d][p]               (lines 204-207)  <- packed row d used as dialogue d
//   (cl,hl) = LSTHM1_l(x_l, cl,hl,zl,hq); hl = drop(hl); same for a   (lines 210-213, 28-44)
//   zl = CrossAttention(cl, ca)  -- rank-1, collapsed:  a_i = cl_i (Wq.ca)/sqrt(128),
//        zl_i = sum_j softmax_j(a_i Wk_j) ca_j                        (lines 215, 59-72)
//   out = [hl | ha | zl | hq]                                         (line 218)
//
// The packed-row coupling makes dialogues of a shard interdependent, so a tile-per-CTA kernel needs two
// device-wide exchanges per step (hq1 rows shifted by n0; the Q gather of the next step).  Both go through
// small global buffers with a split arrive/wait grid barrier: the U,V part of the LSTHM gate product runs
// between arrive and wait of the first, the S part + cell update + attention between arrive and wait of
// the second, so the barrier latency is hidden behind independent work.  Cooperative launch guarantees
// co-residency (grid <= #SMs).
#pragma once
#include "common.cuh"

namespace lsthm {

constexpr int kSpsThreads = 512;
constexpr int kU = 128;          // every cell of lsthm_sps is 128 wide (lsthm_sps.py:302-303)
constexpr int kG4 = 4 * kU;      // gates per cell

struct SpsFwdArgs {
    int T, N;
    // packed k-major, gate-interleaved weight images: WQ[c] [256][512] = [Wih^T ; Whh^T], WL[c] [384][512] = [U^T;V^T;S^T]
    const float *wq_img[2], *wl_img[2];
    const float *bq[2];                  // bias_ih + bias_hh, native order i|f|g|o
    const float *Wq, *Wk;                // crossatt_l2a.Wq / Wk  [128]
    const float *gx;                     // [T][N][2][512]  W x + 4 biases, native f|i|o|g per cell (l, a)
    const float *qmask;                  // [T][N][2]
    const int *pi, *n0;                  // [T][N] dialogue at packed row r ; [T] speaker-0 count
    const float *mq[2], *ml, *ma;        // dropout masks [T][N][128] (scaled) or NULL
    const float *att_mask;               // [T][N][128][128] scaled keep mask, or NULL
    float att_p; unsigned long long att_seed;   // in-kernel attention dropout when att_mask==NULL and att_p>0
    float *Q, *XQ;                       // exchange: Q [2][N][2][128], XQ [2][N][128] (hq1 rows)
    unsigned *bar;                       // monotonic grid-barrier counter (zeroed by the host)
    float *out;                          // [T][N][512]
    // stash (all or none)
    float *sGQ, *sCQ, *sHQ, *sXQ;        // LSTM: gates [T][N][2][512] (i|f|g|o), c, h(post-drop), input  [T][N][2][128]
    float *sGL, *sCL, *sHL;              // LSTHM: gates [T][N][2][512] (f|i|o|g), c, h(post-drop)      [T][N][2][128]
    // ---- MODE 1 (GRU speaker state of lsthm_onlysp / lsthm_nsps; dialogues independent, no exchange) ----
    const float *whh_img;                // k-major [128][128 units x (r,z,n,0)] image of gru_s.weight_hh
    const float *bhh;                    // gru_s.bias_hh [384] (r|z|n)
    const float *gxs;                    // [T][N][384]  W_ih U + b_ih, r|z|n
    const float *ms;                     // dropout mask on the speaker state [T][N][128] or NULL
    int listener;                        // 0: q[p] = q[p](1-m_p) + h_s m_p (onlysp) ; 1: q[p] = q[other](1-m_p) + h_s m_p (nsps)
    float *sGS, *sQS;                    // stash: GRU r|z|n|(W_hn q + b_hn) [T][N][4][128], selected party state [T][N][128]
};

struct SpsBwdArgs {
    int T, N;
    const float *U[2], *V[2], *S[2];     // native [512][128]
    const float *Wih[2], *Whh[2];
    const float *Wq, *Wk;
    const float *qmask;
    const int *pi, *pr, *n0;             // pr[t][d] = packed row of dialogue d
    const float *mq[2], *ml, *ma, *att_mask;
    float att_p; unsigned long long att_seed;
    const float *dout;                   // [T][N][512]
    const float *sGQ, *sCQ, *sGL, *sCL;
    float *GX, *GY;                      // exchange: GX [2][N][128] (d hq rows), GY [2][3][N][128] (gx0, gx1, gh0)
    unsigned *bar;
    float *dGL, *dGQ;                    // [T][N][2][512] adjoints of the gate pre-activations
    float *dWqk;                         // [grid][2][128] per-CTA partial sums of dWq, dWk
    // ---- MODE 1 ----
    const float *Whh_s;                  // gru_s.weight_hh native [384][128]
    const float *ms;
    int listener;
    const float *sGS, *sQS;
    float *dGi, *dGh;                    // [T][N][384] adjoints of (W_ih U + b_ih) and of (W_hh q + b_hh)
};

// ---------------------------------------------------------------------------------------------
// split grid barrier on a monotonic counter
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_arrive(unsigned *bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
    }
}
__device__ __forceinline__ void grid_wait(unsigned *bar, unsigned target) {
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// in-cell attention dropout: PairDrop (common.cuh) keyed by (step, dialogue), row i, column pair j/2 — the same
// stream in forward and backward, no mask tensor

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// MODE 0: lsthm_sps (two speaker LSTM cells on packed rows, cooperative launch, two grid exchanges per step)
// MODE 1: lsthm_onlysp / lsthm_nsps (one GRU on the current speaker's party state, per-dialogue independent):
//   idx = argmax(qmask_t[d]);  qs = q[d][idx];  hs = drop(GRUCell(U_t, qs))         (lsthm_onlysp.py:174-179, lsthm_nsps.py:177-183)
//   onlysp: q[d][p] = q[d][p](1-m_p) + hs m_p ;  nsps: q[d][p] = q[d][1-idx](1-m_p) + hs m_p      (:181-182 / :185-188)
//   LSTHM1 cells with speaker input hs, rank-1 attention, out = [hl | ha | zl | hs]  (identical to MODE 0)
template <int MT, int MODE>
__global__ void __launch_bounds__(kSpsThreads, 1) sps_fwd_kernel(const __grid_constant__ SpsFwdArgs a) {
    constexpr int MTP = (MT + 3) & ~3;
    constexpr int VEC = kU * MTP;             // one k-major state vector
    constexpr int LDA = kU + 4;               // row layout stride for the attention operand
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, T = a.T;
    const int r0 = blockIdx.x * MT;
    const int rows = min(MT, N - r0);
    const unsigned G = gridDim.x;
    float *s_hl = smem, *s_cl = s_hl + VEC, *s_ha = s_cl + VEC, *s_ca = s_ha + VEC, *s_zl = s_ca + VEC;
    float *s_hq = s_zl + VEC, *s_hq0 = s_hq + VEC, *s_cq0 = s_hq0 + VEC, *s_hq1 = s_cq0 + VEC, *s_cq1 = s_hq1 + VEC;
    float *s_x0 = s_cq1 + VEC, *s_x1 = s_x0 + VEC, *s_h0 = s_x1 + VEC;
    float *s_part = s_h0 + VEC;               // [MTP][1024] split-K partials (both cells); MODE 1: [4][MTP][512]
    constexpr int PART = (MODE == 0 ? 2 : 4) * MTP * kG4;
    float *s_car = s_part + PART;             // [MT][LDA] ca in row layout
    float *s_wk = s_car + MTP * LDA, *s_wq = s_wk + kU, *s_sm = s_wq + kU;   // s_sm [MTP]
    float *s_bhh = s_sm + MTP;                // MODE 1: gru_s.bias_hh [384]
    const bool stash = a.sGL != nullptr;
    const int total = 13 * VEC + PART + MTP * LDA + 2 * kU + MTP + (MODE == 1 ? 3 * kU : 0);
    for (int i = tid; i < total; i += nt) smem[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < kU; i += nt) { s_wk[i] = __ldg(a.Wk + i); s_wq[i] = __ldg(a.Wq + i); }
    if constexpr (MODE == 1)
        for (int i = tid; i < 3 * kU; i += nt) s_bhh[i] = __ldg(a.bhh + i);
    __syncthreads();
    float wkmax = -INFINITY, wkmin = INFINITY;
    for (int j = 0; j < kU; ++j) { wkmax = fmaxf(wkmax, s_wk[j]); wkmin = fminf(wkmin, s_wk[j]); }

    // step-invariant roles: (cell, unit, K-half) in both gate products
    const int g_cell = tid >> 8, g_unit = tid & 127, g_half = (tid >> 7) & 1;
    unsigned epoch = 0;                       // completed grid barriers (each worth G arrivals)

    for (int t = 0; t < T; ++t) {
        const size_t tn0 = (size_t)t * N + r0;
        Acc<MT> acc;
        if constexpr (MODE == 1) {
            float *s_q0 = s_hq0, *s_q1 = s_hq1, *s_qs = s_x0;      // party states and the current speaker's selection
            // ---- A': current speaker's party state (argmax of an all-zero padded row is party 0)
            for (int idx = tid; idx < MT * kU; idx += nt) {
                const int m = idx >> 7, k = idx & 127;
                float qs = 0.f;
                if (m < rows) {
                    const float m0 = __ldg(a.qmask + (tn0 + m) * 2), m1 = __ldg(a.qmask + (tn0 + m) * 2 + 1);
                    qs = m1 > m0 ? s_q1[k * MTP + m] : s_q0[k * MTP + m];
                }
                s_qs[k * MTP + m] = qs;
            }
            __syncthreads();
            // ---- B': W_hh qs, K split over the four thread groups
            {
                const int ks = tid >> 7;
                acc.zero();
                mac<MT, MTP>(acc, reinterpret_cast<const float4 *>(a.whh_img) + (size_t)(ks * 32) * kU + g_unit, kU,
                             s_qs + ks * 32 * MTP, 32);
                store_partial<MT, MTP>(s_part, kG4, ks, 4 * g_unit, acc);
            }
            __syncthreads();
            // ---- C': GRU gates, speaker state, party update
            for (int idx = tid; idx < MT * kU; idx += nt) {
                const int m = idx >> 7, u = idx & 127;
                float h = 0.f;
                if (m < rows) {
                    float4 p = *reinterpret_cast<const float4 *>(s_part + (size_t)m * kG4 + 4 * u);
#pragma unroll
                    for (int ks = 1; ks < 4; ++ks) {
                        const float4 q = *reinterpret_cast<const float4 *>(s_part + (size_t)(ks * MTP + m) * kG4 + 4 * u);
                        p.x += q.x; p.y += q.y; p.z += q.z;
                    }
                    const float *gi = a.gxs + (tn0 + m) * 3 * kU + u;
                    const float r = sigmoidf_(__ldg(gi) + p.x + s_bhh[u]);
                    const float z = sigmoidf_(__ldg(gi + kU) + p.y + s_bhh[kU + u]);
                    const float hn = p.z + s_bhh[2 * kU + u];
                    const float n = tanhf_(__ldg(gi + 2 * kU) + r * hn);
                    const float qs = s_qs[u * MTP + m];
                    h = (1.f - z) * n + z * qs;
                    if (a.ms) h *= __ldg(a.ms + (tn0 + m) * kU + u);
                    const float m0 = __ldg(a.qmask + (tn0 + m) * 2), m1 = __ldg(a.qmask + (tn0 + m) * 2 + 1);
                    const float q0 = s_q0[u * MTP + m], q1 = s_q1[u * MTP + m];
                    const float b0 = a.listener ? (m1 > m0 ? q0 : q1) : q0, b1 = a.listener ? b0 : q1;
                    s_q0[u * MTP + m] = b0 * (1.f - m0) + h * m0;
                    s_q1[u * MTP + m] = b1 * (1.f - m1) + h * m1;
                    a.out[(tn0 + m) * 4 * kU + 3 * kU + u] = h;
                    if (stash) {
                        float *gs = a.sGS + (tn0 + m) * 4 * kU + u;
                        gs[0] = r; gs[kU] = z; gs[2 * kU] = n; gs[3 * kU] = hn;
                        a.sQS[(tn0 + m) * kU + u] = qs;
                    }
                }
                s_hq[u * MTP + m] = h;
            }
            __syncthreads();
            // ---- D: LSTHM gate products  [h_c | zl | hs] (K = 384, halves of the three blocks per thread group)
            acc.zero();
            {
                const float4 *wp = reinterpret_cast<const float4 *>(a.wl_img[g_cell]) + g_unit;
                const float *act = g_half == 0 ? (g_cell == 0 ? s_hl : s_ha) : s_zl;
                mac<MT, MTP>(acc, wp + (size_t)(g_half * kU) * kU, kU, act, kU);
            }
        }
        const int n0 = MODE == 0 ? __ldg(a.n0 + t) : 0, n1 = N - n0;
        if constexpr (MODE == 0) {
        const int *pit = a.pi + (size_t)t * N;
        const float *Qprev = a.Q + (size_t)((t + 1) & 1) * N * 2 * kU;
        float *Qcur = a.Q + (size_t)(t & 1) * N * 2 * kU;
        // ---- A: gather speaker-LSTM inputs and h0 from Q_{t-1} (complete after the previous step's 2nd barrier)
        if (t > 0) grid_wait(a.bar, epoch * G);
        for (int idx = tid; idx < MT * kU; idx += nt) {
            const int m = idx >> 7, k = idx & 127, r = r0 + m;
            float x0 = 0.f, x1 = 0.f, h0 = 0.f;
            if (t > 0 && m < rows) {
                if (r < n0) x0 = __ldcg(Qprev + ((size_t)pit[r] * 2 + 0) * kU + k);
                if (r < n1) x1 = __ldcg(Qprev + ((size_t)pit[n0 + r] * 2 + 1) * kU + k);
                h0 = r < n0 ? x0 : __ldcg(Qprev + ((size_t)pit[r] * 2 + 1) * kU + k);
            }
            s_x0[k * MTP + m] = x0; s_x1[k * MTP + m] = x1; s_h0[k * MTP + m] = h0;
        }
        __syncthreads();
        // ---- B: the two speaker LSTM cells on packed rows (skipped entirely when no dialogue has that speaker)
        const bool run_q = g_cell == 0 ? n0 > 0 : n1 > 0;
        if (run_q) {
            acc.zero();
            const float4 *wp = reinterpret_cast<const float4 *>(a.wq_img[g_cell]) + g_unit;
            const float *act = g_half == 0 ? (g_cell == 0 ? s_x0 : s_x1) : (g_cell == 0 ? s_hq0 : s_hq1);
            mac<MT, MTP>(acc, wp + (size_t)(g_half * kU) * kU, kU, act, kU);
            if (g_half == 1) store_partial<MT, MTP>(s_part + g_cell * kG4, 2 * kG4, 0, 4 * g_unit, acc);
        }
        __syncthreads();
        if (g_half == 0) {
            float *s_h = g_cell == 0 ? s_hq0 : s_hq1, *s_c = g_cell == 0 ? s_cq0 : s_cq1;
            float hv[MTP], cv[MTP];
            load_rows<MTP>(hv, s_h + g_unit * MTP);
            load_rows<MTP>(cv, s_c + g_unit * MTP);
            const float *bq = a.bq[g_cell];
            const float bi = __ldg(bq + g_unit), bf = __ldg(bq + kU + g_unit), bg = __ldg(bq + 2 * kU + g_unit),
                        bo = __ldg(bq + 3 * kU + g_unit);
            const float *mk = a.mq[g_cell];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
                if (run_q) {
                    const float4 pp = *reinterpret_cast<const float4 *>(s_part + m * 2 * kG4 + g_cell * kG4 + 4 * g_unit);
                    gi = sigmoidf_(acc.get(0, m) + pp.x + bi);
                    gf = sigmoidf_(acc.get(1, m) + pp.y + bf);
                    gg = tanhf_(acc.get(2, m) + pp.z + bg);
                    go = sigmoidf_(acc.get(3, m) + pp.w + bo);
                    const float c = gf * cv[m] + gi * gg;
                    float h = go * tanhf_(c);
                    if (mk && m < rows) h *= __ldg(mk + (tn0 + m) * kU + g_unit);
                    cv[m] = c; hv[m] = h;
                }
                if (m < rows) {
                    if (g_cell == 1) a.XQ[((size_t)(t & 1) * N + r0 + m) * kU + g_unit] = hv[m];
                    if (stash) {
                        const size_t b = (tn0 + m) * 2 + g_cell;
                        float *gq = a.sGQ + b * kG4 + g_unit;
                        gq[0] = gi; gq[kU] = gf; gq[2 * kU] = gg; gq[3 * kU] = go;
                        a.sCQ[b * kU + g_unit] = cv[m];
                        a.sHQ[b * kU + g_unit] = hv[m];
                        a.sXQ[b * kU + g_unit] = (g_cell == 0 ? s_x0 : s_x1)[g_unit * MTP + m];
                    }
                }
            }
            if (run_q) {
                store_rows<MTP>(s_h + g_unit * MTP, hv);
                store_rows<MTP>(s_c + g_unit * MTP, cv);
            }
        }
        grid_arrive(a.bar);                    // hq1 rows of this step are published
        ++epoch;
        // ---- D1: U h + V zl part of the LSTHM gate products (independent of hq) while the barrier completes
        acc.zero();
        {
            const float4 *wp = reinterpret_cast<const float4 *>(a.wl_img[g_cell]) + g_unit;
            const float *act = g_half == 0 ? (g_cell == 0 ? s_hl : s_ha) : s_zl;
            mac<MT, MTP>(acc, wp + (size_t)(g_half * kU) * kU, kU, act, kU);
        }
        grid_wait(a.bar, epoch * G);
        // ---- C: hq for the dialogue rows of this tile, party-state update Q_t
        for (int idx = tid; idx < MT * kU; idx += nt) {
            const int m = idx >> 7, k = idx & 127, d = r0 + m;
            float hq = 0.f;
            if (m < rows) {
                hq = d < n0 ? s_hq0[k * MTP + m]
                            : (n0 == 0 ? s_hq1[k * MTP + m] : __ldcg(a.XQ + ((size_t)(t & 1) * N + d - n0) * kU + k));
                const float h0 = s_h0[k * MTP + m];
                const float m0 = __ldg(a.qmask + (tn0 + m) * 2), m1 = __ldg(a.qmask + (tn0 + m) * 2 + 1);
                Qcur[((size_t)d * 2 + 0) * kU + k] = h0 * (1.f - m0) + hq * m0;
                Qcur[((size_t)d * 2 + 1) * kU + k] = h0 * (1.f - m1) + hq * m1;
                a.out[(tn0 + m) * 4 * kU + 3 * kU + k] = hq;
            }
            s_hq[k * MTP + m] = hq;
        }
        grid_arrive(a.bar);                    // Q_t is published (waited for at the top of step t+1)
        ++epoch;
        }   // MODE 0
        // ---- D2: S hq part (same thread -> same accumulators), K halves of 64
        {
            const float4 *wp = reinterpret_cast<const float4 *>(a.wl_img[g_cell]) + g_unit;
            mac<MT, MTP>(acc, wp + (size_t)(2 * kU + g_half * 64) * kU, kU, s_hq + g_half * 64 * MTP, 64);
            if (g_half == 1) store_partial<MT, MTP>(s_part + g_cell * kG4, 2 * kG4, 0, 4 * g_unit, acc);
        }
        __syncthreads();
        if (g_half == 0) {
            float *s_h = g_cell == 0 ? s_hl : s_ha, *s_c = g_cell == 0 ? s_cl : s_ca;
            float hv[MTP], cv[MTP];
            load_rows<MTP>(cv, s_c + g_unit * MTP);
            const float *mk = g_cell == 0 ? a.ml : a.ma;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                hv[m] = 0.f;
                if (m < rows) {
                    const float4 pp = *reinterpret_cast<const float4 *>(s_part + m * 2 * kG4 + g_cell * kG4 + 4 * g_unit);
                    const float *gx = a.gx + ((tn0 + m) * 2 + g_cell) * kG4 + g_unit;
                    const float gf = sigmoidf_(acc.get(0, m) + pp.x + __ldg(gx));
                    const float gi = sigmoidf_(acc.get(1, m) + pp.y + __ldg(gx + kU));
                    const float go = sigmoidf_(acc.get(2, m) + pp.z + __ldg(gx + 2 * kU));
                    const float gg = tanhf_(acc.get(3, m) + pp.w + __ldg(gx + 3 * kU));
                    const float c = gf * cv[m] + gi * gg;
                    float h = tanhf_(c) * go;
                    if (mk) h *= __ldg(mk + (tn0 + m) * kU + g_unit);
                    cv[m] = c; hv[m] = h;
                    a.out[(tn0 + m) * 4 * kU + g_cell * kU + g_unit] = h;
                    if (stash) {
                        const size_t b = (tn0 + m) * 2 + g_cell;
                        float *gl = a.sGL + b * kG4 + g_unit;
                        gl[0] = gf; gl[kU] = gi; gl[2 * kU] = go; gl[3 * kU] = gg;
                        a.sCL[b * kU + g_unit] = c;
                        if (MODE == 0) a.sHL[b * kU + g_unit] = h;
                    }
                    if (g_cell == 1) s_car[m * LDA + g_unit] = c;
                } else {
                    cv[m] = 0.f;
                }
            }
#pragma unroll
            for (int m = MT; m < MTP; ++m) { hv[m] = 0.f; cv[m] = 0.f; }
            store_rows<MTP>(s_h + g_unit * MTP, hv);
            store_rows<MTP>(s_c + g_unit * MTP, cv);
        }
        __syncthreads();
        // ---- in-cell cross attention, rank-1 form: s = Wq.ca / sqrt(128) per dialogue
        if (warp < MT) {
            float s = 0.f;
            for (int k = lane; k < kU; k += 32) s += s_wq[k] * s_car[warp * LDA + k];
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) s_sm[warp] = s * 0.08838834764831845f;    // 1/sqrt(128)
        }
        __syncthreads();
        for (int pair = tid; pair < MT * kU; pair += nt) {
            const int m = pair >> 7, i = pair & 127;
            float z = 0.f;
            if (m < rows) {
                const float ai = s_cl[i * MTP + m] * s_sm[m];
                const float mx = ai >= 0.f ? ai * wkmax : ai * wkmin;
                float den = 0.f, num = 0.f;
                const float *car = s_car + m * LDA;
                const int n = r0 + m;
                if (a.att_mask) {
                    const float *am = a.att_mask + (((size_t)t * N + n) * kU + i) * kU;
                    for (int j = 0; j < kU; ++j) {
                        const float e = __expf(ai * s_wk[j] - mx);
                        den += e;
                        num = fmaf(e * __ldg(am + j), car[j], num);
                    }
                } else if (a.att_p > 0.f) {
                    const PairDrop pd(a.att_seed, (uint32_t)t * 65536u + (uint32_t)n, a.att_p);
                    for (int j = 0; j < kU; j += 2) {
                        float s0, s1;
                        pd.pair(i, j >> 1, s0, s1);
                        const float e0 = __expf(ai * s_wk[j] - mx), e1 = __expf(ai * s_wk[j + 1] - mx);
                        den += e0; den += e1;
                        num = fmaf(e0 * s0, car[j], num); num = fmaf(e1 * s1, car[j + 1], num);
                    }
                } else {
#pragma unroll 4
                    for (int j = 0; j < kU; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(s_wk + j);
                        const float4 c4 = *reinterpret_cast<const float4 *>(car + j);
                        const float e0 = __expf(ai * w4.x - mx), e1 = __expf(ai * w4.y - mx);
                        const float e2 = __expf(ai * w4.z - mx), e3 = __expf(ai * w4.w - mx);
                        den += (e0 + e1) + (e2 + e3);
                        num = fmaf(e0, c4.x, num); num = fmaf(e1, c4.y, num);
                        num = fmaf(e2, c4.z, num); num = fmaf(e3, c4.w, num);
                    }
                }
                z = num / den;
                a.out[(tn0 + m) * 4 * kU + 2 * kU + i] = z;
            }
            s_zl[i * MTP + m] = z;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// backward (BPTT).  Adjoint carries: Ghl,Gcl,Gha,Gca,Gzl on dialogue rows; Ghq0,Gcq0,Ghq1,Gcq1 on packed
// rows; the adjoint of Q_t arrives from step t+1 through the GY exchange buffers.
// ---------------------------------------------------------------------------------------------
// MODE 1 carries the adjoint of the party state q_t (both parties) in shared memory instead of the packed-row LSTM carries;
// nothing is exchanged between CTAs.
template <int MT, int MODE>
__global__ void __launch_bounds__(kSpsThreads, 1) sps_bwd_kernel(const __grid_constant__ SpsBwdArgs a) {
    constexpr int MTP = (MT + 3) & ~3;
    constexpr int VEC = kU * MTP;
    constexpr int LDA = kU + 4;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, T = a.T;
    const int r0 = blockIdx.x * MT;
    const int rows = min(MT, N - r0);
    const unsigned G = gridDim.x;
    float *s_Ghl = smem, *s_Gcl = s_Ghl + VEC, *s_Gha = s_Gcl + VEC, *s_Gca = s_Gha + VEC, *s_Gzl = s_Gca + VEC;
    float *s_Ghq0 = s_Gzl + VEC, *s_Gcq0 = s_Ghq0 + VEC, *s_Ghq1 = s_Gcq0 + VEC, *s_Gcq1 = s_Ghq1 + VEC;
    float *s_ghq = s_Gcq1 + VEC, *s_gh0 = s_ghq + VEC;
    float *s_ds = s_gh0 + VEC;                 // [2*512][MTP] k-major gate adjoints (LSTHM, then reused for the LSTMs)
    float *s_part = s_ds + 2 * kG4 * MTP;      // split-K partials, 16384 floats
    float *s_row = s_part + 16384;             // 8 row-layout arrays [MT][LDA] for the attention backward
    float *r_cl = s_row, *r_ca = r_cl + MTP * LDA, *r_gz = r_ca + MTP * LDA, *r_ai = r_gz + MTP * LDA;
    float *r_mx = r_ai + MTP * LDA, *r_id = r_mx + MTP * LDA, *r_out = r_id + MTP * LDA, *r_tmp = r_out + MTP * LDA;
    float *s_wk = r_tmp + MTP * LDA, *s_wq = s_wk + kU, *s_dwq = s_wq + kU, *s_dwk = s_dwq + kU;
    float *s_sm = s_dwk + kU, *s_dsm = s_sm + MTP;     // s_sm, s_dsm [MTP]
    const int total = 11 * VEC + 2 * kG4 * MTP + 16384 + 8 * MTP * LDA + 4 * kU + 2 * MTP;
    for (int i = tid; i < total; i += nt) smem[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < kU; i += nt) { s_wk[i] = __ldg(a.Wk + i); s_wq[i] = __ldg(a.Wq + i); }
    __syncthreads();
    float wkmax = -INFINITY, wkmin = INFINITY;
    for (int j = 0; j < kU; ++j) { wkmax = fmaxf(wkmax, s_wk[j]); wkmin = fminf(wkmin, s_wk[j]); }
    const int c_cell = tid >> 8, c_unit = tid & 127, c_grp = (tid >> 7) & 1;   // cell stages: (cell, unit, group of 4 rows)
    unsigned epoch = 0;

    for (int t = T - 1; t >= 0; --t) {
        const int n0 = MODE == 0 ? __ldg(a.n0 + t) : 0, n1 = N - n0;
        const size_t tn0 = (size_t)t * N + r0;
        // ---- P1: stage c_l, c_a (stash) and the incoming dL/dz_l for the attention backward (row layout)
        for (int idx = tid; idx < MT * kU; idx += nt) {
            const int m = idx >> 7, k = idx & 127;
            float cl = 0.f, ca = 0.f, gz = 0.f;
            if (m < rows) {
                cl = __ldg(a.sCL + ((tn0 + m) * 2 + 0) * kU + k);
                ca = __ldg(a.sCL + ((tn0 + m) * 2 + 1) * kU + k);
                gz = __ldg(a.dout + (tn0 + m) * 4 * kU + 2 * kU + k) + s_Gzl[k * MTP + m];
            }
            r_cl[m * LDA + k] = cl; r_ca[m * LDA + k] = ca; r_gz[m * LDA + k] = gz;
        }
        __syncthreads();
        if (warp < MT) {
            float s = 0.f;
            for (int k = lane; k < kU; k += 32) s += s_wq[k] * r_ca[warp * LDA + k];
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) s_sm[warp] = s * 0.08838834764831845f;
        }
        __syncthreads();
        // ---- P2a: per (dialogue, i): recompute the softmax row statistics, g_i = dL/da_i, dc_l
        for (int pair = tid; pair < MT * kU; pair += nt) {
            const int m = pair >> 7, i = pair & 127;
            float gi_s = 0.f;
            if (m < rows) {
                const float x1 = r_cl[m * LDA + i];
                const float ai = x1 * s_sm[m];
                const float mx = ai >= 0.f ? ai * wkmax : ai * wkmin;
                const float *car = r_ca + m * LDA;
                const int n = r0 + m;
                const float *am = a.att_mask ? a.att_mask + (((size_t)t * N + n) * kU + i) * kU : nullptr;
                float den = 0.f, n1s = 0.f, n2s = 0.f, n3s = 0.f;
                const bool hashed = am == nullptr && a.att_p > 0.f;
                const PairDrop pd(a.att_seed, (uint32_t)t * 65536u + (uint32_t)n, a.att_p);
                for (int j = 0; j < kU; j += 2) {
                    float s0 = 1.f, s1 = 1.f;
                    if (am) { s0 = __ldg(am + j); s1 = __ldg(am + j + 1); }
                    else if (hashed) pd.pair(i, j >> 1, s0, s1);
                    const float wk0 = s_wk[j], wk1 = s_wk[j + 1];
                    const float e0 = __expf(ai * wk0 - mx), e1 = __expf(ai * wk1 - mx);
                    const float ex0 = e0 * s0 * car[j], ex1 = e1 * s1 * car[j + 1];
                    den += e0; n1s += ex0; n2s = fmaf(ex0, wk0, n2s); n3s = fmaf(e0, wk0, n3s);
                    den += e1; n1s += ex1; n2s = fmaf(ex1, wk1, n2s); n3s = fmaf(e1, wk1, n3s);
                }
                const float inv = 1.0f / den, out = n1s * inv, go = r_gz[m * LDA + i];
                const float g = go * (n2s - out * n3s) * inv;
                r_ai[m * LDA + i] = ai; r_mx[m * LDA + i] = mx; r_id[m * LDA + i] = inv; r_out[m * LDA + i] = out;
                s_Gcl[i * MTP + m] += g * s_sm[m];          // dc_l through a_i = c_l_i * s
                gi_s = g * x1;
            }
            r_tmp[m * LDA + i] = gi_s;                       // g_i * c_l_i, reduced below into d s
        }
        __syncthreads();
        if (warp < MT) {
            float s = 0.f;
            for (int k = lane; k < kU; k += 32) s += r_tmp[warp * LDA + k];
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) s_dsm[warp] = s * 0.08838834764831845f;   // dL/d(Wq.ca)
        }
        __syncthreads();
        // ---- P2b: per (dialogue, column pair j, j+1): dc_a (direct + through s), per-row dWk contribution.
        //      A thread owns two adjacent columns so one dropout hash (and one set of row constants) serves both.
        for (int pair = tid; pair < MT * (kU / 2); pair += nt) {
            const int m = pair >> 6, j = (pair & 63) * 2;
            float dwk0 = 0.f, dwk1 = 0.f;
            if (m < rows) {
                const float wk0 = s_wk[j], wk1 = s_wk[j + 1], x20 = r_ca[m * LDA + j], x21 = r_ca[m * LDA + j + 1];
                const int n = r0 + m;
                float dca0 = 0.f, dca1 = 0.f;
                const PairDrop pd(a.att_seed, (uint32_t)t * 65536u + (uint32_t)n, a.att_p);
                const bool hashed = a.att_mask == nullptr && a.att_p > 0.f;
                for (int i = 0; i < kU; ++i) {
                    const float ai = r_ai[m * LDA + i], mxi = r_mx[m * LDA + i], idi = r_id[m * LDA + i];
                    const float p0 = __expf(ai * wk0 - mxi) * idi, p1 = __expf(ai * wk1 - mxi) * idi;
                    float s0 = 1.f, s1 = 1.f;
                    if (a.att_mask) {
                        const float2 mk = __ldg(reinterpret_cast<const float2 *>(a.att_mask + (((size_t)t * N + n) * kU + i) * kU + j));
                        s0 = mk.x; s1 = mk.y;
                    } else if (hashed) pd.pair(i, j >> 1, s0, s1);
                    const float go = r_gz[m * LDA + i], oi = r_out[m * LDA + i], ag = ai * go;
                    dca0 = fmaf(go * p0, s0, dca0);
                    dca1 = fmaf(go * p1, s1, dca1);
                    dwk0 = fmaf(p0 * (s0 * x20 - oi), ag, dwk0);
                    dwk1 = fmaf(p1 * (s1 * x21 - oi), ag, dwk1);
                }
                s_Gca[j * MTP + m] += dca0 + s_wq[j] * s_dsm[m];
                s_Gca[(j + 1) * MTP + m] += dca1 + s_wq[j + 1] * s_dsm[m];
            }
            r_tmp[m * LDA + j] = dwk0;
            r_tmp[m * LDA + j + 1] = dwk1;
        }
        __syncthreads();
        if (tid < kU) {                                       // fixed-order reduction over the tile's dialogues
            float wq = 0.f, wk = 0.f;
            for (int m = 0; m < rows; ++m) { wk += r_tmp[m * LDA + tid]; wq += r_ca[m * LDA + tid] * s_dsm[m]; }
            s_dwk[tid] += wk; s_dwq[tid] += wq;
        }
        // ---- P3: LSTHM cell backward (l and a): ds -> shared (k-major, native f|i|o|g) + global dGL
        {
            const int rb = c_grp * 4;
            if (rb < MTP) {
                float *s_Gh = c_cell == 0 ? s_Ghl : s_Gha, *s_Gc = c_cell == 0 ? s_Gcl : s_Gca;
                const float *mk = c_cell == 0 ? a.ml : a.ma;
                float dsv[4][4], gcn[4];
                const float4 gh4 = *reinterpret_cast<const float4 *>(s_Gh + c_unit * MTP + rb);
                const float4 gc4 = *reinterpret_cast<const float4 *>(s_Gc + c_unit * MTP + rb);
                const float ghv[4] = {gh4.x, gh4.y, gh4.z, gh4.w}, gcv[4] = {gc4.x, gc4.y, gc4.z, gc4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int m = rb + r;
                    dsv[0][r] = dsv[1][r] = dsv[2][r] = dsv[3][r] = 0.f; gcn[r] = 0.f;
                    if (m < rows) {
                        const size_t b = (tn0 + m) * 2 + c_cell;
                        const float *gl = a.sGL + b * kG4 + c_unit;
                        const float f = __ldg(gl), ig = __ldg(gl + kU), og = __ldg(gl + 2 * kU), gg = __ldg(gl + 3 * kU);
                        const float c = c_cell == 0 ? r_cl[m * LDA + c_unit] : r_ca[m * LDA + c_unit];
                        const float cprev = t > 0 ? __ldg(a.sCL + (b - (size_t)2 * N) * kU + c_unit) : 0.f;
                        float gh = ghv[r] + __ldg(a.dout + (tn0 + m) * 4 * kU + c_cell * kU + c_unit);
                        if (mk) gh *= __ldg(mk + (tn0 + m) * kU + c_unit);
                        const float tc = tanhf_(c);
                        const float gc = gcv[r] + gh * og * (1.f - tc * tc);
                        dsv[0][r] = gc * cprev * f * (1.f - f);
                        dsv[1][r] = gc * gg * ig * (1.f - ig);
                        dsv[2][r] = gh * tc * og * (1.f - og);
                        dsv[3][r] = gc * ig * (1.f - gg * gg);
                        gcn[r] = gc * f;
                        float *dg = a.dGL + b * kG4 + c_unit;
                        dg[0] = dsv[0][r]; dg[kU] = dsv[1][r]; dg[2 * kU] = dsv[2][r]; dg[3 * kU] = dsv[3][r];
                    }
                }
                *reinterpret_cast<float4 *>(s_Gc + c_unit * MTP + rb) = make_float4(gcn[0], gcn[1], gcn[2], gcn[3]);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<float4 *>(s_ds + (c_cell * kG4 + g * kU + c_unit) * MTP + rb) =
                        make_float4(dsv[g][0], dsv[g][1], dsv[g][2], dsv[g][3]);
            }
        }
        __syncthreads();
        // ---- P4: [Gh_c' | Gzl_c | ghq_c] = ds_c^T [U_c | V_c | S_c]   (K = 512 split in two, J = 3 x 128)
        Acc<MT> acc;
        if (tid < 384) {
            const int cell = tid / 192, rest = tid % 192, which = rest / 64, rem = rest % 64, quad = rem & 31, sp = rem >> 5;
            const float *W = which == 0 ? a.U[cell] : which == 1 ? a.V[cell] : a.S[cell];
            acc.zero();
            mac<MT, MTP>(acc, reinterpret_cast<const float4 *>(W) + (size_t)(sp * 256) * 32 + quad, 32,
                         s_ds + (cell * kG4 + sp * 256) * MTP, 256);
            // partial slot: [(cell*2+sp)][which*128 + col] per row
            store_partial<MT, MTP>(s_part + (cell * 2 + sp) * MTP * 384, 384, 0, which * kU + 4 * quad, acc);
        }
        __syncthreads();
        if constexpr (MODE == 1) {
            float *s_Gq0 = s_Ghq0, *s_Gq1 = s_Ghq1, *s_gqs = s_gh0;
            // ---- P5'/P6': finish the carries; dL/dhs = S parts + dout + party-update adjoint; GRU cell backward (pointwise)
            for (int idx = tid; idx < MT * kU; idx += nt) {
                const int m = idx >> 7, k = idx & 127;
                float ghl = 0.f, gha = 0.f, gzl = 0.f, nq0 = 0.f, nq1 = 0.f, gqs = 0.f, dgr = 0.f, dgz = 0.f, dhn = 0.f;
                if (m < rows) {
                    const float *p00 = s_part + (0 * MTP + m) * 384, *p01 = s_part + (1 * MTP + m) * 384;
                    const float *p10 = s_part + (2 * MTP + m) * 384, *p11 = s_part + (3 * MTP + m) * 384;
                    ghl = p00[k] + p01[k];
                    gha = p10[k] + p11[k];
                    gzl = (p00[kU + k] + p01[kU + k]) + (p10[kU + k] + p11[kU + k]);
                    float ghs = (p00[2 * kU + k] + p01[2 * kU + k]) + (p10[2 * kU + k] + p11[2 * kU + k]);
                    ghs += __ldg(a.dout + (tn0 + m) * 4 * kU + 3 * kU + k);
                    const float m0 = __ldg(a.qmask + (tn0 + m) * 2), m1 = __ldg(a.qmask + (tn0 + m) * 2 + 1);
                    const float G0 = s_Gq0[k * MTP + m], G1 = s_Gq1[k * MTP + m];
                    ghs += G0 * m0 + G1 * m1;
                    if (a.listener) {       // q_t[p] = q_{t-1}[1 - idx](1 - m_p) + hs m_p
                        const float gql = G0 * (1.f - m0) + G1 * (1.f - m1);
                        if (m1 > m0) nq0 = gql; else nq1 = gql;
                    } else {                // q_t[p] = q_{t-1}[p](1 - m_p) + hs m_p
                        nq0 = G0 * (1.f - m0); nq1 = G1 * (1.f - m1);
                    }
                    if (a.ms) ghs *= __ldg(a.ms + (tn0 + m) * kU + k);
                    const float *gs = a.sGS + (tn0 + m) * 4 * kU + k;
                    const float r = __ldg(gs), z = __ldg(gs + kU), n = __ldg(gs + 2 * kU), hn = __ldg(gs + 3 * kU);
                    const float qs = __ldg(a.sQS + (tn0 + m) * kU + k);
                    const float dnp = ghs * (1.f - z) * (1.f - n * n);
                    dgr = dnp * hn * r * (1.f - r);
                    dgz = ghs * (qs - n) * z * (1.f - z);
                    dhn = dnp * r;
                    gqs = ghs * z;
                    float *gi = a.dGi + (tn0 + m) * 3 * kU + k, *gh = a.dGh + (tn0 + m) * 3 * kU + k;
                    gi[0] = dgr; gi[kU] = dgz; gi[2 * kU] = dnp;
                    gh[0] = dgr; gh[kU] = dgz; gh[2 * kU] = dhn;
                }
                s_Ghl[k * MTP + m] = ghl; s_Gha[k * MTP + m] = gha; s_Gzl[k * MTP + m] = gzl;
                s_Gq0[k * MTP + m] = nq0; s_Gq1[k * MTP + m] = nq1; s_gqs[k * MTP + m] = gqs;
                s_ds[(k) * MTP + m] = dgr; s_ds[(kU + k) * MTP + m] = dgz; s_ds[(2 * kU + k) * MTP + m] = dhn;
            }
            __syncthreads();
            // ---- P7': dL/dqs += d(W_hh qs)^T W_hh   (K = 384 in 12 chunks, J = 128)
            if (tid < 384) {
                const int quad = tid & 31, sp = tid >> 5;
                Acc<MT> acc;
                acc.zero();
                mac<MT, MTP>(acc, reinterpret_cast<const float4 *>(a.Whh_s) + (size_t)(sp * 32) * 32 + quad, 32,
                             s_ds + (sp * 32) * MTP, 32);
                store_partial<MT, MTP>(s_part + sp * MTP * kU, kU, 0, 4 * quad, acc);
            }
            __syncthreads();
            // ---- P8': the selected party receives the adjoint of qs
            for (int idx = tid; idx < MT * kU; idx += nt) {
                const int m = idx >> 7, k = idx & 127;
                if (m < rows) {
                    float g = s_gqs[k * MTP + m];
#pragma unroll
                    for (int sp = 0; sp < 12; ++sp) g += s_part[(sp * MTP + m) * kU + k];
                    const float m0 = __ldg(a.qmask + (tn0 + m) * 2), m1 = __ldg(a.qmask + (tn0 + m) * 2 + 1);
                    if (m1 > m0) s_Gq1[k * MTP + m] += g; else s_Gq0[k * MTP + m] += g;
                }
            }
            __syncthreads();
        } else {
        // ---- P5: finish the carries; total dL/dhq; adjoint of Q_t from step t+1 (second exchange of that step)
        if (t < T - 1) grid_wait(a.bar, epoch * G);
        for (int idx = tid; idx < MT * kU; idx += nt) {
            const int m = idx >> 7, k = idx & 127, d = r0 + m;
            float ghl = 0.f, gha = 0.f, gzl = 0.f, ghq = 0.f, gh0 = 0.f;
            if (m < rows) {
                const float *p00 = s_part + (0 * MTP + m) * 384, *p01 = s_part + (1 * MTP + m) * 384;
                const float *p10 = s_part + (2 * MTP + m) * 384, *p11 = s_part + (3 * MTP + m) * 384;
                ghl = p00[k] + p01[k];
                gha = p10[k] + p11[k];
                gzl = (p00[kU + k] + p01[kU + k]) + (p10[kU + k] + p11[kU + k]);
                ghq = (p00[2 * kU + k] + p01[2 * kU + k]) + (p10[2 * kU + k] + p11[2 * kU + k]);
                ghq += __ldg(a.dout + (tn0 + m) * 4 * kU + 3 * kU + k);
                if (t < T - 1) {
                    // Q_t[d][p] was gathered at step t+1 by the packed row pr_{t+1}(d), party = its speaker then
                    const int n0n = __ldg(a.n0 + t + 1);
                    const int rr = __ldg(a.pr + (size_t)(t + 1) * N + d);
                    const float *GYb = a.GY + (size_t)((t + 1) & 1) * 3 * N * kU;
                    const int p = rr < n0n ? 0 : 1;
                    const float gq = (p == 0 ? __ldcg(GYb + (size_t)rr * kU + k)
                                             : __ldcg(GYb + ((size_t)N + rr - n0n) * kU + k)) +
                                     __ldcg(GYb + ((size_t)2 * N + rr) * kU + k);
                    const float mp = __ldg(a.qmask + (tn0 + m) * 2 + p);
                    ghq += gq * mp;
                    gh0 = gq * (1.f - mp);
                }
                a.GX[((size_t)(t & 1) * N + d) * kU + k] = ghq;
            }
            s_Ghl[k * MTP + m] = ghl; s_Gha[k * MTP + m] = gha; s_Gzl[k * MTP + m] = gzl;
            s_ghq[k * MTP + m] = ghq; s_gh0[k * MTP + m] = gh0;
        }
        grid_arrive(a.bar);                     // exchange 1: dL/dhq rows published
        ++epoch;
        grid_wait(a.bar, epoch * G);
        // ---- P6: speaker LSTM cells backward (pointwise): dgates -> shared (k-major, native i|f|g|o) + global dGQ
        {
            const int rb = c_grp * 4;
            const bool run_q = c_cell == 0 ? n0 > 0 : n1 > 0;
            if (rb < MTP) {
                float *s_Gh = c_cell == 0 ? s_Ghq0 : s_Ghq1, *s_Gc = c_cell == 0 ? s_Gcq0 : s_Gcq1;
                const float *mk = a.mq[c_cell];
                float dgv[4][4], gcn[4], ghn[4];
                const float4 gh4 = *reinterpret_cast<const float4 *>(s_Gh + c_unit * MTP + rb);
                const float4 gc4 = *reinterpret_cast<const float4 *>(s_Gc + c_unit * MTP + rb);
                const float ghv[4] = {gh4.x, gh4.y, gh4.z, gh4.w}, gcv[4] = {gc4.x, gc4.y, gc4.z, gc4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int m = rb + r, row = r0 + m;
                    dgv[0][r] = dgv[1][r] = dgv[2][r] = dgv[3][r] = 0.f; gcn[r] = gcv[r]; ghn[r] = ghv[r];
                    if (m < rows) {
                        // dL/dhq reaches this packed row: cell 0 from dialogue row `row` (< n0), cell 1 from n0 + row
                        float gh = ghv[r];
                        if (c_cell == 0) { if (row < n0) gh += s_ghq[c_unit * MTP + m]; }
                        else if (row < n1) gh += n0 == 0 ? s_ghq[c_unit * MTP + m]
                                                         : __ldcg(a.GX + ((size_t)(t & 1) * N + n0 + row) * kU + c_unit);
                        ghn[r] = gh;
                        const size_t b = (tn0 + m) * 2 + c_cell;
                        float *dg = a.dGQ + b * kG4 + c_unit;
                        if (run_q) {
                            const float *gq = a.sGQ + b * kG4 + c_unit;
                            const float ig = __ldg(gq), f = __ldg(gq + kU), gg = __ldg(gq + 2 * kU), og = __ldg(gq + 3 * kU);
                            const float c = __ldg(a.sCQ + b * kU + c_unit);
                            const float cprev = t > 0 ? __ldg(a.sCQ + (b - (size_t)2 * N) * kU + c_unit) : 0.f;
                            if (mk) gh *= __ldg(mk + (tn0 + m) * kU + c_unit);
                            const float tc = tanhf_(c);
                            const float gc = gcv[r] + gh * og * (1.f - tc * tc);
                            dgv[0][r] = gc * gg * ig * (1.f - ig);
                            dgv[1][r] = gc * cprev * f * (1.f - f);
                            dgv[2][r] = gc * ig * (1.f - gg * gg);
                            dgv[3][r] = gh * tc * og * (1.f - og);
                            gcn[r] = gc * f;
                        }
                        dg[0] = dgv[0][r]; dg[kU] = dgv[1][r]; dg[2 * kU] = dgv[2][r]; dg[3 * kU] = dgv[3][r];
                    }
                }
                *reinterpret_cast<float4 *>(s_Gc + c_unit * MTP + rb) = make_float4(gcn[0], gcn[1], gcn[2], gcn[3]);
                // a skipped cell passes its hidden-state adjoint through unchanged; an executed one gets it from P7
                if (!run_q) *reinterpret_cast<float4 *>(s_Gh + c_unit * MTP + rb) = make_float4(ghn[0], ghn[1], ghn[2], ghn[3]);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<float4 *>(s_ds + (c_cell * kG4 + g * kU + c_unit) * MTP + rb) =
                        make_float4(dgv[g][0], dgv[g][1], dgv[g][2], dgv[g][3]);
            }
        }
        __syncthreads();
        // ---- P7: [gx_c | Ghq_c'] = dg_c^T [Wih_c | Whh_c]   (K = 512 in 4 chunks, J = 2 x 128, both cells)
        {
            const int cell = tid >> 8, rest = tid & 255, which = rest >> 7, rem = rest & 127, quad = rem & 31, sp = rem >> 5;
            const float *W = which == 0 ? a.Wih[cell] : a.Whh[cell];
            acc.zero();
            mac<MT, MTP>(acc, reinterpret_cast<const float4 *>(W) + (size_t)(sp * 128) * 32 + quad, 32,
                         s_ds + (cell * kG4 + sp * 128) * MTP, 128);
            store_partial<MT, MTP>(s_part + (cell * 4 + sp) * MTP * 256, 256, 0, which * kU + 4 * quad, acc);
        }
        __syncthreads();
        // ---- P8: new hidden-state carries; publish the gather adjoints for step t-1 (exchange 2)
        {
            float *GYb = a.GY + (size_t)(t & 1) * 3 * N * kU;
            for (int idx = tid; idx < MT * kU; idx += nt) {
                const int m = idx >> 7, k = idx & 127, row = r0 + m;
                if (m < rows) {
                    float gx0 = 0.f, gx1 = 0.f, gq0 = 0.f, gq1 = 0.f;
#pragma unroll
                    for (int sp = 0; sp < 4; ++sp) {
                        const float *p0 = s_part + ((0 * 4 + sp) * MTP + m) * 256, *p1 = s_part + ((1 * 4 + sp) * MTP + m) * 256;
                        gx0 += p0[k]; gq0 += p0[kU + k]; gx1 += p1[k]; gq1 += p1[kU + k];
                    }
                    if (n0 > 0) s_Ghq0[k * MTP + m] = gq0;
                    if (n1 > 0) s_Ghq1[k * MTP + m] = gq1;
                    GYb[(size_t)row * kU + k] = gx0;
                    GYb[((size_t)N + row) * kU + k] = gx1;
                    GYb[((size_t)2 * N + row) * kU + k] = s_gh0[k * MTP + m];
                }
            }
        }
        grid_arrive(a.bar);                     // exchange 2 published; waited for in P5 of step t-1
        ++epoch;
        }   // MODE 0
    }
    __syncthreads();
    if (tid < kU) {
        a.dWqk[(size_t)blockIdx.x * 2 * kU + tid] = s_dwq[tid];
        a.dWqk[(size_t)blockIdx.x * 2 * kU + kU + tid] = s_dwk[tid];
    }
}

}  // namespace lsthm
