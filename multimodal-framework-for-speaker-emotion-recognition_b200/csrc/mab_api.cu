// Host side of the AT/ATV recurrence (include/lsthm_b200.h, lsthm_mab_*): the sharding plan of a
// CTA group, weight images, cooperative launches.
#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/lsthm_b200.h"
#include "mab_kernels.cuh"

namespace lsthm {

int fail_msg(const char *msg);
int set_error(const char *what, cudaError_t e);

static int cdiv2(int a, int b) { return (a + b - 1) / b; }
constexpr int kSmemLimit = 227 * 1024;

// stage-1 allocation of `G` ranks to the modalities (whole 8-unit chunks, contiguous ranks per modality); returns the
// largest per-rank gate cost
static int stage1_alloc(const M2Plan &P, int G, int ranks[kMaxMod]) {
    int chunks[kMaxMod], cost[kMaxMod];
    for (int m = 0; m < P.nm; ++m) { chunks[m] = P.dh[m] / 8; cost[m] = 32 * (P.dh[m] + P.MH); ranks[m] = 1; }
    for (int used = P.nm; used < G; ++used) {
        int best = -1, bc = -1;
        for (int m = 0; m < P.nm; ++m) {
            if (ranks[m] >= chunks[m]) continue;
            const int c = cost[m] * cdiv2(chunks[m], ranks[m]);
            if (c > bc) { bc = c; best = m; }
        }
        if (best < 0) break;
        ++ranks[best];
    }
    int cmax = 0;
    for (int m = 0; m < P.nm; ++m) cmax = std::max(cmax, cost[m] * cdiv2(chunks[m], ranks[m]));
    return cmax;
}

// Builds the sharding plan.  sms <= 0: assume a B200 (148 SMs) — used by the size queries that must work without a device.
static int build_plan(const lsthm_mab_desc *d, int sms, M2Plan &P) {
    if (!d) return fail_msg("null descriptor");
    if (d->n_mod < 1 || d->n_mod > kMaxMod) return fail_msg("n_mod must be 1..3");
    if (d->n_att != kHeads) return fail_msg("n_att must be 4 (reference: num_atts = 4)");
    if (d->T < 1 || d->N < 1) return fail_msg("T and N must be positive");
    if (d->map_h != 64) return fail_msg("map_h must be 64 (reference: map_h = 64)");
    memset(&P, 0, sizeof(P));
    P.T = d->T; P.N = d->N; P.nm = d->n_mod; P.MH = d->map_h;
    int D = 0;
    for (int m = 0; m < P.nm; ++m) {
        if (d->dh[m] < 16 || d->dh[m] % 16 || d->dh[m] > 128) return fail_msg("cell sizes must be multiples of 16 in 16..128");
        if (d->rd[m] < 1) return fail_msg("reduce sizes must be positive");
        P.dh[m] = d->dh[m]; P.off[m] = D; D += d->dh[m];
    }
    if (D > 256) return fail_msg("sum of cell sizes must be <= 256");
    P.D = D; P.G4 = 4 * D;
    // stage 2: per head, D split into ranges of whole k16 steps, at most 80 features each
    const int k16 = D / 16;
    P.nr = cdiv2(k16, 5);
    const int ns2 = 4 * P.nr;
    if (ns2 > kM2MaxRanks) return fail_msg("too many attention slices for one CTA group");
    // stage 1: every rank owns at most 16 hidden units of one modality (one 8-unit chunk per epilogue warp of a lane quarter);
    // the group has as many ranks as the larger of the two stages needs
    int ranks[kMaxMod], n1 = 0;
    for (int m = 0; m < P.nm; ++m) { ranks[m] = cdiv2(P.dh[m] / 8, 2); n1 += ranks[m]; }
    const int G = std::max(ns2, n1);
    if (G > kM2MaxRanks) return fail_msg("cells too large for one CTA group");
    for (int m = 0; n1 < G; m = (m + 1) % P.nm)                     // spare ranks: spread the widest modality thinner
        if (ranks[m] < P.dh[m] / 8) { ++ranks[m]; ++n1; }
    P.G = G;
    int rank = 0;
    for (int m = 0; m < P.nm; ++m) {
        const int chunks = P.dh[m] / 8, base = chunks / ranks[m], rem = chunks % ranks[m];
        int c0 = 0;
        for (int i = 0; i < ranks[m]; ++i, ++rank) {
            const int nc = base + (i < rem ? 1 : 0);
            if (ranks[m] > 8) return fail_msg("more than 8 ranks per modality");
            P.r[rank].m = m; P.r[rank].u0 = P.off[m] + 8 * c0; P.r[rank].nu = 8 * nc;
            P.r[rank].dhm = P.dh[m]; P.r[rank].offm = P.off[m];
            P.r[rank].mr0 = rank - i; P.r[rank].mr1 = rank - i + ranks[m];
            if (8 * nc > kM2MaxNU) return fail_msg("cell too large for the group plan");
            c0 += nc;
        }
    }
    for (; rank < G; ++rank) return fail_msg("internal: unassigned rank in the group plan");
    for (int r = 0; r < G; ++r) { P.r[r].head = -1; P.r[r].j0 = 0; P.r[r].nj = 0; }
    {
        const int base = k16 / P.nr, rem = k16 % P.nr;
        for (int k = 0; k < kHeads; ++k) {
            int j0 = 0;
            for (int i = 0; i < P.nr; ++i) {
                const int nj = 16 * (base + (i >= P.nr - rem ? 1 : 0));
                M2Rank &R = P.r[k * P.nr + i];
                R.head = k; R.j0 = j0; R.nj = nj;
                if (nj > kM2MaxNJ) return fail_msg("attention slice too wide");
                j0 += nj;
            }
        }
    }
    // blobs
    for (int r = 0; r < G; ++r) {
        P.blob_f = std::max(P.blob_f, m2_fwd_blob(P, P.r[r]).total);
        P.blob_b = std::max(P.blob_b, m2_bwd_blob(P, P.r[r]).total);
    }
    // dialogues per group
    if (sms <= 0) sms = 148;
    const int gmax = sms / G;
    if (gmax < 1) return fail_msg("device has fewer SMs than one CTA group");
    int DG = d->rows_per_cta > 0 ? std::min(d->rows_per_cta, kM2MaxDG) : std::min(kM2MaxDG, m2_align(cdiv2(P.N, gmax), 8));
    DG = std::min(DG, 8 * G);                                   // the combine role covers 8 dialogues per rank
    for (;; DG -= 8) {
        if (DG < 1) return fail_msg("the group plan does not fit in shared memory");
        const int Mr = m2_align(DG, 8);
        int kf = D, kb = 0;
        for (int m = 0; m < P.nm; ++m) kf = std::max(kf, P.dh[m] + P.MH);
        for (int r = 0; r < G; ++r) kb = std::max(kb, P.MH + P.r[r].nj + 4 * P.r[r].nu);
        P.act_f = kf * Mr * 4;
        P.act_b = kb * Mr * 4;
        if (kM2CtrlBytes + P.blob_f + P.act_f + 3072 <= kSmemLimit && kM2CtrlBytes + P.blob_b + P.act_b + 3072 <= kSmemLimit) break;
        if (DG <= 8) return fail_msg("the group plan does not fit in shared memory");
    }
    P.DG = DG; P.Mr = m2_align(DG, 8);
    P.nblocks = cdiv2(P.N, DG);
    P.ngroups = std::min(gmax, P.nblocks);
    P.cd = cdiv2(DG, G);
    // exchange workspace of one group
    const int Mr = P.Mr;
    int o = 0;
    auto take = [&](int bytes) { const int at = o; o += m2_align(bytes, 128); return at; };
    P.ws_xc = take(D * Mr * 4);
    P.ws_xh = take(2 * D * Mr * 4);
    P.ws_xu = take(P.MH * Mr * 4);
    P.ws_xp = take(G * Mr * P.MH * 4);
    P.ws_xst = take(G * Mr * 2 * 4);
    P.ws_xdc = take((G + 4) * D * Mr * 4);
    P.ws_xdu = take(G * Mr * P.MH * 4);
    P.ws_xdh = take(G * 16 * Mr * 32);
    P.ws_xdup = take(P.MH * Mr * 4 + Mr * 16);
    P.ws_group = o;
    return 0;
}

// packed area: [composite weights fp32][rank table][forward blobs][backward blobs]
static size_t comp_floats(const M2Plan &P) { return (size_t)2 * P.MH * P.G4 + P.MH + P.G4; }
static size_t ranktab_off(const M2Plan &P) { return (size_t)m2_align((int)(comp_floats(P) * 4), 128); }
static size_t blobs_off(const M2Plan &P) { return ranktab_off(P) + (size_t)m2_align((int)(kM2MaxRanks * sizeof(M2Rank)), 128); }
static size_t bars_bytes(const M2Plan &P) { return (size_t)std::max(1, 148 / P.G + 1) * 512; }

static int device_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}

template <typename K, typename A>
static int coop_launch2(K kernel, const A &args, int grid, size_t smem_bytes, cudaStream_t st, const char *what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return set_error(what, e);
    void *params[] = {const_cast<A *>(&args)};
    e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(kM2Threads), params, smem_bytes, st);
    return e == cudaSuccess ? 0 : set_error(what, e);
}

}  // namespace lsthm

using namespace lsthm;

extern "C" {

size_t lsthm_mab_pack_bytes(const lsthm_mab_desc *d) {
    M2Plan P;
    if (build_plan(d, 0, P)) return 0;
    return blobs_off(P) + (size_t)P.G * (P.blob_f + P.blob_b) + 256;
}

size_t lsthm_mab_workspace_bytes(const lsthm_mab_desc *d) {
    M2Plan P;
    if (build_plan(d, 0, P)) return 0;
    // sized for the largest group count / dialogue block any device could pick for these dims
    M2Plan Q = P;
    lsthm_mab_desc d2 = *d;
    d2.rows_per_cta = kM2MaxDG;
    if (build_plan(&d2, 0, Q)) return 0;
    const size_t per_group = (size_t)std::max(P.ws_group, Q.ws_group);
    return bars_bytes(P) + per_group * (size_t)std::max(1, 160 / P.G);
}

int lsthm_mab_plan_info(const lsthm_mab_desc *d, int32_t *out, int32_t n_out) {
    // out: G, nr, DG, Mr, ngroups, nblocks, cd, blob_f, blob_b, act_f, act_b, smem_fwd, smem_bwd, ws_group, then per rank
    // (m, u0, nu, head, j0, nj)
    M2Plan P;
    if (build_plan(d, 0, P)) return 1;
    const int32_t head[14] = {P.G, P.nr, P.DG, P.Mr, P.ngroups, P.nblocks, P.cd, P.blob_f, P.blob_b, P.act_f, P.act_b,
                              kM2CtrlBytes + P.blob_f + P.act_f + 3072, kM2CtrlBytes + P.blob_b + P.act_b + 3072, P.ws_group};
    int k = 0;
    for (int i = 0; i < 14 && k < n_out; ++i) out[k++] = head[i];
    for (int r = 0; r < P.G; ++r) {
        const int32_t v[6] = {P.r[r].m, P.r[r].u0, P.r[r].nu, P.r[r].head, P.r[r].j0, P.r[r].nj};
        for (int i = 0; i < 6 && k < n_out; ++i) out[k++] = v[i];
    }
    return 0;
}

int lsthm_mab_set_trace(void *buf) {
#ifndef LSTHM_M2_TRACE
    if (buf) return fail_msg("lsthm_mab_set_trace: the library was built without LSTHM_M2_TRACE (make TRACE=1)");
#endif
    long long *p = reinterpret_cast<long long *>(buf);
    cudaError_t e = cudaMemcpyToSymbol(g_m2_trace, &p, sizeof(p));
    return e == cudaSuccess ? 0 : set_error("lsthm_mab_set_trace", e);
}

int lsthm_mab_pack(const lsthm_mab_desc *d, const lsthm_mab_weights *w, void *packed, void *stream) {
    M2Plan P;
    if (build_plan(d, 0, P)) return 1;
    if (!w || !packed) return fail_msg("null weights/packed pointer");
    if (!w->Watt || !w->batt || !w->Wf1 || !w->bf1 || !w->Wf2 || !w->bf2) return fail_msg("null weight pointer");
    float *comp = reinterpret_cast<float *>(packed);
    M2CompArgs c;
    c.P = P;
    int R = 0;
    for (int m = 0; m < kMaxMod; ++m) {
        if (m < P.nm && (!w->U[m] || !w->V[m] || !w->Wr[m] || !w->br[m])) return fail_msg("null weight pointer");
        c.V[m] = w->V[m]; c.Wr[m] = w->Wr[m]; c.br[m] = w->br[m];
        c.rd[m] = m < P.nm ? d->rd[m] : 0; c.roff[m] = R; R += c.rd[m];
    }
    c.R = R;
    c.Wf1 = w->Wf1; c.bf1 = w->bf1; c.Wf2 = w->Wf2; c.bf2 = w->bf2;
    c.W1 = comp; c.W2 = c.W1 + (size_t)P.MH * P.G4; c.b1 = c.W2 + (size_t)P.G4 * P.MH; c.bv = c.b1 + P.MH;
    mab_compose_kernel<<<148 * 2, 256, 0, (cudaStream_t)stream>>>(c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("lsthm_mab_pack compose launch", e);
    M2ImgArgs ia;
    ia.P = P;
    for (int m = 0; m < kMaxMod; ++m) ia.U[m] = w->U[m];
    ia.Watt = w->Watt; ia.batt = w->batt; ia.W1 = c.W1; ia.W2 = c.W2; ia.b1 = c.b1; ia.bv = c.bv;
    uint8_t *base = reinterpret_cast<uint8_t *>(packed) + blobs_off(P);
    ia.ranktab = reinterpret_cast<M2Rank *>(reinterpret_cast<uint8_t *>(packed) + ranktab_off(P));
    ia.blob_f = base;
    ia.blob_b = base + (size_t)P.G * P.blob_f;
    mab_image_kernel<<<dim3(12, P.G), 256, 0, (cudaStream_t)stream>>>(ia);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_mab_pack image launch", e);
}

int lsthm_mab_launch_info(const lsthm_mab_desc *d, int32_t *grid, int32_t *block, int32_t *group, int32_t *dialogues_per_group,
                           int32_t *smem_fwd, int32_t *smem_bwd, int32_t *padded_rows) {
    M2Plan P;
    if (build_plan(d, device_sms(), P)) return 1;
    if (grid) *grid = P.ngroups * P.G;
    if (block) *block = kM2Threads;
    if (group) *group = P.G;
    if (dialogues_per_group) *dialogues_per_group = P.DG;
    if (smem_fwd) *smem_fwd = kM2CtrlBytes + P.blob_f + P.act_f + 3072;
    if (smem_bwd) *smem_bwd = kM2CtrlBytes + P.blob_b + P.act_b + 3072;
    if (padded_rows) *padded_rows = P.nblocks * P.Mr;
    return 0;
}

int lsthm_mab_fwd(const lsthm_mab_desc *d, const void *packed, const float *gx, const float *drop_mask, float *hz, float *u,
                   float *sC, float *sCp, float *sG, float *sE, float *sMS, float *sP, void *workspace, void *stream) {
    M2FwdArgs a;
    const int sms = device_sms();
    if (sms <= 0) return fail_msg("no CUDA device (there is no CPU path)");
    if (build_plan(d, sms, a.P)) return 1;
    if (!packed || !gx || !hz || !u || !workspace) return fail_msg("null packed/gx/hz/u/workspace pointer");
    const bool any = sC || sCp || sG || sE || sMS || sP, all = sC && sCp && sG && sE && sMS && sP;
    if (any && !all) return fail_msg("stash pointers must be all set or all NULL");
    const M2Plan &P = a.P;
    a.blob = reinterpret_cast<const uint8_t *>(packed) + blobs_off(P);
    a.ranktab = reinterpret_cast<const M2Rank *>(reinterpret_cast<const uint8_t *>(packed) + ranktab_off(P));
    a.gx = gx; a.mask = drop_mask; a.hz = hz; a.sU = u;
    a.sC = sC; a.sCp = sCp; a.sG = sG; a.sE = sE; a.sMS = sMS; a.sP = sP;
    a.bars = reinterpret_cast<unsigned *>(workspace);
    a.ws = reinterpret_cast<uint8_t *>(workspace) + bars_bytes(P);
    cudaError_t e = cudaMemsetAsync(workspace, 0, bars_bytes(P), (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("lsthm_mab_fwd counter reset", e);
    const size_t smem = (size_t)kM2CtrlBytes + P.blob_f + P.act_f + 3072;
    return coop_launch2(mab_fwd_kernel, a, P.ngroups * P.G, smem, (cudaStream_t)stream, "lsthm_mab_fwd launch");
}

int lsthm_mab_bwd(const lsthm_mab_desc *d, const void *packed, const float *dhz, const float *duz, const float *drop_mask,
                   const float *sCp, const float *sG, const float *sE, const float *sMS, const float *sP, const float *u,
                   float *dgx, float *de, float *dup, float *att, void *workspace, void *stream) {
    M2BwdArgs a;
    const int sms = device_sms();
    if (sms <= 0) return fail_msg("no CUDA device (there is no CPU path)");
    if (build_plan(d, sms, a.P)) return 1;
    if (!packed || !dhz || !duz || !sCp || !sG || !sE || !sMS || !sP || !u || !dgx || !de || !dup || !workspace)
        return fail_msg("null pointer argument");
    const M2Plan &P = a.P;
    a.blob = reinterpret_cast<const uint8_t *>(packed) + blobs_off(P) + (size_t)P.G * P.blob_f;
    a.ranktab = reinterpret_cast<const M2Rank *>(reinterpret_cast<const uint8_t *>(packed) + ranktab_off(P));
    a.dhz = dhz; a.duz = duz; a.mask = drop_mask; a.sCp = sCp; a.sG = sG; a.sE = sE; a.sMS = sMS; a.sP = sP; a.sU = u;
    a.dgx = dgx; a.de = de; a.dup = dup; a.att = att;
    a.bars = reinterpret_cast<unsigned *>(workspace);
    a.ws = reinterpret_cast<uint8_t *>(workspace) + bars_bytes(P);
    cudaError_t e = cudaMemsetAsync(workspace, 0, bars_bytes(P), (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("lsthm_mab_bwd counter reset", e);
    const size_t smem = (size_t)kM2CtrlBytes + P.blob_b + P.act_b + 3072;
    return coop_launch2(mab_bwd_kernel, a, P.ngroups * P.G, smem, (cudaStream_t)stream, "lsthm_mab_bwd launch");
}

}  // extern "C"
