// Host side of the speaker-state cell entry points (include/lsthm_b200.h, lsthm_sps_*).
#include <algorithm>
#include <string>

#include "../../include/lsthm_b200.h"
#include "sps_kernels.cuh"

namespace lsthm {

int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);

// weight packing: transposes nn.Linear / nn.LSTMCell matrices into the k-major, gate-interleaved image the FFMA products stream
struct PackJob {
    const float *src;
    int dst, J, K, ld, row_off, gate_dh;
};
struct PackJobs {
    PackJob j[24];
    int n;
};
__global__ void sps_pack_kernel(const __grid_constant__ PackJobs jobs, float *__restrict__ packed) {
    const PackJob &b = jobs.j[blockIdx.y];
    const int total = b.J * b.K;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / b.K, k = idx - j * b.K;
        const int col = b.gate_dh ? 4 * (j % b.gate_dh) + j / b.gate_dh : j;
        packed[b.dst + (size_t)(b.row_off + k) * b.ld + col] = __ldg(b.src + idx);
    }
}
static int launch_pack(const PackJobs &jobs, float *packed, cudaStream_t st) {
    sps_pack_kernel<<<dim3(32, jobs.n), 256, 0, st>>>(jobs, packed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("weight pack launch", e);
}

#define LSTHM_DECL_SPS(n)                                                              \
    int launch_sps_fwd_##n(const SpsFwdArgs &, int, size_t, cudaStream_t);             \
    int launch_sps_bwd_##n(const SpsBwdArgs &, int, size_t, cudaStream_t);             \
    int launch_gsp_fwd_##n(const SpsFwdArgs &, int, size_t, cudaStream_t);             \
    int launch_gsp_bwd_##n(const SpsBwdArgs &, int, size_t, cudaStream_t);
LSTHM_DECL_SPS(1) LSTHM_DECL_SPS(2) LSTHM_DECL_SPS(3) LSTHM_DECL_SPS(4)
LSTHM_DECL_SPS(5) LSTHM_DECL_SPS(6) LSTHM_DECL_SPS(7) LSTHM_DECL_SPS(8)
typedef int (*SpsFwdFn)(const SpsFwdArgs &, int, size_t, cudaStream_t);
typedef int (*SpsBwdFn)(const SpsBwdArgs &, int, size_t, cudaStream_t);
static const SpsFwdFn kSpsFwd[8] = {launch_sps_fwd_1, launch_sps_fwd_2, launch_sps_fwd_3, launch_sps_fwd_4,
                                    launch_sps_fwd_5, launch_sps_fwd_6, launch_sps_fwd_7, launch_sps_fwd_8};
static const SpsBwdFn kSpsBwd[8] = {launch_sps_bwd_1, launch_sps_bwd_2, launch_sps_bwd_3, launch_sps_bwd_4,
                                    launch_sps_bwd_5, launch_sps_bwd_6, launch_sps_bwd_7, launch_sps_bwd_8};

static const SpsFwdFn kGspFwd[8] = {launch_gsp_fwd_1, launch_gsp_fwd_2, launch_gsp_fwd_3, launch_gsp_fwd_4,
                                    launch_gsp_fwd_5, launch_gsp_fwd_6, launch_gsp_fwd_7, launch_gsp_fwd_8};
static const SpsBwdFn kGspBwd[8] = {launch_gsp_bwd_1, launch_gsp_bwd_2, launch_gsp_bwd_3, launch_gsp_bwd_4,
                                    launch_gsp_bwd_5, launch_gsp_bwd_6, launch_gsp_bwd_7, launch_gsp_bwd_8};

static int sps_rows(const lsthm_sps_desc *d) {
    if (d->rows_per_cta >= 1 && d->rows_per_cta <= 8) return d->rows_per_cta;
    return std::min(8, std::max(1, (d->N + 147) / 148));
}
static size_t sps_smem_fwd(int MT) {
    const int MTP = (MT + 3) & ~3;
    return sizeof(float) * (size_t)(13 * kU * MTP + MTP * 2 * kG4 + MTP * (kU + 4) + 2 * kU + MTP);
}
static size_t sps_smem_bwd(int MT) {
    const int MTP = (MT + 3) & ~3;
    return sizeof(float) * (size_t)(11 * kU * MTP + 2 * kG4 * MTP + 16384 + 8 * MTP * (kU + 4) + 4 * kU + 2 * MTP);
}
static int gsp_rows(const lsthm_gsp_desc *d) {
    if (d->rows_per_cta >= 1 && d->rows_per_cta <= 8) return d->rows_per_cta;
    return std::min(8, std::max(1, (d->N + 147) / 148));
}
static size_t gsp_smem_fwd(int MT) {
    const int MTP = (MT + 3) & ~3;
    return sizeof(float) * (size_t)(13 * kU * MTP + MTP * 4 * kG4 + MTP * (kU + 4) + 2 * kU + MTP + 3 * kU);
}
static int check_gsp(const lsthm_gsp_desc *d) {
    if (!d) return fail_msg("null descriptor");
    if (d->T < 1 || d->N < 1) return fail_msg("T and N must be positive");
    if (d->att_p < 0.f || d->att_p >= 1.f) return fail_msg("att_p must be in [0,1)");
    if (d->listener != 0 && d->listener != 1) return fail_msg("listener must be 0 (onlysp) or 1 (nsps)");
    return 0;
}
// workspace layout (floats): [0,4) barrier counters (fwd uses word 0, bwd word 1) | Q | XQ | GX | GY
struct SpsWs { size_t q, xq, gx, gy, total; };
static SpsWs sps_ws(int N) {
    SpsWs w;
    size_t o = 4;
    w.q = o; o += (size_t)2 * N * 2 * kU;
    w.xq = o; o += (size_t)2 * N * kU;
    w.gx = o; o += (size_t)2 * N * kU;
    w.gy = o; o += (size_t)2 * 3 * N * kU;
    w.total = o;
    return w;
}
static int check_desc(const lsthm_sps_desc *d) {
    if (!d) return fail_msg("null descriptor");
    if (d->T < 1 || d->N < 1) return fail_msg("T and N must be positive");
    if (d->att_p < 0.f || d->att_p >= 1.f) return fail_msg("att_p must be in [0,1)");
    return 0;
}

}  // namespace lsthm

using namespace lsthm;

extern "C" {

size_t lsthm_sps_packed_floats(void) { return (size_t)2 * 256 * kG4 + (size_t)2 * 384 * kG4; }

size_t lsthm_sps_workspace_floats(const lsthm_sps_desc *d) {
    if (check_desc(d)) return 0;
    return sps_ws(d->N).total;
}

int lsthm_sps_launch_info(const lsthm_sps_desc *d, int32_t *grid, int32_t *block, int32_t *rows, int32_t *smem_fwd,
                          int32_t *smem_bwd) {
    if (check_desc(d)) return 1;
    const int MT = sps_rows(d);
    if (grid) *grid = (d->N + MT - 1) / MT;
    if (block) *block = kSpsThreads;
    if (rows) *rows = MT;
    if (smem_fwd) *smem_fwd = (int32_t)sps_smem_fwd(MT);
    if (smem_bwd) *smem_bwd = (int32_t)sps_smem_bwd(MT);
    return 0;
}

int lsthm_sps_pack(const lsthm_sps_weights *w, float *packed, void *stream) {
    if (!w || !packed) return fail_msg("null weights/packed pointer");
    PackJobs jobs;
    int n = 0;
    for (int c = 0; c < 2; ++c) {
        if (!w->Wih[c] || !w->Whh[c] || !w->U[c] || !w->V[c] || !w->S[c]) return fail_msg("null weight pointer");
        const int q = c * 256 * kG4, l = 2 * 256 * kG4 + c * 384 * kG4;
        jobs.j[n++] = PackJob{w->Wih[c], q, kG4, kU, kG4, 0, kU};
        jobs.j[n++] = PackJob{w->Whh[c], q, kG4, kU, kG4, kU, kU};
        jobs.j[n++] = PackJob{w->U[c], l, kG4, kU, kG4, 0, kU};
        jobs.j[n++] = PackJob{w->V[c], l, kG4, kU, kG4, kU, kU};
        jobs.j[n++] = PackJob{w->S[c], l, kG4, kU, kG4, 2 * kU, kU};
    }
    jobs.n = n;
    return launch_pack(jobs, packed, (cudaStream_t)stream);
}

int lsthm_sps_fwd(const lsthm_sps_desc *d, const lsthm_sps_weights *w, const float *packed, const float *gx,
                  const float *qmask, const int32_t *pi, const int32_t *n0, const lsthm_sps_masks *masks,
                  float *workspace, float *out, float *sGQ, float *sCQ, float *sHQ, float *sXQ, float *sGL,
                  float *sCL, float *sHL, void *stream) {
    if (check_desc(d)) return 1;
    if (!w || !packed || !gx || !qmask || !pi || !n0 || !workspace || !out) return fail_msg("null pointer argument");
    if (!w->bq[0] || !w->bq[1] || !w->Wq || !w->Wk) return fail_msg("null weight pointer");
    const bool any = sGQ || sCQ || sHQ || sXQ || sGL || sCL || sHL, all = sGQ && sCQ && sHQ && sXQ && sGL && sCL && sHL;
    if (any && !all) return fail_msg("stash pointers must be all set or all NULL");
    const int MT = sps_rows(d);
    const SpsWs ws = sps_ws(d->N);
    SpsFwdArgs a{};
    a.T = d->T; a.N = d->N;
    for (int c = 0; c < 2; ++c) {
        a.wq_img[c] = packed + (size_t)c * 256 * kG4;
        a.wl_img[c] = packed + (size_t)2 * 256 * kG4 + (size_t)c * 384 * kG4;
        a.bq[c] = w->bq[c];
        a.mq[c] = masks ? masks->mq[c] : nullptr;
    }
    a.Wq = w->Wq; a.Wk = w->Wk; a.gx = gx; a.qmask = qmask; a.pi = pi; a.n0 = n0;
    a.ml = masks ? masks->ml : nullptr; a.ma = masks ? masks->ma : nullptr;
    a.att_mask = masks ? masks->att_mask : nullptr;
    a.att_p = a.att_mask ? 0.f : d->att_p; a.att_seed = d->att_seed;
    a.Q = workspace + ws.q; a.XQ = workspace + ws.xq;
    a.bar = reinterpret_cast<unsigned *>(workspace);
    a.out = out; a.sGQ = sGQ; a.sCQ = sCQ; a.sHQ = sHQ; a.sXQ = sXQ; a.sGL = sGL; a.sCL = sCL; a.sHL = sHL;
    cudaError_t e = cudaMemsetAsync(workspace, 0, 16, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("lsthm_sps_fwd barrier reset", e);
    return kSpsFwd[MT - 1](a, (d->N + MT - 1) / MT, sps_smem_fwd(MT), (cudaStream_t)stream);
}

int lsthm_sps_bwd(const lsthm_sps_desc *d, const lsthm_sps_weights *w, const float *qmask, const int32_t *pi,
                  const int32_t *pr, const int32_t *n0, const lsthm_sps_masks *masks, const float *dout,
                  const float *sGQ, const float *sCQ, const float *sGL, const float *sCL, float *workspace,
                  float *dGL, float *dGQ, float *dWqk, void *stream) {
    if (check_desc(d)) return 1;
    if (!w || !qmask || !pi || !pr || !n0 || !dout || !sGQ || !sCQ || !sGL || !sCL || !workspace || !dGL || !dGQ || !dWqk)
        return fail_msg("null pointer argument");
    const int MT = sps_rows(d);
    const SpsWs ws = sps_ws(d->N);
    SpsBwdArgs a{};
    a.T = d->T; a.N = d->N;
    for (int c = 0; c < 2; ++c) {
        if (!w->U[c] || !w->V[c] || !w->S[c] || !w->Wih[c] || !w->Whh[c]) return fail_msg("null weight pointer");
        a.U[c] = w->U[c]; a.V[c] = w->V[c]; a.S[c] = w->S[c]; a.Wih[c] = w->Wih[c]; a.Whh[c] = w->Whh[c];
        a.mq[c] = masks ? masks->mq[c] : nullptr;
    }
    if (!w->Wq || !w->Wk) return fail_msg("null weight pointer");
    a.Wq = w->Wq; a.Wk = w->Wk; a.qmask = qmask; a.pi = pi; a.pr = pr; a.n0 = n0;
    a.ml = masks ? masks->ml : nullptr; a.ma = masks ? masks->ma : nullptr;
    a.att_mask = masks ? masks->att_mask : nullptr;
    a.att_p = a.att_mask ? 0.f : d->att_p; a.att_seed = d->att_seed;
    a.dout = dout; a.sGQ = sGQ; a.sCQ = sCQ; a.sGL = sGL; a.sCL = sCL;
    a.GX = workspace + ws.gx; a.GY = workspace + ws.gy;
    a.bar = reinterpret_cast<unsigned *>(workspace) + 1;
    a.dGL = dGL; a.dGQ = dGQ; a.dWqk = dWqk;
    cudaError_t e = cudaMemsetAsync(workspace, 0, 16, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("lsthm_sps_bwd barrier reset", e);
    return kSpsBwd[MT - 1](a, (d->N + MT - 1) / MT, sps_smem_bwd(MT), (cudaStream_t)stream);
}


/* ---- GRU speaker-state cell (lsthm_onlysp / lsthm_nsps) ---- */

size_t lsthm_gsp_packed_floats(void) { return (size_t)kU * kG4 + (size_t)2 * 384 * kG4; }

int lsthm_gsp_launch_info(const lsthm_gsp_desc *d, int32_t *grid, int32_t *block, int32_t *rows, int32_t *smem_fwd,
                          int32_t *smem_bwd) {
    if (check_gsp(d)) return 1;
    const int MT = gsp_rows(d);
    if (grid) *grid = (d->N + MT - 1) / MT;
    if (block) *block = kSpsThreads;
    if (rows) *rows = MT;
    if (smem_fwd) *smem_fwd = (int32_t)gsp_smem_fwd(MT);
    if (smem_bwd) *smem_bwd = (int32_t)sps_smem_bwd(MT);
    return 0;
}

int lsthm_gsp_pack(const lsthm_gsp_weights *w, float *packed, void *stream) {
    if (!w || !packed) return fail_msg("null weights/packed pointer");
    if (!w->Whh) return fail_msg("null weight pointer");
    // the (r,z,n,0) image: the pad column must be zero
    cudaError_t e = cudaMemsetAsync(packed, 0, sizeof(float) * (size_t)kU * kG4, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("lsthm_gsp_pack memset", e);
    PackJobs jobs;
    int n = 0;
    jobs.j[n++] = PackJob{w->Whh, 0, 3 * kU, kU, kG4, 0, kU};          // j = gate*128 + unit -> column 4*unit + gate
    for (int c = 0; c < 2; ++c) {
        if (!w->U[c] || !w->V[c] || !w->S[c]) return fail_msg("null weight pointer");
        const int l = kU * kG4 + c * 384 * kG4;
        jobs.j[n++] = PackJob{w->U[c], l, kG4, kU, kG4, 0, kU};
        jobs.j[n++] = PackJob{w->V[c], l, kG4, kU, kG4, kU, kU};
        jobs.j[n++] = PackJob{w->S[c], l, kG4, kU, kG4, 2 * kU, kU};
    }
    jobs.n = n;
    return launch_pack(jobs, packed, (cudaStream_t)stream);
}

int lsthm_gsp_fwd(const lsthm_gsp_desc *d, const lsthm_gsp_weights *w, const float *packed, const float *gx,
                  const float *gxs, const float *qmask, const lsthm_gsp_masks *masks, float *out, float *sGS,
                  float *sQS, float *sGL, float *sCL, void *stream) {
    if (check_gsp(d)) return 1;
    if (!w || !packed || !gx || !gxs || !qmask || !out) return fail_msg("null pointer argument");
    if (!w->bhh || !w->Wq || !w->Wk) return fail_msg("null weight pointer");
    const bool any = sGS || sQS || sGL || sCL, all = sGS && sQS && sGL && sCL;
    if (any && !all) return fail_msg("stash pointers must be all set or all NULL");
    const int MT = gsp_rows(d);
    SpsFwdArgs a{};
    a.T = d->T; a.N = d->N;
    a.whh_img = packed;
    for (int c = 0; c < 2; ++c) a.wl_img[c] = packed + (size_t)kU * kG4 + (size_t)c * 384 * kG4;
    a.bhh = w->bhh; a.Wq = w->Wq; a.Wk = w->Wk; a.gx = gx; a.gxs = gxs; a.qmask = qmask;
    a.ms = masks ? masks->ms : nullptr; a.ml = masks ? masks->ml : nullptr; a.ma = masks ? masks->ma : nullptr;
    a.att_mask = masks ? masks->att_mask : nullptr;
    a.att_p = a.att_mask ? 0.f : d->att_p; a.att_seed = d->att_seed;
    a.listener = d->listener;
    a.out = out; a.sGS = sGS; a.sQS = sQS; a.sGL = sGL; a.sCL = sCL;
    return kGspFwd[MT - 1](a, (d->N + MT - 1) / MT, gsp_smem_fwd(MT), (cudaStream_t)stream);
}

int lsthm_gsp_bwd(const lsthm_gsp_desc *d, const lsthm_gsp_weights *w, const float *qmask, const lsthm_gsp_masks *masks,
                  const float *dout, const float *sGS, const float *sQS, const float *sGL, const float *sCL, float *dGL,
                  float *dGi, float *dGh, float *dWqk, void *stream) {
    if (check_gsp(d)) return 1;
    if (!w || !qmask || !dout || !sGS || !sQS || !sGL || !sCL || !dGL || !dGi || !dGh || !dWqk)
        return fail_msg("null pointer argument");
    const int MT = gsp_rows(d);
    SpsBwdArgs a{};
    a.T = d->T; a.N = d->N;
    for (int c = 0; c < 2; ++c) {
        if (!w->U[c] || !w->V[c] || !w->S[c]) return fail_msg("null weight pointer");
        a.U[c] = w->U[c]; a.V[c] = w->V[c]; a.S[c] = w->S[c];
    }
    if (!w->Whh || !w->Wq || !w->Wk) return fail_msg("null weight pointer");
    a.Whh_s = w->Whh; a.Wq = w->Wq; a.Wk = w->Wk; a.qmask = qmask;
    a.ms = masks ? masks->ms : nullptr; a.ml = masks ? masks->ml : nullptr; a.ma = masks ? masks->ma : nullptr;
    a.att_mask = masks ? masks->att_mask : nullptr;
    a.att_p = a.att_mask ? 0.f : d->att_p; a.att_seed = d->att_seed;
    a.listener = d->listener;
    a.dout = dout; a.sGS = sGS; a.sQS = sQS; a.sGL = sGL; a.sCL = sCL;
    a.dGL = dGL; a.dGi = dGi; a.dGh = dGh; a.dWqk = dWqk;
    return kGspBwd[MT - 1](a, (d->N + MT - 1) / MT, sps_smem_bwd(MT), (cudaStream_t)stream);
}

}  // extern "C"
