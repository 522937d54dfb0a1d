// Host side of liblsthm_b200.so: layout planning, weight packing and the extern "C" entry points
// declared in include/lsthm_b200.h.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/lsthm_b200.h"
#include "mab_kernels.cuh"

namespace lsthm {

__global__ void mab_pack_kernel(const __grid_constant__ PackJobs jobs, float *__restrict__ packed) {
    const PackJob &b = jobs.j[blockIdx.y];
    const int total = b.J * b.K;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / b.K, k = idx - j * b.K;
        const int col = b.gate_dh ? 4 * (j % b.gate_dh) + j / b.gate_dh : j;
        packed[b.dst + (size_t)(b.row_off + k) * b.ld + col] = __ldg(b.src + idx);
    }
}


// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
thread_local std::string g_err;
static int fail(const std::string &m) {
    g_err = m;
    return 1;
}

static int build_layout(const lsthm_mab_desc *d, MabLayout &L) {
    if (!d) return fail("null descriptor");
    if (d->n_mod < 1 || d->n_mod > kMaxMod) return fail("n_mod must be 1..3");
    if (d->n_att != kHeads) return fail("n_att must be 4 (reference: num_atts = 4)");
    if (d->T < 1 || d->N < 1) return fail("T and N must be positive");
    if (d->map_h < 4 || d->map_h % 4) return fail("map_h must be a positive multiple of 4");
    memset(&L, 0, sizeof(L));
    L.T = d->T; L.N = d->N; L.nm = d->n_mod; L.MH = d->map_h;
    int D = 0, R = 0;
    for (int m = 0; m < L.nm; ++m) {
        if (d->dh[m] < 4 || d->dh[m] % 4 || d->rd[m] < 4 || d->rd[m] % 4)
            return fail("cell and reduce sizes must be positive multiples of 4");
        L.dh[m] = d->dh[m]; L.rd[m] = d->rd[m];
        L.off[m] = D; L.goff[m] = 4 * D; L.roff[m] = R;
        D += d->dh[m]; R += d->rd[m];
    }
    L.D = D; L.G = 4 * D; L.R = R;
    L.nt = rup(2 * D, 32);
    if (L.nt > kMaxThreads) return fail("sum of cell sizes too large for one CTA (2*D > 448)");
    if (R > L.nt || L.MH > L.nt || L.nt < 64) return fail("unsupported dims (R or map_h exceed the CTA width)");
    L.nwarp = L.nt / 32;
    L.ldr = L.G + ((4 - L.G % 32) + 32) % 32;
    L.ldc = L.D + ((4 - L.D % 32) + 32) % 32;
    L.smchunk = rup(cdiv(D, L.nwarp), 8);
    int o = 0;
    for (int m = 0; m < L.nm; ++m) { L.wg[m] = o; o += (L.dh[m] + D) * 4 * L.dh[m]; }
    L.watt = o; o += D * L.G;
    for (int m = 0; m < L.nm; ++m) { L.wr[m] = o; o += 4 * L.dh[m] * L.rd[m]; }
    L.wf1 = o; o += R * L.MH;
    L.wf2 = o; o += L.MH * D;
    L.batt = o; o += L.G;
    L.br = o; o += R;
    L.bf1 = o; o += L.MH;
    L.bf2 = o; o += D;
    L.vcat = o; o += L.G * D;
    L.total = o;
    // forward split-K plans
    L.s3total = 0;
    for (int m = 0; m < L.nm; ++m) {
        L.s3chunk[m] = L.dh[m] >= 32 ? 32 : L.dh[m];
        if (L.dh[m] % L.s3chunk[m]) return fail("cell size must be <32 or a multiple of 32");
        L.s3ns[m] = 4 * L.dh[m] / L.s3chunk[m];
        L.s3items[m] = (L.rd[m] / 4) * L.s3ns[m];
        L.s3total += L.s3items[m];
    }
    L.s4chunk = 16; L.s4ns = cdiv(R, 16);
    L.s5chunk = 8;  L.s5ns = cdiv(L.MH, 8);
    // backward split-K plans
    L.b1ns = 16; L.b1chunk = cdiv(D, 16);
    L.b2chunk = 16; L.b2ns = cdiv(L.MH, 16);
    L.b3total = 0; L.b5total = 0;
    for (int m = 0; m < L.nm; ++m) {
        L.b3ns[m] = std::min(4, std::max(1, L.rd[m] / 32));
        L.b3chunk[m] = cdiv(L.rd[m], L.b3ns[m]);
        L.b3items[m] = L.dh[m] * L.b3ns[m];
        L.b3total += L.b3items[m];
        L.b5chunk[m] = 64; L.b5ns[m] = cdiv(4 * L.dh[m], 64);
        L.b5items[m] = (L.dh[m] / 4) * L.b5ns[m];
        L.b5total += L.b5items[m];
    }
    L.b4ns = 8; L.b4chunk = cdiv(L.G, 8);
    return 0;
}

static void fwd_smem(const MabLayout &L, int MT, FwdSmem &S) {
    const int MTP = (MT + 3) & ~3;
    int o = 8;  // two mbarriers
    S.h = o; o += L.D * MTP;
    S.z = o; o += L.D * MTP;
    S.c = o; o += L.D * MTP;
    S.km = o; o += L.G * MTP;
    S.row = o; o += MTP * L.ldr;
    S.r = o; o += L.R * MTP;
    S.u = o; o += L.MH * MTP;
    S.red = o; o += L.nwarp * kHeads * MTP * 2;
    S.fin = o; o += kHeads * MTP * 2;
    int part = std::max(L.G * MTP, MTP * L.ldr), p3 = 0;
    for (int m = 0; m < L.nm; ++m) { S.s3pb[m] = p3; p3 += L.s3ns[m] * MTP * L.rd[m]; }
    part = std::max(part, p3);
    part = std::max(part, L.s4ns * MTP * L.MH);
    part = std::max(part, L.s5ns * MTP * L.D);
    S.part = o; o += part;
    S.gx = o; o += 2 * MT * L.G;
    S.mask = o; o += 2 * MT * L.MH;
    S.batt = o; o += L.G;
    S.total = o;
}

static void bwd_smem(const MabLayout &L, int MT, BwdSmem &S) {
    const int MTP = (MT + 3) & ~3;
    int o = 8;  // two mbarriers
    S.dh = o; o += L.D * MTP;
    S.dz = o; o += L.D * MTP;
    S.dc = o; o += L.D * MTP;
    S.gh = o; o += L.D * MTP;
    S.gz = o; o += L.D * MTP;
    S.dup = o; o += L.MH * MTP;
    S.dr = o; o += L.R * MTP;
    S.km = o; o += L.G * MTP;
    S.C = o; o += MTP * L.ldc;
    S.A = o; o += MTP * L.ldr;      // A tile, then (with the dvec rows) the B4/B5 dz partials
    S.row = o; o += MTP * L.ldr;
    int p2 = L.b1ns * MTP * L.MH, p3 = 0, p5 = 0;
    p2 = std::max(p2, L.b2ns * MTP * L.R);
    for (int m = 0; m < L.nm; ++m) {
        S.b3pb[m] = p3; p3 += L.b3ns[m] * MTP * 4 * L.dh[m];
        S.b5pb[m] = p5; p5 += L.b5ns[m] * MTP * L.dh[m];
    }
    p2 = std::max(p2, std::max(p3, p5));
    S.p2 = o; o += p2;
    S.red = o; o += L.nwarp * kHeads * MTP;
    S.fin = o; o += kHeads * MTP;
    S.dhz = o; o += 2 * MT * 2 * L.D;
    S.uh = o; o += 2 * MT * L.MH;
    S.mk = o; o += 2 * MT * L.MH;
    S.total = o;
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
        g_num_sms = p.multiProcessorCount;
    }
    return g_num_sms;
}

static int pick_rows(const lsthm_mab_desc *d, bool need_device) {
    if (d->rows_per_cta >= 1 && d->rows_per_cta <= 8) return d->rows_per_cta;
    int sms = need_device ? num_sms() : 148;
    if (sms <= 0) sms = 148;
    return std::min(8, std::max(1, cdiv(d->N, sms)));
}

static int check_cuda(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}

#define LSTHM_DECL_MT(n)                                                                  \
    int launch_fwd_##n(const FwdArgs &, int, size_t, cudaStream_t);                       \
    int launch_bwd_##n(const BwdArgs &, int, size_t, cudaStream_t);
LSTHM_DECL_MT(1) LSTHM_DECL_MT(2) LSTHM_DECL_MT(3) LSTHM_DECL_MT(4)
LSTHM_DECL_MT(5) LSTHM_DECL_MT(6) LSTHM_DECL_MT(7) LSTHM_DECL_MT(8)
static const FwdLaunchFn kFwd[8] = {launch_fwd_1, launch_fwd_2, launch_fwd_3, launch_fwd_4,
                                    launch_fwd_5, launch_fwd_6, launch_fwd_7, launch_fwd_8};
static const BwdLaunchFn kBwd[8] = {launch_bwd_1, launch_bwd_2, launch_bwd_3, launch_bwd_4,
                                    launch_bwd_5, launch_bwd_6, launch_bwd_7, launch_bwd_8};

int set_error(const char *what, cudaError_t e) { return fail(std::string(what) + ": " + cudaGetErrorString(e)); }
int fail_msg(const char *msg) { return fail(msg); }
int launch_pack(const PackJobs &jobs, float *packed, cudaStream_t st) {
    mab_pack_kernel<<<dim3(32, jobs.n), 256, 0, st>>>(jobs, packed);
    return check_cuda("weight pack launch");
}

constexpr size_t kMaxSmemBytes = 227 * 1024;

}  // namespace lsthm

using namespace lsthm;

extern "C" {

int lsthm_abi_version(void) { return LSTHM_ABI_VERSION; }
const char *lsthm_last_error(void) { return g_err.c_str(); }

size_t lsthm_mab_packed_floats(const lsthm_mab_desc *d) {
    MabLayout L;
    if (build_layout(d, L)) return 0;
    return (size_t)L.total;
}

int lsthm_mab_launch_info(const lsthm_mab_desc *d, int32_t *grid, int32_t *block, int32_t *rows, int32_t *smem_fwd,
                          int32_t *smem_bwd) {
    MabLayout L;
    if (build_layout(d, L)) return 1;
    const int MT = pick_rows(d, false);
    FwdSmem F; BwdSmem B;
    fwd_smem(L, MT, F);
    bwd_smem(L, MT, B);
    if (grid) *grid = cdiv(L.N, MT);
    if (block) *block = L.nt;
    if (rows) *rows = MT;
    if (smem_fwd) *smem_fwd = F.total * 4;
    if (smem_bwd) *smem_bwd = B.total * 4;
    return 0;
}

int lsthm_mab_pack(const lsthm_mab_desc *d, const lsthm_mab_weights *w, float *packed, void *stream) {
    MabLayout L;
    if (build_layout(d, L)) return 1;
    if (!w || !packed) return fail("null weights/packed pointer");
    PackJobs jobs;
    int n = 0;
    auto add = [&](const float *src, int dst, int J, int K, int ld, int row_off, int gate_dh) {
        jobs.j[n++] = PackJob{src, dst, J, K, ld, row_off, gate_dh};
    };
    for (int m = 0; m < L.nm; ++m) {
        if (!w->U[m] || !w->V[m] || !w->Wr[m] || !w->br[m]) return fail("null weight pointer");
        const int dh = L.dh[m];
        add(w->U[m], L.wg[m], 4 * dh, dh, 4 * dh, 0, dh);
        add(w->V[m], L.wg[m], 4 * dh, L.D, 4 * dh, dh, dh);
        add(w->Wr[m], L.wr[m], L.rd[m], 4 * dh, L.rd[m], 0, 0);
        add(w->br[m], L.br + L.roff[m], L.rd[m], 1, L.rd[m], 0, 0);
        add(w->V[m], L.vcat + L.goff[m] * L.D, 1, 4 * dh * L.D, 1, 0, 0);
    }
    if (!w->Watt || !w->batt || !w->Wf1 || !w->bf1 || !w->Wf2 || !w->bf2) return fail("null weight pointer");
    add(w->Watt, L.watt, L.G, L.D, L.G, 0, 0);
    add(w->batt, L.batt, L.G, 1, L.G, 0, 0);
    add(w->Wf1, L.wf1, L.MH, L.R, L.MH, 0, 0);
    add(w->bf1, L.bf1, L.MH, 1, L.MH, 0, 0);
    add(w->Wf2, L.wf2, L.D, L.MH, L.D, 0, 0);
    add(w->bf2, L.bf2, L.D, 1, L.D, 0, 0);
    jobs.n = n;
    mab_pack_kernel<<<dim3(32, n), 256, 0, (cudaStream_t)stream>>>(jobs, packed);
    return check_cuda("lsthm_mab_pack launch");
}

int lsthm_mab_fwd(const lsthm_mab_desc *d, const float *packed, const float *gx, const float *drop_mask, float *hz,
                  float *sC, float *sG, float *sA, float *sR, float *sU, void *stream) {
    FwdArgs a;
    if (build_layout(d, a.L)) return 1;
    if (!packed || !gx || !hz) return fail("null packed/gx/hz pointer");
    const bool any = sC || sG || sA || sR || sU, all = sC && sG && sA && sR && sU;
    if (any && !all) return fail("stash pointers must be all set or all NULL");
    const int MT = pick_rows(d, true);
    fwd_smem(a.L, MT, a.S);
    const size_t bytes = (size_t)a.S.total * 4;
    if (bytes > kMaxSmemBytes) return fail("forward tile does not fit in shared memory");
    a.packed = packed; a.gx = gx; a.mask = drop_mask;
    a.hz = hz; a.sC = sC; a.sG = sG; a.sA = sA; a.sR = sR; a.sU = sU;
    const int grid = cdiv(a.L.N, MT);
    return kFwd[MT - 1](a, grid, bytes, (cudaStream_t)stream);
}

int lsthm_mab_bwd(const lsthm_mab_desc *d, const lsthm_mab_weights *w, const float *packed, const float *dhz,
                  const float *drop_mask, const float *sC, const float *sG, const float *sA, const float *sU,
                  float *dgx, float *de, float *dr, float *dup, float *dzt, float *att, void *stream) {
    BwdArgs a;
    if (build_layout(d, a.L)) return 1;
    if (!w || !packed || !dhz || !sC || !sG || !sA || !sU || !dgx || !de || !dr || !dup || !dzt)
        return fail("null pointer argument");
    const int MT = pick_rows(d, true);
    bwd_smem(a.L, MT, a.S);
    const size_t bytes = (size_t)a.S.total * 4;
    if (bytes > kMaxSmemBytes) return fail("backward tile does not fit in shared memory");
    a.packed = packed;
    for (int m = 0; m < kMaxMod; ++m) { a.U[m] = w->U[m]; a.Wr[m] = w->Wr[m]; }
    a.Watt = w->Watt; a.Wf1 = w->Wf1; a.Wf2 = w->Wf2;
    a.dhz = dhz; a.mask = drop_mask; a.sC = sC; a.sG = sG; a.sA = sA; a.sU = sU;
    a.dgx = dgx; a.de = de; a.dr = dr; a.dup = dup; a.dzt = dzt; a.att = att;
    const int grid = cdiv(a.L.N, MT);
    return kBwd[MT - 1](a, grid, bytes, (cudaStream_t)stream);
}

}  // extern "C"
