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


// Composite weights (fp64 accumulation, rounded once to fp32):
//   W1 = Wf1 . blockdiag(Wr_m)  [MH x 4D]  (columns in the attended order k = head*D + j),   b1 = Wf1 br + bf1
//   W2 = Vcat . Wf2             [4D x MH]  (rows in the native gate order),                  bv = Vcat bf2
// written to every image that uses them (see MabLayout).
__global__ void mab_compose_kernel(const __grid_constant__ ComposeArgs a, float *__restrict__ packed) {
    const MabLayout &L = a.L;
    const int D = L.D, G = L.G, MH = L.MH, R = L.R;
    const int n1 = MH * G, n2 = G * MH, total = n1 + n2 + MH + G;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        if (idx < n1) {                                        // W1[col][k]
            const int col = idx / G, k = idx - col * G, head = k / D, j = k - head * D;
            int m = 0;
            while (m + 1 < L.nm && j >= L.off[m + 1]) ++m;
            const int jl = j - L.off[m], dh = L.dh[m];
            double s = 0.0;
            for (int r = 0; r < L.rd[m]; ++r)
                s += (double)__ldg(a.Wf1 + (size_t)col * R + L.roff[m] + r) * (double)__ldg(a.Wr[m] + (size_t)r * 4 * dh + head * dh + jl);
            packed[L.w1 + (size_t)k * MH + col] = (float)s;
            packed[L.w1n + (size_t)col * G + k] = (float)s;
        } else if (idx < n1 + n2) {                            // W2[g][q], g = goff_m + gate*dh_m + jl
            const int i2 = idx - n1, g = i2 / MH, q = i2 - g * MH;
            int m = 0;
            while (m + 1 < L.nm && g >= L.goff[m + 1]) ++m;
            const int lg = g - L.goff[m], dh = L.dh[m], gate = lg / dh, jl = lg - gate * dh;
            const float *vrow = a.V[m] + (size_t)lg * D;
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += (double)__ldg(vrow + j) * (double)__ldg(a.Wf2 + (size_t)j * MH + q);
            packed[L.w2n + (size_t)g * MH + q] = (float)s;
            packed[L.wg[m] + (size_t)(dh + q) * 4 * dh + 4 * jl + gate] = (float)s;      // forward image: row dh+q, gate-interleaved column
        } else if (idx < n1 + n2 + MH) {                       // b1
            const int col = idx - n1 - n2;
            double s = (double)__ldg(a.bf1 + col);
            for (int m = 0; m < L.nm; ++m)
                for (int r = 0; r < L.rd[m]; ++r) s += (double)__ldg(a.Wf1 + (size_t)col * R + L.roff[m] + r) * (double)__ldg(a.br[m] + r);
            packed[L.b1 + col] = (float)s;
        } else {                                               // bv
            const int g = idx - n1 - n2 - MH;
            int m = 0;
            while (m + 1 < L.nm && g >= L.goff[m + 1]) ++m;
            const float *vrow = a.V[m] + (size_t)(g - L.goff[m]) * D;
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += (double)__ldg(vrow + j) * (double)__ldg(a.bf2 + j);
            packed[L.bvz + g] = (float)s;
        }
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
    for (int m = 0; m < L.nm; ++m) { L.wg[m] = o; o += (L.dh[m] + L.MH) * 4 * L.dh[m]; }
    L.watt = o; o += D * L.G;
    L.w1 = o; o += L.G * L.MH;
    L.w1n = o; o += L.MH * L.G;
    L.w2n = o; o += L.G * L.MH;
    L.batt = o; o += L.G;
    L.b1 = o; o += L.MH;
    L.bvz = o; o += L.G;
    L.total = o;
    if (D + L.MH > L.nt) return fail("unsupported dims (D + map_h exceeds the CTA width)");
    // forward split-K plan of the fused reduce+fc.0 product (K = 4D)
    L.s34chunk = 64; L.s34ns = cdiv(L.G, 64);
    // backward split-K plans
    L.b5total = 0;
    for (int m = 0; m < L.nm; ++m) {
        L.b5chunk[m] = 64; L.b5ns[m] = cdiv(4 * L.dh[m], 64);
        L.b5items[m] = (L.dh[m] / 4) * L.b5ns[m];
        L.b5total += L.b5items[m];
    }
    L.b4ns = 8; L.b4chunk = cdiv(L.G, 8);
    L.b5uchunk = 64; L.b5uns = cdiv(L.G, 64);
    return 0;
}

static void fwd_smem(const MabLayout &L, int MT, FwdSmem &S) {
    const int MTP = (MT + 3) & ~3;
    int o = 8;  // two mbarriers
    S.h = o; o += L.D * MTP;
    S.c = o; o += L.D * MTP;
    S.km = o; o += L.G * MTP;
    S.row = o; o += MTP * L.ldr;
    S.u = o; o += L.MH * MTP;
    S.red = o; o += L.nwarp * kHeads * MTP * 2;
    S.fin = o; o += kHeads * MTP * 2;
    int part = std::max(L.G * MTP, MTP * L.ldr);
    part = std::max(part, L.s34ns * MTP * L.MH);
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
    S.du = o; o += L.MH * MTP;
    S.dc = o; o += L.D * MTP;
    S.gh = o; o += L.D * MTP;
    S.dup = o; o += L.MH * MTP;
    S.km = o; o += L.G * MTP;
    S.C = o; o += MTP * L.ldc;
    S.A = o; o += MTP * L.ldr;      // A tile, then (with the dvec rows) the B4 dc / B5 du partials
    S.row = o; o += MTP * L.ldr;
    int p5 = 0;
    for (int m = 0; m < L.nm; ++m) { S.b5pb[m] = p5; p5 += L.b5ns[m] * MTP * L.dh[m]; }
    S.p2 = o; o += p5;
    S.red = o; o += L.nwarp * kHeads * MTP;
    S.fin = o; o += kHeads * MTP;
    S.dhz = o; o += 2 * MT * 2 * L.D;
    S.duz = o; o += 2 * MT * L.MH;
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
    }
    if (!w->Watt || !w->batt || !w->Wf1 || !w->bf1 || !w->Wf2 || !w->bf2) return fail("null weight pointer");
    add(w->Watt, L.watt, L.G, L.D, L.G, 0, 0);
    add(w->batt, L.batt, L.G, 1, L.G, 0, 0);
    jobs.n = n;
    mab_pack_kernel<<<dim3(32, n), 256, 0, (cudaStream_t)stream>>>(jobs, packed);
    if (check_cuda("lsthm_mab_pack launch")) return 1;
    ComposeArgs c;
    c.L = L;
    for (int m = 0; m < kMaxMod; ++m) { c.V[m] = w->V[m]; c.Wr[m] = w->Wr[m]; c.br[m] = w->br[m]; }
    c.Wf1 = w->Wf1; c.bf1 = w->bf1; c.Wf2 = w->Wf2; c.bf2 = w->bf2;
    mab_compose_kernel<<<148 * 2, 256, 0, (cudaStream_t)stream>>>(c, packed);
    return check_cuda("lsthm_mab_pack compose launch");
}

int lsthm_mab_fwd(const lsthm_mab_desc *d, const float *packed, const float *gx, const float *drop_mask, float *hz,
                  float *u, float *sC, float *sG, float *sA, void *stream) {
    FwdArgs a;
    if (build_layout(d, a.L)) return 1;
    if (!packed || !gx || !hz || !u) return fail("null packed/gx/hz/u pointer");
    const bool any = sC || sG || sA, all = sC && sG && sA;
    if (any && !all) return fail("stash pointers must be all set or all NULL");
    const int MT = pick_rows(d, true);
    fwd_smem(a.L, MT, a.S);
    const size_t bytes = (size_t)a.S.total * 4;
    if (bytes > kMaxSmemBytes) return fail("forward tile does not fit in shared memory");
    a.packed = packed; a.gx = gx; a.mask = drop_mask;
    a.hz = hz; a.sC = sC; a.sG = sG; a.sA = sA; a.sU = u;
    const int grid = cdiv(a.L.N, MT);
    return kFwd[MT - 1](a, grid, bytes, (cudaStream_t)stream);
}

int lsthm_mab_bwd(const lsthm_mab_desc *d, const lsthm_mab_weights *w, const float *packed, const float *dhz,
                  const float *duz, const float *drop_mask, const float *sC, const float *sG, const float *sA, const float *u,
                  float *dgx, float *de, float *dup, float *att, void *stream) {
    BwdArgs a;
    if (build_layout(d, a.L)) return 1;
    if (!w || !packed || !dhz || !duz || !sC || !sG || !sA || !u || !dgx || !de || !dup) return fail("null pointer argument");
    const int MT = pick_rows(d, true);
    bwd_smem(a.L, MT, a.S);
    const size_t bytes = (size_t)a.S.total * 4;
    if (bytes > kMaxSmemBytes) return fail("backward tile does not fit in shared memory");
    a.packed = packed;
    for (int m = 0; m < kMaxMod; ++m) a.U[m] = w->U[m];
    a.Watt = w->Watt;
    a.dhz = dhz; a.duz = duz; a.mask = drop_mask; a.sC = sC; a.sG = sG; a.sA = sA; a.sU = u;
    a.dgx = dgx; a.de = de; a.dup = dup; a.att = att;
    const int grid = cdiv(a.L.N, MT);
    return kBwd[MT - 1](a, grid, bytes, (cudaStream_t)stream);
}

}  // extern "C"
