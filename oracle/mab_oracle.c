/*
 * TEST INFRASTRUCTURE — plain-C restatement of the LSTHM + multi-attention-block recurrence
 * of HybridRNN_AT / HybridRNN_ATV (forward) and its hand-derived BPTT (backward).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the
 * product path (CUDA kernels behind include/lsthm_b200.h) never does.
 *
 * What it follows in the reference (file:line under /root/reference):
 *   LSTHM.forward                         model/HybridRNN_ATV.py:21-37  (gate order f,i,o,g)
 *   MARN.forward time loop                model/HybridRNN_ATV.py:117-143 (AT: HybridRNN_AT.py:107-132)
 *   multi-attention block                 model/HybridRNN_ATV.py:123-128
 *   fc (Linear-ReLU-Dropout-Linear)       model/HybridRNN_ATV.py:66,129
 * The backward has no reference source (the reference relies on autograd); it follows
 * SURVEY.md App. C and is pinned by tests/test_oracle_golden.py against torch autograd of the
 * oracle's torch restatement, which in turn is pinned against the live reference.
 *
 * Boundary (identical to the CUDA library's lsthm_mab_fwd/bwd):
 *   gx   [T][N][4D]  = W_m x_m + bW_m + bU_m + bV_m per cell, native column order
 *                      (cell-major; inside a cell f|i|o|g blocks of dh_m)
 *   hz   [T][N][2D]  = [h_t | z_t]                     (what nn_out consumes, line 139)
 *   stash: C [T][N][D], G [T][N][4D] gates after sigmoid/tanh (native order),
 *          A [T][N][4][D] softmax weights, R [T][N][RD] reduce outputs, UH [T][N][MH] fc hidden
 * All weights are in nn.Linear layout [out][in].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef REAL_FLOAT
typedef float real;
#else
typedef double real;
#endif

#define MAXM 3

typedef struct {
    int T, N, n_mod, n_att, map_h;
    int dh[MAXM];   /* cell sizes          */
    int rd[MAXM];   /* reduce output sizes */
} mab_dims;

typedef struct {
    const real *U[MAXM], *V[MAXM];          /* [4dh][dh], [4dh][D] */
    const real *Watt, *batt;                /* [4D][D], [4D]       */
    const real *Wr[MAXM], *br[MAXM];        /* [rd][4dh], [rd]     */
    const real *Wf1, *bf1, *Wf2, *bf2;      /* [MH][RD],[MH],[D][MH],[D] */
} mab_weights;

typedef struct {
    real *U[MAXM], *V[MAXM], *Watt, *batt, *Wr[MAXM], *br[MAXM], *Wf1, *bf1, *Wf2, *bf2;
} mab_wgrads;

static int sumi(const int *a, int n) { int s = 0; for (int i = 0; i < n; ++i) s += a[i]; return s; }
static real sigm(real x) { return (real)1 / ((real)1 + exp(-x)); }

int mab_oracle_real_bytes(void) { return (int)sizeof(real); }

/* drop_mask: [T][N][MH] already scaled by 1/(1-p), or NULL (eval). */
void mab_oracle_fwd(const mab_dims *d, const mab_weights *w, const real *gx, const real *drop_mask,
                    real *hz, real *sC, real *sG, real *sA, real *sR, real *sUH)
{
    const int D = sumi(d->dh, d->n_mod), RD = sumi(d->rd, d->n_mod), MH = d->map_h, H = d->n_att;
    const int N = d->N, T = d->T;
    real *e = (real *)malloc(sizeof(real) * H * D);
    real *vec = (real *)malloc(sizeof(real) * H * D);
    for (int n = 0; n < N; ++n) {
        for (int t = 0; t < T; ++t) {
            const size_t tn = (size_t)t * N + n, pn = (size_t)(t - 1) * N + n;
            const real *hp = t ? hz + pn * 2 * D : NULL;        /* h_{t-1} */
            const real *zp = t ? hz + pn * 2 * D + D : NULL;    /* z_{t-1} */
            const real *cp = t ? sC + pn * D : NULL;
            real *c = sC + tn * D, *h = hz + tn * 2 * D, *z = h + D, *g = sG + tn * 4 * D;
            /* LSTHM cells, HybridRNN_ATV.py:119-121 -> 21-37 */
            int o = 0, go = 0;
            for (int m = 0; m < d->n_mod; ++m) {
                const int dh = d->dh[m];
                for (int r = 0; r < 4 * dh; ++r) {
                    real s = gx[tn * 4 * D + go + r];
                    if (t) {
                        for (int k = 0; k < dh; ++k) s += w->U[m][(size_t)r * dh + k] * hp[o + k];
                        for (int k = 0; k < D; ++k) s += w->V[m][(size_t)r * D + k] * zp[k];
                    }
                    g[go + r] = (r < 3 * dh) ? sigm(s) : tanh(s);
                }
                for (int j = 0; j < dh; ++j) {
                    const real f = g[go + j], i = g[go + dh + j], og = g[go + 2 * dh + j], ch = g[go + 3 * dh + j];
                    c[o + j] = f * (t ? cp[o + j] : 0) + i * ch;
                    h[o + j] = tanh(c[o + j]) * og;
                }
                o += dh; go += 4 * dh;
            }
            /* attention over the D features, 4 heads: HybridRNN_ATV.py:123-125 */
            real *a = sA + tn * H * D;
            for (int k = 0; k < H; ++k) {
                real mx = -INFINITY, sum = 0;
                for (int j = 0; j < D; ++j) {
                    real s = w->batt[k * D + j];
                    for (int q = 0; q < D; ++q) s += w->Watt[(size_t)(k * D + j) * D + q] * c[q];
                    e[k * D + j] = s; if (s > mx) mx = s;
                }
                for (int j = 0; j < D; ++j) { a[k * D + j] = exp(e[k * D + j] - mx); sum += a[k * D + j]; }
                for (int j = 0; j < D; ++j) { a[k * D + j] /= sum; vec[k * D + j] = a[k * D + j] * c[j]; }
            }
            /* per-modality regroup (head-major) + reduce: HybridRNN_ATV.py:126-128 */
            real *r = sR + tn * RD;
            o = 0; int ro = 0;
            for (int m = 0; m < d->n_mod; ++m) {
                const int dh = d->dh[m];
                for (int q = 0; q < d->rd[m]; ++q) {
                    real s = w->br[m][q];
                    for (int k = 0; k < H; ++k)
                        for (int j = 0; j < dh; ++j)
                            s += w->Wr[m][(size_t)q * H * dh + k * dh + j] * vec[k * D + o + j];
                    r[ro + q] = s;
                }
                o += dh; ro += d->rd[m];
            }
            /* fc: Linear - ReLU - Dropout - Linear, HybridRNN_ATV.py:66,129 */
            real *u = sUH + tn * MH;
            for (int q = 0; q < MH; ++q) {
                real s = w->bf1[q];
                for (int k = 0; k < RD; ++k) s += w->Wf1[(size_t)q * RD + k] * r[k];
                s = s > 0 ? s : 0;
                if (drop_mask) s *= drop_mask[tn * MH + q];
                u[q] = s;
            }
            for (int j = 0; j < D; ++j) {
                real s = w->bf2[j];
                for (int q = 0; q < MH; ++q) s += w->Wf2[(size_t)j * MH + q] * u[q];
                z[j] = s;
            }
        }
    }
    free(e); free(vec);
}

/*
 * BPTT.  dhz [T][N][2D] = dL/d[h_t|z_t] from the head.  Outputs the per-step adjoints the
 * hoisted weight-gradient products need (dgx = d/d(gate pre-activations), de = d/d(att logits),
 * dr, dup = d/d(fc.0 pre-activation), dzt = total d/dz_t) and, for checking the host-side
 * products, the weight gradients themselves accumulated in (n, t-descending) order.
 */
void mab_oracle_bwd(const mab_dims *d, const mab_weights *w, const real *dhz, const real *drop_mask,
                    const real *hz, const real *sC, const real *sG, const real *sA, const real *sR,
                    const real *sUH, real *dgx, real *de_out, real *dr_out, real *dup_out, real *dzt_out,
                    mab_wgrads *gw)
{
    const int D = sumi(d->dh, d->n_mod), RD = sumi(d->rd, d->n_mod), MH = d->map_h, H = d->n_att;
    const int N = d->N, T = d->T;
    real *dh_c = (real *)calloc(D, sizeof(real)), *dz_c = (real *)calloc(D, sizeof(real));
    real *dc_c = (real *)calloc(D, sizeof(real));
    real *gh = (real *)malloc(sizeof(real) * D), *gz = (real *)malloc(sizeof(real) * D);
    real *gc = (real *)malloc(sizeof(real) * D), *dvec = (real *)malloc(sizeof(real) * H * D);
    real *vec = (real *)malloc(sizeof(real) * H * D);
    for (int n = 0; n < N; ++n) {
        memset(dh_c, 0, sizeof(real) * D); memset(dz_c, 0, sizeof(real) * D); memset(dc_c, 0, sizeof(real) * D);
        for (int t = T - 1; t >= 0; --t) {
            const size_t tn = (size_t)t * N + n, pn = (size_t)(t - 1) * N + n;
            const real *c = sC + tn * D, *g = sG + tn * 4 * D, *a = sA + tn * H * D;
            const real *u = sUH + tn * MH, *r = sR + tn * RD;
            real *ds = dgx + tn * 4 * D, *de = de_out + tn * H * D, *dr = dr_out + tn * RD;
            real *dup = dup_out + tn * MH, *dzt = dzt_out + tn * D;
            for (int j = 0; j < D; ++j) {
                gh[j] = dhz[tn * 2 * D + j] + dh_c[j];
                gz[j] = dhz[tn * 2 * D + D + j] + dz_c[j];
                gc[j] = dc_c[j];
                dzt[j] = gz[j];
            }
            /* z_t = Wf2 u + bf2 ; u = relu(Wf1 r + bf1) * mask */
            for (int q = 0; q < MH; ++q) {
                real s = 0;
                for (int j = 0; j < D; ++j) s += w->Wf2[(size_t)j * MH + q] * gz[j];
                s = (u[q] != 0) ? s : 0;               /* relu'(pre) (and mask==0) */
                if (drop_mask) s *= drop_mask[tn * MH + q];
                dup[q] = s;
            }
            for (int k = 0; k < RD; ++k) {
                real s = 0;
                for (int q = 0; q < MH; ++q) s += w->Wf1[(size_t)q * RD + k] * dup[q];
                dr[k] = s;
            }
            /* reduce layers -> d(attended) */
            int o = 0, ro = 0;
            for (int m = 0; m < d->n_mod; ++m) {
                const int dh = d->dh[m];
                for (int k = 0; k < H; ++k)
                    for (int j = 0; j < dh; ++j) {
                        real s = 0;
                        for (int q = 0; q < d->rd[m]; ++q) s += w->Wr[m][(size_t)q * H * dh + k * dh + j] * dr[ro + q];
                        dvec[k * D + o + j] = s;
                        vec[k * D + o + j] = a[k * D + o + j] * c[o + j];
                    }
                o += dh; ro += d->rd[m];
            }
            /* attended = a * cs ; a = softmax(e) ; e = Watt cs + batt */
            for (int k = 0; k < H; ++k) {
                real dot = 0;
                for (int j = 0; j < D; ++j) dot += a[k * D + j] * dvec[k * D + j] * c[j];
                for (int j = 0; j < D; ++j) {
                    de[k * D + j] = a[k * D + j] * (dvec[k * D + j] * c[j] - dot);
                    gc[j] += dvec[k * D + j] * a[k * D + j];
                }
            }
            for (int q = 0; q < D; ++q) {
                real s = 0;
                for (int kj = 0; kj < H * D; ++kj) s += w->Watt[(size_t)kj * D + q] * de[kj];
                gc[q] += s;
            }
            /* cells */
            o = 0; int go = 0;
            for (int m = 0; m < d->n_mod; ++m) {
                const int dh = d->dh[m];
                for (int j = 0; j < dh; ++j) {
                    const real f = g[go + j], i = g[go + dh + j], og = g[go + 2 * dh + j], ch = g[go + 3 * dh + j];
                    const real tc = tanh(c[o + j]);
                    const real cprev = t ? sC[pn * D + o + j] : 0;
                    const real gcj = gc[o + j] + gh[o + j] * og * (1 - tc * tc);
                    ds[go + j] = gcj * cprev * f * (1 - f);
                    ds[go + dh + j] = gcj * ch * i * (1 - i);
                    ds[go + 2 * dh + j] = gh[o + j] * tc * og * (1 - og);
                    ds[go + 3 * dh + j] = gcj * i * (1 - ch * ch);
                    dc_c[o + j] = gcj * f;
                }
                o += dh; go += 4 * dh;
            }
            /* adjoints into step t-1 : U^T ds, V^T ds */
            memset(dz_c, 0, sizeof(real) * D);
            o = 0; go = 0;
            for (int m = 0; m < d->n_mod; ++m) {
                const int dh = d->dh[m];
                for (int k = 0; k < dh; ++k) {
                    real s = 0;
                    for (int rr = 0; rr < 4 * dh; ++rr) s += w->U[m][(size_t)rr * dh + k] * ds[go + rr];
                    dh_c[o + k] = s;
                }
                for (int k = 0; k < D; ++k) {
                    real s = 0;
                    for (int rr = 0; rr < 4 * dh; ++rr) s += w->V[m][(size_t)rr * D + k] * ds[go + rr];
                    dz_c[k] += s;
                }
                o += dh; go += 4 * dh;
            }
            if (!gw) continue;
            /* weight gradients (what the product computes as hoisted products over all (t,n)) */
            o = 0; go = 0; ro = 0;
            for (int m = 0; m < d->n_mod; ++m) {
                const int dh = d->dh[m];
                if (t) {
                    const real *hp = hz + pn * 2 * D, *zp = hp + D;
                    for (int rr = 0; rr < 4 * dh; ++rr) {
                        for (int k = 0; k < dh; ++k) gw->U[m][(size_t)rr * dh + k] += ds[go + rr] * hp[o + k];
                        for (int k = 0; k < D; ++k) gw->V[m][(size_t)rr * D + k] += ds[go + rr] * zp[k];
                    }
                }
                for (int q = 0; q < d->rd[m]; ++q) {
                    gw->br[m][q] += dr[ro + q];
                    for (int k = 0; k < H; ++k)
                        for (int j = 0; j < dh; ++j)
                            gw->Wr[m][(size_t)q * H * dh + k * dh + j] += dr[ro + q] * vec[k * D + o + j];
                }
                o += dh; go += 4 * dh; ro += d->rd[m];
            }
            for (int kj = 0; kj < H * D; ++kj) {
                gw->batt[kj] += de[kj];
                for (int q = 0; q < D; ++q) gw->Watt[(size_t)kj * D + q] += de[kj] * c[q];
            }
            for (int q = 0; q < MH; ++q) {
                gw->bf1[q] += dup[q];
                for (int k = 0; k < RD; ++k) gw->Wf1[(size_t)q * RD + k] += dup[q] * r[k];
            }
            for (int j = 0; j < D; ++j) {
                gw->bf2[j] += gz[j];
                for (int q = 0; q < MH; ++q) gw->Wf2[(size_t)j * MH + q] += gz[j] * u[q];
            }
        }
    }
    free(dh_c); free(dz_c); free(dc_c); free(gh); free(gz); free(gc); free(dvec); free(vec);
}
