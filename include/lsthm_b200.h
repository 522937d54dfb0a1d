/*
 * lsthm_b200 — C ABI of the B200-native LSTHM hybrid-recurrence library (liblsthm_b200.so).
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; what a replacement has to bind
 * is the *operator* each entry point stands in for.  Citations are file:line under the reference
 * tree (MallVilliers/Multimodal-Framework-for-speaker-emotion-recognition).
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless it says "host".
 *   - the caller owns all memory (PyTorch's caching allocator in our host mirror); the library
 *     never allocates or frees device memory and keeps no pointer after a call returns.
 *   - every launch goes to the cudaStream_t passed as `stream` (void* to keep this header free of
 *     cuda_runtime.h); calls are asynchronous; there is no device synchronisation inside.
 *   - return 0 on success; non-zero -> lsthm_last_error() (thread-local, host string).
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails.
 *   - tensors are contiguous, time-major [T][N][...], fp32.
 */
#ifndef LSTHM_B200_H_
#define LSTHM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSTHM_ABI_VERSION 7
#define LSTHM_MAX_MOD 3

int lsthm_abi_version(void);
const char *lsthm_last_error(void);

/* ------------------------------------------------------------------------------------------
 * AT / ATV: LSTHM cells + multi-attention block (MAB) recurrence.
 * Replaces the body of the time loop of MARN.forward
 *     model/HybridRNN_ATV.py:117-143   (AT: model/HybridRNN_AT.py:107-132)
 * i.e. LSTHM.forward (model/HybridRNN_ATV.py:21-37) for every modality, the 4-head softmax
 * attention over the concatenated cell states (123-125), the per-modality reduce layers
 * (126-128) and fc = Linear-ReLU-Dropout-Linear (66, 129) — and the BPTT autograd runs through it.
 * The input projections W_m x_m and the per-step head nn_out (68-73,139-141) are time-parallel and
 * stay on the host side.
 *
 * Execution model (sm_100a): GROUPS of co-resident thread blocks (cooperative launch; 13 blocks for ATV,
 * 9 for AT) each own a block of up to 96 dialogues for all T steps.  The chain weights are sharded over
 * the ranks of a group and stay RESIDENT in shared memory for the whole launch as bf16 hi/lo operand
 * images; every product runs on tcgen05 tensor cores (M = the group's dialogues, three-term split,
 * fp32 accumulation in TMEM); the ranks exchange c_t/h_t, partial products and u_t through L2.
 * The kernels use composite weights — the chain has no nonlinearity between reduce_dim_nn_m and fc.0
 * (HybridRNN_ATV.py:126-129), nor between fc.3 and the V term of the next step's gates (:129 -> :25):
 *     W1 = Wf1 . blockdiag(Wr_m)  [map_h x 4D],  b1 = Wf1 br + bf1      W2 = Vcat . Wf2  [4D x map_h],  bv = Vcat bf2
 * (composed in fp64 at pack time): three dependent products per step instead of five.
 * Constraints: map_h = 64, n_att = 4, every dh_m a multiple of 16 in 16..128, sum dh_m <= 256.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t T;                 /* padded dialogue length                                   */
    int32_t N;                 /* dialogues in the batch                                   */
    int32_t n_mod;             /* 2 (AT: text, audio) or 3 (ATV: + visual)                 */
    int32_t n_att;             /* attention heads (reference: 4)                           */
    int32_t map_h;             /* fc hidden width (reference: 64)                          */
    int32_t dh[LSTHM_MAX_MOD]; /* cell sizes   (ATV 128,16,64)                             */
    int32_t rd[LSTHM_MAX_MOD]; /* reduce sizes (ATV 16,128,100)                            */
    int32_t rows_per_cta;      /* dialogues per block group: 0 = auto, else 1..96          */
} lsthm_mab_desc;

/* Weights in nn.Linear layout [out][in], exactly the module's parameter storage. */
typedef struct {
    const float *U[LSTHM_MAX_MOD];    /* lsthm_m.U.weight [4dh_m][dh_m]   HybridRNN_ATV.py:18 */
    const float *V[LSTHM_MAX_MOD];    /* lsthm_m.V.weight [4dh_m][D]      HybridRNN_ATV.py:19 */
    const float *Watt, *batt;         /* att.0            [4D][D],[4D]    HybridRNN_ATV.py:60 */
    const float *Wr[LSTHM_MAX_MOD];   /* reduce_dim_nn_m.0 [rd_m][4dh_m]  HybridRNN_ATV.py:62-64 */
    const float *br[LSTHM_MAX_MOD];
    const float *Wf1, *bf1;           /* fc.0 [map_h][R]                  HybridRNN_ATV.py:66 */
    const float *Wf2, *bf2;           /* fc.3 [D][map_h]                                      */
} lsthm_mab_weights;

/* bytes of the packed weight area (composite weights + per-rank operand images for forward and backward) */
size_t lsthm_mab_pack_bytes(const lsthm_mab_desc *d);
/* bytes of the exchange workspace a launch needs (group exchange buffers + barrier counters); the library resets
 * the counters itself (one cudaMemsetAsync on `stream` per launch), the rest needs no initialisation */
size_t lsthm_mab_workspace_bytes(const lsthm_mab_desc *d);
/* compose W1, W2 (fp64 accumulation) and split every rank's slices into bf16 hi/lo images in the canonical K-major UMMA
 * layout (call after every optimizer step, before fwd/bwd) */
int lsthm_mab_pack(const lsthm_mab_desc *d, const lsthm_mab_weights *w, void *packed, void *stream);

/*
 * Forward recurrence over all T steps (cooperative launch).
 *   gx        [T][N][4D]   W_m x_m + bW_m + bU_m + bV_m, cell-major, f|i|o|g inside a cell
 *   drop_mask [T][N][map_h] keep/(1-p) mask of fc's Dropout, or NULL (eval mode)
 *   hz        [T][N][2D]   out: the h_t half of [h_t | z_t] (what nn_out consumes, HybridRNN_ATV.py:139); the z_t half is
 *                          NOT written: z_t = fc.3(u_t) = u_t Wf2^T + bf2 is one time-parallel product over all T*N rows
 *                          that the caller forms from `u` (nothing on the serial path needs z_t)
 *   u         [T][N][map_h] out: fc hidden after ReLU and dropout (HybridRNN_ATV.py:66)
 * stash (all out, may ALL be NULL for inference):
 *   sC [T][N][D] cell states, row-major (the caller's weight-gradient products read it)
 * and the PRIVATE stash of the kernel pair, piece-major inside a dialogue block so that every warp access is one contiguous
 * run:  [T][block][column / 4][row][4]  with `padded_rows / blocks` rows per block (lsthm_mab_launch_info):
 *   sCp (width D) cell states        sG (width 4D) gates after sigmoid/tanh, column order of gx
 *   sE (width 4D) attention logits (att.0 output incl. bias, BEFORE the softmax), column = head * D + feature
 *   sMS [T][block][4][row][2] per head (max logit, 1 / sum exp(e - max)): softmax weights are a = exp(e - max) * inv
 *   sP (width 4 * map_h) per head  W1[:, head block] . (a_head * c), column = head * map_h + q
 *      (the softmax backward needs <dup, P_head>)
 * each private tensor holds T * padded_rows * width floats (sMS: T * padded_rows * 8).
 */
int lsthm_mab_fwd(const lsthm_mab_desc *d, const void *packed, const float *gx, const float *drop_mask, float *hz, float *u,
                  float *sC, float *sCp, float *sG, float *sE, float *sMS, float *sP, void *workspace, void *stream);

/*
 * BPTT (cooperative launch).
 *   dhz  [T][N][2D]  dL/d[h_t|z_t] from the head (the kernel reads the h half)
 *   duz  [T][N][map_h]  (dL/dz_t from the head) . Wf2 — the head's gradient pulled through fc.3, one time-parallel
 *                    product formed by the caller
 * out (the adjoints the time-parallel weight-gradient products consume):
 *   dgx  [T][N][4D]  dL/d(gate pre-activations) ds_t  -> dW,dU,dV,db and dx; also the carried part of dL/dz:
 *                    dL/dz_t(total) = dL/dz_t(head) + ds_{t+1} . Vcat  -> d fc.3
 *   de   [T][N][4D]  dL/d(att logits)            -> d att.0
 *   dup  [T][N][map_h] dL/d(fc.0 pre-activation) -> d fc.0, and dr = dup . Wf1 -> d reduce_dim_nn_*
 *   att  [T][N][4D]  (optional, may be NULL) the attended features a * c of the forward, regrouped per modality and
 *                    head-major inside a modality (columns 4*off_m + head*dh_m + j): column block m is the operand
 *                    of d reduce_dim_nn_m and of the recomputed reduce outputs r_m = att_m Wr_m^T + br_m
 */
int lsthm_mab_bwd(const lsthm_mab_desc *d, const void *packed, const float *dhz, const float *duz, const float *drop_mask,
                  const float *sCp, const float *sG, const float *sE, const float *sMS, const float *sP, const float *u,
                  float *dgx, float *de, float *dup, float *att, void *workspace, void *stream);

/* Launch geometry on the current device: grid = groups * group size; padded_rows = blocks * rows per block (multiple of 8). */
int lsthm_mab_launch_info(const lsthm_mab_desc *d, int32_t *grid, int32_t *block, int32_t *group,
                          int32_t *dialogues_per_group, int32_t *smem_fwd, int32_t *smem_bwd, int32_t *padded_rows);
/* The sharding plan for a 148-SM device (host only, no device needed): 14 header ints (G, ranges per head, dialogues
 * per group, padded rows, groups, blocks, combine share, blob_f, blob_b, act_f, act_b, smem_fwd, smem_bwd, ws_group)
 * followed by (modality, first unit, units, head, first feature, features) per rank. */
int lsthm_mab_plan_info(const lsthm_mab_desc *d, int32_t *out, int32_t n_out);
/* Profiling aid (library built with `make TRACE=1`): buf = device buffer of [T][2][16] int64 (or NULL to switch off);
 * block 0 of the kernels records clock64() at the phase boundaries of its control warp (role 0) and of its first
 * epilogue warp (role 1). */
int lsthm_mab_set_trace(void *buf);

/* ------------------------------------------------------------------------------------------
 * lsthm_sps: speaker-state LSTHM cell (one direction of the bidirectional MARN1_sps).
 * Replaces MARN_cell.forward, model/lsthm_sps.py:156-221: the per-speaker nn.LSTMCell pair on
 * "packed" rows (_select_parties, :238-259), the party-state update (:204-207), LSTHM1.forward for
 * text and audio (:28-44, with the speaker term S), dropout on every recurrent state and the in-cell
 * rank-1 CrossAttention (:59-72).  The input projections W x, the bidirectional wrapper
 * (_reverse_seq, :396-410), the sequence-level CrossAttention2/3 and the heads stay on the host side.
 *
 * Dialogues of a shard are coupled through the packed rows, so the kernels are
 * cooperative launches over the whole shard: N <= rows_per_cta * #SMs (N <= 1184 on a B200).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t T, N;
    int32_t rows_per_cta;      /* 0 = auto; else 1..8                                          */
    float att_p;               /* in-kernel attention dropout prob when att_mask == NULL; 0=off */
    uint64_t att_seed;
} lsthm_sps_desc;

typedef struct {               /* all [512][128] in nn.Linear / nn.LSTMCell layout; index 0/1 = l/a or q0/q1 */
    const float *U[2], *V[2], *S[2];       /* marn_cell.lsthm_{l,a}.{U,V,S}.weight   lsthm_sps.py:17-19 */
    const float *Wih[2], *Whh[2];          /* marn_cell.lstm_q{0,1}.weight_{ih,hh}   lsthm_sps.py:146-147 */
    const float *bq[2];                    /* bias_ih + bias_hh [512] (summed by the caller)          */
    const float *Wq, *Wk;                  /* marn_cell.crossatt_l2a.{Wq,Wk} [128]   lsthm_sps.py:52-53 */
} lsthm_sps_weights;

typedef struct {               /* dropout masks already scaled by 1/(1-p); any may be NULL (= no dropout) */
    const float *mq[2];        /* on hq0 / hq1 after the speaker cells  [T][N][128]   lsthm_sps.py:184,189 */
    const float *ml, *ma;      /* on h_l / h_a                          [T][N][128]   lsthm_sps.py:211,213 */
    const float *att_mask;     /* on the attention weights              [T][N][128][128] lsthm_sps.py:69  */
} lsthm_sps_masks;

size_t lsthm_sps_packed_floats(void);
int lsthm_sps_pack(const lsthm_sps_weights *w, float *packed, void *stream);
/* floats of caller-provided scratch for the inter-CTA exchange buffers + barrier counter */
size_t lsthm_sps_workspace_floats(const lsthm_sps_desc *d);

/*
 *   gx    [T][N][2][512]  W_c x_c + (bW+bU+bV+bS)_c for c = l, a; gate order f|i|o|g
 *   qmask [T][N][2]       one-hot current speaker (zero rows on padding)
 *   pi    [T][N] int32    dialogue occupying packed row r (speaker-0 dialogues first, ascending ids)
 *   n0    [T]    int32    number of speaker-0 dialogues
 *   out   [T][N][512]     [h_l | h_a | z_l | h_q]                        (lsthm_sps.py:218)
 *   stash (all or none): sGQ [T][N][2][512] LSTMCell gates i|f|g|o, sCQ/sHQ/sXQ [T][N][2][128] cell state,
 *         hidden state after dropout, cell input;  sGL [T][N][2][512] LSTHM gates f|i|o|g, sCL/sHL [T][N][2][128]
 */
int lsthm_sps_fwd(const lsthm_sps_desc *d, const lsthm_sps_weights *w, const float *packed, const float *gx,
                  const float *qmask, const int32_t *pi, const int32_t *n0, const lsthm_sps_masks *masks,
                  float *workspace, float *out, float *sGQ, float *sCQ, float *sHQ, float *sXQ, float *sGL,
                  float *sCL, float *sHL, void *stream);

/*
 *   pr   [T][N] int32  packed row of dialogue d (inverse of pi)
 *   dout [T][N][512]   dL/d out
 *   dGL  [T][N][2][512] dL/d(LSTHM gate pre-activations)   -> W,U,V,S grads and dx
 *   dGQ  [T][N][2][512] dL/d(LSTMCell gate pre-activations) -> weight_ih, weight_hh, biases
 *   dWqk [grid][2][128] per-CTA partial sums of dL/dWq, dL/dWk (sum over the first axis); grid from
 *        lsthm_sps_launch_info
 */
int lsthm_sps_bwd(const lsthm_sps_desc *d, const lsthm_sps_weights *w, const float *qmask, const int32_t *pi,
                  const int32_t *pr, const int32_t *n0, const lsthm_sps_masks *masks, const float *dout,
                  const float *sGQ, const float *sCQ, const float *sGL, const float *sCL, float *workspace,
                  float *dGL, float *dGQ, float *dWqk, void *stream);

int lsthm_sps_launch_info(const lsthm_sps_desc *d, int32_t *grid, int32_t *block, int32_t *rows, int32_t *smem_fwd,
                          int32_t *smem_bwd);

/* ------------------------------------------------------------------------------------------
 * lsthm_gsp: GRU speaker-state LSTHM cell — one direction of MARN1_onlysp (listener = 0) and of
 * MARN1_nsps (listener = 1).  Replaces MARN_cell.forward of model/lsthm_onlysp.py:156-188 and of
 * model/lsthm_nsps.py:160-198 (the default model of train.py):
 *   idx = argmax(qmask_t[d]) (an all-zero padded row selects party 0);  qs = q[d][idx]
 *   hs  = dropout(GRUCell(U_t, qs))     gate order r|z|n;  n = tanh(W_in U + b_in + r (W_hn qs + b_hn))
 *   listener = 0:  q[d][p] = q[d][p] (1 - m_p) + hs m_p           (lsthm_onlysp.py:180-182)
 *   listener = 1:  q[d][p] = q[d][1 - idx] (1 - m_p) + hs m_p     (lsthm_nsps.py:184-188)
 *   (c_l,h_l) = LSTHM1_l(x_l, c_l,h_l,z_l, hs), same for a; z_l = CrossAttention(c_l, c_a) (rank-1 form)
 *   out = [h_l | h_a | z_l | hs]
 * Dialogues are independent: ordinary launch, any N, no workspace.  The input-side GRU product W_ih U + b_ih is
 * time-parallel and supplied by the caller (gxs), like gx.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t T, N;
    int32_t rows_per_cta;      /* 0 = automatic (ceil(N/148) clamped to 1..8) */
    int32_t listener;          /* 0 = lsthm_onlysp party update, 1 = lsthm_nsps (listener state copied from the other party) */
    float att_p;               /* in-kernel dropout on the in-cell attention weights (0 = off) */
    uint64_t att_seed;
} lsthm_gsp_desc;

typedef struct {
    const float *U[2], *V[2], *S[2];   /* marn_cell.lsthm_{l,a}.{U,V,S}.weight [512][128]   lsthm_onlysp.py:17-20 */
    const float *Whh;                  /* marn_cell.gru_s.weight_hh [384][128]              lsthm_onlysp.py:152  */
    const float *bhh;                  /* marn_cell.gru_s.bias_hh [384]                                          */
    const float *Wq, *Wk;              /* marn_cell.crossatt_l2a.{Wq,Wk} [128]              lsthm_onlysp.py:50-51 */
} lsthm_gsp_weights;

typedef struct {               /* dropout masks already scaled by 1/(1-p); any may be NULL (= no dropout) */
    const float *ms;           /* on the speaker state hs   [T][N][128]       lsthm_onlysp.py:176 */
    const float *ml, *ma;      /* on h_l / h_a              [T][N][128]       lsthm_onlysp.py:184,186 */
    const float *att_mask;     /* on the attention weights  [T][N][128][128]  lsthm_onlysp.py:63 */
} lsthm_gsp_masks;

size_t lsthm_gsp_packed_floats(void);
int lsthm_gsp_pack(const lsthm_gsp_weights *w, float *packed, void *stream);

/*
 *   gx    [T][N][2][512]  W_c x_c + (bW+bU+bV+bS)_c for c = l, a; gate order f|i|o|g
 *   gxs   [T][N][384]     gru_s.weight_ih U_t + bias_ih; gate order r|z|n
 *   qmask [T][N][2]       one-hot current speaker (zero rows on padding)
 *   out   [T][N][512]     [h_l | h_a | z_l | hs]
 *   stash (all or none): sGS [T][N][4][128] r|z|n|(W_hn qs + b_hn), sQS [T][N][128] qs,
 *         sGL [T][N][2][512] LSTHM gates f|i|o|g, sCL [T][N][2][128] cell states
 */
int lsthm_gsp_fwd(const lsthm_gsp_desc *d, const lsthm_gsp_weights *w, const float *packed, const float *gx,
                  const float *gxs, const float *qmask, const lsthm_gsp_masks *masks, float *out, float *sGS,
                  float *sQS, float *sGL, float *sCL, void *stream);

/*
 *   dout [T][N][512]    dL/d out
 *   dGL  [T][N][2][512] dL/d(LSTHM gate pre-activations)       -> W,U,V,S grads and dx
 *   dGi  [T][N][384]    dL/d(W_ih U + b_ih)                    -> weight_ih, bias_ih, dU
 *   dGh  [T][N][384]    dL/d(W_hh qs + b_hh)                   -> weight_hh (with sQS), bias_hh
 *   dWqk [grid][2][128] per-CTA partial sums of dL/dWq, dL/dWk (grid from lsthm_gsp_launch_info)
 */
int lsthm_gsp_bwd(const lsthm_gsp_desc *d, const lsthm_gsp_weights *w, const float *qmask, const lsthm_gsp_masks *masks,
                  const float *dout, const float *sGS, const float *sQS, const float *sGL, const float *sCL, float *dGL,
                  float *dGi, float *dGh, float *dWqk, void *stream);

int lsthm_gsp_launch_info(const lsthm_gsp_desc *d, int32_t *grid, int32_t *block, int32_t *rows, int32_t *smem_fwd,
                          int32_t *smem_bwd);

/* ------------------------------------------------------------------------------------------
 * fp32-accurate dense GEMM on the tcgen05 tensor cores (split-bf16, three UMMAs per k-step) for the
 * time-parallel products of the path: the hoisted input projections W x (LSTHM.forward line 23 of
 * model/HybridRNN_ATV.py, LSTHM1.forward line 29 of model/lsthm_sps.py), the nn.Linear layers of
 * model/encoder.py and of the heads, and the weight-gradient products autograd forms for them.
 *   mode 0 (NT): C[M][N] = A[M][K] . B[N][K]^T (+ bias[N])     y  = x W^T + b   (torch.nn.functional.linear)
 *   mode 1 (NN): C[M][N] = A[M][K] . B[K][N]                   dx = dy W
 *   mode 2 (TN): C[M][N] = A[K][M]^T . B[K][N]                 dW = dy^T x      (split-K, deterministic reduce)
 *   mode 3     : mode 0 followed by ReLU in the epilogue       relu(x W^T + b)  (encoder.py:108, w_1 of the FFN)
 * All matrices fp32 row-major, 16-byte aligned, lda/ldb multiples of 4.  `workspace` (may be NULL) holds
 * the split-K partials: lsthm_gemm3_workspace_floats(mode, M, N, K) floats.
 * ------------------------------------------------------------------------------------------ */
/* OR this into `mode` for the bf16 mode of BASELINE.json (stated separately from fp32 parity): operands are rounded
 * to bf16 while they are staged and each k-step is ONE UMMA (fp32 accumulation) instead of the three split terms. */
#define LSTHM_GEMM_BF16 0x10
/* OR this into `mode` (lsthm_gemm3 only) for the SIX-term product: operands split three ways (24 mantissa bits), terms
 * hi.hi + hi.mid + mid.hi + hi.lo + lo.hi + mid.mid — fp32-grade accuracy even when the sum cancels structurally, as in
 * the all-ones projections of a LayerNorm output in CrossAttention2/3 (model/lsthm_sps.py:82-84, 91-93). */
#define LSTHM_GEMM_X6 0x20
size_t lsthm_gemm3_workspace_floats(int32_t mode, int32_t M, int32_t N, int32_t K);
int lsthm_gemm3(int32_t mode, int32_t M, int32_t N, int32_t K, const float *A, int32_t lda, const float *B, int32_t ldb,
                const float *bias, float *C, int32_t ldc, float *workspace, size_t workspace_floats, void *stream);

/* The same product for modes 0, 1 and 3 when B is a layer WEIGHT W and M = T*N is large: W is split once into
 * bf16 hi/lo tile images in the UMMA shared-memory layout (`pack`, lsthm_gemm3w_pack_bytes(N, K) bytes, written by
 * the call), the main kernel converts only the activation operand, fetches weight tiles with bulk async copies and
 * owns 128 x 256 output tiles.  Same numerics as lsthm_gemm3. */
size_t lsthm_gemm3w_pack_bytes(int32_t N, int32_t K);
int lsthm_gemm3w(int32_t mode, int32_t M, int32_t N, int32_t K, const float *A, int32_t lda, const float *W, int32_t ldw,
                 const float *bias, float *C, int32_t ldc, void *pack, size_t pack_bytes, void *stream);

/* MARN1_sps._reverse_seq, model/lsthm_sps.py:396-410 (same in lsthm_onlysp.py / lsthm_nsps.py): per-dialogue flip over its own
 * length with zero padding, X [L][B][w] -> out [L][B][w], len [B] int32 (= sum of umask rows); w even.  Self-adjoint: the
 * backward is the same call on the gradient. */
int lsthm_reverse_seq(int32_t L, int32_t B, int32_t w, const float *X, const int32_t *len, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * MaskedLoss.forward, loss.py:13-21 (weight = None):  loss = sum_r L(pred[r] * mask[r], target[r]) / sum(mask), L = cross
 * entropy (kind 0: nn.CrossEntropyLoss, the train.py default) or NLL (kind 1), and its autograd backward.  A padded row
 * contributes the constant log C to the cross entropy exactly as in the reference.  pred [R][C] fp32 contiguous (C <= 32),
 * target [R] int64, mask [R] fp32.  out2[0] = loss, out2[1] = sum(mask) (kept for the backward); gout = dL/d loss (device
 * scalar); workspace: lsthm_masked_loss_workspace_floats(R) floats.  Row sums are reduced in a fixed order (deterministic).
 * ------------------------------------------------------------------------------------------ */
size_t lsthm_masked_loss_workspace_floats(int64_t R);
int lsthm_masked_loss_fwd(int64_t R, int32_t C, int32_t kind, const float *pred, const int64_t *target, const float *mask,
                          float *workspace, float *out2, void *stream);
int lsthm_masked_loss_bwd(int64_t R, int32_t C, int32_t kind, const float *pred, const int64_t *target, const float *mask,
                          const float *out2, const float *gout, float *dpred, void *stream);

/* ------------------------------------------------------------------------------------------
 * Sequence-level cross-modal attention core (tcgen05, split-bf16, fp32 accuracy): what CrossAttention2.forward and
 * CrossAttention3.forward do after their three projections — model/lsthm_sps.py:94-99, 122-127 (same code in
 * lsthm_onlysp.py) and model/lsthm_nsps.py:100-104:
 *     out = dropout(softmax(q k^T * scale)) v     per dialogue, unmasked over its L <= 128 utterances, one head,
 * width D (multiple of 4, <= 128), and its autograd backward.  Scores, probabilities and the dropout mask stay on the SM.
 * Row i of dialogue b of a matrix X with row stride ldx: X + (b*row_stride_b + i*row_stride_i)*ldx (both 0 = batch-major).
 * dq/dk/dv use the strides of q/k/v; out/dout use ldo.  lse [B][L] (optional, may be NULL): each query row's log-sum-exp of the
 * scaled scores in log2 units, a diagnostic output of the forward.  The backward needs neither `out` nor `lse`: it recomputes
 * the row statistics from its own score product so that p, delta and dS come from the same rounded values.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t B, L, D;
    int32_t ldq, ldk, ldv, ldo;
    float scale;               /* 1/sqrt(d_k)                                   lsthm_sps.py:97 */
    float p_drop;              /* dropout on the attention weights (0 = off)    lsthm_sps.py:98 */
    uint64_t seed;
    int64_t row_stride_b, row_stride_i;
} lsthm_xattn_desc;

int lsthm_xattn_fwd(const lsthm_xattn_desc *d, const float *q, const float *k, const float *v, float *out, float *lse,
                    void *stream);
int lsthm_xattn_bwd(const lsthm_xattn_desc *d, const float *q, const float *k, const float *v, const float *dout, float *dq,
                    float *dk, float *dv, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-head self-attention of the utterance encoder (tcgen05, split-bf16, fp32 accuracy).
 * Replaces ScaledDotProductAttention.forward, model/encoder.py:71-86, as called by
 * MultiHeadAttention.forward (:27-60) with mask=None: per head  softmax(q k^T * scale) -> dropout -> . v
 * over the L <= 128 utterances of a dialogue, d_k = d_v = 40; and its autograd backward.
 * Row i of dialogue b, head h of a matrix X with row stride ldx (floats):
 *   X + (b*row_stride_b + i*row_stride_i)*ldx + h*40      (both strides 0 = batch-major [B][L]: b*L + i;
 *   the reference's time-major activations [L][B][.] are row_stride_b = 1, row_stride_i = B: no permute copy).
 * q/k/v may be three column blocks of one fused projection output (ldq = ldk = ldv = 3*H*40).
 * Dropout is generated in-kernel from (seed, b, h, i, j); the same call arguments regenerate it in bwd.
 * `lse` [B*H][L] receives each query row's log-sum-exp of the scaled scores (log2 units) in the forward and
 * must be handed back to the backward, which then needs a single pass over the score tile.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t B, L, H, d_head;          /* d_head must be 40                                  */
    int32_t ldq, ldk, ldv, ldo;       /* row strides (floats), multiples of 4               */
    float scale;                      /* 1 / temperature = 1/sqrt(d_k)  (encoder.py:22)     */
    float p_drop;                     /* attention dropout (encoder.py:66), 0 in eval mode  */
    uint64_t seed;
    int64_t row_stride_b, row_stride_i; /* in rows; see above                                */
    int32_t precision;                /* 0: fp32-accurate (operands split in bf16 hi + lo, three UMMAs per k-step);
                                         1: bf16 mode (operands rounded to bf16, one UMMA per k-step, fp32 softmax) */
    int32_t reserved;
} lsthm_attn_desc;

int lsthm_attn_fwd(const lsthm_attn_desc *d, const float *q, const float *k, const float *v, float *out, float *lse,
                   void *stream);
/* dq/dk/dv use the row strides ldq/ldk/ldv; out/dout use ldo */
int lsthm_attn_bwd(const lsthm_attn_desc *d, const float *q, const float *k, const float *v, const float *out,
                   const float *lse, const float *dout, float *dq, float *dk, float *dv, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fused  out = LayerNorm(dropout(y + bias) + residual)  over R rows of width d (4 <= d <= 512, d % 4 == 0), forward and
 * backward.  Replaces the tail of MultiHeadAttention.forward, model/encoder.py:54-58
 * (`q = self.dropout(self.fc(q)); q += residual; q = self.layer_norm(q)`) and of PositionwiseFeedForward.forward,
 * model/encoder.py:106-112.  All matrices fp32 with unit inner stride and row strides (floats) that are
 * multiples of 4; rows may be in any order (batch- or time-major).
 *   fwd:  v = dropout(y + bias) + res (kept for the backward, may be NULL in eval; bias [d] may be NULL),  out = LN(v) * gamma + beta
 *   bwd:  dres = dL/dres (= dL/dv),  dy = dL/dy (= dres * dropout scale; not written when p_drop == 0: it equals
 *         dres),  dgamma, dbeta [d],  dbias [d] = column sums of dy (bias gradient of the Linear that made y);
 *         any of the three may be NULL.  `workspace`: lsthm_dln_workspace_floats(d) floats.
 * Dropout is regenerated in the backward from (seed, row, column).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t R;
    int32_t d;
    float eps;                        /* LayerNorm eps (encoder.py:25: 1e-6)                 */
    float p_drop;
    uint64_t seed;
} lsthm_dln_desc;

size_t lsthm_dln_workspace_floats(int32_t d);
int lsthm_dln_fwd(const lsthm_dln_desc *d, const float *y, int32_t ldy, const float *bias, const float *res, int32_t ldres, const float *gamma,
                  const float *beta, float *v, int32_t ldv, float *out, int32_t ldo, void *stream);
int lsthm_dln_bwd(const lsthm_dln_desc *d, const float *dout, int32_t lddo, const float *v, int32_t ldv, const float *gamma,
                  float *dy, int32_t lddy, float *dres, int32_t lddres, float *dgamma, float *dbeta, float *dbias,
                  float *workspace, size_t workspace_floats, void *stream);

/* Column sums out[C] = sum_r A[r][c] of a row-major fp32 matrix (C % 4 == 0, ld % 4 == 0): the bias gradients
 * autograd forms as dy.sum(0) for every nn.Linear on the path, over all T*N rows.  Deterministic (fixed order).
 * `workspace`: lsthm_colsum_workspace_floats(R, C) floats. */
size_t lsthm_colsum_workspace_floats(int64_t R, int32_t C);
int lsthm_colsum(int64_t R, int32_t C, const float *A, int32_t ld, float *out, float *workspace, size_t workspace_floats,
                 void *stream);

/* Batch assembly of the reference trainer, model_trainer.py:104-105 (`textf = (r1 + r2 + r3 + r4) / 4` followed by
 * `torch.cat((textf, acouf), dim=-1)`), in one pass over R = L*B utterances: r1..r4 [R][d_text] are the four RoBERTa
 * layers of dataloader.py:29-32, acouf [R][d_audio], x [R][d_text + d_audio].  All contiguous fp32, widths % 4 == 0.
 * Same operation order as the reference, so the result is bit-identical. */
int lsthm_assemble_input(int64_t R, int32_t d_text, int32_t d_audio, const float *r1, const float *r2, const float *r3, const float *r4,
                         const float *acouf, float *x, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fused Adam step on a flat fp32 buffer.  Replaces `self.optim.step()` of the reference's trainer
 * (model_trainer.py:82,120: torch.optim.Adam(lr, weight_decay=2e-5), L2-style decay, no amsgrad) for one
 * gradient bucket; `step` counts from 1 (bias correction).  Parameters whose gradient is None in the
 * reference (never-used ones) must simply not be part of the buffers.
 * ------------------------------------------------------------------------------------------ */
int lsthm_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, size_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LSTHM_B200_H_ */
