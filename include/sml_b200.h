/* sml_b200.h — C-ABI of libsml_b200.so: the B200 (sm_100a) kernels behind the
 * Subspace-Multimodal-Learning multiple-instance fusion hot path.
 *
 * The reference is pure Python (no FFI of its own); each entry point below is
 * the unit a reference-side binding replaces, cited as /root/reference file:line.
 * Conventions:
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller owns all memory (inputs, outputs, workspaces); nothing is allocated
 *     or freed behind the ABI and no call synchronises the device;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, non-zero = error, message via sml_last_error()
 *     (thread-local); no exceptions cross the boundary;
 *   - entries are re-entrant (forward runs on the caller thread, backward on the
 *     autograd thread).
 */
#ifndef SML_B200_H
#define SML_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* sml_last_error(void);
int sml_version(void);
/* number of kernels launched through this library by the calling process (bench `gpu_launches`) */
int64_t sml_launch_count(void);
/* Dropout salt for CUDA-graph capture of a train step (train_test.py:57-85 as one replayable graph).  Every dropout mask is a
 * counter-based hash of a (seed, offset) pair passed by value, which a captured graph would freeze; with a salt installed each
 * mask kernel xors its seed with the 64-bit word salt2[1] read on the device at run time.  sml_seed_salt_advance moves
 * {salt2[0] counter, salt2[1] salt} to the next value (captured as the first node of the step graph: fresh masks per replay, the
 * forward and backward kernels of one step still agree).  sml_set_seed_salt(NULL) (default) leaves seeds as passed. */
int sml_set_seed_salt(const void* salt2);
int sml_seed_salt_advance(void* salt2, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense contractions.  One descriptor drives both engines:
 *   sml_gemm_f32 : SIMT fp32 FFMA path (exact-fp32 reference path on the device; any strides)
 *   sml_gemm_tc  : tcgen05.mma + TMEM accumulators + TMA-fed smem operands (bf16 operands,
 *                  fp32 accumulate; optional bf16x3 split operands for fp32-grade results)
 * C[b1,b2] = alpha * A[b1,b2] (MxK) * B[b1,b2] (KxN) + beta * C + diag * I  (+ bias[n]) (relu)
 * Replaces every torch matmul / einsum / nn.Linear / 1x1 conv on the path:
 *   models/NystromAttention.py:89,122-125,140,150 ; :20-35 (pinv chain, `diag`)
 *   models/DeformableAttention1D.py:175,199,211,230,233 ; models/DeformableAttention2D.py:246,279,292,311,314
 *   models/DeformCrossTransMIL.py:83,35-37 ; models/MultiheadAttention.py:125-182,282,303-306 ; models/mil.py:214,219
 * ------------------------------------------------------------------------------------- */
typedef struct sml_gemm_desc {
  int32_t M, N, K;
  int32_t nb1, nb2;                 /* two-level batch (e.g. slides x heads) */
  const void* A;  int64_t sa_m, sa_k, sa_b1, sa_b2;   /* element strides; A is [M,K] */
  const void* B;  int64_t sb_k, sb_n, sb_b1, sb_b2;   /* B is [K,N] */
  void*       C;  int64_t sc_m, sc_n, sc_b1, sc_b2;   /* fp32 output (may be NULL for tc if Cb given) */
  float alpha, beta, diag;
  const float* bias;                /* [N] or NULL */
  int32_t relu;
  /* tensor-core engine only: */
  const void* A_lo; const void* B_lo;  /* bf16 low parts for the bf16x3 split (NULL = single pass) */
  void* Cb_hi; void* Cb_lo;            /* optional bf16 copies of the result (same strides as C) */
  /* optional row-segment broadcast term added in the epilogue: + rowtab_scale * rowtab[b1,b2][m / rowdiv][n]
   * (the landmark-mean backward dq[t] += dq_landmark[t / l] / l of models/NystromAttention.py:102-118, fused) */
  const float* rowtab; int64_t rt_m, rt_b1, rt_b2; int32_t rowdiv; float rowtab_scale;
} sml_gemm_desc;

int sml_gemm_f32(const sml_gemm_desc* d, void* stream);
int sml_gemm_tc(const sml_gemm_desc* d, void* stream);
/* n DEPENDENT bf16x3 GEMMs (each step may read what earlier steps wrote: the Moore-Penrose chain of
 * models/NystromAttention.py:31-33 and its backward) in one persistent cooperative launch with in-kernel grid barriers;
 * counter = 4 bytes of device scratch */
int sml_gemm_tc_program(const sml_gemm_desc* descs, int32_t n, void* counter, void* stream);
/* fp32 -> bf16 hi (and lo = bf16(x - hi) when lo != NULL), n elements */
int sml_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream);
/* same on rows: x [B, n, D] fp32 -> hi / lo [B, npad = n + pad, D] with `pad` zero rows in front of every batch (the front
 * padding of models/NystromAttention.py:79-82 applied to the incoming gradient without a padded fp32 copy) */
int sml_split_bf16_pad(const float* x, void* hi, void* lo, int64_t B, int64_t n, int64_t npad, int64_t pad, int32_t D, void* stream);

/* ---------------------------------------------------------------------------------------
 * Row-wise kernels shared by the path.
 * ------------------------------------------------------------------------------------- */
/* LayerNorm over the last dim D of x[R,D] -> y rows written with row stride y_ld; saves mean/rstd.
 * (models/mil.py:186, models/DeformCrossTransMIL.py:66,127, nn.LayerNorm eps=1e-5) */
/* row r of x ([R,D], R = batches*n) is written to row (r/n)*npad + pad + r%n of y (front-padded layout of
 * models/NystromAttention.py:79-82; pass n=R, npad=R, pad=0 for a plain LayerNorm) */
int sml_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd,
                      int64_t R, int32_t D, float eps, int64_t n, int64_t npad, int64_t pad, void* y_hi, void* y_lo,
                      void* stream);   /* y_hi/y_lo: optional bf16 (hi, residual) copies of y for the tcgen05 engine */
/* dx = LN'(dy); dw_part/db_part are [nblk,D] partial sums (nblk = sml_layernorm_bwd_blocks(R)) */
int sml_layernorm_bwd_blocks(int64_t R);
int sml_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                      float* dx, float* dw_part, float* db_part, int64_t R, int32_t D, int64_t n, int64_t npad,
                      int64_t pad, const float* res, void* stream);   /* dy is read in the padded layout, dx written dense;
                      res (optional, [R,D]) is added to dx: the gradient of the skip connection of x + f(LN(x)) (mil.py:186) */
/* y[r,:] = softmax(x[r,:]) in place over C columns (models/NystromAttention.py:137) */
/* hi/lo: optional bf16 copies; write_f32 = 0 leaves x untouched (probabilities kept in bf16 only) */
int sml_softmax_fwd(float* x, int64_t R, int64_t C, void* hi, void* lo, int32_t write_f32, void* stream);
/* in place on dA: dS = A * (dA - sum(dA*A)) */
/* probabilities from A (fp32) or A_bf16; dS to dA in place (write_f32) and/or to the bf16 copies */
int sml_softmax_bwd(const float* A, const void* A_bf16, float* dA, int64_t R, int64_t C, void* hi, void* lo, int32_t write_f32,
                    void* stream);
/* y = x + dropout(u, p) with a counter-based generator (seed, offset); p == 0 -> plain add.
 * Same call with grad tensors regenerates the mask in backward (models/mil.py:183 dropout=0.1) */
int sml_dropout_add(const float* x, const float* u, float* y, int64_t n, float p, uint64_t seed, uint64_t offset,
                    void* stream);
int sml_axpby(float* y, const float* x, float a, float b, int64_t n, void* hi, void* lo, void* stream);   /* y = a*x + b*y */
/* out = a*x + dg*I over nmat square m x m matrices (identity terms of models/NystromAttention.py:33) */
int sml_diag_axpby(float* out, const float* x, float a, float dg, int64_t nmat, int32_t m, void* hi, void* lo, void* stream);
/* dx = dy where y > 0 else 0 (ReLU of models/mil.py:214, models/DeformCrossTransMIL.py:83) */
int sml_relu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream);
/* y[b,t,:] = (x ? x[b,t,:] : 0) + a*vec[b,:]  — the rank-1 half of FusionNet (models/DeformCrossTransMIL.py:104,33-38:
 * the repeated omic vector is never materialised) and the backward of the token mean (Pooler :193) */
int sml_add_rowvec(float* y, const float* x, const float* vec, float a, int64_t B, int64_t T, int32_t C, void* stream);
/* 1-D bias MLP as an exact piecewise-linear table of t = sign(dx) log(|dx| + 1) (models/DeformableAttention1D.py:60-102; cpb =
 * the packed parameters of sml_deform_attn_fwd with CD = 1): brk [sml_cpb1d_table_slots()] ascending breakpoints (+inf
 * padded), coef [(slots + 1), OUT, 2] = (slope, intercept) per segment and output, nseg [1].  eval: out [n, OUT]. */
int sml_cpb1d_table_slots(void);
int sml_cpb1d_table_build(const float* cpb, int32_t OUT, float* brk, float* coef, int32_t* nseg, void* stream);
int sml_cpb1d_table_eval(const float* t, int64_t n, const float* brk, const float* coef, const int32_t* nseg, int32_t OUT,
                         float* out, void* stream);
/* 1-D fused deformable attention with the bias taken from that table instead of the per-pair MLP (CD = 1;
 * models/DeformableAttention1D.py:211-231 with :60-102 tabulated).  q fp32 [B, n, heads 64]; kh, vh = bf16 planes of K and V
 * [B, nkv, heads 64] (the "bf16" mode's operand rounding, applied once by the caller); vs [(B G), nkv] key positions;
 * krange [sml_deform_pwl_krange_floats(cfg8)] = (min, max) key position of every 32-key tile, filled by
 * sml_deform_pwl_key_ranges once per forward and reused by the backward. */
int64_t sml_deform_pwl_krange_floats(const int32_t* cfg8);
int sml_deform_pwl_key_ranges(const int32_t* cfg8, const float* vs, float* krange, void* stream);
int sml_deform_attn_fwd_pwl(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const void* kh, const void* vh,
                            const float* vs, const float* krange, const float* brk, const float* coef, const int32_t* nseg,
                            float* out, float* lse, void* stream);
/* Backward.  dq is written; dk, dv, dvs and gseg [OUT, sml_cpb1d_seg_rows(), 2] are accumulated into (zero them first).
 * gseg[o][s] = (sum dS, sum dS p) over the pairs of the heads with output o whose signed-log distance p lies in segment s
 * of the table; sml_cpb1d_table_bwd maps the sums to the gradient of the MLP parameters (dcpb [sml_deform_cpb_nparams(1,
 * OUT)], overwritten; specification oracle/pwl.py:param_grads_from_segments). */
int sml_cpb1d_seg_rows(void);
int sml_deform_attn_bwd_pwl(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const void* kh, const void* vh,
                            const float* vs, const float* krange, const float* brk, const float* coef, const int32_t* nseg,
                            const float* out, const float* lse, const float* dout, float* dq, float* dk, float* dv, float* dvs,
                            float* gseg, void* stream);
int sml_cpb1d_table_bwd(const float* cpb, int32_t OUT, const float* brk, const int32_t* nseg, const float* gseg, float* dcpb,
                        void* stream);
/* Gradient modulation of the fusion classifier after backward (train_test.py:87-184; diag2021 / grade / subtype tasks):
 * W, Wgrad [L, 2 hs], bias [L], feat_t / feat_i [B, hs] (the two branch embeddings), labels [B] int64.  Rewrites the
 * conflicting half of every row of Wgrad in place; info (optional) [4] = (score_t, score_i, #rows re-projected on the tumor
 * half, on the immune half).  One kernel, no host synchronisation (the reference syncs once per classifier row). */
int sml_grad_modulate(const float* W, const float* bias, float* Wgrad, const float* feat_t, const float* feat_i,
                      const int64_t* labels, int32_t B, int32_t L, int32_t hs, float* info, void* stream);
/* torch.nn.AlphaDropout of the MaxNet encoders (models/model.py:152-170), same counter-based mask as sml_dropout_add;
 * bwd = 0: y = alpha_dropout(x); bwd = 1: y = d x given x = d y (same seed / offset) */
int sml_alpha_dropout(const float* x, float* y, int64_t n, float p, int32_t bwd, uint64_t seed, uint64_t offset, void* stream);
/* pointwise heads: op 0 tanh (Pooler :200), 1 sigmoid, 2 relu (fusion.py:38-50), 3 elu (model.py:150) */
int sml_unary_fwd(int32_t op, const float* x, float* y, int64_t n, void* stream);
int sml_unary_bwd(int32_t op, const float* x, const float* y, const float* dy, float* dx, int64_t n, void* stream);
/* 1 if any tcgen05 kernel hit its bounded mbarrier wait (protocol error); reads a device flag (synchronises) */
int sml_tc_timeout_flag(void);
/* out[c] = sum_r x[r,c] + beta*out[c]; part is a [sml_colsum_parts(R), C] workspace */
int sml_colsum_parts(int64_t R);
int sml_colsum(const float* x, float* out, float* part, int64_t R, int64_t C, float beta, void* stream);

/* ---------------------------------------------------------------------------------------
 * Nystrom attention pieces (models/NystromAttention.py:74-157, :20-35).
 * qkv layout: [B, n', 3, h, d] fp32 (row = token, q|k|v chunks, head-major inside a chunk).
 * ------------------------------------------------------------------------------------- */
/* landmark segment means of q and k: ql, kl [B,h,m,d]; landmark id of token t is t / l (:102-118).
 * Here and in the two resconv entries: qkv == NULL -> read the bf16 copy `qkv_bf16` instead (bf16 mode writes no fp32 qkv). */
int sml_nystrom_landmarks_fwd(const float* qkv, float* ql, float* kl, int32_t B, int32_t npad, int32_t h, int32_t d,
                              int32_t m, const void* qkv_bf16, void* stream);
/* dqkv[q,k chunks][t] += d{ql,kl}[t / l] / l */
int sml_nystrom_landmarks_bwd(float* dqkv, const float* dql, const float* dkl, int32_t B, int32_t npad, int32_t h,
                              int32_t d, int32_t m, void* dqkv_hi, void* dqkv_lo, void* stream);
/* integer landmark ids (bit-exact contract): ids[t] = t / (npad / m) */
int sml_nystrom_landmark_ids(int32_t* ids, int32_t npad, int32_t m, void* stream);
/* scal[0] = max over all (b,h,i) of sum_j |x|, scal[1] = max over all (b,h,j) of sum_i |x| (global, F8);
 * arg[0] = flat (bh*m + i) of the row max, arg[1] = flat (bh*m + j) of the column max.  x [BH,m,m] (:22-26) */
int sml_pinv_scale(const float* x, float* scal, int32_t* arg, void* packed_ws /*16 B*/, int32_t BH, int32_t m,
                   void* stream);
/* z0 = x^T / (scal[0]*scal[1]) (:26) */
int sml_pinv_init(const float* x, const float* scal, float* z0, int32_t BH, int32_t m, void* stream);
/* backward of pinv_init: dx += dz0^T / s ; then the scalar path d s -> row/col max positions */
int sml_pinv_init_bwd(const float* x, const float* scal, const int32_t* arg, const float* dz0, float* dx,
                      float* tmp /*4 B*/, int32_t BH, int32_t m, void* stream);
/* depthwise conv along tokens on the v chunk, per head, shared over d: out[B,n',h*d] += conv(v) (:144-145) */
int sml_nystrom_resconv_fwd(const float* qkv, const float* w, float* out, int32_t B, int32_t npad, int32_t h,
                            int32_t d, int32_t ksize, void* out_hi, void* out_lo, const void* qkv_bf16, void* stream);
/* dqkv[v chunk] += conv^T(dout); dw[h,k] = sum dout*v (dw_part [nblk,h,k]; nblk returned by ..._blocks) */
int sml_nystrom_resconv_bwd_blocks(int32_t B, int32_t npad, int32_t hd);
int sml_nystrom_resconv_bwd(const float* qkv, const float* w, const float* dout, float* dqkv, float* dw_part,
                            int32_t B, int32_t npad, int32_t h, int32_t d, int32_t ksize, void* dqkv_hi, void* dqkv_lo,
                            const void* qkv_bf16, void* stream);
/* Fused backward of out = softmax(q k~^T) W for the token-by-landmark path (models/NystromAttention.py:122,137,140):
 * per 128-token tile dA1 = dout W^T (TMEM), dS = A1 (dA1 - rowsum(dA1 A1)) in registers, dq = dS k~ (+ dql[t / l] / l),
 * dk~ += dS^T q, without materialising dA1 / dS.  bf16 operands (dout [B,n',h*d], W [B,h,m,d], A1 [B,h,n',m],
 * k~ [B,h,m,d], qkv [B,n',3,h,d]); writes the q chunk of dqkv (fp32 and / or bf16, either may be NULL) and atomically adds into dkl.
 * Specialised to d = 64, m = 256. */
int sml_nystrom_a1_bwd(const void* dout_bf16, const void* wm_bf16, const void* a1_bf16, const void* kl_bf16,
                       const void* qkv_bf16, const float* dql, float* dqkv, void* dqkv_hi, float* dkl, int32_t B,
                       int32_t npad, int32_t h, int32_t d, int32_t m, void* stream);
/* Landmark-by-token product without its probability matrix (models/NystromAttention.py:124, 130-137, 140: sim3, attn3,
 * attn3 @ v), "bf16" mode, 256 landmarks, head dim 64.  cfg11 = (B, H, 256, 64, n', row stride, batch stride lo, hi, k offset,
 * v offset, nchunks): K / V are bf16 views into the packed qkv tensor [B, n', 3, H, 64] (element strides / offsets).
 * fwd: opart [B H nchunks 256 64] and ml [B H nchunks 256 2] fp32 scratch; a3v [B, H, 256, 64] fp32 (+ optional bf16 copy),
 * lse [B, H, 256] row log-sum-exp (saved for the backward).
 * bwd: dql [B, H, 256, 64] accumulated; dk -> bf16 K slice of dqkv_bf16 (+ dkl[t / lsz] / lsz if dkl is given, else call
 * sml_nystrom_a3_dk_landmarks once dkl is complete); dv -> fp32 V slice of dqkv_f32. */
int sml_nystrom_a3_fwd(const int32_t* cfg11, const void* ql_bf16, const void* qkv_bf16, float* opart, float* ml, float* a3v,
                       void* a3v_hi, float* lse, void* stream);
int sml_nystrom_a3_bwd(const int32_t* cfg11, const void* ql_bf16, const void* qkv_bf16, const float* da3v, const float* a3v,
                       const float* lse, const float* dkl, int32_t lsz, float* dql, void* dqkv_bf16, float* dqkv_f32, void* stream);
int sml_nystrom_a3_dk_landmarks(const int32_t* cfg11, const float* dkl, int32_t lsz, void* dqkv_bf16, void* stream);
/* Inference only (no saved probabilities): out [B, n', H 64] fp32 = softmax_j(q_t . kl_j) wm_j per head, q = the packed bf16 qkv
 * tensor at offset 0, kl / wm bf16 [B, H, m, 64], m a multiple of 64 (models/NystromAttention.py:122, 130, 140). */
int sml_nystrom_a1_fwd_nograd(const int32_t* cfg11, const void* qkv_bf16, const void* kl_bf16, const void* wm_bf16, float* out,
                              void* stream);
/* PPEG parameter folding and its backward: w7 [C,49], w5 [C,25], w3 [C,9] depthwise kernels + biases [C] (models/mil.py:195-198)
 * -> wf [49,C] tap-major stencil incl. the identity, bf [C]; unfold scatters d wf back to the three kernels (d bias is shared) */
int sml_ppeg_fold(const float* w7, const float* w5, const float* w3, const float* b7, const float* b5, const float* b3,
                  float* wf, float* bf, int32_t C, void* stream);
int sml_ppeg_unfold(const float* dwf, float* dw7, float* dw5, float* dw3, int32_t C, void* stream);
/* PPEG: identity + depthwise 7x7 + 5x5 + 3x3 folded into one 7x7 stencil, token-major [B,1+H*W,C];
 * the cls row is copied through (models/mil.py:192-206).  w7 [49,C] folded tap-major weights, bias [C] folded.
 * Backward: dw_part [nblk,49,C], db_part [nblk,C]. */
int sml_ppeg_fwd(const float* x, const float* w7, const float* bias, float* y, int32_t B, int32_t H, int32_t W,
                 int32_t C, void* stream);
int sml_ppeg_bwd_blocks(int32_t B, int32_t H, int32_t W);
int sml_ppeg_bwd(const float* x, const float* w7, const float* dy, float* dx, float* dw_part, float* db_part,
                 int32_t B, int32_t H, int32_t W, int32_t C, void* stream);


/* ---------------------------------------------------------------------------------------
 * Deformable cross-attention 1-D / 2-D (models/DeformableAttention1D.py:156-240,
 * models/DeformableAttention2D.py:223-325).  Token-major tensors: q/k/v/out [b, tokens, heads*64],
 * x2/kv [b, tokens, C].  A 1-D sequence is the 2-D case with H = 1 (offset net) / W = 1 (sampling, F5).
 * ------------------------------------------------------------------------------------- */
/* cfg13 = {B, G, gd, inner, H, W, KH, KW, SH, SW, PH, PW, OC}: depthwise conv(k,s,p) -> GELU -> 1x1 (gd->OC) -> tanh
 * -> *scale -> + grid -> normalise (1D.py:139-146,182-188; 2D.py:207-213,255-265).
 * off [(b g), nkv, OC], vgrid [(b g), OC, HK, WK] (returned to the caller as in the reference), vs [(b g), nkv, OC] */
int sml_deform_offsets_fwd(const int32_t* cfg13, float scale, const float* q, const float* wc, const float* bc,
                           const float* w2, float* off, float* vgrid, float* vs, void* stream);
int sml_deform_offsets_bwd_blocks(const int32_t* cfg13);
/* dvs [(b g), nkv, OC] and optional dvgrid (grad of the returned vgrid); dpre [(b g), nkv, gd] workspace;
 * part [nblk, gd*KH*KW + gd + OC*gd] partial parameter grads; dq [b, H*W, inner] is accumulated (+=) atomic-free */
int sml_deform_offsets_bwd(const int32_t* cfg13, float scale, const float* q, const float* wc, const float* bc,
                           const float* w2, const float* off, const float* dvs, const float* dvgrid, float* dpre,
                           float* part, float* dq, void* stream);
/* cfg8 = {B, G, cg, H, W, nkv, CD, mode}: bilinear grid_sample, zeros padding, align_corners=False
 * (1D.py:36-43 with mode 0 = the reference's (x=vgrid, y=0) pad order on an [n,1] image, mode 1 = upstream order;
 * 2D.py:268-271).  taps (optional) [(b g), nkv, 2] int32 = floor of the pixel coordinate (x0, y0): bit-exact contract */
int sml_deform_gather_fwd(const int32_t* cfg8, const float* x2, const float* vs, float* kv, int32_t* taps, void* stream);
/* dx2 must be zeroed by the caller; dvs_out = dvs_in (may be NULL) + sampling-path gradient.  d x2 is scattered without
 * atomics (per-(b,g) shared-memory sort by destination + one warp per run of equal destinations: deterministic) when
 * 4 nkv <= 4096 entries (2-D lattices up to 128 x 128), with fp32 atomics otherwise */
int sml_deform_gather_bwd(const int32_t* cfg8, const float* x2, const float* vs, const float* dkv, float* dx2,
                          const float* dvs_in, float* dvs_out, void* stream);
/* cfg8 = {B, heads, n, nkv, G, CD, W, 64}: out = softmax(q k^T + CPB(q_pos - k_pos)) v with the bias MLP evaluated per
 * pair on the fly (1D.py:60-102,211-231; 2D.py:120-157,292-312); q pre-scaled; dropout on the probabilities with a
 * counter-based mask (seed); cpb = packed [W1 32*CD][b1 32][W2 32*32][b2 32][W3 OUT*32][b3 OUT]; lse [b, heads, n] */
int sml_deform_attn_fwd(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k, const float* v,
                        const float* vs, const float* cpb, float* out, float* lse, void* stream);
int sml_deform_attn_bwd_blocks(const int32_t* cfg8);
int sml_deform_cpb_nparams(int32_t CD, int32_t OUT);
/* dk, dv, dvs zero-initialised by the caller; part [nblk, nparams] CPB partial grads */
int sml_deform_attn_bwd(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k, const float* v,
                        const float* vs, const float* cpb, const float* out, const float* lse, const float* dout,
                        float* dq, float* dk, float* dv, float* dvs, float* part, void* stream);
/* Tensor-core variants of the two entries above (same arguments and contracts + `split`): the 32x32 layer of the bias MLP,
 * q k^T, p v and every backward contraction run as warp-level bf16 MMAs on register fragments.  split = 1: bf16 hi/lo 3-pass
 * products (fp32-grade, 1e-4 tolerance); split = 0: plain bf16 operands (2e-2 tolerance; the "lean" kernels: distances from
 * shared-memory tables, layer 3 and the small bias gradients on the MMA pipe); split = 2: plain bf16, the round-1 kernels. */
int sml_deform_attn_fwd_tc(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k, const float* v,
                           const float* vs, const float* cpb, float* out, float* lse, int32_t split, void* stream);
int sml_deform_attn_bwd_tc(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k, const float* v,
                           const float* vs, const float* cpb, const float* out, const float* lse, const float* dout,
                           float* dq, float* dk, float* dv, float* dvs, float* part, int32_t split, void* stream);
/* Blackwell-native forward of the 2-D case (CD = 2, one head per offset group): same arguments, outputs and dropout mask as
 * sml_deform_attn_fwd_tc(split = 0); the bias MLP, q k^T and p v run as tcgen05.mma with accumulators and the ReLU'd
 * activations in tensor memory (csrc/deform_tc5.cu).  Replaces models/DeformableAttention2D.py:292-312 (+ CPB :120-157). */
int sml_deform_attn_fwd_tc5(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k, const float* v,
                            const float* vs, const float* cpb, float* out, float* lse, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fusion heads, subspace losses and the batch-norm statistics the ranks exchange.
 * ------------------------------------------------------------------------------------- */
int sml_mul(const float* a, const float* b, float* y, int64_t n, void* stream);     /* gates of models/fusion.py:43,50 */
/* sums [2C] = (sum x, sum x^2) over the R local rows: the message of utils/sync_batchnorm/batchnorm.py:104-133
 * and torch.nn.SyncBatchNorm (main.py:118); all-reduced by the caller, then normalised with the global mean/invstd */
int sml_bn_stats(const float* x, float* sums, int32_t R, int32_t C, void* stream);
int sml_bn_fwd(const float* x, const float* mean, const float* invstd, const float* w, const float* b, float* y, int32_t R,
               int32_t C, void* stream);
int sml_bn_bwd_stats(const float* dy, const float* x, const float* mean, const float* invstd, float* sums, int32_t R, int32_t C,
                     void* stream);                                              /* (sum dy, sum dy*xhat) */
int sml_bn_bwd(const float* dy, const float* x, const float* mean, const float* invstd, const float* w, const float* sums,
               float inv_n, float* dx, int32_t R, int32_t C, void* stream);
/* BatchLoss epilogue on Gram matrices (utils/loss.py:26-38): S [N,N] = omic Gram, T [NS,N,N] = vgrid slab Grams;
 * loss [N,N] = (rownorm(S) - mean_s rownorm(T_s))^2 / N; rn [(1+NS),N] saved row norms */
int sml_batchloss_fwd(const float* S, const float* T, float* loss, float* rn, int32_t N, int32_t NS, void* stream);
int sml_batchloss_bwd(const float* S, const float* T, const float* rn, const float* dloss, float* dS, float* dT, int32_t N,
                      int32_t NS, void* stream);
/* OrthogonalLoss (models/cmta_utils.py:1217-1228): loss [B]; stats [B,9] saved; barred operands are detached */
int sml_orthogonal_loss_fwd(const float* P, const float* Ph, const float* G, const float* Gh, float gamma, float* loss,
                            float* stats, int32_t B, int32_t D, void* stream);
int sml_orthogonal_loss_bwd(const float* P, const float* Ph, const float* G, const float* Gh, float gamma, const float* stats,
                            const float* dloss, float* dP, float* dPh, float* dG, float* dGh, int32_t B, int32_t D, void* stream);

#ifdef __cplusplus
}
#endif
#endif
