/* C ABI of the MMF training step (BASELINE config #5): the operators that replace torch autograd's
 * forward and backward of MultiModalFlowBridge.loss (reference multimodal_flows/model/MMF.py:138-170) for the ParticleFormer /
 * FusedParticleFormer encoders (networks/ParticleTransformers.py:62-122, 177-210), the Adam update of configure_optimizers
 * (model/MMF.py:77-78) and Lightning's gradient_clip_val = 1.0 (scripts/train_mmf.py:166).
 *
 * The host side (multimodal-flows_b200/mmf_b200/training.py) sequences these calls; every call queues kernels on `stream`
 * of the calling thread's current device and returns without synchronising.  All pointers are DEVICE pointers.  Status
 * convention as in mmf_b200.h (0 ok, message in mmf_last_error()).
 *
 * Layout: PACKED rows.  Row r is one real particle; the rows of jet b are jet_off[b] .. jet_off[b + 1]; row_jet[r] = b;
 * row_slot[r] = b * D + d in the reference's padded (B, D) tensors.  Activations are bf16 (GEMM operands) or fp32 (the
 * residual stream, LayerNorm inputs, losses), parameters and gradients fp32 in the reference's state_dict shapes.
 * `ld*` are row pitches in ELEMENTS. */
#ifndef MMF_B200_TRAIN_H
#define MMF_B200_TRAIN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* nn.Linear forward / data gradient / weight gradient on tensor cores (tcgen05, fp32 accumulation):
 *   C[M x N] (+)= A[M x K] B[N x K]^T (+ bias[N]); A, B bf16 with K contiguous (lda, ldb multiples of 8, 16-byte aligned bases).
 *   mode 0: C bf16;  mode 1: C fp32;  mode 2: C fp32 += (TMA reduce-add; K is split over `ksplit` CTAs, bias added once);
 *   mode 3: C bf16 and aux = GELU(C) bf16 (MLP.c_fc + nn.GELU in one pass, utils/models.py:15-16);
 *   mode 4: C = bf16(product * GELU'(aux)), aux = the bf16 pre-activations (data gradient of MLP.c_proj through the GELU);
 *   mode 5: C fp32 = resid + product + bias (+ tadd[row_jet[row]]): `x = x + attn(...)` / `x = x + ffw(...)` (+ time embedding)
 *           of SelfAttnBlock (attention.py:24-25, ParticleTransformers.py:88-89) written out of place by the projection.
 * aux [M x N] bf16 with row pitch ldaux (modes 3, 4; else null); resid [M x N] fp32 / tadd / row_jet (mode 5; else null).
 * replaces F.linear and its autograd (attention.py:44-45, utils/models.py:15-17). */
int mmf_tr_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                const float* bias, int32_t mode, int32_t ksplit, void* aux, int64_t ldaux, const float* resid, int64_t ldr, const float* tadd,
                int64_t ldt, const int32_t* row_jet, void* stream);
/* SelfAttention.c_attn with the per-head LayerNorm of q and k in its epilogue (attention.py:57-64): qkv [M x 3C] bf16 (kept for the
 * backward pass) and the normalised q | k as qkn [M x 2C] bf16; A [M x K] bf16, W [3C x K] bf16, bias [3C] or null */
int mmf_tr_gemm_qkv(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* qkv, int64_t ldq, void* qkn, int64_t ldn,
                    int32_t M, int32_t C, int32_t K, int32_t H, const float* qg, const float* qb, const float* kg, const float* kb, void* stream);
/* weight gradient of nn.Linear straight from the row-major activations: C[M x N] += A^T B with A = dy [K x M], B = x [K x N]
 * (bf16, K = tokens; both are read as MN-major tcgen05 operands, so no transposed copy exists); K split over `ksplit` CTAs */
int mmf_tr_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                   int32_t ksplit, void* stream);
/* small fp32 product with general strides, C[m, n] = sum_k A[m sam + k sak] B[k sbk + n sbn] (+ bias[n]) (+ C):
 * the per-jet linears (time_expand ParticleTransformers.py:109, MultiTaskLoss.uncertainty_net MMF.py:212) and their gradients */
int mmf_tr_sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc, int32_t M,
                 int32_t N, int32_t K, const float* bias, int32_t accumulate, void* stream);
/* in [rows x cols] (fp32 if in_f32 else bf16) -> optional bf16 copy, bf16 transpose [cols x ldT] (operands of the weight-gradient
 * GEMM) and fp32 column sums added to colsum[cols] (bias gradients) */
int mmf_tr_cast_transpose(const void* in, int64_t ld_in, int32_t in_f32, int32_t rows, int32_t cols, void* out_bf16, int64_t ld_out,
                          void* outT_bf16, int64_t ldT, float* colsum, void* stream);
/* bf16 transposed copies of the 2-D weights of the flat fp32 parameter buffer, one launch.  jobs: n_jobs records of
 * {int64 src_offset, int64 dst_offset, int32 rows, int32 cols, int32 first_tile, int32 0} in device memory, tiles of 32 x 32 */
int mmf_tr_weights_transpose(const float* params, void* paramsT_bf16, const void* jobs, int32_t n_jobs, int32_t n_tiles, void* stream);

/* gather of the real particles of a training batch: xs = xt, ks = kt, tgt = x1 - x0 (CFM.py:186-193), k1p = k1; err as in mmf_b200.h */
int mmf_tr_pack(const float* xt, const int64_t* kt, const float* x0, const float* x1, const int64_t* k1, const int32_t* row_slot, int32_t M,
                int32_t V, float* xs, int32_t* ks, float* tgt, int32_t* k1p, int32_t* err, void* stream);
/* transformer_timestep_embedding (utils/models.py:62-75) of t[perm[b]] (perm null: t[b]) into out[b, 0:dim]; dup != 0 repeats it
 * in out[b, dim:2 dim].  perm: packed jet -> jet of the batch (the planner packs jets into attention tiles in its own order) */
int mmf_tr_time_embed(const float* t, const int32_t* perm, int32_t B, int32_t dim, int32_t dup, float* out, int64_t ld, void* stream);
/* wxe.0 + GELU (ParticleTransformers.py:28-30) and its gradient (dh = gradient w.r.t. the GELU output, bf16) */
int mmf_tr_embed_x_fwd(const float* xs, int32_t M, const float* w0, const float* b0, int32_t E, void* h_bf16, int64_t ld, void* stream);
int mmf_tr_embed_x_bwd(const void* dh_bf16, int64_t ld, const float* xs, int32_t M, const float* w0, const float* b0, int32_t E,
                       float* dw0, float* db0, void* stream);
/* wye.0 (nn.Embedding) + GELU (ParticleTransformers.py:31-33) and its gradient */
int mmf_tr_embed_y_fwd(const int32_t* ks, int32_t M, const float* emb, int32_t E, int32_t V, void* g_bf16, int64_t ld, void* stream);
int mmf_tr_embed_y_bwd(const void* dg_bf16, int64_t ld, const int32_t* ks, int32_t M, const float* emb, int32_t E, int32_t V, float* demb,
                       void* stream);

/* y = LayerNorm(x (+ add)) g + b (+ tadd[row_jet]) (utils/models.py:36-37, eps 1e-5), C = 128 or 256; outputs optional;
 * mean / rstd [M] saved for the backward call */
int mmf_tr_ln_fwd(const float* x, int64_t ldx, const float* add, int64_t lda, const float* g, const float* b, const float* tadd,
                  int64_t ldt, const int32_t* row_jet, int32_t M, int32_t C, void* out_bf16, int64_t ld16, float* out_f32, int64_t ld32,
                  float* mean, float* rstd, void* stream);
/* dx (+)= dLN/dx, dg += , db += (atomic); optionally the final dx also leaves as bf16 (dx_bf16, the operand of the next linear's
 * gradient products) and its column sums are added to dxsum [C] (that linear's bias gradient) */
int mmf_tr_ln_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* add, int64_t lda, const float* mean,
                  const float* rstd, const float* g, int32_t M, int32_t C, float* dx, int64_t lddx, int32_t accumulate, float* dg, float* db,
                  void* dx_bf16, int64_t ld16, float* dxsum, void* stream);
/* per-head LayerNorm of q and k (attention.py:62-64): qkv bf16 [M, 3C] -> qn, kn bf16 [M, C]; the backward call turns
 * dqkv[:, 0:2C] (gradients of qn | kn) into the gradients of q | k in place */
int mmf_tr_qkln_fwd(const void* qkv, int64_t ld, int32_t M, int32_t C, int32_t H, const float* qg, const float* qb, const float* kg,
                    const float* kb, void* qn, void* kn, int64_t ldn, void* stream);
int mmf_tr_qkln_bwd(void* dqkv, int64_t ldd, const void* qkv, int64_t ld, int32_t M, int32_t C, int32_t H, const float* qg, const float* kg,
                    float* dqg, float* dqb, float* dkg, float* dkb, void* stream);
/* masked self-attention of whole jets, scale 1/sqrt(hs) (attention.py:53-74), two kernel families:
 *  (1) tensor cores (tcgen05): work items = runs of consecutive whole jets with at most 128 rows in total, items[i] = (first row,
 *      rows) as int32 pairs; *n_items (device) of the grid_items launched CTAs work, the rest exit (fixed grids under CUDA graphs).
 *      The forward call keeps 2 floats per (row, head) in stats [M, H, 2]; the backward call recomputes the probabilities from
 *      them and writes dqn | dkn | dv into dqkv [M, 3C].  Jets of more than 128 particles are not in any item: they take
 *  (2) CUDA cores: one CTA per (jet, head) for the jets of MORE than min_n particles (min_n = 0: all jets); P (bf16,
 *      H * sum n_b^2 elements, jet b at p_off[b] * H) is kept for the backward call. */
int mmf_tr_attn_tc_fwd(const void* qn, int64_t ldq, const void* kn, int64_t ldk, const void* v, int64_t ldv, int32_t M, int32_t C, int32_t hs,
                       const int32_t* items, const int32_t* n_items, int32_t grid_items, const int32_t* row_jet, const int32_t* jet_off,
                       float* stats, void* o, int64_t ldo, void* stream);
int mmf_tr_attn_tc_bwd(const void* dO, int64_t lddo, const void* qn, int64_t ldq, const void* kn, int64_t ldk, const void* v, int64_t ldv,
                       int32_t M, int32_t C, int32_t hs, const int32_t* items, const int32_t* n_items, int32_t grid_items,
                       const int32_t* row_jet, const int32_t* jet_off, const float* stats, void* dqkv, int64_t ldd, void* stream);
int mmf_tr_attn_fwd(const void* qn, int64_t ldq, const void* kn, int64_t ldk, const void* v, int64_t ldv, const int32_t* jet_off,
                    const int64_t* p_off, int32_t B, int32_t H, int32_t hs, int32_t nmax, int32_t min_n, void* o, int64_t ldo, void* P,
                    void* stream);
int mmf_tr_attn_bwd(const void* dO, int64_t lddo, const void* o, int64_t ldo, const void* P, const void* qn, int64_t ldq, const void* kn,
                    int64_t ldk, const void* v, int64_t ldv, const int32_t* jet_off, const int64_t* p_off, int32_t B, int32_t H, int32_t hs,
                    int32_t nmax, int32_t min_n, void* dqkv, int64_t ldd, int32_t C, void* stream);
/* exact-erf GELU (nn.GELU(), utils/models.py:16) on n contiguous elements (bf16, or fp32 if f32) and its gradient */
int mmf_tr_gelu_fwd(const void* z, void* h, int64_t n, int32_t f32, void* stream);
int mmf_tr_gelu_bwd(const void* dh, const void* z, void* dz, int64_t n, int32_t f32, void* stream);
/* residual stream: out = a (+ y) (+ tadd[row_jet]) on [M x C] fp32 (attention.py:24-25, ParticleTransformers.py:84-89) */
int mmf_tr_add(float* out, int64_t ldo, const float* a, int64_t lda, const float* y, int64_t ldy, const float* tadd, int64_t ldt,
               const int32_t* row_jet, int32_t M, int32_t C, void* stream);
/* out[b, :] (+)= sum of the rows of jet b (gradient of a per-jet broadcast: the time embedding) */
int mmf_tr_jet_sum(const float* g, int64_t ld, const int32_t* jet_off, int32_t B, int32_t C, float* out, int64_t ldo, int32_t accumulate,
                   void* stream);
/* head_x.2 / head_y.2 (ParticleTransformers.py:48-53) on h = [hx | hy] bf16 [M, 2 I]; the backward call returns the gradient
 * w.r.t. the PRE-activation z of head_*.0 (bf16) and accumulates the gradients of the two small linears */
int mmf_tr_head_fwd(const void* h, int64_t ldh, int32_t I, const float* wx, const float* bx, const float* wy, const float* by, int32_t V,
                    int32_t M, float* vt, float* logits, void* stream);
int mmf_tr_head_bwd(const float* dvt, const float* dlog, const void* h, const void* z, int64_t ldh, int32_t I, const float* wx,
                    const float* wy, int32_t V, int32_t M, void* dz, float* dwx, float* dbx, float* dwy, float* dby, void* stream);
/* per-jet masked MSE and cross entropy with ignore_index 0 (MMF.py:152-165) on packed rows */
int mmf_tr_loss_fwd(const float* vt, const float* logits, const float* tgt, const int32_t* k1, const int32_t* jet_off, int32_t B, int32_t V,
                    float* loss_mse, float* loss_ce, void* stream);
/* MultiTaskLoss (MMF.py:203-233): u = null -> "sum", else "time-weighted" with u [B, 2] from the uncertainty net.
 * out5 = batch means of (loss, l_mse, l_ce, w_mse, w_ce); gl1 / gl2 [B] = d loss / d l_mse, d l_ce; du [B, 2] = d loss / d u */
int mmf_tr_loss_combine(const float* loss_mse, const float* loss_ce, const float* u, int32_t B, float* out5, float* gl1, float* gl2, float* du,
                        void* stream);
/* d loss / d (vt, logits) per row; rows at and beyond jet_off[B] (a batch padded to a fixed row capacity M) get zeros */
int mmf_tr_loss_bwd(const float* vt, const float* logits, const float* tgt, const int32_t* k1, const int32_t* row_jet, const int32_t* jet_off,
                    const float* gl1, const float* gl2, int32_t M, int32_t B, int32_t V, float* dvt, float* dlog, void* stream);
/* out[0] = sum g^2 (the squared gradient norm of clip_grad_norm_); `out` points to 2048 floats (out[1..] holds per-block partial
 * sums of a two-stage, fixed-order - hence bit-reproducible - reduction: data-parallel replicas compute identical clip coefficients) */
int mmf_tr_sumsq(const float* g, int64_t n, float* out, void* stream);
/* torch.optim.Adam step `step` (1-based) on flat buffers; the gradient is first multiplied by grad_scale and, when sumsq is
 * given and max_norm > 0, by min(1, max_norm / (grad_scale sqrt(sumsq) + 1e-6)); p16 (optional) receives the bf16 copy */
int mmf_tr_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, int32_t step,
                const float* sumsq, float max_norm, float grad_scale, void* p16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMF_B200_TRAIN_H */
