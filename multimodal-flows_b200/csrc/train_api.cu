// extern "C" entry points of the training-step operators (include/mmf_b200_train.h): argument checks + the launchers of
// kernels_traingemm.cu / kernels_trainops.cu.  No state: the host side owns every buffer.
#include "../../include/mmf_b200_train.h"

#include "mmf_train.h"

using namespace mmf;

#define S_(stream) static_cast<cudaStream_t>(stream)
#define BF(p) static_cast<bf16*>(p)
#define CBF(p) static_cast<const bf16*>(p)

extern "C" {

int mmf_tr_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                const float* bias, int32_t mode, int32_t ksplit, void* aux, int64_t ldaux, const float* resid, int64_t ldr, const float* tadd,
                int64_t ldt, const int32_t* row_jet, void* stream) {
    TrGemmResid rs{};
    rs.resid = resid; rs.ldr = ldr; rs.tadd = tadd; rs.ldt = ldt; rs.row_jet = row_jet; rs.M = M;
    return launch_tr_gemm(A, lda, B, ldb, C, ldc, M, N, K, bias, mode, ksplit, aux, ldaux, &rs, S_(stream));
}

int mmf_tr_gemm_qkv(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* qkv, int64_t ldq, void* qkn, int64_t ldn,
                    int32_t M, int32_t C, int32_t K, int32_t H, const float* qg, const float* qb, const float* kg, const float* kb, void* stream) {
    MMF_REQUIRE(H > 0 && C % H == 0, "gemm_qkv: bad head count");
    return launch_tr_gemm_qkv(A, lda, W, ldw, bias, qkv, ldq, qkn, ldn, M, C, K, C / H, qg, qb, kg, kb, S_(stream));
}

int mmf_tr_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                   int32_t ksplit, void* stream) {
    return launch_tr_gemm_tn(A, lda, B, ldb, C, ldc, M, N, K, ksplit, S_(stream));
}

int mmf_tr_sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc, int32_t M,
                 int32_t N, int32_t K, const float* bias, int32_t accumulate, void* stream) {
    if (M <= 0 || N <= 0) return 0;
    MMF_REQUIRE(A && B && C, "sgemm: null operand");
    return launch_tr_sgemm(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, accumulate, S_(stream));
}

int mmf_tr_cast_transpose(const void* in, int64_t ld_in, int32_t in_f32, int32_t rows, int32_t cols, void* out_bf16, int64_t ld_out,
                          void* outT_bf16, int64_t ldT, float* colsum, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    MMF_REQUIRE(in, "cast_transpose: null input");
    return launch_tr_cast_transpose(in, ld_in, in_f32, rows, cols, BF(out_bf16), ld_out, BF(outT_bf16), ldT, colsum, S_(stream));
}

int mmf_tr_weights_transpose(const float* params, void* paramsT_bf16, const void* jobs, int32_t n_jobs, int32_t n_tiles, void* stream) {
    MMF_REQUIRE(params && paramsT_bf16 && jobs, "weights_transpose: null argument");
    static_assert(sizeof(TrTransposeJob) == 32, "job record layout is part of the ABI");
    return launch_tr_weights_transpose(params, BF(paramsT_bf16), static_cast<const TrTransposeJob*>(jobs), n_jobs, n_tiles, S_(stream));
}

int mmf_tr_pack(const float* xt, const int64_t* kt, const float* x0, const float* x1, const int64_t* k1, const int32_t* row_slot, int32_t M,
                int32_t V, float* xs, int32_t* ks, float* tgt, int32_t* k1p, int32_t* err, void* stream) {
    MMF_REQUIRE(M == 0 || (xt && kt && x0 && x1 && k1 && row_slot && xs && ks && tgt && k1p && err), "pack: null argument");
    return launch_tr_pack(xt, reinterpret_cast<const long long*>(kt), x0, x1, reinterpret_cast<const long long*>(k1), row_slot, M, V, xs, ks,
                          tgt, k1p, err, S_(stream));
}

int mmf_tr_time_embed(const float* t, const int32_t* perm, int32_t B, int32_t dim, int32_t dup, float* out, int64_t ld, void* stream) {
    MMF_REQUIRE(t && out && dim >= 4 && dim % 2 == 0, "time_embed: bad argument");
    return launch_tr_time_embed(t, perm, B, dim, dup, out, ld, S_(stream));
}

int mmf_tr_embed_x_fwd(const float* xs, int32_t M, const float* w0, const float* b0, int32_t E, void* h_bf16, int64_t ld, void* stream) {
    return launch_tr_embed_x_fwd(xs, M, w0, b0, E, BF(h_bf16), ld, S_(stream));
}
int mmf_tr_embed_x_bwd(const void* dh_bf16, int64_t ld, const float* xs, int32_t M, const float* w0, const float* b0, int32_t E,
                       float* dw0, float* db0, void* stream) {
    return launch_tr_embed_x_bwd(CBF(dh_bf16), ld, xs, M, w0, b0, E, dw0, db0, S_(stream));
}
int mmf_tr_embed_y_fwd(const int32_t* ks, int32_t M, const float* emb, int32_t E, int32_t V, void* g_bf16, int64_t ld, void* stream) {
    return launch_tr_embed_y_fwd(ks, M, emb, E, V, BF(g_bf16), ld, S_(stream));
}
int mmf_tr_embed_y_bwd(const void* dg_bf16, int64_t ld, const int32_t* ks, int32_t M, const float* emb, int32_t E, int32_t V, float* demb,
                       void* stream) {
    return launch_tr_embed_y_bwd(CBF(dg_bf16), ld, ks, M, emb, E, V, demb, S_(stream));
}

int mmf_tr_ln_fwd(const float* x, int64_t ldx, const float* add, int64_t lda, const float* g, const float* b, const float* tadd,
                  int64_t ldt, const int32_t* row_jet, int32_t M, int32_t C, void* out_bf16, int64_t ld16, float* out_f32, int64_t ld32,
                  float* mean, float* rstd, void* stream) {
    MMF_REQUIRE(M == 0 || (x && g), "layernorm: null argument");
    TrLnArgs a{x, ldx, add, lda, g, b, tadd, ldt, row_jet, M, C, BF(out_bf16), ld16, out_f32, ld32, mean, rstd};
    return launch_tr_ln_fwd(a, S_(stream));
}
int mmf_tr_ln_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* add, int64_t lda, const float* mean,
                  const float* rstd, const float* g, int32_t M, int32_t C, float* dx, int64_t lddx, int32_t accumulate, float* dg, float* db,
                  void* dx_bf16, int64_t ld16, float* dxsum, void* stream) {
    MMF_REQUIRE(M == 0 || (dy && x && mean && rstd && g && dx && dg), "layernorm backward: null argument");
    TrLnBwdArgs a{dy, lddy, x, ldx, add, lda, mean, rstd, g, M, C, dx, lddx, accumulate, dg, db, BF(dx_bf16), ld16, dxsum};
    return launch_tr_ln_bwd(a, S_(stream));
}

int mmf_tr_qkln_fwd(const void* qkv, int64_t ld, int32_t M, int32_t C, int32_t H, const float* qg, const float* qb, const float* kg,
                    const float* kb, void* qn, void* kn, int64_t ldn, void* stream) {
    return launch_tr_qkln_fwd(CBF(qkv), ld, M, C, H, qg, qb, kg, kb, BF(qn), BF(kn), ldn, S_(stream));
}
int mmf_tr_qkln_bwd(void* dqkv, int64_t ldd, const void* qkv, int64_t ld, int32_t M, int32_t C, int32_t H, const float* qg, const float* kg,
                    float* dqg, float* dqb, float* dkg, float* dkb, void* stream) {
    return launch_tr_qkln_bwd(BF(dqkv), ldd, CBF(qkv), ld, M, C, H, qg, kg, dqg, dqb, dkg, dkb, S_(stream));
}

int mmf_tr_attn_fwd(const void* qn, int64_t ldq, const void* kn, int64_t ldk, const void* v, int64_t ldv, const int32_t* jet_off,
                    const int64_t* p_off, int32_t B, int32_t H, int32_t hs, int32_t nmax, int32_t min_n, void* o, int64_t ldo, void* P,
                    void* stream) {
    return launch_tr_attn_fwd(CBF(qn), ldq, CBF(kn), ldk, CBF(v), ldv, jet_off, reinterpret_cast<const long long*>(p_off), B, H, hs, nmax,
                              min_n, BF(o), ldo, BF(P), S_(stream));
}
int mmf_tr_attn_bwd(const void* dO, int64_t lddo, const void* o, int64_t ldo, const void* P, const void* qn, int64_t ldq, const void* kn,
                    int64_t ldk, const void* v, int64_t ldv, const int32_t* jet_off, const int64_t* p_off, int32_t B, int32_t H, int32_t hs,
                    int32_t nmax, int32_t min_n, void* dqkv, int64_t ldd, int32_t C, void* stream) {
    return launch_tr_attn_bwd(CBF(dO), lddo, CBF(o), ldo, CBF(P), CBF(qn), ldq, CBF(kn), ldk, CBF(v), ldv, jet_off,
                              reinterpret_cast<const long long*>(p_off), B, H, hs, nmax, min_n, BF(dqkv), ldd, C, S_(stream));
}

int mmf_tr_attn_tc_fwd(const void* qn, int64_t ldq, const void* kn, int64_t ldk, const void* v, int64_t ldv, int32_t M, int32_t C, int32_t hs,
                       const int32_t* items, const int32_t* n_items, int32_t grid_items, const int32_t* row_jet, const int32_t* jet_off,
                       float* stats, void* o, int64_t ldo, void* stream) {
    MMF_REQUIRE(M == 0 || (qn && kn && v && items && n_items && row_jet && jet_off && stats && o), "attention: null argument");
    TrAttnTcArgs a{};
    a.items = reinterpret_cast<const int2*>(items); a.n_items = n_items; a.row_jet = row_jet; a.jet_off = jet_off; a.stats = stats;
    a.o = BF(o); a.ldo = ldo;
    return launch_tr_attn_tc_fwd(CBF(qn), ldq, CBF(kn), ldk, CBF(v), ldv, M, C, hs, grid_items, a, S_(stream));
}
int mmf_tr_attn_tc_bwd(const void* dO, int64_t lddo, const void* qn, int64_t ldq, const void* kn, int64_t ldk, const void* v, int64_t ldv,
                       int32_t M, int32_t C, int32_t hs, const int32_t* items, const int32_t* n_items, int32_t grid_items,
                       const int32_t* row_jet, const int32_t* jet_off, const float* stats, void* dqkv, int64_t ldd, void* stream) {
    MMF_REQUIRE(M == 0 || (dO && qn && kn && v && items && n_items && row_jet && jet_off && stats && dqkv), "attention: null argument");
    TrAttnTcArgs a{};
    a.items = reinterpret_cast<const int2*>(items); a.n_items = n_items; a.row_jet = row_jet; a.jet_off = jet_off;
    a.stats = const_cast<float*>(stats); a.dqkv = BF(dqkv); a.ldd = ldd;
    return launch_tr_attn_tc_bwd(CBF(dO), lddo, CBF(qn), ldq, CBF(kn), ldk, CBF(v), ldv, M, C, hs, grid_items, a, S_(stream));
}

int mmf_tr_gelu_fwd(const void* z, void* h, int64_t n, int32_t f32, void* stream) { return launch_tr_gelu_fwd(z, h, n, f32, S_(stream)); }
int mmf_tr_gelu_bwd(const void* dh, const void* z, void* dz, int64_t n, int32_t f32, void* stream) {
    return launch_tr_gelu_bwd(dh, z, dz, n, f32, S_(stream));
}

int mmf_tr_add(float* out, int64_t ldo, const float* a, int64_t lda, const float* y, int64_t ldy, const float* tadd, int64_t ldt,
               const int32_t* row_jet, int32_t M, int32_t C, void* stream) {
    return launch_tr_add(out, ldo, a, lda, y, ldy, tadd, ldt, row_jet, M, C, S_(stream));
}
int mmf_tr_jet_sum(const float* g, int64_t ld, const int32_t* jet_off, int32_t B, int32_t C, float* out, int64_t ldo, int32_t accumulate,
                   void* stream) {
    return launch_tr_jet_sum(g, ld, jet_off, B, C, out, ldo, accumulate, S_(stream));
}

int mmf_tr_head_fwd(const void* h, int64_t ldh, int32_t I, const float* wx, const float* bx, const float* wy, const float* by, int32_t V,
                    int32_t M, float* vt, float* logits, void* stream) {
    return launch_tr_head_fwd(CBF(h), ldh, I, wx, bx, wy, by, V, M, vt, logits, S_(stream));
}
int mmf_tr_head_bwd(const float* dvt, const float* dlog, const void* h, const void* z, int64_t ldh, int32_t I, const float* wx,
                    const float* wy, int32_t V, int32_t M, void* dz, float* dwx, float* dbx, float* dwy, float* dby, void* stream) {
    return launch_tr_head_bwd(dvt, dlog, CBF(h), CBF(z), ldh, I, wx, wy, V, M, BF(dz), dwx, dbx, dwy, dby, S_(stream));
}

int mmf_tr_loss_fwd(const float* vt, const float* logits, const float* tgt, const int32_t* k1, const int32_t* jet_off, int32_t B, int32_t V,
                    float* loss_mse, float* loss_ce, void* stream) {
    return launch_tr_loss_fwd(vt, logits, tgt, k1, jet_off, B, V, loss_mse, loss_ce, S_(stream));
}
int mmf_tr_loss_combine(const float* loss_mse, const float* loss_ce, const float* u, int32_t B, float* out5, float* gl1, float* gl2, float* du,
                        void* stream) {
    MMF_REQUIRE(loss_mse && loss_ce && out5 && gl1 && gl2 && (!u || du), "loss_combine: null argument");
    return launch_tr_loss_combine(loss_mse, loss_ce, u, B, out5, gl1, gl2, du, S_(stream));
}
int mmf_tr_loss_bwd(const float* vt, const float* logits, const float* tgt, const int32_t* k1, const int32_t* row_jet, const int32_t* jet_off,
                    const float* gl1, const float* gl2, int32_t M, int32_t B, int32_t V, float* dvt, float* dlog, void* stream) {
    return launch_tr_loss_bwd(vt, logits, tgt, k1, row_jet, jet_off, gl1, gl2, M, B, V, dvt, dlog, S_(stream));
}

int mmf_tr_sumsq(const float* g, int64_t n, float* out, void* stream) { return launch_tr_sumsq(g, n, out, S_(stream)); }
int mmf_tr_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, int32_t step,
                const float* sumsq, float max_norm, float grad_scale, void* p16, void* stream) {
    MMF_REQUIRE(n == 0 || (p && g && m && v), "adam: null argument");
    return launch_tr_adam(p, g, m, v, n, lr, beta1, beta2, eps, step, sumsq, max_norm, grad_scale, BF(p16), S_(stream));
}

}  // extern "C"
