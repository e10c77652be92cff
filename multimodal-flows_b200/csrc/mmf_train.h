// Launchers of the training-step kernels (kernels_traingemm.cu, kernels_trainops.cu); C ABI in train_api.cu.
// Everything works on the PACKED row layout of the encoder: row r = one real particle, rows of a jet are contiguous,
// jet_off[b] .. jet_off[b + 1] are the rows of jet b, row_jet[r] its jet (reference masks are prefix masks, utils/aoj.py:882-883).
#pragma once
#include "mmf_internal.h"

namespace mmf {

// Programmatic dependent launch: every training kernel begins with griddepcontrol.wait (its predecessor in the stream has
// completed and flushed) followed by griddepcontrol.launch_dependents (its successor may be scheduled now and run its own
// prologue up to the wait), so launch latency and CTA ramp-up of kernel N + 1 hide under kernel N - also as programmatic
// edges of the captured CUDA graph.  MMF_TRAIN_PDL=0 launches without the attribute (the two instructions are then no-ops).
bool tr_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t tr_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tr_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// C[M x N] (+)= A[M x K] B[N x K]^T (+ bias); mode 0: bf16 store, 1: fp32 store, 2: fp32 reduce-add (K split over `ksplit` CTAs),
// 3: bf16 store + aux = GELU(C) (bf16), 4: bf16 store of the product times GELU'(aux) (aux: bf16 pre-activations),
// 5: fp32 store of resid + product + bias (+ tadd[row_jet]) - the residual stream written out of place by the projection
struct TrGemmResid {
    const float* resid; long long ldr;       // [M x N] fp32 input of the residual connection
    const float* tadd; long long ldt;        // optional per-jet row (time embedding)
    const int* row_jet;
    int M;
    // mode 6 (c_attn + per-head LayerNorm of q and k)
    int C, hs;
    const float *qg, *qb, *kg, *kb;
};
int launch_tr_gemm_qkv(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* qkv, long long ldq, void* qkn,
                       long long ldn, int M, int C, int K, int hs, const float* qg, const float* qb, const float* kg, const float* kb,
                       cudaStream_t s);
int launch_tr_gemm(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N, int K,
                   const float* bias, int mode, int ksplit, void* aux, long long ldaux, const TrGemmResid* resid, cudaStream_t s);

// C[M x N] += A^T B, A [K x M], B [K x N] row-major bf16 (weight gradient from row-major activations, MN-major operands)
int launch_tr_gemm_tn(const void* A, long long lda, const void* B, long long ldb, float* C, long long ldc, int M, int N, int K, int ksplit,
                      cudaStream_t s);

// small fp32 products on CUDA cores (per-jet operands: time_expand, uncertainty net): C = A B (+ bias) (+ C), general strides
int launch_tr_sgemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                    long long ldc, int M, int N, int K, const float* bias, int accumulate, cudaStream_t s);

// in [rows x cols] (fp32 or bf16, row pitch ld_in) -> bf16 copy, bf16 transpose [cols x ldT], column sums (atomic, fp32); any output may be null
int launch_tr_cast_transpose(const void* in, long long ld_in, int in_f32, int rows, int cols, bf16* out, long long ld_out,
                             bf16* outT, long long ldT, float* colsum, cudaStream_t s);

struct TrTransposeJob { long long src, dst; int rows, cols, tile0, pad; };   // element offsets into the fp32 / bf16 flat buffers
int launch_tr_weights_transpose(const float* p, bf16* pT, const TrTransposeJob* jobs_dev, int n_jobs, int n_tiles, cudaStream_t s);

int launch_tr_pack(const float* xt, const long long* kt, const float* x0, const float* x1, const long long* k1, const int* row_slot,
                   int M, int V, float* xs, int* ks, float* tgt, int* k1p, int* err, cudaStream_t s);
int launch_tr_time_embed(const float* t, const int* perm, int B, int dim, int dup, float* out, long long ld, cudaStream_t s);
int launch_tr_embed_x_fwd(const float* xs, int M, const float* w0, const float* b0, int E, bf16* h, long long ld, cudaStream_t s);
int launch_tr_embed_x_bwd(const bf16* dh, long long ld, const float* xs, int M, const float* w0, const float* b0, int E, float* dw0,
                          float* db0, cudaStream_t s);
int launch_tr_embed_y_fwd(const int* ks, int M, const float* emb, int E, int V, bf16* g, long long ld, cudaStream_t s);
int launch_tr_embed_y_bwd(const bf16* dg, long long ld, const int* ks, int M, const float* emb, int E, int V, float* demb, cudaStream_t s);

struct TrLnArgs {
    const float* x; long long ldx;           // input rows
    const float* add; long long lda;         // optional second summand (the skip connection): LN(x + add)
    const float *g, *b;                      // affine (b may be null)
    const float* tadd; long long ldt;        // optional per-jet row added AFTER the LayerNorm (time embedding)
    const int* row_jet;
    int M, C;                                // C = 128 or 256
    bf16* out16; long long ld16;             // optional outputs
    float* out32; long long ld32;
    float *mean, *rstd;                      // [M] saved for the backward pass
};
int launch_tr_ln_fwd(const TrLnArgs& a, cudaStream_t s);
struct TrLnBwdArgs {
    const float* dy; long long lddy;
    const float* x; long long ldx;
    const float* add; long long lda;
    const float *mean, *rstd, *g;
    int M, C;
    float* dx; long long lddx; int accumulate;
    float *dg, *db;                          // atomically accumulated (db may be null)
    bf16* dx16; long long ld16;              // optional: bf16 copy of the final dx (operand of the next linear's products)
    float* dxsum;                            // optional: += column sums of the final dx (that linear's bias gradient)
};
int launch_tr_ln_bwd(const TrLnBwdArgs& a, cudaStream_t s);

// per-head LayerNorm of q and k (reference networks/attention.py:62-64)
int launch_tr_qkln_fwd(const bf16* qkv, long long ld, int M, int C, int H, const float* qg, const float* qb, const float* kg,
                       const float* kb, bf16* qn, bf16* kn, long long ldn, cudaStream_t s);
int launch_tr_qkln_bwd(bf16* dqkv, long long ldd, const bf16* qkv, long long ld, int M, int C, int H, const float* qg, const float* kg,
                       float* dqg, float* dqb, float* dkg, float* dkb, cudaStream_t s);

// masked self-attention of whole jets (reference attention.py:53-74) on CUDA cores: one CTA per (jet, head); P is kept for the
// backward pass; jets of at most min_n particles are skipped (they run on the tensor-core kernels below)
int launch_tr_attn_fwd(const bf16* qn, long long ldq, const bf16* kn, long long ldk, const bf16* v, long long ldv, const int* jet_off,
                       const long long* p_off, int B, int H, int hs, int nmax, int min_n, bf16* o, long long ldo, bf16* P, cudaStream_t s);
int launch_tr_attn_bwd(const bf16* dO, long long lddo, const bf16* o, long long ldo, const bf16* P, const bf16* qn, long long ldq,
                       const bf16* kn, long long ldk, const bf16* v, long long ldv, const int* jet_off, const long long* p_off, int B,
                       int H, int hs, int nmax, int min_n, bf16* dqkv, long long ldd, int C, cudaStream_t s);

// the same on tensor cores for tiles of whole jets with at most 128 rows in total (kernels_trainattn.cu); items[i] = (first row,
// rows); CTAs with blockIdx.x >= *n_items exit (fixed launch grids under CUDA graphs); stats [M, H, 2] fp32 replaces P
struct TrAttnTcArgs {
    const int2* items;
    const int* n_items;
    const int* row_jet;
    const int* jet_off;
    float* stats;
    bf16* o; long long ldo;          // forward output
    bf16* dqkv; long long ldd;       // backward output: dq | dk | dv sections of width C
    int H, C;
    float scale, scale_log2e;
};
int launch_tr_attn_tc_fwd(const bf16* qn, long long ldq, const bf16* kn, long long ldk, const bf16* v, long long ldv, int M, int C, int hs,
                          int grid_items, TrAttnTcArgs a, cudaStream_t s);
int launch_tr_attn_tc_bwd(const bf16* dO, long long lddo, const bf16* qn, long long ldq, const bf16* kn, long long ldk, const bf16* v,
                          long long ldv, int M, int C, int hs, int grid_items, TrAttnTcArgs a, cudaStream_t s);

int launch_tr_gelu_fwd(const void* z, void* h, long long n, int f32, cudaStream_t s);
int launch_tr_gelu_bwd(const void* dh, const void* z, void* dz, long long n, int f32, cudaStream_t s);
// out = a + y + tadd[row_jet]  (y, tadd optional)
int launch_tr_add(float* out, long long ldo, const float* a, long long lda, const float* y, long long ldy, const float* tadd,
                  long long ldt, const int* row_jet, int M, int C, cudaStream_t s);
int launch_tr_jet_sum(const float* g, long long ld, const int* jet_off, int B, int C, float* out, long long ldo, int accumulate, cudaStream_t s);

int launch_tr_head_fwd(const bf16* h, long long ldh, int I, const float* wx, const float* bx, const float* wy, const float* by, int V,
                       int M, float* vt, float* logits, cudaStream_t s);
int launch_tr_head_bwd(const float* dvt, const float* dlog, const bf16* h, const bf16* z, long long ldh, int I, const float* wx,
                       const float* wy, int V, int M, bf16* dz, float* dwx, float* dbx, float* dwy, float* dby, cudaStream_t s);

int launch_tr_loss_fwd(const float* vt, const float* logits, const float* tgt, const int* k1, const int* jet_off, int B, int V,
                       float* loss_mse, float* loss_ce, cudaStream_t s);
// MultiTaskLoss (reference model/MMF.py:203-233) and its derivatives: out5 = means of (loss, l_mse, l_ce, w_mse, w_ce);
// gl1 / gl2 = d loss / d l_mse[b], d l_ce[b]; du = d loss / d u[b, 0:2] (time-weighted only)
int launch_tr_loss_combine(const float* loss_mse, const float* loss_ce, const float* u, int B, float* out5, float* gl1, float* gl2,
                           float* du, cudaStream_t s);
int launch_tr_loss_bwd(const float* vt, const float* logits, const float* tgt, const int* k1, const int* row_jet, const int* jet_off,
                       const float* gl1, const float* gl2, int M, int B, int V, float* dvt, float* dlog, cudaStream_t s);

constexpr int kSumsqScratch = 2048;      // floats behind `out`: out[0] the result, the rest per-block partial sums
int launch_tr_sumsq(const float* g, long long n, float* out, cudaStream_t s);
// torch.optim.Adam (reference model/MMF.py:77-78) with Lightning's gradient_clip_val norm clipping (scripts/train_mmf.py:166) folded in
int launch_tr_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, int step,
                   const float* sumsq, float max_norm, float grad_scale, bf16* p16, cudaStream_t s);

}  // namespace mmf
