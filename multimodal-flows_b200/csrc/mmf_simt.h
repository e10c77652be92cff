// Launchers of the CUDA-core kernels (kernels_simt.cu).
#pragma once
#include "mmf_internal.h"

namespace mmf {

struct EmbedFinishArgs {
    int rows;
    float* resid;            // [rows,256] fp32: x-half holds the raw wxe.2 output on entry
    float* skip;             // [rows,256] fp32 copy kept for the stream-level residuals
    bf16* act;               // [rows,256] bf16: LayerNorm of the first block
    const int* ks;           // packed tokens
    const float* ytab;       // [V,128] = LN_ln1y(wye.2(GELU(wye.0[k]))), precomputed per checkpoint
    const float *ln1x_g, *ln1x_b;
    const float* temb;       // [*,temb_ld], 256 wide (x-half | y-half)
    int temb_ld;
    const int* row_jet;      // null: every row uses temb row 0
    int next_ln_width;       // 128 (two groups) or 256
    const float *next_g, *next_b;   // [256]
};

struct AddLnArgs {
    int rows;
    float* resid;
    const float* skip;
    bf16* act;
    int ln1_width;
    const float *ln1_g, *ln1_b;     // [256]
    const float* temb;              // added after LN_1, or null
    int temb_ld;
    const int* row_jet;
    int write_resid;
    int ln2_width;
    const float *ln2_g, *ln2_b;     // null: bf16 output is LN_1 result (+temb)
};

struct HeadOutArgs {
    int rows;
    const bf16* hidden;      // [rows, ld_hidden]: head_x hidden in cols [0,512), head_y hidden in [512,1024)
    int ld_hidden;
    const float *wx, *bx;    // head_x.2 [3,512], [3]
    const float *wy, *by;    // head_y.2 [V,512], [V]
    const int* row_slot;     // packed row -> b*D + d
    float* vt_out;           // padded (B,D,3) or null
    float* logits_out;       // padded (B,D,V)
    int do_step;
    StepLaunch sl;
    float w, coef;           // thermostat constants of this step (time is uniform inside the sampler)
    float* xs;               // packed state, updated in place
    int* ks;
    const unsigned char* forced;   // (B*D) tokens forced after the step, or null
    float* rates_out;        // padded (B,D,V) or null
    int argmax_out;          // tokens <- argmax_v rates (use_final_max_rates)
};

int launch_pack(const float* x0, const long long* k0, const int* row_slot, int rows, int V, float* xs, int* ks,
                int* err_flag, cudaStream_t stream);
int launch_unpack(const float* xs, const int* ks, const int* row_slot, int rows, float* x_out, long long* k_out,
                  cudaStream_t stream);
int sample_record_bytes(int D);
int launch_sample_pack(const float* x, const long long* k, const long long* mask, const float* mean, const float* std_, long long B,
                       int D, unsigned char* rec, cudaStream_t stream);
int launch_sample_unpack(const unsigned char* rec, long long B, int D, float* x, long long* k, long long* mask, cudaStream_t stream);
// forward half of the training step (kernels_train.cu)
int launch_bridge_sample(const float* x0, const float* x1, const long long* k0, const long long* k1, const float* t, float sigma,
                         float beta, int V, const float* z, const float* u, unsigned long long seed, unsigned long long slot0,
                         long long B, int D, float* xt, long long* kt, int* err, cudaStream_t s);
int launch_multitask_loss(const float* vt, const float* logits, const float* x0, const float* x1, const long long* k1,
                          const long long* mask, int B, int D, int V, float* loss_mse, float* loss_ce, cudaStream_t s);
int launch_ema_update(float* ema, const float* p, double decay, long long n, cudaStream_t s);
int launch_loss_combine(const float* t, const float* loss_mse, const float* loss_ce, const float* w_fc, const float* b_fc,
                        const float* w_pr, const float* b_pr, int E, int mode, int B, float* out5, cudaStream_t s);
int launch_force_tokens(const unsigned char* forced, const int* row_slot, int rows, int* ks, cudaStream_t stream);
int launch_embed_x(const float* xs, int rows, const float* w0, const float* b0, int E, int apply_gelu, bf16* out,
                   int ld_out, cudaStream_t stream);
int launch_embed_finish(const EmbedFinishArgs& a, cudaStream_t stream);
int launch_add_ln(const AddLnArgs& a, cudaStream_t stream);
int launch_head_out(const HeadOutArgs& a, int V, cudaStream_t stream);

}  // namespace mmf
