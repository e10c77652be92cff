/*
 * mmf_b200.h -- C ABI of libmmf_b200.so: the B200-native (sm_100a) generation hot path of
 * dfaroughy/Multimodal-flows.  Plain pointers and sizes only; no torch types cross this boundary.
 *
 * Every entry point replaces one piece of the reference's PyTorch hot path (paths relative to
 * /root/reference/multimodal_flows):
 *
 *   mmf_model_create / destroy   MODEL_REGISTRY[config.model](config) + load_state_dict
 *                                networks/registry.py:4-9, model/MMF.py:30, model/CFM.py:23
 *   mmf_encoder_forward          ParticleFormer.forward       networks/ParticleTransformers.py:62-122
 *                                FusedParticleFormer.forward  networks/ParticleTransformers.py:177-210
 *                                EPiC.forward                 networks/EPiC.py:38-62
 *   mmf_hybrid_step              HybridSolver.tauleap_step    model/solvers.py:22-60   (after the model call)
 *                                RandomTelegraphBridge.rate   model/MJB.py:163-195
 *   mmf_euler_step               ContinuousSolver.euler_step  model/solvers.py:139-143 (after the model call)
 *   mmf_generate[_host]          MultiModalFlowBridge.simulate_dynamics  model/MMF.py:172-200
 *                                ConditionalFlowMatching.simulate_dynamics  model/CFM.py:133-154
 *   mmf_make_source              _make_source_dataloader (noise, masks)     scripts/sample_mmf.py:70-92
 *                                sample_from_empirical_masks                utils/aoj.py:875-890
 *   mmf_jet_observables          ParticleClouds / JetFeatures kinematics    utils/aoj.py:333-368, 452-471, 514-521
 *                                flavor_mutliplicities                      utils/metrics.py:10-33
 *                                de-standardisation of the sample           utils/callbacks.py:52-56
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on failure; mmf_last_error() gives a thread-local message.
 *   - "device" pointers are CUDA device pointers on the model's device, "host" pointers are host memory.
 *   - tensors are contiguous, row-major:  x (B,D,3) float32,  k (B,D) int64,  mask (B,D) int64 (non-zero = real
 *     particle),  t (B,) float32,  logits / rates / u (B,D,V) float32.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are asynchronous with respect
 *     to the host except where noted; they never allocate or free caller memory.
 *   - outputs at padded slots (mask == 0) are zero.  The reference lets padded slots evolve and zeroes them in
 *     FlowGeneratorCallback (utils/callbacks.py:57); real slots never depend on them.
 *   - a handle is bound to one device and is not thread-safe; different handles are independent.
 */
#ifndef MMF_B200_H
#define MMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMF_ABI_VERSION 2

typedef struct MmfModel MmfModel;

enum MmfArch { MMF_ARCH_PARTICLEFORMER = 0, MMF_ARCH_FUSED_PARTICLEFORMER = 1, MMF_ARCH_EPIC = 2 };

/* Hyper-parameters read by the hot path (reference scripts/train_mmf.py:42-56). */
typedef struct MmfModelDesc {
    int32_t arch;               /* enum MmfArch */
    int32_t vocab_size;         /* V, 9 */
    int32_t dim_continuous;     /* 3 */
    int32_t n_embd;             /* 256 */
    int32_t n_inner;            /* 512 */
    int32_t n_head;             /* 4 */
    int32_t n_layer;            /* 5 */
    int32_t n_layer_fused;      /* 6 (ParticleFormer only) */
    int32_t n_embd_glob;        /* 16 (EPiC only) */
    int32_t qk_layernorm;       /* 1 */
    int32_t max_num_particles;  /* D, 150 */
} MmfModelDesc;

/* One named fp32 parameter of the reference state_dict (host memory, row-major). */
typedef struct MmfWeightRef {
    const char* name;           /* e.g. "transformer.blocks_x.0.attn.c_attn.weight" */
    const float* data;
    int32_t ndim;
    int64_t shape[4];
} MmfWeightRef;

/* Per-call sampling options (reference config.temperature/top_k/top_p/beta/use_final_max_rates). */
typedef struct MmfStepOptions {
    float temperature;          /* logits / T when T != 1 */
    float beta;                 /* telegraph stochasticity */
    int32_t top_k;              /* <= 0: off */
    float top_p;                /* <= 0: off */
    int32_t use_final_max_rates;
    int32_t method;             /* mmf_hybrid_step only: 0 tau-leap (model/solvers.py:22-60), 1 categorical Euler (:62-91) */
    uint64_t seed;              /* Philox key when no uniforms are supplied */
    uint64_t first_global_jet;  /* global index of jet 0 of this call: draws depend on (seed, global slot, step) only */
} MmfStepOptions;

int mmf_abi_version(void);
const char* mmf_last_error(void);

/* Packs the checkpoint for the device (bf16 K-major matrices, folded tables) and allocates nothing else until
 * the first forward.  Fails on hyper-parameters outside the accelerated envelope (n_embd 256, n_inner 512,
 * n_head 4, D <= 152, V <= 16); there is no fallback path.  Synchronous. */
int mmf_model_create(const MmfModelDesc* desc, const MmfWeightRef* weights, int32_t n_weights, int32_t device,
                     MmfModel** out);
void mmf_model_destroy(MmfModel* model);

/* One encoder forward.  Device pointers.  t is per jet.  logits_out may be NULL for EPiC (k is ignored then).
 * Synchronises internally (mask and times are read back to plan the packed layout; the token flag at the end). */
int mmf_encoder_forward(MmfModel* model, const float* x, const int64_t* k, const int64_t* mask, const float* t,
                        int32_t B, int32_t D, float* vt_out, float* logits_out, void* stream);

/* The fused hybrid step on device tensors, in place on x and k: Euler update of x, telegraph tau-leap of k
 * (opts->method 0) or the categorical Euler jump of HybridSolver.euler_step (opts->method 1, model/solvers.py:62-91:
 * k' ~ Categorical(dp), dp_v = min(rate_v dt, 1) off the diagonal, dp_k = max(1 - sum, 0), filters on dp; it consumes
 * u[..., 0] only, by inverse CDF in channel order).
 * u: (B,D,V) uniforms in [0,1) or NULL (Philox from opts->seed, opts->first_global_jet, step_index).
 * rates_out: (B,D,V) or NULL.  No model handle needed; `device` selects the GPU. */
int mmf_hybrid_step(const float* vt, const float* logits, float* x, int64_t* k, const float* t, float dt,
                    const MmfStepOptions* opts, const float* u, uint32_t step_index, int32_t B, int32_t D,
                    int32_t V, float* rates_out, int32_t device, void* stream);

/* The reference asserts 0 <= k < V before every rate evaluation (model/MJB.py:177-182, two host synchronisations per
 * step).  mmf_hybrid_step instead clamps an out-of-range token to 0 and raises a per-device flag; this call waits for
 * `stream`, returns 3 (message in mmf_last_error) when the flag was raised since the last query, and clears it. */
int mmf_hybrid_step_status(int32_t device, void* stream);

/* x += vt * dt on n floats (EPiC / ContinuousSolver carrier). */
int mmf_euler_step(const float* vt, float* x, float dt, int64_t n, int32_t device, void* stream);

/* Jet-level observables of a sample in one fused pass over device tensors (the parity / quality report of a run):
 *   utils/callbacks.py:52-56   de-standardisation continuous * std + mean (mean, std: HOST float[3], NULL = identity)
 *   utils/aoj.py:333-346       ParticleClouds: px, py, pz, E of every unmasked particle from (pT, eta_rel, phi_rel)
 *   utils/aoj.py:358-368       particle charge from the token (+1: 4, 6, 8; -1: 3, 5, 7)
 *   utils/aoj.py:452-463       JetFeatures: summed four-momentum, pt, m, eta, phi
 *   utils/aoj.py:514-521       jet charge for kappa = 0 and kappa = 1
 *   utils/metrics.py:10-33     flavor multiplicities: counts of each token per jet (the derived sums are host arithmetic)
 * x (B,D,3) f32 standardised, k (B,D) i64 or NULL, mask (B,D) i64 (a slot counts when mask > 0).
 * kin_out (B, MMF_OBS_NKIN) f32 = px py pz E pt m eta phi charge jet_charge multiplicity m2 (m = sqrt(m2); sums in fp64);
 * counts_out (B,V) i32 or NULL.  Empty jets give the reference's values (0 sums, NaN eta / jet_charge). */
#define MMF_OBS_NKIN 12
int mmf_jet_observables(const float* x, const int64_t* k, const int64_t* mask, const float* mean, const float* std_,
                        int32_t B, int32_t D, int32_t V, float* kin_out, int32_t* counts_out, int32_t device,
                        void* stream);

/* ---- forward half of the training step (model/MMF.py:138-170).  The encoder call in the middle is
 * mmf_encoder_forward (per-jet times); there are no backward kernels yet, so this is the loss, not a trainer. ----
 *
 * mmf_bridge_sample: the intermediate state of both bridges at per-jet times t (B,):
 *   model/CFM.py:171-184   UniformFlow.sample            xt = t x1 + (1 - t) x0 + sigma z
 *   model/MJB.py:197-257   RandomTelegraphBridge.sample  kt ~ Categorical(P), P(k) = p(k->k1; t,1) p(k0->k; 0,t) / p(k0->k1; 0,1),
 *                                                        p(a->b; s,t) = 1/V + w (delta_ab - 1/V), w = exp(-V beta (t - s))
 * z (B,D,3) normals and u (B,D) uniforms may be supplied (parity: the categorical draw is the inverse CDF of u in channel
 * order) or NULL (Philox4x32-10 keyed on (seed, first_global_jet D + slot)).  Outputs xt (B,D,3) f32, kt (B,D) i64.  No masking,
 * like the reference.  Out-of-range tokens raise the flag read by mmf_hybrid_step_status. */
int mmf_bridge_sample(const float* x0, const float* x1, const int64_t* k0, const int64_t* k1, const float* t, float sigma, float beta,
                      int32_t V, const float* z, const float* u, uint64_t seed, uint64_t first_global_jet, int32_t B, int32_t D,
                      float* xt, int64_t* kt, int32_t device, void* stream);
/* mmf_multitask_loss: model/MMF.py:152-168 + MultiTaskLoss :203-233 after the encoder returned vt (B,D,3), logits (B,D,V):
 *   loss_mse[b] = sum_{d,c} mask (vt - (x1 - x0))^2 / max(n_b, 1)          (conditional drift of model/CFM.py:186-193)
 *   loss_ce[b]  = sum_d mask CE(logits, k1; ignore_index 0) / max(n_b, 1)
 *   mode 0 "sum": loss = mean(loss_mse + loss_ce);  mode 1 "time-weighted": (u1, u2) = uncertainty_net(time features of t),
 *   loss = mean(0.5 (u1 + e^-u1 loss_mse) + 0.5 (u2 + e^-u2 loss_ce)); w_fc (E,E), b_fc (E), w_proj (2,E), b_proj (2): the
 *   checkpoint's loss_combine.uncertainty_net.{c_fc,c_proj} (device pointers, fp32; ignored for mode 0).
 * per_jet: device (2,B) scratch that receives loss_mse | loss_ce; out5: device float[5] = loss, mean loss_mse, mean loss_ce,
 * mean w_mse, mean w_ce (what MultiModalFlowBridge.loss returns). */
int mmf_multitask_loss(const float* vt, const float* logits, const float* x0, const float* x1, const int64_t* k1, const int64_t* mask,
                       const float* t, int32_t B, int32_t D, int32_t V, int32_t mode, int32_t n_embd, const float* w_fc, const float* b_fc,
                       const float* w_proj, const float* b_proj, float* per_jet, float* out5, int32_t device, void* stream);

/* EMA of the weights: what timm's ModelEmaV2.update does for every state_dict entry when the reference's
 * EMACallback.on_train_batch_end fires (utils/callbacks.py:152-226):  ema <- decay * ema + (1 - decay) * p  on n fp32 values
 * (device pointers), every operation rounded on its own like the torch expression, so the result is bit-identical. */
int mmf_ema_update(float* ema, const float* p, double decay, int64_t n, int32_t device, void* stream);

/* The generated sample as one narrow record per jet - the device-side half of FlowGeneratorCallback:
 *   utils/callbacks.py:52-56   sample.continuous = sample.continuous * std + mean     (mean, std: HOST float[3], NULL = identity)
 *   utils/callbacks.py:57      sample.apply_mask()                                    (padded slots zeroed)
 * fused with the narrowing of the int64 token / mask tensors to one byte per slot.  Record of jet b, R =
 * mmf_sample_record_bytes(D) = round_up(13 D, 16) bytes:  [D][3] f32 kinematics | [D] u8 (token | mask << 7) | zero padding.
 * The records of a rank's shard are what the single end-of-run collective moves (utils/callbacks.py:27-58 writes one temp
 * file per rank and re-reads them on rank 0) and what mmf_b200.writer lays out as generated_sample.h5.
 * x (B,D,3) f32, k (B,D) i64 or NULL (EPiC), mask (B,D) i64; records: B * R bytes (device).  mmf_unpack_sample is the inverse
 * (k / mask may be NULL), producing the reference's dtypes again. */
int64_t mmf_sample_record_bytes(int32_t D);
int mmf_pack_sample(const float* x, const int64_t* k, const int64_t* mask, const float* mean, const float* std_, int64_t B,
                    int32_t D, uint8_t* records, int32_t device, void* stream);
int mmf_unpack_sample(const uint8_t* records, int64_t B, int32_t D, float* x, int64_t* k, int64_t* mask, int32_t device,
                      void* stream);

/* The source state of the sampler, built on the device (no host RNG, no H2D copy of the batch):
 *   scripts/sample_mmf.py:82-84   noise_continuous = randn * pad_mask, noise_discrete = randint(1, vocab_size) * pad_mask
 *   utils/aoj.py:875-890          sample_from_empirical_masks: multiplicity ~ Categorical(histogram), prefix masks
 * mult_probs: HOST float[D+1], weight of multiplicity 0..D (the density histogram of aoj.py:877; need not be normalised).
 * Every draw is a function of (seed, first_global_jet + b, slot) only - Philox4x32-10, Box-Muller - so the sample is
 * independent of batch size and of the sharding over GPUs.  The reference draws from torch's generators, so parity is in
 * distribution; masks, multiplicities and tokens are bit-exact against oracle/source_oracle.py (same counters).
 * Device outputs: x0 (B,D,3) f32, k0 (B,D) i64 or NULL (EPiC), mask (B,D) i64, n_out (B) i32 (the multiplicities).
 * D <= 255. */
int mmf_make_source(const float* mult_probs, int32_t B, int32_t D, int32_t V, uint64_t seed, uint64_t first_global_jet,
                    float* x0, int64_t* k0, int64_t* mask, int32_t* n_out, int32_t device, void* stream);

/* The whole N-step sampler on device tensors.
 *   t_grid      host array of the N time points (the caller builds torch.linspace(eps, 1-eps, N) so that it is
 *               bit-identical to the reference), dt = (t[N-1]-t[0])/(N-1) likewise
 *   u           device (N,B,D,V) supplied uniforms or NULL
 *   forced_k    device (N,B,D) uint8 teacher-forced tokens applied after each step, or NULL
 *   rates_out   device (B,D,V) rates of the last step, or NULL
 * k0 / k_out / opts may be NULL for EPiC.  x_out may alias x0, k_out may alias k0.
 * Synchronises twice: the mask is copied to the host to plan the tiles (use mmf_generate_n to avoid it), and the
 * out-of-range-token flag is read back at the end (status 3, the reference's assert of model/MJB.py:177-182). */
int mmf_generate(MmfModel* model, const float* x0, const int64_t* k0, const int64_t* mask, int32_t B, int32_t D,
                 const float* t_grid, int32_t N, float dt, const MmfStepOptions* opts, const float* u,
                 const uint8_t* forced_k, float* x_out, int64_t* k_out, float* rates_out, void* stream);

/* Fully asynchronous form of mmf_generate for the reference's prefix masks (utils/aoj.py:882-883: mask[i, :n_i] = 1):
 * n_per_jet is a HOST array of the B multiplicities, so nothing is copied back and the host never waits - the per-call
 * tables travel through pinned memory and the call returns as soon as the kernel is queued on `stream`.  Out-of-range
 * tokens are reported by mmf_model_status, not by this call. */
int mmf_generate_n(MmfModel* model, const float* x0, const int64_t* k0, const int32_t* n_per_jet, int32_t B, int32_t D,
                   const float* t_grid, int32_t N, float dt, const MmfStepOptions* opts, const float* u,
                   const uint8_t* forced_k, float* x_out, int64_t* k_out, float* rates_out, void* stream);

/* Waits for `stream` and returns 3 (message in mmf_last_error) if a kernel of this handle met a token outside
 * [0, vocab_size) since the last query (reference model/MJB.py:177-182); clears the flag. */
int mmf_model_status(MmfModel* model, void* stream);

/* mmf_generate with HOST buffers (pinned or pageable): stages inputs to the device on `stream`, runs, copies the results
 * back and waits for `stream` (the results are on the host when it returns). */
int mmf_generate_host(MmfModel* model, const float* x0, const int64_t* k0, const int64_t* mask, int32_t B, int32_t D,
                      const float* t_grid, int32_t N, float dt, const MmfStepOptions* opts, float* x_out,
                      int64_t* k_out, void* stream);

/* Number of kernels this handle has launched so far (bench.py reports it as gpu_launches). */
int64_t mmf_launch_count(const MmfModel* model);

/* Optional per-kernel-class profile: when enabled every launch is bracketed by CUDA events on its stream.
 * mmf_profile_read synchronises the device and returns, per class, accumulated milliseconds, launch counts and
 * the algorithmic FLOPs (real particles only) handed to those launches.  Arrays hold mmf_profile_num_classes(). */
int mmf_profile_enable(MmfModel* model, int32_t on);
int32_t mmf_profile_num_classes(void);
const char* mmf_profile_class_name(int32_t cls);
int mmf_profile_read(MmfModel* model, double* ms, int64_t* launches, double* flops, int32_t reset);

/* ---- diagnostics: the building-block kernels, exposed so that tests can check each against torch ---- */
/* out = epilogue(A[M,K] * W[N,K]^T + bias); device pointers; A, W bf16; M multiple of 128, K multiple of 64.
 * mode 0: bf16 out (act 0 none / 1 GELU); mode 1: fp32 out (N multiple of 128) */
int mmf_dbg_gemm(const void* A_bf16, const void* W_bf16, const float* bias, int32_t M, int32_t N, int32_t K,
                 int32_t mode, int32_t act, void* out, int32_t device, void* stream);
/* residual[M,C] += A*W^T + bias (+ temb[0,:]); act_out[M,C] = bf16(LayerNorm(residual)); C in {128,256} */
int mmf_dbg_gemm_resln(const void* A_bf16, const void* W_bf16, const float* bias, const float* temb,
                       const float* ln_g, const float* ln_b, int32_t M, int32_t C, int32_t K, float* residual,
                       void* act_out_bf16, int32_t device, void* stream);
/* fused QKV projection: q,k [M,C] bf16 with per-head LayerNorm, vT [C,M] bf16 */
int mmf_dbg_gemm_qkv(const void* A_bf16, const void* W_bf16, const float* bias, const float* q_g, const float* q_b,
                     const float* k_g, const float* k_b, int32_t M, int32_t C, int32_t hs, void* q_out, void* k_out,
                     void* vT_out, int32_t device, void* stream);
/* masked attention over packed jets: q,k [M,C] bf16, vT [C,M] bf16, jet_n host array of n_jets multiplicities
 * (rows are the jets back to back); out [M,C] bf16.  Synchronous. */
int mmf_dbg_attention(const void* q, const void* k, const void* vT, const int32_t* jet_n, int32_t n_jets, int32_t M,
                      int32_t C, int32_t hs, void* out, int32_t device, void* stream);

/* host only: placement of n weight tiles (sizes in KB, consumption order, repeated every timestep) in the 64 KB
 * shared-memory ring of the persistent tile kernel.  dst_kb[i] = offset in KB; dep[i]: tile g of the launch may be written
 * once all tiles up to g - dep[i] have been consumed.  Fails when a tile cannot be placed. */
int mmf_dbg_ring_plan(const int32_t* tile_kb, int32_t n, int32_t* dst_kb, int32_t* dep);

#ifdef __cplusplus
}
#endif
#endif /* MMF_B200_H */
