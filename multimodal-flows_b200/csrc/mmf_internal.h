// Internal interfaces between the translation units of libmmf_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmf_common.cuh"
#include "step_math.cuh"

namespace mmf {

typedef __nv_bfloat16 bf16;

constexpr int kTileM = 128;     // rows of one tcgen05 accumulator tile (= TMEM lanes)
constexpr int kBK = 64;         // bf16 elements per 128-byte swizzled smem row
constexpr int kMaxKeys = 160;   // attention keys per work item (jets have <= 150 particles)

// ------------------------------------------------------------------ tensor-core GEMM
enum GemmEpilogue { EPI_STORE_BF16 = 0, EPI_STORE_F32 = 1, EPI_QKV = 2, EPI_RESLN = 3 };

struct GemmArgs {
    int kblocks;                 // K / 64 (per group)
    int a_col_group_stride;      // column offset of group g inside the A matrix (elements)
    int w_rows_per_group;        // N per group = rows of the stacked weight per group
    const float* bias;           // [G * N] or null
    int act;                     // 0 none, 1 exact-erf GELU           (STORE_*)
    int out_col_group_stride;    // column offset of group g in the output (STORE_*)
    int sect_width;              // C: width of the q / k / v sections (QKV) and of the LN row (RESLN)
    int hs;                      // head size (QKV)
    const float *q_g, *q_b, *k_g, *k_b;   // per-head LayerNorm affine, [G * hs] (QKV); null = no q/k LN
    bf16* vt;                    // V^T [G*C rows][vt_ld tokens] (QKV)
    long long vt_ld;
    const float* temb;           // additive time embedding rows [*, temb_ld] or null (RESLN)
    int temb_ld;
    const int* row_jet;          // row -> temb row; null = every row uses temb row 0 (RESLN)
    const float *ln_g, *ln_b;    // LayerNorm applied to the updated residual -> bf16 out (RESLN); null = skip
};

// tmA: bf16 activations [rows, cols] box {64,128};  tmB: stacked bf16 weight [G*N, K] box {64, BN}
// tmOut0: STORE_BF16/RESLN: bf16 out box {64,128};  STORE_F32: fp32 out box {32,128};  QKV: Q
// tmOut1: QKV: K;  RESLN: fp32 residual box {32,128} (loaded, updated in place, stored)
int launch_gemm(int epilogue, int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut0,
                const CUtensorMap& tmOut1, const GemmArgs& args, int m_tiles, int n_tiles, int groups,
                cudaStream_t stream);
int gemm_smem_bytes(int epilogue, int BN, int kblocks);

// ------------------------------------------------------------------ tensor-core attention
struct AttnItem {       // one CTA worth of work: <=128 query rows against <=160 key rows of the packed layout
    int q_row0, nq, k_row0, nk;
};
struct AttnArgs {
    const AttnItem* items;
    const int* seg_beg;      // per packed row: first row of its jet
    const int* seg_end;      // per packed row: one past the last row of its jet
    bf16* out;               // [rows, ld_out] attention output (heads side by side)
    int ld_out;
    int hs;                  // 32 or 64; each CTA owns a 64-column slab = 64/hs heads
    float scale_log2e;       // (1/sqrt(hs)) * log2(e)
};
// tmQ: box {64,128}; tmK: box {64,32}; tmVT: V^T [C rows, tokens] box {64 tokens, 64 rows}
int launch_attention(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmVT, const AttnArgs& args,
                     int n_items, int n_slabs, cudaStream_t stream);

// ------------------------------------------------------------------ tensor maps
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_cols, uint32_t box_rows);

// ------------------------------------------------------------------ SIMT kernels (kernels_simt.cu)
struct StepLaunch {
    StepParams sp;
    const float* u;          // (B*D, V) supplied uniforms or null -> Philox
    uint64_t seed;
    uint64_t slot0;          // global particle slot of element 0 (first_global_jet * D)
    uint32_t step;
    int* err_flag;           // device int, bit1 set on out-of-range token
};
// padded (B, D) layout, in place: x (B,D,3) k (B,D) int64, t (B,) per jet
int launch_hybrid_step(const float* vt, const float* logits, float* x, long long* k, const float* t, int B, int D,
                       const StepLaunch& sl, float* rates_out, cudaStream_t stream);
// continuous-only Euler update x += vt dt (EPiC carrier)
int launch_euler(const float* vt, float* x, float dt, long long n, cudaStream_t stream);

// jet observables (kernels_simt.cu 1c): kin (B, kObsKin) = px py pz E pt m eta phi charge jet_charge multiplicity m2
constexpr int kObsKin = 12;
struct ObsArgs {
    long long B;
    int D, V;
    float mean[3], std[3];   // de-standardisation x * std + mean
    float* kin;              // (B, kObsKin)
    int* counts;             // (B, V) tokens among unmasked particles, or null
};
int launch_jet_observables(const float* x, const long long* k, const long long* mask, const ObsArgs& a, cudaStream_t stream);

// source state on the device (kernels_simt.cu 1d)
constexpr int kSrcMaxD = 255;
struct SourceArgs {
    long long B;
    int D, V;
    unsigned long long seed, first_jet, div_magic;   // div_magic = floor(2^64 / D) + 1
    float cdf[kSrcMaxD + 1];                         // cdf[m] = P(multiplicity <= m), m = 0..D
};
int launch_make_source(const SourceArgs& a, float* x0, long long* k0, long long* mask, int* n_out, cudaStream_t stream);

}  // namespace mmf
