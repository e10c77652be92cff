// The tensor-core GEMM of the training step: one tcgen05 kernel serves the forward linears, the
// data gradients and the weight gradients of every nn.Linear on the encoder path (reference networks/attention.py:44-45,
// utils/models.py:15-17, networks/ParticleTransformers.py:28-53 under autograd).
//
//   C[M x N] (+)= A[M x K] * B[N x K]^T (+ bias[N])      A, B bf16 with K contiguous, fp32 accumulation in TMEM
//
//     forward      y  = x  W^T + b      A = x  [tokens, in]     B = W    [out, in]
//     data grad    dx = dy W            A = dy [tokens, out]    B = W^T  [in, out]     (bf16 transposed copy of the weight)
//     weight grad  dW = dy^T x          A = dy [tokens, out]    B = x    [tokens, in]  both read as MN-MAJOR operands (no transposed
//                                       copies exist anywhere); K = tokens, split over blockIdx.z, the partial tiles meet in a
//                                       TMA reduce-add
//
// One CTA = one 128 x 128 output tile: a TMA producer thread and an MMA issuer thread run a 4-stage ring of 128-byte-swizzled
// [128 x 64] operand boxes (MN-major operands: two [64 tokens x 64 features] boxes each); the ring is as deep as the K loop is
// long (2 ... 4 stages), so short-K products run three CTAs per SM and hide each other's prologue / epilogue; four epilogue warps move the accumulator TMEM -> registers -> swizzled shared memory and one thread
// stores (or reduce-adds) the tile with TMA.  Tensor maps carry the exact extents, so ragged M / N / K edges are zero-filled on
// load and clipped on store - no padding rules for the callers.
#include "mmf_internal.h"
#include "mmf_ptx.cuh"
#include "mmf_tile.cuh"
#include "mmf_train.h"

namespace mmf {
namespace {

constexpr int kTrStages = 3;              // 98 KB of operand ring at most: two CTAs per SM even for the long split-K loops
constexpr int kTrBox = kTileM * 128;             // one operand box: 128 rows x 128 B
constexpr int kTrStage = 2 * kTrBox;             // A + B
constexpr int tr_smem_bytes(int stages) { return 1024 + stages * kTrStage + 1024; }

// MN-major operand, 128-byte swizzle: [8 K-rows][64 elements] atoms of 1024 B; SBO = 1024 B between 8-row groups along K,
// LBO = 8192 B between the two 64-element groups along M / N (two [64 x 64] TMA boxes side by side)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(8192 >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

struct TrGemmBars {
    uint64_t full[kTrStages], empty[kTrStages], acc_full, aux_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_exact(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
template <int N>
__device__ __forceinline__ void head_layernorm(float* v, const float* g, const float* b) {     // reference utils/models.py:36-37, eps 1e-5
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) s += v[i];
    const float mean = s * (1.0f / N);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(q * (1.0f / N) + 1e-5f);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fmaf((v[i] - mean) * rstd, __ldg(g + i), b ? __ldg(b + i) : 0.f);
}
__device__ __forceinline__ void stage_bf16_32(uint8_t* chunk, int r, int half, const float* v) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
        st_shared_v4(chunk + sw128_offset(r, half * 4 + u), pack_bf16x2(v[u * 8 + 0], v[u * 8 + 1]), pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]),
                     pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]), pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]));
}

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}

// MODE 0: bf16 store, 1: fp32 store, 2: fp32 reduce-add, 3: bf16 store of z AND of GELU(z) (second output through tmD),
// 4: bf16 store of acc * GELU'(z), z tile loaded through tmD, 5: fp32 store of resid + acc + bias (+ tadd[row_jet]): the residual
// stream written out of place by the projection itself, 6: c_attn: bf16 store of q | k | v AND of the per-head LayerNorm of the q
// and k sections (through tmD, a [M x 2C] buffer); TN: A [K x M], B [K x N] row-major (MN-major operands)
template <int MODE, bool TN>
__global__ void __launch_bounds__(192, 3)
tr_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const float* __restrict__ bias,
               const int N, const int kb_total, const int kb_per_split, const int stages, const TrGemmResid rs) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    TrGemmBars* bars = reinterpret_cast<TrGemmBars*>(smem);
    uint8_t* tiles = smem + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTileM, n0 = blockIdx.y * 128;
    const int kb0 = blockIdx.z * kb_per_split;
    const int nkb = min(kb_per_split, kb_total - kb0);

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        if (MODE >= 3) tma_prefetch_desc(&tmD);
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int i = 0; i < kTrStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
            mbar_init(&bars->acc_full, 1);
            mbar_init(&bars->aux_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&bars->tmem_base, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    grid_dep_wait();                 // PDL: everything above overlapped the previous kernel; its results are visible from here
    grid_dep_launch();

    uint8_t* aux = tiles + stages * kTrStage;                 // MODE 4: the [128 x 128] bf16 tile of pre-activations
    if (warp == 4) {
        if (lane == 0) {
            if constexpr (MODE == 4) {
                mbar_expect_tx(&bars->aux_full, kTrStage);
                tma_load_2d(aux, &tmD, &bars->aux_full, n0, m0);
                tma_load_2d(aux + kTrBox, &tmD, &bars->aux_full, n0 + 64, m0);
            }
            for (int i = 0; i < nkb; ++i) {
                const int s = i % stages, it = i / stages;
                if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
                mbar_expect_tx(&bars->full[s], kTrStage);
                uint8_t* sa = tiles + s * kTrStage;
                if constexpr (TN) {
                    tma_load_2d(sa, &tmA, &bars->full[s], m0, (kb0 + i) * kBK);
                    tma_load_2d(sa + kTrBox / 2, &tmA, &bars->full[s], m0 + 64, (kb0 + i) * kBK);
                    tma_load_2d(sa + kTrBox, &tmB, &bars->full[s], n0, (kb0 + i) * kBK);
                    tma_load_2d(sa + kTrBox + kTrBox / 2, &tmB, &bars->full[s], n0 + 64, (kb0 + i) * kBK);
                } else {
                    tma_load_2d(sa, &tmA, &bars->full[s], (kb0 + i) * kBK, m0);
                    tma_load_2d(sa + kTrBox, &tmB, &bars->full[s], (kb0 + i) * kBK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, 128) | (TN ? (1u << 15) | (1u << 16) : 0u);   // bits 15 / 16: A / B MN-major
            constexpr uint32_t kstep = TN ? (2048u >> 4) : 2u;          // K = 16: 16 rows of 128 B (MN-major) or 32 bytes (K-major)
            for (int i = 0; i < nkb; ++i) {
                const int s = i % stages, it = i / stages;
                mbar_wait(&bars->full[s], it & 1);
                tc_fence_after();
                const uint32_t aa = smem_u32(tiles + s * kTrStage), ab = aa + kTrBox;
                const uint64_t da = TN ? umma_desc_sw128_mn(aa) : umma_desc_sw128(aa);
                const uint64_t db = TN ? umma_desc_sw128_mn(ab) : umma_desc_sw128(ab);
#pragma unroll
                for (int ks = 0; ks < kBK / 16; ++ks) umma_bf16(tmem_base, da + kstep * ks, db + kstep * ks, idesc, (i | ks) != 0 ? 1u : 0u);
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->acc_full);
        }
        __syncwarp();
    } else {
        // epilogue: thread = accumulator row; the whole 128 x 128 tile is staged in the (now idle) operand ring
        const int r = warp * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const bool add_bias = bias != nullptr && (MODE != 2 || blockIdx.z == 0);
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();
        if constexpr (MODE == 6) {
            const int sect = n0 / rs.C;                                 // 0: q, 1: k, 2: v (C is a multiple of the 128-column tile)
            const float* hg = sect == 0 ? rs.qg : rs.kg;
            const float* hb = sect == 0 ? rs.qb : rs.kb;
            for (int cc = 0; cc < 2; ++cc) {
                float v[64];
                tmem_ld32(taddr + cc * 64, v);
                tmem_ld32(taddr + cc * 64 + 32, v + 32);
                tmem_ld_wait();
                if (add_bias) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) v[i] += __ldg(bias + n0 + cc * 64 + i);
                }
                stage_bf16_32(tiles + cc * kTrBox, r, 0, v);
                stage_bf16_32(tiles + cc * kTrBox, r, 1, v + 32);
                if (sect < 2) {                                          // LayerNorm over each head of the bf16 values the backward pass reads
#pragma unroll
                    for (int i = 0; i < 64; ++i) v[i] = bf16_round(v[i]);
                    if (rs.hs == 32) { head_layernorm<32>(v, hg, hb); head_layernorm<32>(v + 32, hg, hb); }
                    else head_layernorm<64>(v, hg, hb);
                    stage_bf16_32(tiles + (2 + cc) * kTrBox, r, 0, v);
                    stage_bf16_32(tiles + (2 + cc) * kTrBox, r, 1, v + 32);
                }
            }
        } else if constexpr (MODE == 0 || MODE == 3 || MODE == 4) {
            if constexpr (MODE == 4) mbar_wait(&bars->aux_full, 0);
            for (int c = 0; c < 4; ++c) {
                float v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                if (add_bias) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const int col = n0 + c * 32 + i; v[i] += col < N ? __ldg(bias + col) : 0.f; }
                }
                if constexpr (MODE == 4) {                              // dz = dh GELU'(z)
                    const uint8_t* zc = aux + (c >> 1) * kTrBox;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 w = ld_shared_f4(zc + sw128_offset(r, (c & 1) * 4 + u));
                        const uint32_t ww[4] = {__float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            v[u * 8 + 2 * e] *= gelu_grad_exact(__uint_as_float(ww[e] << 16));
                            v[u * 8 + 2 * e + 1] *= gelu_grad_exact(__uint_as_float(ww[e] & 0xffff0000u));
                        }
                    }
                }
                stage_bf16_32(tiles + (c >> 1) * kTrBox, r, c & 1, v);   // 64 bf16 columns per [128 x 128 B] staging chunk
                if constexpr (MODE == 3) {                              // h = GELU(z) of the bf16 z the backward pass will read
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = gelu_exact(bf16_round(v[i]));
                    stage_bf16_32(tiles + (2 + (c >> 1)) * kTrBox, r, c & 1, v);
                }
            }
        } else {
            for (int c = 0; c < 4; ++c) {
                float v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                if (add_bias) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const int col = n0 + c * 32 + i; v[i] += col < N ? __ldg(bias + col) : 0.f; }
                }
                if constexpr (MODE == 5) {
                    const int row = m0 + r;
                    if (row < rs.M) {                                   // (rows beyond M are clipped by the store)
                        const float4* rp = reinterpret_cast<const float4*>(rs.resid + row * rs.ldr + n0 + c * 32);
                        const float4* tp = rs.tadd ? reinterpret_cast<const float4*>(rs.tadd + static_cast<long long>(rs.row_jet ? rs.row_jet[row] : 0) * rs.ldt + n0 + c * 32) : nullptr;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 a4 = rp[u];
                            v[u * 4] += a4.x; v[u * 4 + 1] += a4.y; v[u * 4 + 2] += a4.z; v[u * 4 + 3] += a4.w;
                            if (tp) { const float4 t4 = __ldg(tp + u); v[u * 4] += t4.x; v[u * 4 + 1] += t4.y; v[u * 4 + 2] += t4.z; v[u * 4 + 3] += t4.w; }
                        }
                    }
                }
                stage_row_f32(tiles + c * kTrBox, r, v);
            }
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (threadIdx.x == 0) {
            if constexpr (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 6) {
                for (int cc = 0; cc < 2; ++cc) tma_store_2d(&tmC, tiles + cc * kTrBox, n0 + cc * 64, m0);
                if (MODE == 3 || (MODE == 6 && n0 < 2 * rs.C))
                    for (int cc = 0; cc < 2; ++cc) tma_store_2d(&tmD, tiles + (2 + cc) * kTrBox, n0 + cc * 64, m0);
            } else if constexpr (MODE == 1 || MODE == 5) {
                for (int c = 0; c < 4; ++c) tma_store_2d(&tmC, tiles + c * kTrBox, n0 + c * 32, m0);
            } else {
                for (int c = 0; c < 4; ++c) tma_reduce_add_2d(&tmC, tiles + c * kTrBox, n0 + c * 32, m0);
            }
            tma_store_commit();
            tma_store_wait_read<0>();        // the staging memory must outlive the reads; the writes are complete at grid end
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

template <int MODE, bool TN>
int launch_mode(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmD, const float* bias, int N,
                int kb_total, int kb_per_split, dim3 grid, cudaStream_t s, const TrGemmResid& rs = TrGemmResid{}) {
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        MMF_CUDA_OK(cudaFuncSetAttribute(tr_gemm_kernel<MODE, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem_bytes(kTrStages + 1)));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    // the epilogue stages the whole output tile in the ring: two stages hold it (64 KB fp32)
    const int stages = kb_per_split < 2 ? 2 : (kb_per_split > kTrStages ? kTrStages : kb_per_split);
    const int smem = tr_smem_bytes(stages + (MODE == 4 ? 1 : 0));           // + the pre-activation tile
    MMF_CUDA_OK(tr_launch(tr_gemm_kernel<MODE, TN>, grid, dim3(192), smem, s, tmA, tmB, tmC, tmD, bias, N, kb_total, kb_per_split, stages, rs));
    return 0;
}

}  // namespace

int launch_tr_gemm(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N, int K,
                   const float* bias, int mode, int ksplit, void* aux, long long ldaux, const TrGemmResid* resid, cudaStream_t s) {
    if (M <= 0 || N <= 0) return 0;                      // an empty batch: nothing to do (empty torch tensors have null pointers)
    MMF_REQUIRE(A && B && C, "gemm: null operand");
    MMF_REQUIRE(mode >= 0 && mode <= 5, "gemm: mode is 0 (bf16), 1 (fp32), 2 (fp32 reduce-add), 3 (bf16 + GELU copy), 4 (bf16 times GELU'(aux)), 5 (fp32 residual + product)");
    MMF_REQUIRE((mode != 3 && mode != 4) || aux, "gemm: modes 3 and 4 need the auxiliary [M x N] bf16 tensor");
    MMF_REQUIRE(mode != 5 || (resid && resid->resid && N % 128 == 0), "gemm: mode 5 needs the residual input and N a multiple of 128");
    if (M <= 0 || N <= 0) return 0;
    MMF_REQUIRE(K > 0, "gemm: K must be positive");
    const int kb_total = (K + kBK - 1) / kBK;
    if (ksplit < 1 || mode != 2) ksplit = 1;
    if (ksplit > kb_total) ksplit = kb_total;
    const int per = (kb_total + ksplit - 1) / ksplit;
    ksplit = (kb_total + per - 1) / per;                  // no empty split
    CUtensorMap tmA, tmB, tmC, tmD;
    if (make_tmap_2d(&tmA, A, 2, M, K, lda, 64, 128)) return 1;
    if (make_tmap_2d(&tmB, B, 2, N, K, ldb, 64, 128)) return 1;
    if (mode == 0 || mode == 3 || mode == 4) { if (make_tmap_2d(&tmC, C, 2, M, N, ldc, 64, 128)) return 1; }
    else { if (make_tmap_2d(&tmC, C, 4, M, N, ldc, 32, 128)) return 1; }
    tmD = tmC;
    if ((mode == 3 || mode == 4) && make_tmap_2d(&tmD, aux, 2, M, N, ldaux, 64, 128)) return 1;
    const dim3 grid((M + kTileM - 1) / kTileM, (N + 127) / 128, ksplit);
    switch (mode) {
        case 0: return launch_mode<0, false>(tmA, tmB, tmC, tmD, bias, N, kb_total, per, grid, s);
        case 1: return launch_mode<1, false>(tmA, tmB, tmC, tmD, bias, N, kb_total, per, grid, s);
        case 2: return launch_mode<2, false>(tmA, tmB, tmC, tmD, bias, N, kb_total, per, grid, s);
        case 3: return launch_mode<3, false>(tmA, tmB, tmC, tmD, bias, N, kb_total, per, grid, s);
        case 4: return launch_mode<4, false>(tmA, tmB, tmC, tmD, bias, N, kb_total, per, grid, s);
        default: {
            TrGemmResid rs = *resid;
            rs.M = M;
            return launch_mode<5, false>(tmA, tmB, tmC, tmD, bias, N, kb_total, per, grid, s, rs);
        }
    }
}

// SelfAttention.c_attn + q / k LayerNorm (reference attention.py:57-64): qkv [M x 3C] bf16 and the normalised q | k [M x 2C] bf16
int launch_tr_gemm_qkv(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* qkv, long long ldq, void* qkn,
                       long long ldn, int M, int C, int K, int hs, const float* qg, const float* qb, const float* kg, const float* kb,
                       cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_REQUIRE(A && W && qkv && qkn && qg && kg, "gemm_qkv: null argument");
    MMF_REQUIRE(C % 128 == 0 && (hs == 32 || hs == 64), "gemm_qkv: width a multiple of 128, head size 32 or 64");
    const int kb_total = (K + kBK - 1) / kBK;
    CUtensorMap tmA, tmB, tmC, tmD;
    if (make_tmap_2d(&tmA, A, 2, M, K, lda, 64, 128) || make_tmap_2d(&tmB, W, 2, 3 * C, K, ldw, 64, 128) ||
        make_tmap_2d(&tmC, qkv, 2, M, 3 * C, ldq, 64, 128) || make_tmap_2d(&tmD, qkn, 2, M, 2 * C, ldn, 64, 128)) return 1;
    TrGemmResid rs{};
    rs.M = M; rs.C = C; rs.hs = hs; rs.qg = qg; rs.qb = qb; rs.kg = kg; rs.kb = kb;
    const dim3 grid((M + kTileM - 1) / kTileM, 3 * C / 128, 1);
    return launch_mode<6, false>(tmA, tmB, tmC, tmD, bias, 3 * C, kb_total, kb_total, grid, s, rs);
}

// C[M x N] += A^T B with A [K x M], B [K x N] row-major bf16 (the weight gradient dW += dy^T x straight from the row-major
// activations: both operands MN-major); K split over `ksplit` CTAs per output tile, fp32 TMA reduce-add
int launch_tr_gemm_tn(const void* A, long long lda, const void* B, long long ldb, float* C, long long ldc, int M, int N, int K, int ksplit,
                      cudaStream_t s) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    MMF_REQUIRE(A && B && C, "gemm_tn: null operand");
    const int kb_total = (K + kBK - 1) / kBK;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > kb_total) ksplit = kb_total;
    const int per = (kb_total + ksplit - 1) / ksplit;
    ksplit = (kb_total + per - 1) / per;
    CUtensorMap tmA, tmB, tmC;
    if (make_tmap_2d(&tmA, A, 2, K, M, lda, 64, 64)) return 1;
    if (make_tmap_2d(&tmB, B, 2, K, N, ldb, 64, 64)) return 1;
    if (make_tmap_2d(&tmC, C, 4, M, N, ldc, 32, 128)) return 1;
    const dim3 grid((M + kTileM - 1) / kTileM, (N + 127) / 128, ksplit);
    return launch_mode<2, true>(tmA, tmB, tmC, tmC, nullptr, N, kb_total, per, grid, s);
}

}  // namespace mmf
