// Masked self-attention of the training step on tensor cores (reference networks/attention.py:53-74 and its autograd):
// forward and backward as tcgen05 kernels over TILES of whole jets.
//
// A work item is a run of consecutive whole jets with at most 128 particles in total (= the 128 TMEM lanes); attention inside
// an item is block diagonal: row r attends the rows of its own jet, [jet_off[row_jet[r]], jet_off[row_jet[r] + 1]).  One CTA
// handles one (item, 64-column slab) = one head of 64 or two heads of 32.  Thread r of the four warps owns query row r.
//
//   forward    S = Q K^T -> softmax in registers (two passes over the score tile in TMEM) -> P (bf16, shared memory) -> O = P V.
//              Only the row statistics (max * scale * log2e, 1 / sum) are kept for the backward pass: 8 bytes per (row, head).
//   backward   S = Q K^T and dP = dO V^T again on the tensor cores; P and dS = P (dP - sum_j P dP) * scale are rebuilt row by
//              row (the row sum uses the very P that multiplies it, so sum_j dS_ij = 0 holds to fp32 rounding); then
//              dQ = dS K, dK = dS^T Q, dV = P^T dO.  P and dS take turns in one pair of shared-memory boxes and the gradients
//              reuse the accumulator columns of dP, so two CTAs fit on an SM (96 KB, 256 TMEM columns each).
//
// No operand is ever transposed in memory: V, K, Q and dO are read as MN-major B operands ([rows = contraction index][64
// features]) where the contraction runs over particles, P and dS as MN-major A operands for the "transposed" products.
// Jets of more than 128 particles (two tiles sharing keys) stay on the CUDA-core kernels of kernels_trainops.cu.
#include "mmf_internal.h"
#include "mmf_ptx.cuh"
#include "mmf_tile.cuh"
#include "mmf_train.h"

namespace mmf {
namespace {

constexpr int kBox = kTileM * 128;               // [128 rows][128 B] = 16 KB: one 64-column bf16 slab of 128 rows
constexpr uint32_t kRowStep16 = 2048u >> 4;      // 16 rows of 128 B: one K = 16 step of an MN-major operand

struct AttnTcBars {
    uint64_t loaded, mma_a, mma_b, mma_c;
    uint32_t tmem_base;
};

// MN-major operand made of 64-element-wide [rows][128 B] boxes `lbo_bytes` apart (one box: lbo ignored)
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
constexpr uint32_t kAMn = 1u << 15, kBMn = 1u << 16;       // instruction-descriptor bits: A / B operand is MN-major

__device__ __forceinline__ void row_segment(const TrAttnTcArgs& a, int row0, int r, int nrows, int* kb, int* ke) {
    *kb = 0; *ke = 0;
    if (r < nrows) {
        const int jet = a.row_jet[row0 + r];
        *kb = a.jet_off[jet] - row0;
        *ke = a.jet_off[jet + 1] - row0;
    }
}

__device__ __forceinline__ void store_bf16_row32(uint8_t* chunk, int r, int half, const float* v) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
        st_shared_v4(chunk + sw128_offset(r, half * 4 + u), pack_bf16x2(v[u * 8 + 0], v[u * 8 + 1]), pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]),
                     pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]), pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]));
}

// ------------------------------------------------------------------------------------------------ forward
template <int HS>
__global__ void __launch_bounds__(128)
tr_attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const TrAttnTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    AttnTcBars* bars = reinterpret_cast<AttnTcBars*>(smem);
    uint8_t* Qs = smem + 1024;
    uint8_t* Ks = Qs + kBox;
    uint8_t* Vs = Ks + kBox;
    uint8_t* Ps = Vs + kBox;                     // two boxes: keys [0,64) and [64,128)
    if (static_cast<int>(blockIdx.x) >= *a.n_items) return;        // (n_items comes from a host copy, not from the previous kernel)
    const int2 item = a.items[blockIdx.x];
    const int row0 = item.x, nrows = item.y;
    const int col0 = blockIdx.y * 64;
    const int warp = threadIdx.x >> 5, r = threadIdx.x;
    constexpr int NH = 64 / HS, KS = HS / 16;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(&bars->loaded, 1); mbar_init(&bars->mma_a, 1); mbar_init(&bars->mma_b, 1);
        fence_mbar_init();
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(&bars->tmem_base, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    grid_dep_wait();                             // PDL: the prologue above overlapped the previous kernel
    grid_dep_launch();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars->loaded, 3 * kBox);
        tma_load_2d(Qs, &tmQ, &bars->loaded, col0, row0);
        tma_load_2d(Ks, &tmK, &bars->loaded, col0, row0);
        tma_load_2d(Vs, &tmV, &bars->loaded, col0, row0);
    }
    int kb, ke;
    row_segment(a, row0, r, nrows, &kb, &ke);
    const bool valid = r < nrows;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

    for (int h = 0; h < NH; ++h) {
        if (threadIdx.x == 0) {
            if (h == 0) mbar_wait(&bars->loaded, 0);
            tc_fence_after();
            const uint64_t dq = umma_desc_sw128(smem_u32(Qs)) + 2 * (h * KS), dk = umma_desc_sw128(smem_u32(Ks)) + 2 * (h * KS);
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) umma_bf16(tmem_base, dq + 2 * ks, dk + 2 * ks, idesc, ks != 0 ? 1u : 0u);
            umma_commit(&bars->mma_a);
        }
        __syncwarp();
        mbar_wait(&bars->mma_a, h & 1);
        tc_fence_after();
        float mx = -INFINITY;
        for (int c = 0; c < 4; ++c) {
            float s[32];
            tmem_ld32(taddr + c * 32, s);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { const int col = c * 32 + j; if (col >= kb && col < ke) mx = fmaxf(mx, s[j]); }
        }
        const float mscaled = valid ? mx * a.scale_log2e : 0.f;
        float sum = 0.f;
        for (int c = 0; c < 4; ++c) {
            float s[32];
            tmem_ld32(taddr + c * 32, s);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c * 32 + j;
                const float p = (col >= kb && col < ke) ? exp2f(fmaf(s[j], a.scale_log2e, -mscaled)) : 0.f;
                s[j] = p;
                sum += p;
            }
            store_bf16_row32(Ps + (c >> 1) * kBox, r, c & 1, s);
        }
        const float inv = 1.0f / (sum > 0.f ? sum : 1.f);
        if (valid) {
            float2* st = reinterpret_cast<float2*>(a.stats) + (static_cast<long long>(row0 + r) * a.H + blockIdx.y * NH + h);
            *st = make_float2(mscaled, inv);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            // O = P V: A = P K-major (keys contiguous, two boxes of 64 keys), B = V MN-major ([key rows][64 features])
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64) | kBMn;
            const uint64_t dv = umma_desc_sw128(smem_u32(Vs));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint64_t dp = umma_desc_sw128(smem_u32(Ps + (kk >> 2) * kBox)) + 2 * (kk & 3);
                umma_bf16(tmem_base + 128, dp, dv + kRowStep16 * kk, idesc, kk != 0 ? 1u : 0u);
            }
            umma_commit(&bars->mma_b);
        }
        __syncwarp();
        mbar_wait(&bars->mma_b, h & 1);
        tc_fence_after();
        {
            float o[HS];
#pragma unroll
            for (int c = 0; c < HS / 32; ++c) tmem_ld32(taddr + 128 + h * HS + c * 32, o + c * 32);
            tmem_ld_wait();
            if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(a.o + static_cast<long long>(row0 + r) * a.ldo + col0 + h * HS);
#pragma unroll
                for (int u = 0; u < HS / 8; ++u)
                    dst[u] = make_uint4(pack_bf16x2(o[u * 8 + 0] * inv, o[u * 8 + 1] * inv), pack_bf16x2(o[u * 8 + 2] * inv, o[u * 8 + 3] * inv),
                                        pack_bf16x2(o[u * 8 + 4] * inv, o[u * 8 + 5] * inv), pack_bf16x2(o[u * 8 + 6] * inv, o[u * 8 + 7] * inv));
            }
        }
        tc_fence_before();
        __syncthreads();                         // the next head's S overwrites the score columns, its P the probability boxes
    }
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ backward
// Two CTAs share an SM (96 KB of shared memory, 256 TMEM columns each): P and dS take turns in ONE pair of boxes and the
// gradients land in the accumulator columns of dP once that is consumed.  Per head:
//   S = Q K^T -> [0,128), dP = dO V^T -> [128,256)          | rows: delta = sum_j P dP, dS -> boxes
//   dQ = dS K -> [128,192), dK = dS^T Q -> [192,256)        | rows: store dQ, dK; P (from S, still intact) -> boxes
//   dV = P^T dO -> [128,192)                                | rows: store dV
template <int HS>
__global__ void __launch_bounds__(128, 2)
tr_attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmD, const TrAttnTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    AttnTcBars* bars = reinterpret_cast<AttnTcBars*>(smem);
    uint8_t* Qs = smem + 1024;
    uint8_t* Ks = Qs + kBox;
    uint8_t* Vs = Ks + kBox;
    uint8_t* Ds = Vs + kBox;                     // dO
    uint8_t* Xs = Ds + kBox;                     // dS, then P: two boxes of 64 keys
    if (static_cast<int>(blockIdx.x) >= *a.n_items) return;        // (n_items comes from a host copy, not from the previous kernel)
    const int2 item = a.items[blockIdx.x];
    const int row0 = item.x, nrows = item.y;
    const int col0 = blockIdx.y * 64;
    const int warp = threadIdx.x >> 5, r = threadIdx.x;
    constexpr int NH = 64 / HS, KS = HS / 16;
    constexpr uint32_t cS = 0, cP = 128, cQ = 128, cK = 192, cV = 128;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmD);
        mbar_init(&bars->loaded, 1); mbar_init(&bars->mma_a, 1); mbar_init(&bars->mma_b, 1); mbar_init(&bars->mma_c, 1);
        fence_mbar_init();
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(&bars->tmem_base, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    grid_dep_wait();
    grid_dep_launch();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars->loaded, 4 * kBox);
        tma_load_2d(Qs, &tmQ, &bars->loaded, col0, row0);
        tma_load_2d(Ks, &tmK, &bars->loaded, col0, row0);
        tma_load_2d(Vs, &tmV, &bars->loaded, col0, row0);
        tma_load_2d(Ds, &tmD, &bars->loaded, col0, row0);
    }
    int kb, ke;
    row_segment(a, row0, r, nrows, &kb, &ke);
    const bool valid = r < nrows;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    auto store_out = [&](uint32_t col, int w, int h) {        // dq | dk | dv sections of dqkv; row r is a query (dq) or a key (dk, dv)
        float v[HS];
#pragma unroll
        for (int c = 0; c < HS / 32; ++c) tmem_ld32(taddr + col + h * HS + c * 32, v + c * 32);
        tmem_ld_wait();
        if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(a.dqkv + static_cast<long long>(row0 + r) * a.ldd + w * a.C + col0 + h * HS);
#pragma unroll
            for (int u = 0; u < HS / 8; ++u)
                dst[u] = make_uint4(pack_bf16x2(v[u * 8 + 0], v[u * 8 + 1]), pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]),
                                    pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]), pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]));
        }
    };

    for (int h = 0; h < NH; ++h) {
        if (threadIdx.x == 0) {
            if (h == 0) mbar_wait(&bars->loaded, 0);
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
            const uint32_t sub = 2 * (h * KS);                           // this head's 16-element steps inside the 128-byte rows
            const uint64_t dq = umma_desc_sw128(smem_u32(Qs)) + sub, dk = umma_desc_sw128(smem_u32(Ks)) + sub;
            const uint64_t dd = umma_desc_sw128(smem_u32(Ds)) + sub, dv = umma_desc_sw128(smem_u32(Vs)) + sub;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) umma_bf16(tmem_base + cS, dq + 2 * ks, dk + 2 * ks, idesc, ks != 0 ? 1u : 0u);   // S  = Q K^T
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) umma_bf16(tmem_base + cP, dd + 2 * ks, dv + 2 * ks, idesc, ks != 0 ? 1u : 0u);   // dP = dO V^T
            umma_commit(&bars->mma_a);
        }
        __syncwarp();
        float2 st = make_float2(0.f, 0.f);
        if (valid) st = *(reinterpret_cast<const float2*>(a.stats) + (static_cast<long long>(row0 + r) * a.H + blockIdx.y * NH + h));
        mbar_wait(&bars->mma_a, h & 1);
        tc_fence_after();
        // delta = sum_j P_ij dP_ij with the P that multiplies it below
        float delta = 0.f;
        for (int c = 0; c < 4; ++c) {
            float s[32], dp[32];
            tmem_ld32(taddr + cS + c * 32, s);
            tmem_ld32(taddr + cP + c * 32, dp);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c * 32 + j;
                const float p = (col >= kb && col < ke) ? exp2f(fmaf(s[j], a.scale_log2e, -st.x)) * st.y : 0.f;
                delta = fmaf(p, dp[j], delta);
            }
        }
        for (int c = 0; c < 4; ++c) {
            float s[32], dp[32];
            tmem_ld32(taddr + cS + c * 32, s);
            tmem_ld32(taddr + cP + c * 32, dp);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c * 32 + j;
                const float p = (col >= kb && col < ke) ? exp2f(fmaf(s[j], a.scale_log2e, -st.x)) * st.y : 0.f;
                dp[j] = p * (dp[j] - delta) * a.scale;
            }
            store_bf16_row32(Xs + (c >> 1) * kBox, r, c & 1, dp);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();                         // dS is in the boxes, dP is consumed: its columns take dQ and dK
        if (threadIdx.x == 0) {
            tc_fence_after();
            // dK = dS^T Q: A MN-major ([query rows][keys], two boxes of 64 keys), B MN-major ([query rows][64 features])
            constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64) | kAMn | kBMn;
            const uint64_t as = desc_mn(smem_u32(Xs), kBox), bq = umma_desc_sw128(smem_u32(Qs));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_bf16(tmem_base + cK, as + kRowStep16 * kk, bq + kRowStep16 * kk, idesc_t, kk != 0 ? 1u : 0u);
            // dQ = dS K: A K-major (keys contiguous), B = K MN-major ([key rows][64 features])
            constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64) | kBMn;
            const uint64_t bk = umma_desc_sw128(smem_u32(Ks));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint64_t da = umma_desc_sw128(smem_u32(Xs + (kk >> 2) * kBox)) + 2 * (kk & 3);
                umma_bf16(tmem_base + cQ, da, bk + kRowStep16 * kk, idesc_q, kk != 0 ? 1u : 0u);
            }
            umma_commit(&bars->mma_b);
        }
        __syncwarp();
        mbar_wait(&bars->mma_b, h & 1);           // the products have read dS: the boxes are free for P
        tc_fence_after();
        store_out(cQ, 0, h);
        store_out(cK, 1, h);
        for (int c = 0; c < 4; ++c) {            // P again, from the scores that still sit in [0,128)
            float s[32];
            tmem_ld32(taddr + cS + c * 32, s);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c * 32 + j;
                s[j] = (col >= kb && col < ke) ? exp2f(fmaf(s[j], a.scale_log2e, -st.x)) * st.y : 0.f;
            }
            store_bf16_row32(Xs + (c >> 1) * kBox, r, c & 1, s);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();                         // P is in the boxes, dQ / dK have left their columns
        if (threadIdx.x == 0) {
            tc_fence_after();
            // dV = P^T dO: A MN-major, B MN-major
            constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64) | kAMn | kBMn;
            const uint64_t ap = desc_mn(smem_u32(Xs), kBox), bd = umma_desc_sw128(smem_u32(Ds));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_bf16(tmem_base + cV, ap + kRowStep16 * kk, bd + kRowStep16 * kk, idesc_t, kk != 0 ? 1u : 0u);
            umma_commit(&bars->mma_c);
        }
        __syncwarp();
        mbar_wait(&bars->mma_c, h & 1);
        tc_fence_after();
        store_out(cV, 2, h);
        tc_fence_before();
        __syncthreads();                         // the next head's S / dP overwrite the columns, its dS the boxes
    }
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

constexpr int kFwdSmem = 1024 + 5 * kBox + 1024;
constexpr int kBwdSmem = 1024 + 6 * kBox + 1024;      // 98 KB: two CTAs per SM

template <int TAG, typename K>
int configure_once(K kernel, int bytes) {
    static bool configured[64] = {false};
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        MMF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    return 0;
}

}  // namespace

int launch_tr_attn_tc_fwd(const bf16* qn, long long ldq, const bf16* kn, long long ldk, const bf16* v, long long ldv, int M, int C, int hs,
                          int grid_items, TrAttnTcArgs a, cudaStream_t s) {
    if (grid_items <= 0 || M <= 0) return 0;
    MMF_REQUIRE((hs == 32 || hs == 64) && C % 64 == 0, "attention: head size 32 or 64");
    CUtensorMap tq, tk, tv;
    if (make_tmap_2d(&tq, qn, 2, M, C, ldq, 64, 128) || make_tmap_2d(&tk, kn, 2, M, C, ldk, 64, 128) || make_tmap_2d(&tv, v, 2, M, C, ldv, 64, 128)) return 1;
    a.H = C / hs; a.C = C;
    a.scale = 1.0f / sqrtf(static_cast<float>(hs));
    a.scale_log2e = a.scale * 1.4426950408889634f;
    const dim3 grid(grid_items, C / 64);
    if (hs == 32) {
        if (configure_once<0>(tr_attn_tc_fwd_kernel<32>, kFwdSmem)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_tc_fwd_kernel<32>, grid, dim3(128), kFwdSmem, s, tq, tk, tv, a));
    } else {
        if (configure_once<1>(tr_attn_tc_fwd_kernel<64>, kFwdSmem)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_tc_fwd_kernel<64>, grid, dim3(128), kFwdSmem, s, tq, tk, tv, a));
    }
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_attn_tc_bwd(const bf16* dO, long long lddo, const bf16* qn, long long ldq, const bf16* kn, long long ldk, const bf16* v,
                          long long ldv, int M, int C, int hs, int grid_items, TrAttnTcArgs a, cudaStream_t s) {
    if (grid_items <= 0 || M <= 0) return 0;
    MMF_REQUIRE((hs == 32 || hs == 64) && C % 64 == 0, "attention: head size 32 or 64");
    CUtensorMap tq, tk, tv, td;
    if (make_tmap_2d(&tq, qn, 2, M, C, ldq, 64, 128) || make_tmap_2d(&tk, kn, 2, M, C, ldk, 64, 128) ||
        make_tmap_2d(&tv, v, 2, M, C, ldv, 64, 128) || make_tmap_2d(&td, dO, 2, M, C, lddo, 64, 128)) return 1;
    a.H = C / hs; a.C = C;
    a.scale = 1.0f / sqrtf(static_cast<float>(hs));
    a.scale_log2e = a.scale * 1.4426950408889634f;
    const dim3 grid(grid_items, C / 64);
    if (hs == 32) {
        if (configure_once<2>(tr_attn_tc_bwd_kernel<32>, kBwdSmem)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_tc_bwd_kernel<32>, grid, dim3(128), kBwdSmem, s, tq, tk, tv, td, a));
    } else {
        if (configure_once<3>(tr_attn_tc_bwd_kernel<64>, kBwdSmem)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_tc_bwd_kernel<64>, grid, dim3(128), kBwdSmem, s, tq, tk, tv, td, a));
    }
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace mmf
