// Tensor-core kernels for sm_100a: tcgen05.mma with TMEM accumulators, operands staged by TMA.
//
//   gemm_kernel<BN, EPI>   D[128 x BN] = A[128 x K] * W[BN x K]^T  (bf16 in, fp32 accumulate) with the
//                          layer's elementwise tail fused into the TMEM->register epilogue:
//        EPI_STORE_BF16    bias (+ exact GELU) -> bf16                       (c_fc, head hidden)
//        EPI_STORE_F32     bias -> fp32                                      (wxe.2 pre-LayerNorm output)
//        EPI_QKV           bias, per-head LayerNorm on q and k -> Q, K ; V stored transposed (V^T)
//        EPI_RESLN         residual += acc + bias + time-embedding (in place, through smem) and
//                          LayerNorm of the new residual -> bf16 operand of the next GEMM
//   attn_kernel            masked softmax(Q K^T / sqrt(hs)) V for a run of whole jets (block-diagonal mask)
//
// reference semantics: networks/attention.py:23-26 (block), :53-74 (attention), utils/models.py:20-37.
// Layout rules used everywhere: operands are K-major, 128-byte rows, SWIZZLE_128B; one CTA owns one
// 128-row tile (= the 128 TMEM lanes); thread i of the four epilogue warps owns row i.
#include "mmf_internal.h"
#include "mmf_ptx.cuh"
#include "mmf_tile.cuh"

namespace mmf {

namespace {

constexpr int kABytes = kTileM * 128;          // one A stage: 128 rows x 128 B
constexpr int kChunkBytes = kTileM * 128;      // one staging chunk: 128 rows x 128 B
constexpr int kBarBytes = 1024;

template <int N>
__device__ __forceinline__ void layernorm_inplace(float* v, const float* g, const float* b) {
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

// ---------------------------------------------------------------------------------------------
// GEMM
// ---------------------------------------------------------------------------------------------
struct GemmBars {
    uint64_t full[4], empty[4], acc_full, rfull[3], rdone[3], ofull[2], oempty[2];
    uint32_t tmem_base;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(192, 2)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmOut0, const __grid_constant__ CUtensorMap tmOut1,
            const GemmArgs a, const int stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem);
    uint8_t* tiles = smem + kBarBytes;
    constexpr int kBBytes = BN * 128;
    constexpr int kStage = kABytes + kBBytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTileM, n0 = blockIdx.y * BN, g = blockIdx.z;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut0);
        tma_prefetch_desc(&tmOut1);
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int i = 0; i < 4; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
            mbar_init(&bars->acc_full, 1);
            for (int i = 0; i < 3; ++i) { mbar_init(&bars->rfull[i], 1); mbar_init(&bars->rdone[i], 128); }
            for (int i = 0; i < 2; ++i) { mbar_init(&bars->ofull[i], 128); mbar_init(&bars->oempty[i], 1); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&bars->tmem_base, BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    grid_dep_wait();            // PDL: inputs written by the previous kernel are visible from here on

    // epilogue staging aliases the pipeline stages (all MMAs have retired when acc_full fires)
    uint8_t* rbuf = tiles;                       // 3 x 16 KB fp32 residual chunks (RESLN)
    uint8_t* obuf = (EPI == EPI_RESLN) ? tiles + 3 * kChunkBytes : tiles;   // 2 x 16 KB output chunks

    if (warp == 4) {
        if (lane == 0) {
            // ------------------------------ TMA producer: mainloop -------------------------------
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int s = kb % stages, it = kb / stages;
                if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
                mbar_expect_tx(&bars->full[s], kStage);
                tma_load_2d(tiles + s * kStage, &tmA, &bars->full[s], g * a.a_col_group_stride + kb * kBK, m0);
                tma_load_2d(tiles + s * kStage + kABytes, &tmB, &bars->full[s], kb * kBK, g * a.w_rows_per_group + n0);
            }
            // ------------------------------ TMA side of the epilogue -----------------------------
            mbar_wait(&bars->acc_full, 0);
            if constexpr (EPI == EPI_RESLN) {
                const int C = BN, nch = BN / 32;
                for (int c = 0; c < 3 && c < nch; ++c) {
                    mbar_expect_tx(&bars->rfull[c], kChunkBytes);
                    tma_load_2d(rbuf + c * kChunkBytes, &tmOut1, &bars->rfull[c], g * C + c * 32, m0);
                }
                for (int c = 0; c < nch; ++c) {
                    const int b = c % 3;
                    mbar_wait(&bars->rdone[b], (c / 3) & 1);
                    tma_store_2d(&tmOut1, rbuf + b * kChunkBytes, g * C + c * 32, m0);
                    tma_store_commit();
                    if (c + 3 < nch) {
                        tma_store_wait_read<0>();
                        mbar_expect_tx(&bars->rfull[b], kChunkBytes);
                        tma_load_2d(rbuf + b * kChunkBytes, &tmOut1, &bars->rfull[b], g * C + (c + 3) * 32, m0);
                    }
                }
                if (a.ln_g != nullptr) {
                    const int nout = BN / 64;
                    for (int cc = 0; cc < nout; ++cc) {
                        const int ob = cc & 1;
                        mbar_wait(&bars->ofull[ob], (cc >> 1) & 1);
                        tma_store_2d(&tmOut0, obuf + ob * kChunkBytes, g * C + cc * 64, m0);
                        tma_store_commit();
                        if (cc + 2 < nout) { tma_store_wait_read<0>(); mbar_arrive(&bars->oempty[ob]); }
                    }
                }
            } else if constexpr (EPI == EPI_QKV) {
                const int C = a.sect_width, sect = n0 / C, cbase = g * C + (n0 % C);
                if (sect < 2) {
                    for (int cc = 0; cc < BN / 64; ++cc) {
                        mbar_wait(&bars->ofull[cc], 0);
                        tma_store_2d(sect == 0 ? &tmOut0 : &tmOut1, obuf + cc * kChunkBytes, cbase + cc * 64, m0);
                        tma_store_commit();
                    }
                }
            } else {
                const int width = (EPI == EPI_STORE_BF16) ? 64 : 32;
                const int nout = BN / width;
                const int col0 = g * a.out_col_group_stride + n0;
                for (int cc = 0; cc < nout; ++cc) {
                    const int ob = cc & 1;
                    mbar_wait(&bars->ofull[ob], (cc >> 1) & 1);
                    tma_store_2d(&tmOut0, obuf + ob * kChunkBytes, col0 + cc * width, m0);
                    tma_store_commit();
                    if (cc + 2 < nout) { tma_store_wait_read<0>(); mbar_arrive(&bars->oempty[ob]); }
                }
            }
            tma_store_wait_all();
        }
        __syncwarp();
    } else if (warp == 5) {
        if (lane == 0) {
            // ------------------------------ MMA issuer -------------------------------------------
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int s = kb % stages, it = kb / stages;
                mbar_wait(&bars->full[s], it & 1);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_u32(tiles + s * kStage));
                const uint64_t db = umma_desc_sw128(smem_u32(tiles + s * kStage + kABytes));
#pragma unroll
                for (int ks = 0; ks < kBK / 16; ++ks)
                    umma_bf16(tmem_base, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->acc_full);
        }
        __syncwarp();
    } else {
        // ---------------------------------- epilogue warps 0..3 ----------------------------------
        const int r = warp * 32 + lane;                 // row inside the tile == TMEM lane
        const int row = m0 + r;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();

        if constexpr (EPI == EPI_RESLN) {
            const int C = BN;
            const float* bias = a.bias ? a.bias + g * C : nullptr;
            const float* tb = nullptr;
            if (a.temb) tb = a.temb + static_cast<size_t>(a.row_jet ? a.row_jet[row] : 0) * a.temb_ld + g * C;
            const bool ln = a.ln_g != nullptr;
            float sum = 0.f;
            for (int c = 0; c < BN / 32; ++c) {
                const int b = c % 3;
                float acc[32];
                tmem_ld32(taddr + c * 32, acc);
                mbar_wait(&bars->rfull[b], (c / 3) & 1);
                tmem_ld_wait();
                uint8_t* rb = rbuf + b * kChunkBytes;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 rv = ld_shared_f4(rb + sw128_offset(r, u));
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), tv = bv;
                    if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + c * 32) + u);
                    if (tb) tv = __ldg(reinterpret_cast<const float4*>(tb + c * 32) + u);
                    acc[u * 4 + 0] += rv.x + bv.x + tv.x;
                    acc[u * 4 + 1] += rv.y + bv.y + tv.y;
                    acc[u * 4 + 2] += rv.z + bv.z + tv.z;
                    acc[u * 4 + 3] += rv.w + bv.w + tv.w;
                    sum += (acc[u * 4 + 0] + acc[u * 4 + 1]) + (acc[u * 4 + 2] + acc[u * 4 + 3]);
                }
                stage_row_f32(rb, r, acc);
                if (ln) tmem_st32(taddr + c * 32, acc);
                fence_proxy_async();
                mbar_arrive(&bars->rdone[b]);
            }
            if (ln) {
                tmem_st_wait();
                const float mean = sum * (1.0f / C);
                float ss = 0.f;
                for (int c = 0; c < BN / 32; ++c) {
                    float v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; ss = fmaf(d, d, ss); }
                }
                const float rstd = rsqrtf(ss * (1.0f / C) + 1e-5f);
                const float* lg = a.ln_g + g * C;
                const float* lb = a.ln_b ? a.ln_b + g * C : nullptr;
                for (int cc = 0; cc < BN / 64; ++cc) {
                    const int ob = cc & 1;
                    float v[64];
                    tmem_ld32(taddr + cc * 64, v);
                    tmem_ld32(taddr + cc * 64 + 32, v + 32);
                    if (cc >= 2) mbar_wait(&bars->oempty[ob], ((cc >> 1) - 1) & 1);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        v[i] = fmaf((v[i] - mean) * rstd, __ldg(lg + cc * 64 + i), lb ? __ldg(lb + cc * 64 + i) : 0.f);
                    stage_row_bf16(obuf + ob * kChunkBytes, r, v);
                    fence_proxy_async();
                    mbar_arrive(&bars->ofull[ob]);
                }
            }
        } else if constexpr (EPI == EPI_QKV) {
            const int C = a.sect_width, sect = n0 / C, cbase = g * C + (n0 % C);
            const float* bias = a.bias ? a.bias + g * a.w_rows_per_group + n0 : nullptr;
            if (sect < 2) {
                const float* hg = (sect == 0 ? a.q_g : a.k_g);
                const float* hb = (sect == 0 ? a.q_b : a.k_b);
                if (hg) hg += g * a.hs;
                if (hb) hb += g * a.hs;
                for (int cc = 0; cc < BN / 64; ++cc) {
                    float v[64];
                    tmem_ld32(taddr + cc * 64, v);
                    tmem_ld32(taddr + cc * 64 + 32, v + 32);
                    tmem_ld_wait();
                    if (bias) {
#pragma unroll
                        for (int i = 0; i < 64; ++i) v[i] += __ldg(bias + cc * 64 + i);
                    }
                    if (hg) {
                        if (a.hs == 32) {
                            layernorm_inplace<32>(v, hg, hb);
                            layernorm_inplace<32>(v + 32, hg, hb);
                        } else {
                            layernorm_inplace<64>(v, hg, hb);
                        }
                    }
                    stage_row_bf16(obuf + cc * kChunkBytes, r, v);
                    fence_proxy_async();
                    mbar_arrive(&bars->ofull[cc]);
                }
            } else {
                for (int c = 0; c < BN / 32; ++c) {
                    float v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
                    bf16* dst = a.vt + static_cast<size_t>(cbase + c * 32) * a.vt_ld + row;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        dst[static_cast<size_t>(i) * a.vt_ld] = __float2bfloat16_rn(v[i] + (bias ? __ldg(bias + c * 32 + i) : 0.f));
                }
            }
        } else {
            const float* bias = a.bias ? a.bias + g * a.w_rows_per_group + n0 : nullptr;
            constexpr int width = (EPI == EPI_STORE_BF16) ? 64 : 32;
            for (int cc = 0; cc < BN / width; ++cc) {
                const int ob = cc & 1;
                float v[width];
                tmem_ld32(taddr + cc * width, v);
                if constexpr (width == 64) tmem_ld32(taddr + cc * width + 32, v + 32);
                if (cc >= 2) mbar_wait(&bars->oempty[ob], ((cc >> 1) - 1) & 1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < width; ++i) {
                    float x = v[i] + (bias ? __ldg(bias + cc * width + i) : 0.f);
                    v[i] = a.act == 1 ? gelu_erf(x) : x;
                }
                if constexpr (EPI == EPI_STORE_BF16) stage_row_bf16(obuf + ob * kChunkBytes, r, v);
                else stage_row_f32(obuf + ob * kChunkBytes, r, v);
                fence_proxy_async();
                mbar_arrive(&bars->ofull[ob]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, BN);
}

template <int BN, int EPI>
int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut0, const CUtensorMap& tmOut1,
                  const GemmArgs& args, int m_tiles, int n_tiles, int groups, cudaStream_t stream) {
    const int stages = args.kblocks < (BN == 128 ? 3 : 2) ? args.kblocks : (BN == 128 ? 3 : 2);
    const int smem = gemm_smem_bytes(EPI, BN, args.kblocks);
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        MMF_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    gemm_kernel<BN, EPI><<<dim3(m_tiles, n_tiles, groups), 192, smem, stream>>>(tmA, tmB, tmOut0, tmOut1, args, stages);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------------------------
struct AttnBars {
    uint64_t loaded, s_done, o_done;
    uint32_t tmem_base;
};
constexpr int kAttnQ = kTileM * 128;                 // 16 KB
constexpr int kAttnK = kMaxKeys * 128;               // 20 KB
constexpr int kAttnVTChunk = 64 * 128;               // 8 KB : 64 feature rows x 64 keys
constexpr int kAttnPChunk = kTileM * 128;            // 16 KB: 128 query rows x 64 keys
constexpr int kAttnKeyChunks = (kMaxKeys + 63) / 64; // 3
constexpr int kAttnSmem = kBarBytes + kAttnQ + kAttnK + kAttnKeyChunks * (kAttnVTChunk + kAttnPChunk) + 1024;
constexpr int kAttnTmemCols = 256;                   // S: [0,160)  O: [160,224)

__global__ void __launch_bounds__(128, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmVT, const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    AttnBars* bars = reinterpret_cast<AttnBars*>(smem);
    uint8_t* Qs = smem + kBarBytes;
    uint8_t* Ks = Qs + kAttnQ;
    uint8_t* VTs = Ks + kAttnK;
    uint8_t* Ps = VTs + kAttnKeyChunks * kAttnVTChunk;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = blockIdx.y * 64;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmVT);
        mbar_init(&bars->loaded, 1);
        mbar_init(&bars->s_done, 1);
        mbar_init(&bars->o_done, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&bars->tmem_base, kAttnTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    grid_dep_wait();

    const AttnItem item = a.items[blockIdx.x];
    const int nk16 = (item.nk + 15) & ~15;
    const int kchunks = (nk16 + 63) >> 6;
    const int kboxes = (nk16 + 31) >> 5;
    const int hs = a.hs, nheads = 64 / hs, ksteps = hs / 16;

    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars->loaded, kAttnQ + kboxes * 32 * 128 + kchunks * kAttnVTChunk);
        tma_load_2d(Qs, &tmQ, &bars->loaded, col0, item.q_row0);
        for (int i = 0; i < kboxes; ++i) tma_load_2d(Ks + i * 32 * 128, &tmK, &bars->loaded, col0, item.k_row0 + i * 32);
        for (int c = 0; c < kchunks; ++c) tma_load_2d(VTs + c * kAttnVTChunk, &tmVT, &bars->loaded, item.k_row0 + c * 64, col0);
    }

    const int r = threadIdx.x;
    const int row = item.q_row0 + r;
    const bool valid = r < item.nq;
    int kb = 0, ke = 0;                       // this row attends key columns [kb, ke) of the item
    if (valid) {
        kb = a.seg_beg[row] - item.k_row0;
        ke = a.seg_end[row] - item.k_row0;
        kb = kb < 0 ? 0 : kb;
        ke = ke > item.nk ? item.nk : ke;
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

    for (int h = 0; h < nheads; ++h) {
        if (threadIdx.x == 0) {
            if (h == 0) mbar_wait(&bars->loaded, 0);
            tc_fence_after();
            const uint64_t dq = umma_desc_sw128(smem_u32(Qs)) + 2 * (h * ksteps);
            const uint64_t dk = umma_desc_sw128(smem_u32(Ks)) + 2 * (h * ksteps);
            const uint32_t idesc = umma_idesc_bf16(kTileM, nk16);
            for (int ks = 0; ks < ksteps; ++ks) umma_bf16(tmem_base, dq + 2 * ks, dk + 2 * ks, idesc, ks != 0 ? 1u : 0u);
            umma_commit(&bars->s_done);
        }
        __syncwarp();
        mbar_wait(&bars->s_done, h & 1);
        tc_fence_after();

        // ---- masked softmax, one query row per thread; scores stay in TMEM between the two passes
        float mx = -INFINITY;
        for (int c = 0; c * 32 < nk16; ++c) {
            float s[32];
            tmem_ld32(taddr + c * 32, s);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c * 32 + j;
                if (col >= kb && col < ke) mx = fmaxf(mx, s[j]);
            }
        }
        const float mscaled = (mx == -INFINITY) ? 0.f : mx * a.scale_log2e;
        float sum = 0.f;
        for (int c = 0; c * 32 < nk16; ++c) {
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
            uint8_t* pc = Ps + (c >> 1) * kAttnPChunk;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                st_shared_v4(pc + sw128_offset(r, (c & 1) * 4 + u), pack_bf16x2(s[u * 8 + 0], s[u * 8 + 1]),
                             pack_bf16x2(s[u * 8 + 2], s[u * 8 + 3]), pack_bf16x2(s[u * 8 + 4], s[u * 8 + 5]),
                             pack_bf16x2(s[u * 8 + 6], s[u * 8 + 7]));
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();

        if (threadIdx.x == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(kTileM, hs);
            for (int kk = 0; kk * 16 < nk16; ++kk) {
                const int c = kk >> 2;
                const uint64_t dp = umma_desc_sw128(smem_u32(Ps + c * kAttnPChunk)) + 2 * (kk & 3);
                const uint64_t dv = umma_desc_sw128(smem_u32(VTs + c * kAttnVTChunk + h * hs * 128)) + 2 * (kk & 3);
                umma_bf16(tmem_base + kMaxKeys + h * hs, dp, dv, idesc, kk != 0 ? 1u : 0u);
            }
            umma_commit(&bars->o_done);
        }
        __syncwarp();
        mbar_wait(&bars->o_done, h & 1);
        tc_fence_after();

        const float inv = 1.0f / (sum > 0.f ? sum : 1.f);
        for (int c = 0; c * 32 < hs; ++c) {
            float o[32];
            tmem_ld32(taddr + kMaxKeys + h * hs + c * 32, o);
            tmem_ld_wait();
            if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(a.out + static_cast<size_t>(row) * a.ld_out + col0 + h * hs + c * 32);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    dst[u] = make_uint4(pack_bf16x2(o[u * 8 + 0] * inv, o[u * 8 + 1] * inv),
                                        pack_bf16x2(o[u * 8 + 2] * inv, o[u * 8 + 3] * inv),
                                        pack_bf16x2(o[u * 8 + 4] * inv, o[u * 8 + 5] * inv),
                                        pack_bf16x2(o[u * 8 + 6] * inv, o[u * 8 + 7] * inv));
            }
        }
        // the next head's S MMA overwrites TMEM columns [0,160): every thread is past its last S read
        // (the __syncthreads above) and P is free again because o_done has fired.
        tc_fence_before();
        __syncthreads();
    }

    if (warp == 0) tmem_dealloc(tmem_base, kAttnTmemCols);
}

}  // namespace

int gemm_smem_bytes(int epilogue, int BN, int kblocks) {
    const int max_stages = BN == 128 ? 3 : 2;
    const int stages = kblocks < max_stages ? kblocks : max_stages;
    int pipe = stages * (kABytes + BN * 128);
    int epi = (epilogue == EPI_RESLN) ? 5 * kChunkBytes : 2 * kChunkBytes;
    return kBarBytes + (pipe > epi ? pipe : epi) + 1024;
}

int launch_gemm(int epilogue, int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut0,
                const CUtensorMap& tmOut1, const GemmArgs& args, int m_tiles, int n_tiles, int groups,
                cudaStream_t stream) {
    MMF_REQUIRE(args.kblocks >= 1 && args.kblocks <= 16, "gemm: K must be a multiple of 64 up to 1024");
    if (BN == 128) {
        switch (epilogue) {
            case EPI_STORE_BF16: return launch_gemm_t<128, EPI_STORE_BF16>(tmA, tmB, tmOut0, tmOut1, args, m_tiles, n_tiles, groups, stream);
            case EPI_STORE_F32: return launch_gemm_t<128, EPI_STORE_F32>(tmA, tmB, tmOut0, tmOut1, args, m_tiles, n_tiles, groups, stream);
            case EPI_QKV: return launch_gemm_t<128, EPI_QKV>(tmA, tmB, tmOut0, tmOut1, args, m_tiles, n_tiles, groups, stream);
            case EPI_RESLN: return launch_gemm_t<128, EPI_RESLN>(tmA, tmB, tmOut0, tmOut1, args, m_tiles, n_tiles, groups, stream);
        }
    } else if (BN == 256 && epilogue == EPI_RESLN) {
        return launch_gemm_t<256, EPI_RESLN>(tmA, tmB, tmOut0, tmOut1, args, m_tiles, n_tiles, groups, stream);
    }
    set_last_error("gemm: unsupported (epilogue, BN) combination");
    return 2;
}

int launch_attention(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmVT, const AttnArgs& args,
                     int n_items, int n_slabs, cudaStream_t stream) {
    MMF_REQUIRE(args.hs == 32 || args.hs == 64, "attention: head size must be 32 or 64");
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        MMF_CUDA_OK(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (n_items == 0) return 0;
    attn_kernel<<<dim3(n_items, n_slabs), 128, kAttnSmem, stream>>>(tmQ, tmK, tmVT, args);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// tensor maps (driver entry point fetched through the runtime; no -lcuda needed)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_cols, uint32_t box_rows) {
    static PFN_tmapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MMF_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MMF_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
    }
    MMF_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "tensor map: bf16 or fp32 only");
    MMF_REQUIRE(box_cols * elem_bytes == 128, "tensor map: inner box must be one 128-byte swizzle row");
    MMF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * elem_bytes) % 16 == 0,
                "tensor map: base and row pitch must be 16-byte aligned");
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {ld_elems * static_cast<uint64_t>(elem_bytes)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = encode(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                               const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(rc)));
        return 1;
    }
    return 0;
}

}  // namespace mmf
