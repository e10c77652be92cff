// EPiC sampler as ONE persistent kernel: a CTA owns a 128-row tile of whole jets for all N timesteps.
//
//   local features  loc  [128 x 256] fp32 live in TMEM columns [0,256)   (the residual stream of EPiCLayer)
//   scratch accum        [128 x 256] fp32        TMEM columns [256,512)
//   GEMM A operand       [128 x 256] bf16 in shared memory (4 SWIZZLE_128B chunks), rewritten by each epilogue
//   weights              bf16 16 KB tiles streamed from L2 in consumption order by 1-D bulk copies (4-stage ring)
//   per-jet global path  masked mean/sum pooling + the two tiny global linears on CUDA cores, overlapped with the
//                        fc_loc1 tensor-core GEMM of the same layer (it does not depend on them)
//
// warps 0..7: epilogue / SIMT (thread = row r, column half hf); warp 8 lane 0: weight producer; warp 9 lane 0:
// tcgen05.mma issuer.  Reference: networks/EPiC.py:38-62 (forward), :110-124 (projection), :152-173 (layer),
// :65-72 (pooling); Euler step model/solvers.py:139-143.
#include "mmf_epic.h"
#include "mmf_ptx.cuh"
#include "mmf_tile.cuh"

namespace mmf {

namespace {

constexpr int kEpi = 256;                 // epilogue threads
constexpr int kThreads = 320;
constexpr int kStages = 4;
constexpr int kTile = 128 * 128;          // bytes of one 128 x 64 bf16 tile
constexpr int kPoolLd = 528;              // pooled vector: mean(256) | 0.01 sum(256) | glob(16)

struct EpicBars {
    uint64_t full[kStages], empty[kStages], acc_full, a_ready, xchg;
    uint32_t tmem_base;
};

// float offsets of the fp32 scratch that follows the operand buffers
constexpr int oXs = 0, oPool = oXs + 384, oPart = oPool + kEpicMaxJets * kPoolLd, oHid = oPart + 2 * kEpicMaxJets * 256,
              oJb = oHid + kEpicMaxJets * 256, oGlob = oJb + kEpicMaxJets * 256, oGpre = oGlob + 128, oGskip = oGpre + 128,
              oHeadp = oGskip + 128, oXchg = oHeadp + 512, oMeta = oXchg + 512, oRowJet = oMeta + 32,
              oConst = oRowJet + 32 /* a3 [256][4] | b_loc2p [256] | bl2 [5][256] */, oTb = oConst + 1024 + 256 + kEpicLayers * 256,
              oEnd = oTb + kEpicTbLd;
constexpr int kSmemBytes = 1024 /*align*/ + 1024 /*bars*/ + 4 * kTile + kStages * kTile + oEnd * 4;

__device__ __forceinline__ void epi_bar() { named_bar_sync(1, kEpi); }

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// hidden layer of a global MLP for all jets of the tile:  hid[j][o] = act(sum_k Wt[k][o] * pool[j][k] + bias[o])
// (Wt bf16 [K][256], transposed on the host).  The matrix is streamed from L2 once per call with 16-byte loads, 16 in flight
// per thread (the loop is bound by bytes in flight, not by FMAs): a warp covers 32 outputs (4 lanes x 8) x 8 k-slices, the
// k-slices are reduced with shuffles.  Jets are processed four at a time (tiles rarely hold more).
// NJ jets (1..4) at a time; NJ is a template parameter so that no predicated-off FMAs are issued for missing jets
template <bool GELU, int NJ>
__device__ __forceinline__ void global_hidden_jets(const uint4* __restrict__ wp, int kn, const float* s_pool_k0, float* s_hid,
                                                   const float* bias0, int bias_jet_stride, const int* jet_tb, int j0, int o0, int ks) {
    // accumulators as float2 pairs: the 3-register FFMA issues every other cycle per scheduler on this part, the packed
    // fma.rn.f32x2 (FFMA2) does two of them per instruction - the loop is FMA-pipe-bound (528 x 256 x NJ FMAs per tile)
    float2 acc2[NJ][4];
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[jj][e] = make_float2(0.f, 0.f);
#pragma unroll 16
    for (int k = 0; k < kn; ++k) {
        const uint4 w = __ldg(wp + k * 256);                // row 8 k + ks, outputs o0 .. o0+7
        const float2 we[4] = {make_float2(bf16_lo(w.x), bf16_hi(w.x)), make_float2(bf16_lo(w.y), bf16_hi(w.y)),
                              make_float2(bf16_lo(w.z), bf16_hi(w.z)), make_float2(bf16_lo(w.w), bf16_hi(w.w))};
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const float p = s_pool_k0[(j0 + jj) * kPoolLd + k * 8];   // the 8 slices of a warp read 8 consecutive words
            const float2 pp = make_float2(p, p);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc2[jj][e] = __ffma2_rn(we[e], pp, acc2[jj][e]);
        }
    }
    float acc[NJ][8];
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj)
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[jj][2 * e] = acc2[jj][e].x; acc[jj][2 * e + 1] = acc2[jj][e].y; }
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float v = acc[jj][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[jj][e] = v;
        }
    if (ks == 0) {
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const float* b = bias0 + static_cast<size_t>(bias_jet_stride ? jet_tb[j0 + jj] : 0) * bias_jet_stride + o0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float v = acc[jj][e] + __ldg(b + e);
                s_hid[(j0 + jj) * 256 + o0 + e] = GELU ? gelu_erf(v) : leaky_relu(v);
            }
        }
    }
}

template <bool GELU>
__device__ __forceinline__ void global_hidden(const bf16* __restrict__ Wt, int K, const float* s_pool, float* s_hid,
                                              const float* bias0, int bias_jet_stride, const int* jet_tb, int njets, int tid) {
    const int lane = tid & 31, ks = lane >> 2, o0 = ((tid >> 5) * 4 + (lane & 3)) * 8;
    const int kn = K >> 3, k0 = ks;                            // slice ks takes rows ks, ks + 8, ... (64 or 66 of them)
    const uint4* wp = reinterpret_cast<const uint4*>(Wt + static_cast<size_t>(k0) * 256 + o0);
    for (int j0 = 0; j0 < njets; j0 += 4) {
        const int nj = njets - j0;                             // warp-uniform
        if (nj >= 4) global_hidden_jets<GELU, 4>(wp, kn, s_pool + k0, s_hid, bias0, bias_jet_stride, jet_tb, j0, o0, ks);
        else if (nj == 3) global_hidden_jets<GELU, 3>(wp, kn, s_pool + k0, s_hid, bias0, bias_jet_stride, jet_tb, j0, o0, ks);
        else if (nj == 2) global_hidden_jets<GELU, 2>(wp, kn, s_pool + k0, s_hid, bias0, bias_jet_stride, jet_tb, j0, o0, ks);
        else global_hidden_jets<GELU, 1>(wp, kn, s_pool + k0, s_hid, bias0, bias_jet_stride, jet_tb, j0, o0, ks);
    }
    epi_bar();
}

// W2 sits in shared memory as [16][64] float4 with the float4 index of row q XOR-ed by (2q + (c >> 5)) & 7 (stage_w2): the 8
// lanes of a quarter warp (4 rows x 2 halves) then read 8 different 16-byte bank groups instead of one.
__device__ __forceinline__ int w2_slot(int q, int c) { return q * 64 + (c ^ ((2 * q + (c >> 5)) & 7)); }
__device__ __forceinline__ float global_out16(const float* W2 /* shared memory, swizzled */, const float* __restrict__ b2,
                                              const float* s_hid, int tid) {
    const int j = tid >> 5, q = (tid >> 1) & 15, half = tid & 1;
    const float4* w = reinterpret_cast<const float4*>(W2);
    const float4* h = reinterpret_cast<const float4*>(s_hid + j * 256 + half * 128);
    float2 a01 = f2dup(0.f), a23 = f2dup(0.f);                // four independent chains in two packed accumulators
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        const float4 a = w[w2_slot(q, half * 32 + o)], b = h[o];
        a01 = f2fma(make_float2(a.x, a.y), make_float2(b.x, b.y), a01);
        a23 = f2fma(make_float2(a.z, a.w), make_float2(b.z, b.w), a23);
    }
    float acc = (a01.x + a01.y) + (a23.x + a23.y);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    return acc + __ldg(b2 + q);
}

__global__ void __launch_bounds__(kThreads, 1) epic_tile_kernel(const EpicLaunch a) {
    // The dynamic shared window starts 1024-byte aligned (no static shared memory in this kernel); it is used directly so
    // that the compiler keeps the shared address space (LDS/STS instead of generic loads).  SWIZZLE_128B needs the alignment.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    EpicBars* bars = reinterpret_cast<EpicBars*>(smem);
    uint8_t* Abuf = smem + 1024;
    uint8_t* ring = Abuf + 4 * kTile;
    float* fb = reinterpret_cast<float*>(ring + kStages * kTile);
    float *s_xs = fb + oXs, *s_pool = fb + oPool, *s_part = fb + oPart, *s_hid = fb + oHid, *s_jb = fb + oJb,
          *s_glob = fb + oGlob, *s_gpre = fb + oGpre, *s_gskip = fb + oGskip, *s_headp = fb + oHeadp, *s_xchg = fb + oXchg;
    EpicTileMeta* s_meta = reinterpret_cast<EpicTileMeta*>(fb + oMeta);
    uint8_t* s_rowjet = reinterpret_cast<uint8_t*>(fb + oRowJet);
    float *s_a3 = fb + oConst, *s_b2p = s_a3 + 1024, *s_bl2 = s_b2p + 256, *s_tb = fb + oTb;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = a.tile0 + blockIdx.x;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->a_ready, kEpi);
        mbar_init(&bars->xchg, kEpi);
        fence_mbar_init();
    }
    if (warp == 9) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    if (tid < static_cast<int>(sizeof(EpicTileMeta) / 4))
        reinterpret_cast<int*>(s_meta)[tid] = reinterpret_cast<const int*>(a.meta + tile)[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();                     // the peer's barriers are initialised before any remote arrive
    const uint32_t tmem_base = bars->tmem_base;
    const int nrows = s_meta->nrows, njets = s_meta->njets;

    // producer and MMA warps run converged (uniform loop state); one elected lane issues the asynchronous instructions
    if (warp == 8) {
        // ------------------------------------------------ weight producer -----------------------------------
        uint32_t it = 0;
        for (int step = 0; step < a.nsteps; ++step) {
            for (int t = 0; t < kEpicTilesPerStep; ++t, ++it) {
                const uint32_t s = it % kStages, round = it / kStages;
                if (round > 0) mbar_wait(&bars->empty[s], (round - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(&bars->full[s], kTile);
                    bulk_load_1d(ring + s * kTile, a.p.wstream + static_cast<size_t>(t) * kTile, kTile, &bars->full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ------------------------------------------------ MMA issuer ----------------------------------------
        constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
        uint32_t it = 0, pa = 0;
        const uint32_t a_u32 = smem_u32(Abuf), ring_u32 = smem_u32(ring);
        auto gemm = [&](uint32_t dcol, uint32_t accumulate) {     // D[:, dcol..dcol+256) (+)= A[128x256] W^T
            mbar_wait(&bars->a_ready, pa);
            pa ^= 1;
            for (int kb = 0; kb < 4; ++kb) {
                for (int nh = 0; nh < 2; ++nh, ++it) {
                    const uint32_t s = it % kStages;
                    mbar_wait(&bars->full[s], (it / kStages) & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(a_u32 + kb * kTile);
                    const uint64_t db = umma_desc_sw128(ring_u32 + s * kTile);
                    if (elect_one()) {
                        umma_bf16(tmem_base + dcol + nh * 128, da, db, idesc, (accumulate | kb) != 0 ? 1u : 0u);
                        umma_bf16(tmem_base + dcol + nh * 128, da + 2, db + 2, idesc, 1u);
                        umma_bf16(tmem_base + dcol + nh * 128, da + 4, db + 4, idesc, 1u);
                        umma_bf16(tmem_base + dcol + nh * 128, da + 6, db + 6, idesc, 1u);
                        umma_commit(&bars->empty[s]);
                        if (kb == 3 && nh == 1) umma_commit(&bars->acc_full);
                    }
                }
            }
        };
        for (int step = 0; step < a.nsteps; ++step) {
            gemm(256, 0);                                   // proj.mlp_local.2
            for (int l = 0; l < kEpicLayers; ++l) {
                gemm(256, 0);                               // fc_loc1 (local part)
                gemm(0, 1);                                 // fc_loc2, accumulated onto the residual in TMEM
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ epilogue / SIMT warps ------------------------------
        const int r = (warp & 3) * 32 + lane, hf = warp >> 2;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t peer = cluster_ctarank() ^ 1u;
        uint32_t pf = 0, px = 0;
        if (tid < 128) {
            int j = 0;
            for (int q = 0; q < njets; ++q) j = (tid >= s_meta->jet_begin[q]) ? q : j;
            s_rowjet[tid] = static_cast<uint8_t>(j);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                s_xs[tid * 3 + c] = tid < nrows ? a.xs0[(static_cast<size_t>(tile) * 128 + tid) * 3 + c] : 0.f;
        }
        for (int i = tid; i < 256; i += kEpi) {
            s_a3[i * 4] = a.p.a3[i * 3]; s_a3[i * 4 + 1] = a.p.a3[i * 3 + 1]; s_a3[i * 4 + 2] = a.p.a3[i * 3 + 2]; s_a3[i * 4 + 3] = 0.f;
            s_b2p[i] = a.p.b_loc2p[i];
            for (int l = 0; l < kEpicLayers; ++l) s_bl2[l * 256 + i] = a.p.bl2[l][i];
        }
        epi_bar();
        const int jrow = s_rowjet[r];
        // skip stream (bf16), layout per tile [col / 8][row][8]: a warp reads / writes 512 contiguous bytes per instruction
        uint4* skip8 = reinterpret_cast<uint4*>(a.loc_skip) + static_cast<size_t>(tile) * 32 * 128 + r;   // + (col / 8) * 128

        int mark_i = 0;
        auto mark = [&](int step) {
            if (a.trace && blockIdx.x == 0 && tid == 0 && step < 2 && mark_i < 64) a.trace[step * 64 + mark_i] = clock64();
            ++mark_i;
        };
        // masked sum pooling of the bf16 local features in Abuf, one pass over the rows whatever the number of jets: warp =
        // (64-column chunk, row half), lane = a column pair; a thread walks its 64 rows in order and flushes its partial sums
        // at every jet boundary (rows of a jet are contiguous, the boundary test is warp-uniform) into s_acc[half][jet][col]
        // (the hidden-layer buffers, free at this point); a second pass, thread = column, adds the two halves.  Fixed order,
        // so the result is deterministic.
        // The fc_loc1 GEMM of the layer is released (a_ready) only AFTER the rows have been read: an N = 128 tcgen05.mma with both
        // operands in shared memory takes the whole 128 B/clk of the shared-memory pipe, and pooling under it ran 5x slower.
        // The GEMM is not needed before the global path (~12 k cycles) is through, so nothing is lost by starting it here.
        auto pool = [&](bool with_glob) {
            const int chunk = warp & 3, half = warp >> 2;
            const uint8_t* ch = Abuf + chunk * kTile + (lane & 3) * 4;
            const uint32_t u = lane >> 2;
            float* s_acc = s_hid;                                            // [2][kEpicMaxJets][256] floats = s_hid | s_jb
            const int r0 = half * 64, r1 = nrows < r0 + 64 ? nrows : r0 + 64;
            float* mine = s_acc + half * kEpicMaxJets * 256 + chunk * 64 + 2 * lane;
            for (int j = 0; j < njets; ++j)                                  // jets without a row in this half contribute zero
                if (s_meta->jet_begin[j + 1] <= r0 || s_meta->jet_begin[j] >= r1) { mine[j * 256] = 0.f; mine[j * 256 + 1] = 0.f; }
            if (r0 < r1) {
                int j = s_rowjet[r0], rr = r0;
                float a0 = 0.f, a1 = 0.f;
                while (true) {
                    const int jend = s_meta->jet_begin[j + 1], lim = jend < r1 ? jend : r1;       // rows [rr, lim) belong to jet j
                    for (; rr + 8 <= lim; rr += 8) {                     // eight loads in flight, tree sums
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = *reinterpret_cast<const uint32_t*>(ch + sw128_offset(rr + i, u));
                        float2 t[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) t[i] = make_float2(bf16_lo(w[i]), bf16_hi(w[i]));
                        const float2 s = f2add(f2add(f2add(t[0], t[1]), f2add(t[2], t[3])), f2add(f2add(t[4], t[5]), f2add(t[6], t[7])));   // same tree
                        a0 += s.x;
                        a1 += s.y;
                    }
                    for (; rr < lim; ++rr) {
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(ch + sw128_offset(rr, u));
                        a0 += bf16_lo(w);
                        a1 += bf16_hi(w);
                    }
                    mine[j * 256] = a0; mine[j * 256 + 1] = a1;
                    if (rr >= r1) break;
                    a0 = 0.f; a1 = 0.f;
                    ++j;
                }
            }
            mbar_arrive(&bars->a_ready);                         // fc_loc1 of this layer may start (operand fenced by its writers)
            epi_bar();
            const int col = tid;
            for (int j = 0; j < njets; ++j) {
                float s = s_acc[j * 256 + col] + s_acc[(kEpicMaxJets + j) * 256 + col];
                if (s_meta->pair) {                              // the other half of the jet lives in the peer CTA
                    dsmem_st_f32(dsmem_addr(s_xchg + px * 256 + col, peer), s);
                    mbar_arrive_remote(dsmem_addr(&bars->xchg, peer));
                    mbar_wait_cluster(&bars->xchg, px);
                    s += s_xchg[px * 256 + col];
                    px ^= 1;
                }
                s_pool[j * kPoolLd + col] = s / static_cast<float>(s_meta->jet_ntot[j]);
                s_pool[j * kPoolLd + 256 + col] = s * 0.01f;
            }
            if (with_glob && tid < njets * 16) s_pool[(tid >> 4) * kPoolLd + 512 + (tid & 15)] = s_glob[tid];
            epi_bar();
        };

        // second linear of a global MLP ([16][256] fp32 = 16 KB): fetched into shared memory while pooling / the hidden layer run
        auto stage_w2 = [&](const float* W2) {
            const float4* src = reinterpret_cast<const float4*>(W2);
            float4* dst = reinterpret_cast<float4*>(s_part);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = tid + i * kEpi;
                dst[w2_slot(idx >> 6, idx & 63)] = __ldg(src + idx);
            }
        };

        for (int step = 0; step < a.nsteps; ++step) {
            mark_i = 0;
            mark(step);                                       // 0: step start
            const int tb_row = a.per_jet_time ? s_meta->jet_tb[jrow] : step;
            const float* tbr = a.tbias + static_cast<size_t>(tb_row) * kEpicTbLd;
            if (!a.per_jet_time) {                            // one row for the whole tile: stage it in shared memory
                for (int i = tid; i < kEpicTbLd; i += kEpi) s_tb[i] = tbr[i];
                epi_bar();
                tbr = s_tb;
            }

            // ---- proj.mlp_local.0 with wxe folded in: h1 = GELU(A3 x + c1(t)), K = 3 on CUDA cores
            {
                const float x0 = s_xs[r * 3], x1 = s_xs[r * 3 + 1], x2 = s_xs[r * 3 + 2];
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int chunk = hf * 2 + cc;
                    float v[64];
#pragma unroll
                    for (int i = 0; i < 64; i += 4) {
                        const int col = chunk * 64 + i;
                        const float4 c1 = *reinterpret_cast<const float4*>(tbr + col);
                        const float4 wa = *reinterpret_cast<const float4*>(s_a3 + col * 4), wb = *reinterpret_cast<const float4*>(s_a3 + col * 4 + 4),
                                     wc = *reinterpret_cast<const float4*>(s_a3 + col * 4 + 8), wd = *reinterpret_cast<const float4*>(s_a3 + col * 4 + 12);
                        v[i] = gelu_tile(fmaf(wa.z, x2, fmaf(wa.y, x1, fmaf(wa.x, x0, c1.x))));
                        v[i + 1] = gelu_tile(fmaf(wb.z, x2, fmaf(wb.y, x1, fmaf(wb.x, x0, c1.y))));
                        v[i + 2] = gelu_tile(fmaf(wc.z, x2, fmaf(wc.y, x1, fmaf(wc.x, x0, c1.z))));
                        v[i + 3] = gelu_tile(fmaf(wd.z, x2, fmaf(wd.y, x1, fmaf(wd.x, x0, c1.w))));
                    }
                    stage_row_bf16(Abuf + chunk * kTile, r, v);
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&bars->a_ready);
            }
            mark(step);                                       // 1: embedding written
            // ---- proj.mlp_local.2: loc = GELU(acc + b); keep fp32 in TMEM, skip copy in global, bf16 operand in Abuf
            mbar_wait(&bars->acc_full, pf);
            pf ^= 1;
            tc_fence_after();
            mark(step);                                       // 2: proj GEMM done
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = hf * 128 + c * 32;
                float v[32];
                tmem_ld32(taddr + 256 + col0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = gelu_tile(v[i] + s_b2p[col0 + i]);
                tmem_st32(taddr + col0, v);
                uint8_t* ch = Abuf + (col0 >> 6) * kTile;
                const uint32_t u0 = (col0 & 63) >> 3;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    skip8[((col0 >> 3) + u) * 128] = make_uint4(pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                                                               pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    st_shared_v4(ch + sw128_offset(r, u0 + u), pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                                 pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
            }
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            // (layer 0's fc_loc1, which only needs loc, is released inside pool() below)
            epi_bar();
            mark(step);                                       // 3: proj epilogue done
            // ---- proj.mlp_global: pooled(512) ++ temb -> 256 (GELU) -> 16 (GELU)
            stage_w2(a.p.wg2p);
            pool(false);
            mark(step);                                       // 4: pooled
            global_hidden<true>(a.p.wg0t, 512, s_pool, s_hid, a.tbias + 256 + (a.per_jet_time ? 0 : static_cast<size_t>(step) * kEpicTbLd),
                                a.per_jet_time ? kEpicTbLd : 0, s_meta->jet_tb, njets, tid);
            const int gidx = (tid >> 5) * 16 + ((tid >> 1) & 15);      // (jet, output) owned by this thread pair
            const bool gown = (tid & 1) == 0 && (tid >> 5) < njets;
            {
                const float g = gelu_erf(global_out16(s_part, a.p.bg2p, s_hid, tid));
                if (gown) { s_glob[gidx] = g; s_gskip[gidx] = g; }
            }
            epi_bar();
            mark(step);                                       // 5: proj global MLP done

            float hp0 = 0.f, hp1 = 0.f, hp2 = 0.f;
#pragma unroll 1
            for (int l = 0; l < kEpicLayers; ++l) {
                // the fc_loc1 GEMM of this layer is already running on the tensor core; meanwhile the global path:
                if (l == 0) stage_w2(a.p.wg2[0]);             // (later layers: staged under the previous layer's fc_loc2 GEMM)
                // fc_loc1's global columns for this thread's output: loaded now, used after the global MLP
                const float4* wg = reinterpret_cast<const float4*>(a.p.wl1g[l] + tid * 16);
                const float4 w0 = __ldg(wg), w1 = __ldg(wg + 1), w2 = __ldg(wg + 2), w3 = __ldg(wg + 3);
                if (l > 0) pool(true);
                else {
                    if (tid < njets * 16) s_pool[(tid >> 4) * kPoolLd + 512 + (tid & 15)] = s_glob[tid];
                    epi_bar();
                }
                mark(step);                                   // 6+6l: pooled
                global_hidden<false>(a.p.wg1t[l], 528, s_pool, s_hid, a.p.bg1[l], 0, s_meta->jet_tb, njets, tid);
                mark(step);                                   // 7+6l: global hidden done
                {
                    const float g2 = global_out16(s_part, a.p.bg2[l], s_hid, tid);
                    if (gown) s_gpre[gidx] = s_glob[gidx] + g2;
                }
                epi_bar();
                {   // per-jet bias of fc_loc1: time part (table) + global part
                    for (int j = 0; j < njets; ++j) {
                        const float* g = s_gpre + j * 16;
                        float acc = a.per_jet_time ? __ldg(a.tbias + static_cast<size_t>(s_meta->jet_tb[j]) * kEpicTbLd + 512 + l * 256 + tid)
                                                   : s_tb[512 + l * 256 + tid];
                        acc = fmaf(w0.x, g[0], fmaf(w0.y, g[1], fmaf(w0.z, g[2], fmaf(w0.w, g[3], acc))));
                        acc = fmaf(w1.x, g[4], fmaf(w1.y, g[5], fmaf(w1.z, g[6], fmaf(w1.w, g[7], acc))));
                        acc = fmaf(w2.x, g[8], fmaf(w2.y, g[9], fmaf(w2.z, g[10], fmaf(w2.w, g[11], acc))));
                        acc = fmaf(w3.x, g[12], fmaf(w3.y, g[13], fmaf(w3.z, g[14], fmaf(w3.w, g[15], acc))));
                        s_jb[j * 256 + tid] = acc;
                    }
                }
                epi_bar();
                mark(step);                                   // 8+6l: global path done
                // ---- fc_loc1 epilogue: hidden = leaky_relu(acc + jet bias) -> bf16 operand (overwrites loc in Abuf)
                mbar_wait(&bars->acc_full, pf);
                pf ^= 1;
                tc_fence_after();
                mark(step);                                   // 9+6l: fc_loc1 GEMM done (normally long before)
                {
                    const float* jb = s_jb + jrow * 256;
#pragma unroll 1
                    for (int cc = 0; cc < 2; ++cc) {
                        const int chunk = hf * 2 + cc;
                        float v[64];
                        tmem_ld32(taddr + 256 + chunk * 64, v);
                        tmem_ld32(taddr + 256 + chunk * 64 + 32, v + 32);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 64; i += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(jb + chunk * 64 + i);
                            MMF_SET2(v, i, leaky_relu2(f2add(MMF_V2(v, i), make_float2(b.x, b.y))));
                            MMF_SET2(v, i + 2, leaky_relu2(f2add(MMF_V2(v, i + 2), make_float2(b.z, b.w))));
                        }
                        stage_row_bf16(Abuf + chunk * kTile, r, v);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&bars->a_ready);                 // fc_loc2 accumulates onto loc in TMEM
                mark(step);                                   // 10+6l: fc_loc1 epilogue done
                // ---- fc_loc2 epilogue: loc = leaky_relu(loc_pre + b) + loc_skip.  The skip values of this thread's 128 columns
                //      are fetched while the GEMM runs.
                uint4 sk[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) sk[u] = skip8[(hf * 16 + u) * 128];
                if (l + 1 < kEpicLayers) stage_w2(a.p.wg2[l + 1]);   // this layer's global output (the last reader of s_part) is done
                mbar_wait(&bars->acc_full, pf);
                pf ^= 1;
                tc_fence_after();
                mark(step);                                   // 11+6l: fc_loc2 GEMM done
                const bool last = l + 1 == kEpicLayers;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col0 = hf * 128 + c * 32;
                    float v[32];
                    tmem_ld32(taddr + col0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint4 s8 = sk[c * 4 + u];
                        const float4 b0 = *reinterpret_cast<const float4*>(s_bl2 + l * 256 + col0 + 8 * u);
                        const float4 b1 = *reinterpret_cast<const float4*>(s_bl2 + l * 256 + col0 + 8 * u + 4);
                        MMF_SET2(v, 8 * u, f2add(leaky_relu2(f2add(MMF_V2(v, 8 * u), make_float2(b0.x, b0.y))), make_float2(bf16_lo(s8.x), bf16_hi(s8.x))));
                        MMF_SET2(v, 8 * u + 2, f2add(leaky_relu2(f2add(MMF_V2(v, 8 * u + 2), make_float2(b0.z, b0.w))), make_float2(bf16_lo(s8.y), bf16_hi(s8.y))));
                        MMF_SET2(v, 8 * u + 4, f2add(leaky_relu2(f2add(MMF_V2(v, 8 * u + 4), make_float2(b1.x, b1.y))), make_float2(bf16_lo(s8.z), bf16_hi(s8.z))));
                        MMF_SET2(v, 8 * u + 6, f2add(leaky_relu2(f2add(MMF_V2(v, 8 * u + 6), make_float2(b1.z, b1.w))), make_float2(bf16_lo(s8.w), bf16_hi(s8.w))));
                    }
                    if (!last) {
                        tmem_st32(taddr + col0, v);
                        uint8_t* ch = Abuf + (col0 >> 6) * kTile;
                        const uint32_t u0 = (col0 & 63) >> 3;
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            st_shared_v4(ch + sw128_offset(r, u0 + u), pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                                         pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
                    } else {                                  // head: Linear(528,3), local part
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            hp0 = fmaf(__ldg(a.p.wh_loc + col0 + i), v[i], hp0);
                            hp1 = fmaf(__ldg(a.p.wh_loc + 256 + col0 + i), v[i], hp1);
                            hp2 = fmaf(__ldg(a.p.wh_loc + 512 + col0 + i), v[i], hp2);
                        }
                    }
                }
                if (tid < njets * 16) s_glob[tid] = leaky_relu(s_gpre[tid]) + s_gskip[tid];
                if (!last) {
                    tmem_st_wait();
                    fence_proxy_async();
                    tc_fence_before();
                    // (the next layer's fc_loc1 is released inside its pool())
                }
                epi_bar();
            }
            mark(step);                                       // 36: layers done
            // ---- head and Euler update
            if (hf == 1) { s_headp[r * 4] = hp0; s_headp[r * 4 + 1] = hp1; s_headp[r * 4 + 2] = hp2; }
            epi_bar();
            if (hf == 0 && r < nrows) {
                const float* g = s_glob + jrow * 16;
                float vt[3] = {hp0 + s_headp[r * 4], hp1 + s_headp[r * 4 + 1], hp2 + s_headp[r * 4 + 2]};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float acc = vt[d] + tbr[1792 + d];
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc = fmaf(__ldg(a.p.wh_glob + d * 16 + q), g[q], acc);
                    vt[d] = acc;
                }
                if (a.vt_out) {
                    const long long slot = a.row_slot[static_cast<size_t>(tile) * 128 + r];
#pragma unroll
                    for (int d = 0; d < 3; ++d) a.vt_out[slot * 3 + d] = vt[d];
                } else {
#pragma unroll
                    for (int d = 0; d < 3; ++d) s_xs[r * 3 + d] = euler_update(s_xs[r * 3 + d], vt[d], a.dt);
                }
            }
            epi_bar();
        }
        if (a.x_out && tid < nrows) {
            const long long slot = a.row_slot[static_cast<size_t>(tile) * 128 + tid];
#pragma unroll
            for (int d = 0; d < 3; ++d) a.x_out[slot * 3 + d] = s_xs[tid * 3 + d];
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // a pair CTA must not exit while its peer can still reach its smem
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// tbias[tb][m*256 + o] = cst[m][o] + sum_k wt[m][k][o] * temb[tb][k]   (7 folded linears) ; head: 3 outputs
__global__ void __launch_bounds__(256) epic_time_bias_kernel(const EpicTimeFold f, const float* __restrict__ temb,
                                                             float* __restrict__ tbias) {
    __shared__ float s_t[256];
    const int tb = blockIdx.x, o = threadIdx.x;
    s_t[o] = temb[static_cast<size_t>(tb) * 256 + o];
    __syncthreads();
    float* out = tbias + static_cast<size_t>(tb) * kEpicTbLd;
    for (int m = 0; m < 7; ++m) {
        const float* w = f.wt + static_cast<size_t>(m) * 65536 + o;
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < 256; ++k) acc = fmaf(__ldg(w + k * 256), s_t[k], acc);
        out[m * 256 + o] = acc + f.cst[m * 256 + o];
    }
    if (o < 3) {
        float acc = 0.f;
        for (int k = 0; k < 256; ++k) acc = fmaf(f.wht[k * 3 + o], s_t[k], acc);
        out[1792 + o] = acc + f.bh[o];
    }
}

}  // namespace

int epic_smem_bytes() { return kSmemBytes; }

int launch_epic_tiles(const EpicLaunch& a, int n_tiles, int cluster, cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        MMF_CUDA_OK(cudaFuncSetAttribute(epic_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MMF_REQUIRE(cluster == 1 || (cluster == 2 && n_tiles % 2 == 0), "epic: pair launches need an even tile count");
    MMF_CUDA_OK(cudaLaunchKernelEx(&cfg, epic_tile_kernel, a));
    return 0;
}

int launch_epic_time_bias(const EpicTimeFold& f, const float* temb, int n, float* tbias, cudaStream_t stream) {
    if (n == 0) return 0;
    epic_time_bias_kernel<<<n, 256, 0, stream>>>(f, temb, tbias);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace mmf
