// Persistent per-tile ParticleFormer / FusedParticleFormer sampler (see mmf_tftile.h for the data layout).
//
// warps 0..7   epilogue / SIMT: thread = (row r, column half hf); runs the hand-written per-timestep program
// warp  8      parameter producer: streams the per-stage parameter blobs into the double buffer
// warp  9      tcgen05.mma issuer of the weight GEMMs (every op that reads the ring) - walks the op table, one elected lane issues
// warp  10     weight producer: streams the weight tiles into the ring with 1-D bulk copies
// warp  11     tcgen05.mma issuer of the attention products S = Q K^T and O = P V (both operands in the arena): they are
//              short and latency-critical, and no longer queue behind the streaming GEMM of the next unit
//
// Synchronisation: `go` (256 arrivals) epilogue -> MMA issuer, consumed in order by the ops flagged `wait`;
// `done[0/1]` (tcgen05.commit) MMA -> epilogue; full/empty ring barriers producer <-> MMA issuer;
// pfull/pempty parameter double buffer producer <-> epilogue.
#include <cstdio>
#include <cstring>
#ifndef MMF_TILE_TRACE
#define MMF_TILE_TRACE 0
#endif
// -DMMF_PROD_DIAG=1 (MMF_EXTRA_NVCC of build.py): the PRODUCTION kernel with the time-out records of mbar_wait; the process
// prints them when it exits (debugging aid for a launch that trapped: which warp waited for which barrier).
#ifndef MMF_PROD_DIAG
#define MMF_PROD_DIAG 0
#endif
#define MMF_WAIT_DIAG (MMF_TILE_TRACE || MMF_PROD_DIAG)
#include "mmf_ptx.cuh"
#include "mmf_tftile.h"
#include "mmf_tile.cuh"

namespace mmf {

namespace {

constexpr int kEpi = 256;
constexpr int kThreads = 384;
constexpr int kProducers = 1;              // warp 10 (one warp keeps up: the ring is latency-bound, ~500 cycles per copy suffice)
constexpr int kTile = 16384;                // k-tile of a B operand in the arena: 64 keys x 128 bytes (V rows of a P V k-tile)
constexpr int kBars = kTfRingBars;          // weight-ring barrier pairs, used round robin by tile index
#ifndef MMF_COPY_SPLIT
#define MMF_COPY_SPLIT 1
#endif
constexpr int kCopySplit = MMF_COPY_SPLIT;   // bulk copies per weight-tile slice
// operand arena: TfLay<PAIR> in mmf_tftile.h (plain tiles / pair tiles)

// Hand-offs epilogue -> weight-GEMM issuer rotate over kGoBars barriers: a waiter tells phases apart by parity only, and
// with two issuer warps the GEMM issuer can be a hand-off behind (an unsignalled projection waiting for its weight tile
// while the attention issuer keeps the epilogue going), so one barrier could complete two phases unobserved.
constexpr int kGoBars = 4;
struct TfBars {
    uint64_t full[kBars], empty[kBars], done[4], go[kGoBars], go_attn[kGoBars], pfull[2], pempty[2];
    // pair tiles: kfull / vfull - the partner's K / V rows of a unit have landed in this CTA (one expect_tx arrival + the
    // bytes of one bulk copy each, issued by the partner's attention issuer); kfree / vfree - the partner's MMAs have read ITS
    // K / V buffers for the last time in the unit (tcgen05.commit multicast to this CTA): the next unit's rows may go there,
    // and - since the partner's products ran - this CTA's own outgoing copy of the unit has long left its local rows
    uint64_t kfull, vfull, kfree, vfree;
    uint32_t tmem_base;
};

// fp32 scratch after the two parameter buffers (float offsets)
// (mOut, the per-row head partial sums, is only live at the very end of a timestep and aliases the LayerNorm / softmax exchange)
// (mRed / mSum have two slots: the two heads of a 64-column unit are in flight together)
constexpr int mXs = 0, mKs = mXs + 384, mStat = mKs + 128, mRed = mStat + 1024, mSum = mRed + 512, mOut = mStat,
              mTemb = mSum + 512, mEnd = mTemb + 512;
static_assert(128 * 12 <= 1024 + 512 + 512, "head partial sums alias the exchange buffers");
template <bool PAIR> constexpr int smem_bytes() { return 1024 + 1024 + static_cast<int>(TfLay<PAIR>::kArena) + 2 * kTfParamFloats * 4 + mEnd * 4; }
static_assert(smem_bytes<false>() <= 232448 && smem_bytes<true>() <= 232448, "shared memory budget");

constexpr uint32_t kScr = 256;                   // first scratch column in TMEM

struct Epi {
    uint8_t* arena;
    float* pbuf;
    float* misc;
    TfBars* bars;
    const float* P;          // current parameter blob
    uint32_t taddr;          // TMEM base + this warp's lane quarter
    int r, hf, tid;
    uint32_t pd0, pd1, pd2, pd3, pc;
    uint32_t gc;             // hand-offs to the weight-GEMM issuer so far (barrier gc % kGoBars, parity (gc / kGoBars) & 1)
    uint32_t ac;             // hand-offs to the attention issuer so far (same rotation over go_attn[])
    uint32_t kmask, kfull, kpart;   // 16-key groups of this thread's key half: attended by any row of the warp / in full by
                                    // every row / cut by a jet boundary of some row
    uint32_t kc, vc;                // pair tiles: K / V rows of how many units staged so far
    uint32_t nomax;                 // scores are bounded (TfLaunch.softmax_nomax): the softmax needs no row maximum
    unsigned long long* trace;   // clock stamps of CTA 0 / thread 0 for the first two timesteps (debugging aid) or null
    int mark_i, step;
};
// This file is compiled twice (build.py): the production object has no clock stamps at all; the object built with
// -DMMF_TILE_TRACE=1 exports launch_tf_tiles_trace, which tftile_launch uses when MMF_TRACE is set.
#ifndef MMF_TILE_TRACE
#define MMF_TILE_TRACE 0
#endif
// stamp = clock (low 56 bits) | tag << 56.  tag 0: a point of the epilogue program; 1: about to wait; 2 + b: done[b] arrived
// (so the stamp after a tag-1 stamp measures a pure wait, everything else is epilogue work) - tools/tf_trace.py sums them up
__device__ __forceinline__ void mark(Epi& e, unsigned tag = 0) {
#if MMF_TILE_TRACE
    if (e.trace && e.step < 2 && e.mark_i < 512)
        e.trace[e.step * 512 + e.mark_i] = (static_cast<unsigned long long>(clock64()) & 0x00ffffffffffffffull) | (static_cast<unsigned long long>(tag) << 56);
    ++e.mark_i;
#endif
}

__device__ __forceinline__ void epi_bar() { named_bar_sync(1, kEpi); }
__device__ __forceinline__ void wait_done(Epi& e, int b) {
    mark(e, 1);
    if (b == 0) { mbar_wait(&e.bars->done[0], e.pd0, e.mark_i); e.pd0 ^= 1; }     // (tag for the time-out diagnostics)
    else if (b == 1) { mbar_wait(&e.bars->done[1], e.pd1, e.mark_i); e.pd1 ^= 1; }
    else if (b == 2) { mbar_wait(&e.bars->done[2], e.pd2, e.mark_i); e.pd2 ^= 1; }
    else { mbar_wait(&e.bars->done[3], e.pd3, e.mark_i); e.pd3 ^= 1; }
    tc_fence_after();
    mark(e, 2 + b);
}
// hand-offs epilogue -> issuers: `go` releases the next waiting op of the weight-GEMM issuer, `go_attn` of the attention issuer
__device__ __forceinline__ void go(Epi& e) {
    mark(e);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&e.bars->go[e.gc % kGoBars]);
    ++e.gc;
}
__device__ __forceinline__ void go_attn(Epi& e, bool also_gemm = false) {
    mark(e);
    fence_proxy_async();
    tc_fence_before();
    // Rotating barriers, as for `go`: the pair-tile program hands the attention issuer two ops in a row without an MMA result
    // in between (the K rows of a unit, then its V rows + score product; P V, then the next unit's K rows).  With ONE barrier
    // an issuer warp that woke up late - cold instruction cache in the first launches of a process - found it two phases on
    // and waited for ever: an intermittent dead-lock of the pair tiles in 5 ... 15 % of fresh processes (tools/tile_stress.py).
    mbar_arrive(&e.bars->go_attn[e.ac % kGoBars]);
    ++e.ac;
    if (also_gemm) { mbar_arrive(&e.bars->go[e.gc % kGoBars]); ++e.gc; }
}
// blob `ahead` stages past the oldest one still held (0 or 1: two slots)
__device__ __forceinline__ const float* param_acquire(Epi& e, uint32_t ahead = 0) {
    const uint32_t c = e.pc + ahead, p = c & 1;
    mbar_wait(&e.bars->pfull[p], (c >> 1) & 1);
    e.P = e.pbuf + p * kTfParamFloats;
    return e.P;
}
__device__ __forceinline__ void param_release(Epi& e) {
    mbar_arrive(&e.bars->pempty[e.pc & 1]);
    ++e.pc;
}

__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Pair tiles: operand rows 80...127 do not exist in shared memory (TfLay<true>), so the threads of those rows run the
// whole program (tcgen05.ld / barriers are collective) but never store an operand row.
template <bool PAIR>
__device__ __forceinline__ bool row_ok(int r) { return !PAIR || r < static_cast<int>(TfLay<true>::kRows); }
// Pair tiles, issued by one thread of the attention issuer warp: the rows [0, 80) of this CTA's K and V operands are
// byte-for-byte the rows [80, 160) of the partner's (80 is a multiple of the 8-row swizzle atom), so each goes there as ONE
// 10 KB bulk copy shared::cta -> shared::cluster whose bytes complete on the partner's kvfull barrier.
__device__ __forceinline__ void mbar_expect_tx_remote(uint32_t cluster_bar_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t cluster_dst, const void* local_src, uint32_t bytes, uint32_t cluster_bar_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(cluster_dst), "r"(smem_u32(local_src)), "r"(bytes), "r"(cluster_bar_addr) : "memory");
}

// Reductions over register arrays with four independent chains: the epilogue runs two warps per scheduler, so a
// 32-deep dependent chain would leave the issue slots empty.
template <int N>
__device__ __forceinline__ float sum_regs(const float* v) {
    float2 s01 = MMF_V2(v, 0), s23 = MMF_V2(v, 2);
#pragma unroll
    for (int i = 4; i < N; i += 4) { s01 = f2add(s01, MMF_V2(v, i)); s23 = f2add(s23, MMF_V2(v, i + 2)); }
    return (s01.x + s01.y) + (s23.x + s23.y);
}
template <int N>
__device__ __forceinline__ float sqdev_regs(const float* v, float mean) {
    float2 q01 = f2dup(0.f), q23 = f2dup(0.f);
    const float2 nm = f2dup(-mean);
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const float2 d01 = f2add(MMF_V2(v, i), nm), d23 = f2add(MMF_V2(v, i + 2), nm);
        q01 = f2fma(d01, d01, q01); q23 = f2fma(d23, d23, q23);
    }
    return (q01.x + q01.y) + (q23.x + q23.y);
}
// LayerNorm statistics gathered in the same pass that produces the row: sums of (v - c) and (v - c)^2 around a pivot c
// taken from the row itself (its first value), so the variance does not suffer the cancellation of E[x^2] - mean^2.
struct RowStat { float c, s1, s2; };
template <int N>
__device__ __forceinline__ void stat_regs(const float* v, RowStat& st) {
    float2 a01 = f2dup(0.f), a23 = f2dup(0.f), q01 = f2dup(0.f), q23 = f2dup(0.f);
    const float2 nc = f2dup(-st.c);
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const float2 d01 = f2add(MMF_V2(v, i), nc), d23 = f2add(MMF_V2(v, i + 2), nc);
        a01 = f2add(a01, d01); a23 = f2add(a23, d23);
        q01 = f2fma(d01, d01, q01); q23 = f2fma(d23, d23, q23);
    }
    st.s1 += (a01.x + a01.y) + (a23.x + a23.y);
    st.s2 += (q01.x + q01.y) + (q23.x + q23.y);
}
template <int N>
__device__ __forceinline__ float max_regs(const float* v) {
    float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
    for (int i = 4; i < N; i += 4) { m0 = fmaxf(m0, v[i]); m1 = fmaxf(m1, v[i + 1]); m2 = fmaxf(m2, v[i + 2]); m3 = fmaxf(m3, v[i + 3]); }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// 32 fp32 values -> bf16 into the operand chunk that holds absolute column col0 (multiple of 32) of a 256-wide row
template <bool PAIR>
__device__ __forceinline__ void stage32(uint8_t* abase, int r, int col0, const float* v) {
    if (!row_ok<PAIR>(r)) return;
    uint8_t* ch = abase + (col0 >> 6) * TfLay<PAIR>::kChunk;
    const uint32_t u0 = (col0 & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        st_shared_v4(ch + sw128_offset(r, u0 + u), pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                     pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
}

// ---- residual-stream passes; every thread owns columns [hf*128, hf*128+128) of its row -----------------------
// v = resid + add0 (+ add1) (+ skip); stored back; returns the statistics of the thread's 128 columns
__device__ __forceinline__ RowStat resid_update(Epi& e, const float* add0, const float* add1, const float* skipc) {
    RowStat st{0.f, 0.f, 0.f};
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = cc * 32;
        float v[32];
        tmem_ld32(e.taddr + e.hf * 128 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 a = ldf4(add0 + c0 + 4 * u);
            MMF_SET2(v, 4 * u, f2add(MMF_V2(v, 4 * u), make_float2(a.x, a.y)));
            MMF_SET2(v, 4 * u + 2, f2add(MMF_V2(v, 4 * u + 2), make_float2(a.z, a.w)));
        }
        if (add1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4 a = ldf4(add1 + c0 + 4 * u);
                MMF_SET2(v, 4 * u, f2add(MMF_V2(v, 4 * u), make_float2(a.x, a.y)));
                MMF_SET2(v, 4 * u + 2, f2add(MMF_V2(v, 4 * u + 2), make_float2(a.z, a.w)));
            }
        }
        if (skipc) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += skipc[(c0 + i) * 128];
        }
        if (cc == 0) st.c = v[0];
        stat_regs<32>(v, st);
        tmem_st32(e.taddr + e.hf * 128 + c0, v);
    }
    tmem_st_wait();
    return st;
}
// LayerNorm mean / rstd from the row statistics: over the thread's 128 columns, or over all 256 (WIDE: the two threads of
// a row exchange (mean, M2) through shared memory and merge them)
template <bool WIDE>
__device__ __forceinline__ void ln_stats(Epi& e, const RowStat& st, int slot, float& mean, float& rstd) {
    const float dm = st.s1 * (1.0f / 128.0f);
    mean = st.c + dm;
    float m2 = fmaxf(fmaf(-st.s1, dm, st.s2), 0.f);
    if (WIDE) {
        float* ex = e.misc + mStat + slot * 512;
        ex[(e.hf * 128 + e.r) * 2] = mean;
        ex[(e.hf * 128 + e.r) * 2 + 1] = m2;
        epi_bar();
        const float om = ex[((e.hf ^ 1) * 128 + e.r) * 2], o2 = ex[((e.hf ^ 1) * 128 + e.r) * 2 + 1];
        const float d = mean - om;
        m2 = m2 + o2 + d * d * 64.0f;
        mean = 0.5f * (mean + om);
        rstd = rsqrtf(m2 * (1.0f / 256.0f) + 1e-5f);
    } else {
        rstd = rsqrtf(m2 * (1.0f / 128.0f) + 1e-5f);
    }
}
// normalised row -> bf16 GEMM operand (Abuf).  The affine part of every LayerNorm that feeds a linear layer is folded into
// that layer on the host (W' = W diag(g), b' = b + W beta: tftile_model.cu fold_ln), so the operand is (v - mean) rstd itself:
// one packed fma per two elements and no parameter loads.
template <bool PAIR>
__device__ __forceinline__ void ln_to_abuf(Epi& e, float mean, float rstd) {
    const float2 rs = f2dup(rstd), nmrs = f2dup(-mean * rstd);
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = cc * 32;
        float v[32];
        tmem_ld32(e.taddr + e.hf * 128 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) MMF_SET2(v, i, f2fma(MMF_V2(v, i), rs, nmrs));
        stage32<PAIR>(e.arena + TfLay<PAIR>::oA, e.r, e.hf * 128 + c0, v);
    }
}
// normalised row (+ post) -> back into the residual stream; returns the new row statistics (stream junction of ParticleFormer)
__device__ __forceinline__ RowStat ln_to_resid(Epi& e, float mean, float rstd, const float* g, const float* b, const float* post) {
    RowStat st{0.f, 0.f, 0.f};
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = cc * 32;
        float v[32];
        tmem_ld32(e.taddr + e.hf * 128 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 gg = ldf4(g + c0 + 4 * u), bb = ldf4(b + c0 + 4 * u), pp = ldf4(post + c0 + 4 * u);
            v[4 * u] = fmaf((v[4 * u] - mean) * rstd, gg.x, bb.x) + pp.x;
            v[4 * u + 1] = fmaf((v[4 * u + 1] - mean) * rstd, gg.y, bb.y) + pp.y;
            v[4 * u + 2] = fmaf((v[4 * u + 2] - mean) * rstd, gg.z, bb.z) + pp.z;
            v[4 * u + 3] = fmaf((v[4 * u + 3] - mean) * rstd, gg.w, bb.w) + pp.w;
        }
        if (cc == 0) st.c = v[0];
        stat_regs<32>(v, st);
        tmem_st32(e.taddr + e.hf * 128 + c0, v);
    }
    tmem_st_wait();
    return st;
}

// LayerNorm over N consecutive register values with affine parameters in shared memory
template <int N>
__device__ __forceinline__ void ln_regs(float* v, const float* g, const float* b) {
    const float mean = sum_regs<N>(v) * (1.0f / N);
    const float q = sqdev_regs<N>(v, mean);
    const float rstd = rsqrtf(q * (1.0f / N) + 1e-5f);
#pragma unroll
    const float2 rs = f2dup(rstd), nmrs = f2dup(-mean * rstd);
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const float4 gg = ldf4(g + i), bb = ldf4(b + i);
        MMF_SET2(v, i, f2fma(f2fma(MMF_V2(v, i), rs, nmrs), make_float2(gg.x, gg.y), make_float2(bb.x, bb.y)));
        MMF_SET2(v, i + 2, f2fma(f2fma(MMF_V2(v, i + 2), rs, nmrs), make_float2(gg.z, gg.w), make_float2(bb.z, bb.w)));
    }
}

// The same for rows whose mean is zero by construction: LayerNorm over a head does not see a constant added to the head's
// outputs, so the host centres the q / k rows (and biases) of c_attn per head (tftile_model.cu centre_heads) and the sum of a
// head's pre-activations vanishes up to the bf16 rounding of the centred weights (~2e-4 of the row's standard deviation).
// Variance = mean of squares, no mean pass: three packed instructions per pair instead of five.
template <int N>
__device__ __forceinline__ void ln_regs_centred(float* v, const float* g, const float* b) {
    float2 q01 = f2dup(0.f), q23 = f2dup(0.f);
#pragma unroll
    for (int i = 0; i < N; i += 4) { q01 = f2fma(MMF_V2(v, i), MMF_V2(v, i), q01); q23 = f2fma(MMF_V2(v, i + 2), MMF_V2(v, i + 2), q23); }
    const float q = (q01.x + q01.y) + (q23.x + q23.y);
    const float2 rs = f2dup(rsqrtf(q * (1.0f / N) + 1e-5f));
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const float4 gg = ldf4(g + i), bb = ldf4(b + i);
        MMF_SET2(v, i, f2fma(f2mul(MMF_V2(v, i), rs), make_float2(gg.x, gg.y), make_float2(bb.x, bb.y)));
        MMF_SET2(v, i + 2, f2fma(f2mul(MMF_V2(v, i + 2), rs), make_float2(gg.z, gg.w), make_float2(bb.z, bb.w)));
    }
}

// ---- attention epilogues -----------------------------------------------------------------------------------------
// QKV of one 64-column unit sits in scratch (q | k | v, 64 columns each).  hf 0: q and v[0,32); hf 1: k and v[32,64).
// bq/bk point at the unit's 64 bias values (the v bias lives in the projection bias, tftile_model.cu); qg.. are the per-head
// LayerNorm parameters ([HS]).
template <int HS, bool PAIR>
struct QkvCols {     // scratch columns of q|k and v (see the op emission in tftile_model.cu)
    static constexpr uint32_t cQK = HS == 64 ? TfLay<PAIR>::cQkv64 : kScr + 128, cV = HS == 64 ? TfLay<PAIR>::cQkv64 + 128 : kScr + 64;
};
// q and k: bias, per-head LayerNorm, bf16 -> the Q / K operand chunks.  The score product only needs these.
// Pair tiles: k of row r is key r here and key 80 + r in the partner CTA (copied there by the attention issuer).
// `prew` (32 registers, optional): also fetch this thread's v columns, for a v_epilogue that must not touch TMEM any more
template <int HS, bool PAIR>
__device__ __forceinline__ void qk_epilogue(Epi& e, const float* bq, const float* bk, const float* qg, const float* qb,
                                            const float* kg, const float* kb, float* prew = nullptr) {
    using L = TfLay<PAIR>;
    float v[64];
    tmem_ld32(e.taddr + QkvCols<HS, PAIR>::cQK + e.hf * 64, v);
    tmem_ld32(e.taddr + QkvCols<HS, PAIR>::cQK + e.hf * 64 + 32, v + 32);
    if (prew) tmem_ld32(e.taddr + QkvCols<HS, PAIR>::cV + e.hf * 32, prew);   // v leaves TMEM in the same round trip
    tmem_ld_wait();
    const float* bias = e.hf ? bk : bq;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        const float4 a = ldf4(bias + i);
        MMF_SET2(v, i, f2add(MMF_V2(v, i), make_float2(a.x, a.y)));
        MMF_SET2(v, i + 2, f2add(MMF_V2(v, i + 2), make_float2(a.z, a.w)));
    }
    const float* g = e.hf ? kg : qg;
    const float* b = e.hf ? kb : qb;
    if (g) {
        if (HS == 64) ln_regs_centred<64>(v, g, b);
        else { ln_regs_centred<32>(v, g, b); ln_regs_centred<32>(v + 32, g, b); }
    }
    if (row_ok<PAIR>(e.r)) {
        stage_row_bf16(e.arena + (e.hf ? L::oK : L::oQ), e.r, v);
    }
}
// v: bf16 -> V[key = r][d] (row r of a [keys][128 B] swizzled chunk: the MN-major B operand of P V, so no transpose);
// runs under the score MMA (plain tiles) / before it (pair tiles: the 160 score columns cover the v accumulator)
template <int HS, bool PAIR>
__device__ __forceinline__ void v_epilogue(Epi& e, float* prew = nullptr) {
    using L = TfLay<PAIR>;
    float wloc[32];
    float* w = prew ? prew : wloc;
    if (!prew) {
        tmem_ld32(e.taddr + QkvCols<HS, PAIR>::cV + e.hf * 32, w);
        tmem_ld_wait();
    }
    // (no bias: the probabilities of a row sum to one, so P (V + b_v) = P V + b_v and b_v travels through the projection
    // into its bias on the host - tftile_model.cu fold_v_bias)
    if (!row_ok<PAIR>(e.r)) return;
    uint8_t* vb = e.arena + L::oVT;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        st_shared_v4(vb + sw128_offset(e.r, e.hf * 4 + u), pack_bf16x2(w[8 * u], w[8 * u + 1]), pack_bf16x2(w[8 * u + 2], w[8 * u + 3]),
                     pack_bf16x2(w[8 * u + 4], w[8 * u + 5]), pack_bf16x2(w[8 * u + 6], w[8 * u + 7]));
}

// scores of one head in scratch columns [scol, scol + 2 NK); thread handles keys [hf*NK, +NK) of its row (NK = 64: plain
// tiles; NK = 80: pair tiles, hf 0 = the CTA's own rows, hf 1 = the partner's).
// e.kmask: which of the thread's 16-key groups hold a key of ANY row of this warp (warp-uniform, fixed for the
// launch): the other groups are outside every jet of these 32 rows, so their probabilities are exact zeros.
// softmax_probs leaves the unnormalised probabilities in s[] and returns their sum; softmax_store writes them as the bf16
// P operand (plain: chunk hf of the Q|K staging area, which the previous P V product must have finished reading; pair: its
// own region).  `slot` selects the exchange buffers.
// `release_gemm`: the scores read here are the last live scratch columns the next unit's QKV product overwrites, so the
// weight-GEMM issuer is handed that product as soon as they sit in registers (a whole softmax earlier than the P V hand-off;
// it reads only Abuf and the ring, no shared memory written by this epilogue, hence no proxy fence).
// Key j of this thread is valid iff (unsigned)(j - lo) < span (plain: the row's jet; pair: the real rows of that CTA).
template <int NK>
__device__ __forceinline__ float softmax_probs(Epi& e, uint32_t scol, int lo, uint32_t span, int slot, float* s,
                                               bool release_gemm = false) {
    constexpr int NG = NK / 16;
    const uint32_t km = e.kmask;
    tmem_ld32(e.taddr + scol + e.hf * NK, s);
    tmem_ld32(e.taddr + scol + e.hf * NK + 32, s + 32);
    if (NK == 80) tmem_ld16(e.taddr + scol + e.hf * NK + 64, s + 64);
    tmem_ld_wait();
    if (release_gemm) {
        tc_fence_before();
        mbar_arrive(&e.bars->go[e.gc % kGoBars]);
        ++e.gc;
    }
    // keys outside the row's jet get -inf: they drop out of the max and ex2(-inf) = 0 removes them from the sum.
    // lo / span are made opaque here: otherwise the compiler hoists the comparisons out of the timestep loop into a bit
    // mask and then serialises on predicate registers.
    // Per 16-key group (all warp-uniform, fixed for the launch): e.kfull - inside the jet of every row of the warp, no
    // masking; e.kpart - a jet boundary of some row falls inside the group, per-key masks; otherwise every row is either
    // all-in or all-out and one per-lane predicate covers the 16 keys.
    asm volatile("" : "+r"(lo), "+r"(span));
    const int hi = lo + static_cast<int>(span);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (km & (1u << g)) {
            if (e.kpart & (1u << g)) {
#pragma unroll
                for (int j = 16 * g; j < 16 * g + 16; ++j) s[j] = static_cast<uint32_t>(j - lo) < span ? s[j] : -INFINITY;
            } else if (!(e.kfull & (1u << g))) {
                const bool in = lo <= 16 * g && hi >= 16 * g + 16;
#pragma unroll
                for (int j = 16 * g; j < 16 * g + 16; ++j) s[j] = in ? s[j] : -INFINITY;
            }
        }
    }
    // The row maximum only guards the exponentials against overflow; softmax itself is shift-invariant.  q and k are
    // per-head LayerNorm outputs, so |q.k| <= (max|g_q| sqrt(HS) + |b_q|)(max|g_k| sqrt(HS) + |b_k|): when the host finds that
    // bound (in the units of the exponent) small for every block of the checkpoint (e.nomax, CTA-uniform), the maximum, its
    // exchange between the two halves of a row and the barrier are skipped and the exponent is taken of the raw score.
    float msc = 0.f;
    if (!e.nomax) {
        float mx = -INFINITY;
#pragma unroll
        for (int g = 0; g < NG; ++g)
            if (km & (1u << g)) mx = fmaxf(mx, max_regs<16>(s + 16 * g));
        float* red = e.misc + mRed + slot * 256;
        red[e.hf * 128 + e.r] = mx;
        epi_bar();
        mx = fmaxf(mx, red[(e.hf ^ 1) * 128 + e.r]);
        msc = (mx == -INFINITY) ? 0.f : mx;
    }
    // The scores arrive in the units of the exponent: log2(e) / sqrt(HS) is folded into the affine part of the q LayerNorm on
    // the host (tftile_model.cu), so without the maximum the exponential is taken of the accumulator as it is.
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (km & (1u << g)) {
            if (e.nomax) {
#pragma unroll
                for (int j = 16 * g; j < 16 * g + 16; ++j) s[j] = ex2_approx(s[j]);
            } else {
#pragma unroll
                for (int j = 16 * g; j < 16 * g + 16; j += 2) {
                    const float2 a = f2add(MMF_V2(s, j), f2dup(-msc));
                    s[j] = ex2_approx(a.x); s[j + 1] = ex2_approx(a.y);
                }
            }
            sum += sum_regs<16>(s + 16 * g);
        } else {
#pragma unroll
            for (int j = 16 * g; j < 16 * g + 16; ++j) s[j] = 0.f;
        }
    }
    return sum;
}
template <bool PAIR>
__device__ __forceinline__ void softmax_store(Epi& e, int slot, const float* s, float sum) {
    using L = TfLay<PAIR>;
    if (!PAIR) {
        stage_row_bf16(e.arena + (e.hf ? L::oK : L::oQ), e.r, s);          // P chunk hf (keys hf*64..)
    } else if (row_ok<PAIR>(e.r)) {
        // keys [hf*80, +80) = 16-byte units hf*10 .. hf*10 + 9 of the row; 8 units per 64-key chunk, chunks kChunk apart
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const int U = e.hf * 10 + i;
            st_shared_v4(e.arena + L::oP + (U >> 3) * L::kChunk + sw128_offset(e.r, U & 7), pack_bf16x2(s[8 * i], s[8 * i + 1]),
                         pack_bf16x2(s[8 * i + 2], s[8 * i + 3]), pack_bf16x2(s[8 * i + 4], s[8 * i + 5]), pack_bf16x2(s[8 * i + 6], s[8 * i + 7]));
        }
    }
    e.misc[mSum + slot * 256 + e.hf * 128 + e.r] = sum;
}
template <bool PAIR>
__device__ __forceinline__ void softmax_epilogue(Epi& e, uint32_t scol, int lo, uint32_t span, int slot, bool release_gemm = false) {
    float s[PAIR ? 80 : 64];
    const float sum = softmax_probs<PAIR ? 80 : 64>(e, scol, lo, span, slot, s, release_gemm);
    softmax_store<PAIR>(e, slot, s, sum);
}

// O = P V of one head in scratch columns [ocol, ocol+HS) -> normalised bf16 into Os columns [ucol, ucol+HS) of the unit
template <int HS, bool PAIR>
__device__ __forceinline__ void o_epilogue(Epi& e, uint32_t ocol, int ucol, int slot) {
    const float tot = e.misc[mSum + slot * 256 + e.r] + e.misc[mSum + slot * 256 + 128 + e.r];
    const float inv = rcp_approx(tot > 0.f ? tot : 1.f);
    constexpr int W = HS / 2;
    float o[W];
    if (W == 32) tmem_ld32(e.taddr + ocol + e.hf * W, o);
    else tmem_ld16(e.taddr + ocol + e.hf * W, o);
    tmem_ld_wait();
    if (!row_ok<PAIR>(e.r)) return;
    const uint32_t u0 = (ucol + e.hf * W) >> 3;
    uint8_t* os = e.arena + TfLay<PAIR>::oO;
    // (the products below are written o[i] * inv: the compiler does not pair them, so scale the row with packed multiplies first)
#pragma unroll
    for (int i = 0; i < W; i += 2) MMF_SET2(o, i, f2mul(MMF_V2(o, i), f2dup(inv)));
#pragma unroll
    for (int u = 0; u < W / 8; ++u)
        st_shared_v4(os + sw128_offset(e.r, u0 + u), pack_bf16x2(o[8 * u], o[8 * u + 1]), pack_bf16x2(o[8 * u + 2], o[8 * u + 3]),
                     pack_bf16x2(o[8 * u + 4], o[8 * u + 5]), pack_bf16x2(o[8 * u + 6], o[8 * u + 7]));
}

// MLP hidden quarter q in scratch half (q&1): 2 GELU(acc + bias) -> bf16 H(q&1) (the down-projection weights carry the 0.5,
// tftile_model.cu); `bias` points at the quarter's 128 values
template <bool PAIR>
__device__ __forceinline__ void fc_epilogue(Epi& e, int q, const float* bias) {
    using L = TfLay<PAIR>;
    float v[64];
    const uint32_t col = kScr + (q & 1) * 128 + e.hf * 64;
    tmem_ld32(e.taddr + col, v);
    tmem_ld32(e.taddr + col + 32, v + 32);
    tmem_ld_wait();
    if (!row_ok<PAIR>(e.r)) return;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        const float4 a = ldf4(bias + e.hf * 64 + i);
        MMF_SET2(v, i, gelu2x_tile2(f2add(MMF_V2(v, i), make_float2(a.x, a.y))));
        MMF_SET2(v, i + 2, gelu2x_tile2(f2add(MMF_V2(v, i + 2), make_float2(a.z, a.w))));
    }
    stage_row_bf16(e.arena + ((q & 1) ? L::oH1 : L::oH0) + e.hf * L::kChunk, e.r, v);
}

// head hidden quarter: GELU(acc + bias) dotted with NO output rows of W2 (row stride ld), accumulated into out[]
template <int NO>
__device__ __forceinline__ void head_epilogue(Epi& e, int q, const float* bias, const float* w2, int ld, float* out) {
    float v[64];
    const uint32_t col = kScr + (q & 1) * 128 + e.hf * 64;
    tmem_ld32(e.taddr + col, v);
    tmem_ld32(e.taddr + col + 32, v + 32);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        const float4 a = ldf4(bias + e.hf * 64 + i);
        MMF_SET2(v, i, gelu2x_tile2(f2add(MMF_V2(v, i), make_float2(a.x, a.y))));
        MMF_SET2(v, i + 2, gelu2x_tile2(f2add(MMF_V2(v, i + 2), make_float2(a.z, a.w))));
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        const float* w = w2 + o * ld + e.hf * 64;
        float2 acc = f2dup(0.f);                              // even / odd hidden units, packed
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            const float4 a = ldf4(w + i);
            acc = f2fma(make_float2(a.x, a.y), MMF_V2(v, i), acc);
            acc = f2fma(make_float2(a.z, a.w), MMF_V2(v, i + 2), acc);
        }
        out[o] += acc.x + acc.y;
    }
}

// one attention unit (64 q-columns = 64/HS heads) of the current block, epilogue side
// pair tiles: q / k of a unit -> Q / K operand rows, then the hand-off that lets the attention issuer ship the K rows to the
// partner.  Runs for the first unit of a block at its top, for every later unit under the last P V product of the unit
// before it (its QKV accumulator is ready by then), so that the exchange - DSMEM moves ~17 B/clk, 10 KB take ~1200 cycles
// each way - is over when the unit starts.  `G`: the parameter group of that unit (q | k bias at G[qoff], G[koff]).
template <int HS>
__device__ __forceinline__ void pair_stage_k(Epi& e, const float* bq, const float* bk, const float* qg, const float* qb, const float* kg,
                                             const float* kb) {
    wait_done(e, 1);                                  // QKV of that unit
    // kfree of the previous exchange: the partner's score products ran, so this CTA's outgoing K copy left the rows long ago
    if (e.kc > 0) { mark(e, 1); mbar_wait_cluster(&e.bars->kfree, (e.kc - 1) & 1); mark(e, 6); }
    ++e.kc;
    qk_epilogue<HS, true>(e, bq, bk, qg, qb, kg, kb);
    go_attn(e);                                       // -> the issuer sends the K rows
}

// nbq ... nkb: the q / k parameters of the NEXT unit of the block (pair tiles stage its K rows ahead; unused otherwise)
template <int HS, bool PAIR>
__device__ __forceinline__ void attention_unit(Epi& e, bool first, bool more, const float* bq, const float* bk, const float* qg,
                                               const float* qb, const float* kg, const float* kb, int lo, uint32_t span,
                                               const float* nbq = nullptr, const float* nbk = nullptr, const float* nqg = nullptr,
                                               const float* nqb = nullptr, const float* nkg = nullptr, const float* nkb = nullptr) {
    using L = TfLay<PAIR>;
    if (!PAIR) wait_done(e, 1);                       // QKV of this unit (issued under the previous unit's epilogue)
    if (!PAIR) {
        // 32-wide units: the score product of head 0 writes [256,384), over the v accumulator [320,384): v must be in
        // registers before the hand-off (until round 2 it was read right after it - a race the timing happened to hide)
        float vr[32];
        qk_epilogue<HS, false>(e, bq, bk, qg, qb, kg, kb, HS == 32 ? vr : nullptr);
        go_attn(e);                                       // -> S
        v_epilogue<HS, false>(e, HS == 32 ? vr : nullptr);   // under the score MMA; P V is only issued after the next hand-off
        if (HS == 64) {
            wait_done(e, 0);
            softmax_epilogue<false>(e, kScr, lo, span, 0, more);   // (S in registers -> the QKV GEMM of the next unit)
            go_attn(e);                                       // -> P V
            wait_done(e, 0);
            if (!first) wait_done(e, 2);                      // the previous unit's projection (other issuer warp) has read oO
            o_epilogue<64, false>(e, kScr + 192, 0, 0);
            go(e);                                            // -> projection
        } else {
            wait_done(e, 0);                                  // both heads' scores: [256,384) and [384,512)
            softmax_epilogue<false>(e, kScr, lo, span, 0);
            go_attn(e);                                       // -> P V of head 0
            float s[64];                                      // head 1's probabilities are computed under that product ...
            const float sum = softmax_probs<64>(e, kScr + 128, lo, span, 1, s, more);   // (both S in registers -> next QKV GEMM)
            wait_done(e, 3);                                  // ... and stored once it has finished reading head 0's
            softmax_store<false>(e, 1, s, sum);
            go_attn(e);                                       // -> P V of head 1
            if (!first) wait_done(e, 2);                      // the previous unit's projection (other issuer warp) has read oO
            o_epilogue<32, false>(e, kScr, 0, 0);             // O of head 0 in scratch [0,32), under P V of head 1
            wait_done(e, 0);                                  // O of head 1 in scratch [32,64)
            o_epilogue<32, false>(e, kScr + 32, 32, 1);
            go(e);
        }
    } else {
        // ---- pair tile: this CTA holds one half of a 129...160-particle jet; keys = own 80 rows | partner's 80 rows ----
        // The epilogue only fills this CTA's own rows; the attention issuer ships them to the partner (K ahead of time, V at the
        // start of the unit, arriving under the softmax) and waits for the partner's before S / P V (see the issuer loop).
        if (first) pair_stage_k<HS>(e, bq, bk, qg, qb, kg, kb);
        if (e.vc > 0) { mark(e, 1); mbar_wait_cluster(&e.bars->vfree, (e.vc - 1) & 1); mark(e, 6); }   // (as kfree, for the V rows)
        ++e.vc;
        v_epilogue<HS, true>(e);                      // (before S: its 160 columns cover the v accumulator)
        go_attn(e);                                       // -> V rows to the partner, then S once the partner's K rows are here
        constexpr uint32_t cS = L::cS;
        if (HS == 64) {
            wait_done(e, 0);
            softmax_epilogue<true>(e, cS, lo, span, 0, more);
            go_attn(e);                                       // -> P V (O in scratch [0,64))
            if (more) pair_stage_k<HS>(e, nbq, nbk, nqg, nqb, nkg, nkb);   // next unit's q / k under this unit's P V
            wait_done(e, 0);
            if (!first) wait_done(e, 2);
            o_epilogue<64, true>(e, L::cO64, 0, 0);
            go(e);                                            // -> projection
        } else {
            // the two heads take turns on the score columns: S0, softmax 0, then S1 under the P store of head 0
            wait_done(e, 0);                                  // S of head 0
            float s[80];
            float sum = softmax_probs<80>(e, cS, lo, span, 0, s);
            go_attn(e);                                       // S0 in registers -> S of head 1 may overwrite its columns
            softmax_store<true>(e, 0, s, sum);                // (P is free: the previous unit's P V products were waited for)
            go_attn(e);                                       // -> P V of head 0 (done[3])
            wait_done(e, 0);                                  // S of head 1
            sum = softmax_probs<80>(e, cS, lo, span, 1, s, more);   // (all scores in registers -> next QKV GEMM)
            wait_done(e, 3);                                  // P V of head 0 has read P
            softmax_store<true>(e, 1, s, sum);
            go_attn(e);                                       // -> P V of head 1
            if (!first) wait_done(e, 2);
            o_epilogue<32, true>(e, kScr, 0, 0);              // O of head 0 in scratch [0,32), under P V of head 1
            if (more) pair_stage_k<HS>(e, nbq, nbk, nqg, nqb, nkg, nkb);   // next unit's q / k under P V of head 1
            wait_done(e, 0);                                  // O of head 1 in scratch [32,64)
            o_epilogue<32, true>(e, kScr + 32, 32, 1);
            go(e);
        }
    }
}

template <int V, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) tf_tile_kernel(const __grid_constant__ TfLaunch a, const __grid_constant__ TfOpTable optab,
                                                              const __grid_constant__ TfProdTable prodtab) {
    using L = TfLay<PAIR>;
    // The dynamic shared window starts 1024-byte aligned (no static shared memory in this kernel); it is used directly so
    // that the compiler keeps the shared address space (LDS/STS instead of generic loads).  SWIZZLE_128B needs the alignment.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    TfBars* bars = reinterpret_cast<TfBars*>(smem);
    uint8_t* arena = smem + 1024;
    float* pbuf = reinterpret_cast<float*>(arena + L::kArena);
    float* misc = pbuf + 2 * kTfParamFloats;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = a.tile0 + blockIdx.x;
    const TfTileMeta* meta = a.meta + tile;

    // CTAs of a cluster run the same op sequence on different tiles and share every weight tile: each loads 1/cs of it
    // and multicasts the slice to all of them, so a ring stage is free only when all cs MMA issuers released it.
    const uint32_t cs = cluster_nctarank(), crank = cluster_ctarank();
    const uint16_t cmask = static_cast<uint16_t>((1u << cs) - 1u);
    if (tid == 0) {
        for (int i = 0; i < kBars; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], cs); }
        mbar_init(&bars->done[0], 1);
        mbar_init(&bars->done[1], 1);
        mbar_init(&bars->done[2], 1);
        mbar_init(&bars->done[3], 1);
        for (int i = 0; i < kGoBars; ++i) mbar_init(&bars->go[i], kEpi);
        for (int i = 0; i < kGoBars; ++i) mbar_init(&bars->go_attn[i], kEpi);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->pfull[i], 1); mbar_init(&bars->pempty[i], kEpi); }
        mbar_init(&bars->kfull, 1);                       // pair tiles (see TfBars)
        mbar_init(&bars->vfull, 1);
        mbar_init(&bars->kfree, 1);
        mbar_init(&bars->vfree, 1);
        fence_mbar_init();
    }
    if (warp == 9) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();                      // every CTA's barriers exist before any multicast copy / remote arrive
    const uint32_t tmem_base = bars->tmem_base;

    // The producer and MMA warps run their loops with all 32 lanes converged (loop state and op fields stay in uniform
    // registers, the tables sit in the constant bank); one elected lane issues the asynchronous instructions.
    if (warp == 8) {
        // ---------------------------------------------------- parameter producer ------------------------------------
        // Blobs are consumed in index order; blob c goes to slot c & 1 as soon as the epilogue warps released blob c - 2.
        uint32_t pcount = 0;
        for (int step = 0; step < a.nsteps; ++step) {
            for (int i = 0; i < a.n_blobs; ++i) {
                const uint32_t p = pcount & 1;
                if (pcount >= 2) mbar_wait(&bars->pempty[p], ((pcount >> 1) - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(&bars->pfull[p], kTfParamFloats * 4);
                    bulk_load_1d(pbuf + p * kTfParamFloats, a.params + static_cast<size_t>(i) * kTfParamFloats, kTfParamFloats * 4, &bars->pfull[p]);
                }
                __syncwarp();
                ++pcount;
            }
        }
    } else if (warp == 10) {
        // ---------------------------------------------------- weight producers --------------------------------------
        // Tile g of the launch may be written once every tile up to g - dep has been consumed (host plan); `known` counts
        // the tiles this warp has seen consumed, waiting on their barriers strictly in order.
        const uint32_t pw = static_cast<uint32_t>(warp - 10);
        uint32_t known = 0, gbase = 0;
        for (int step = 0; step < a.nsteps; ++step, gbase += a.n_prod) {
            for (int i = pw; i < a.n_prod; i += kProducers) {
                const uint2 ent = prodtab.e[i];
                const uint32_t g = gbase + i, bytes = (ent.x >> 24) * 1024u;
                const int need = static_cast<int>(g) - static_cast<int>((ent.y >> 8) & 0xffu) + 1;
                if (static_cast<int>(known) < need) {
                    // consumption is in order, so waiting for tile need - 1 covers all earlier ones; its barrier cannot be a
                    // phase behind, because need advances by far fewer than kBars tiles between two waits of this warp
                    const uint32_t t = static_cast<uint32_t>(need - 1);
                    mbar_wait(&bars->empty[t % kBars], (t / kBars) & 1);
                    known = static_cast<uint32_t>(need);
                }
                if (elect_one()) {
                    uint64_t* full = &bars->full[g % kBars];
                    mbar_expect_tx(full, bytes);
                    const uint32_t slice = cs == 1 ? bytes : (cs == 2 ? bytes >> 1 : bytes >> 2), part = slice / kCopySplit;
                    uint8_t* dst = arena + L::oRing + (ent.y & 0xffu) * 1024u + crank * slice;
                    const uint8_t* src = a.wstream + static_cast<size_t>(ent.x & 0xffffffu) * 128u + crank * slice;
#pragma unroll
                    for (int c = 0; c < kCopySplit; ++c) {
                        if (cs == 1) bulk_load_1d(dst + c * part, src + c * part, part, full);
                        else bulk_load_1d_multicast(dst + c * part, src + c * part, part, full, cmask);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 9 || warp == 11) {
        // ---------------------------------------------------- MMA issuers -------------------------------------------
        // Both warps walk the whole table; warp 11 issues the ops flagged kTfOpAttn, warp 9 all the others.
        const bool attn_issuer = warp == 11;
        // All 32 lanes run the loop converged (op fields stay in uniform registers, the table sits in the constant bank);
        // one elected lane issues the asynchronous instructions.  Descriptors come precomputed from the host.
        uint32_t pg = 0, gbase = 0, nK = 0, nV = 0;           // hand-offs consumed so far; pair tiles: K / V exchanges started
        const uint32_t base16 = smem_u32(arena) >> 4;
        const uint32_t ring16 = base16 + (L::oRing >> 4) + (1u << 16);
        constexpr uint64_t kDescHi = static_cast<uint64_t>(0x40004040u) << 32;   // SBO 1024 B | version 1 | SWIZZLE_128B
        for (int step = 0; step < a.nsteps; ++step, gbase += a.n_prod) {
            TfOp nx = optab.ops[0];
            uint32_t ti = 0, nt = prodtab.e[0].y;             // next weight tile of this timestep and its ring placement
            for (int i = 0; i < a.n_ops; ++i) {
                const TfOp op = nx;
                if (i + 1 < a.n_ops) nx = optab.ops[i + 1];   // fetched one op ahead
                const uint32_t fl = op.flags;
                if (((fl & kTfOpAttn) != 0) != attn_issuer) continue;       // the other issuer's op (attention ops never touch the ring)
                if (fl & kTfOpWait) {
                    if (attn_issuer) mbar_wait(&bars->go_attn[pg % kGoBars], (pg / kGoBars) & 1, static_cast<uint32_t>(i));
                    else mbar_wait(&bars->go[pg % kGoBars], (pg / kGoBars) & 1, static_cast<uint32_t>(i));
                    ++pg;
                }
                const uint32_t pfl = PAIR ? (static_cast<uint32_t>(op.dcol) >> kTfPairShift) : 0u;
                if (PAIR && pfl) {
                    // K / V exchange of a pair tile.  The rows [0, 80) of this CTA's K (V) operand are byte-for-byte the rows
                    // [80, 160) of the partner's, so each goes there as ONE 10 KB bulk copy whose bytes complete on the
                    // partner's kfull (vfull); a copy may start once the partner's MMAs are done with that buffer (kfree / vfree).
                    constexpr uint32_t kHalf = L::kRows * 128u;
                    const uint32_t peer = crank ^ 1u;
                    if (pfl & kTfPairSendK) {
                        if (nK > 0) mbar_wait_cluster(&bars->kfree, (nK - 1) & 1);
                        if (elect_one()) {
                            const uint32_t kbar = dsmem_addr(&bars->kfull, peer);
                            mbar_expect_tx_remote(kbar, kHalf);
                            bulk_copy_to_peer(dsmem_addr(arena + L::oK + kHalf, peer), arena + L::oK, kHalf, kbar);
                        }
                        __syncwarp();
                        ++nK;
                    }
                    if (pfl & kTfPairSendVWaitK) {
                        if (nV > 0) mbar_wait_cluster(&bars->vfree, (nV - 1) & 1);
                        if (elect_one()) {
                            const uint32_t vbar = dsmem_addr(&bars->vfull, peer);
                            mbar_expect_tx_remote(vbar, kHalf);
                            bulk_copy_to_peer(dsmem_addr(arena + L::oVT + kHalf, peer), arena + L::oVT, kHalf, vbar);
                        }
                        __syncwarp();
                        mbar_wait_cluster(&bars->kfull, nV & 1);          // the partner's K rows of this unit
                        ++nV;
                    }
                    if (pfl & kTfPairWaitV) mbar_wait_cluster(&bars->vfull, (nV - 1) & 1);
                }
                const bool ring = (fl & kTfOpRing) != 0;
                uint32_t a_lo = op.a_lo + base16, b_lo = op.b_lo + base16, acc = fl & kTfOpAcc;
                const uint32_t d = tmem_base + (op.dcol & kTfDcolMask);
                const uint32_t nkt = op.nkt & 0x07u, sig = ((fl >> 4) & 3u) | ((op.nkt & 0x80u) >> 5);
                for (uint32_t kt = 0; kt < nkt; ++kt) {
                    const uint32_t g = gbase + ti;
                    if (ring) {
                        b_lo = ring16 + (nt & 0xffu) * (1024u >> 4);
                        ++ti;
                        nt = prodtab.e[ti < static_cast<uint32_t>(a.n_prod) ? ti : 0].y;
                        mbar_wait(&bars->full[g % kBars], (g / kBars) & 1, 0x8000u | static_cast<uint32_t>(i));
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = kDescHi | a_lo, db = kDescHi | b_lo;
                        const uint32_t bs = (fl & kTfOpBMn) ? (2048u >> 4) : 2u;   // B per K = 16: 16 rows (MN-major) or 32 bytes
                        umma_bf16(d, da, db, op.idesc, acc);
                        umma_bf16(d, da + 2, db + bs, op.idesc, 1u);
                        if (!(fl & kTfOpHalfK)) {
                            umma_bf16(d, da + 4, db + 2 * bs, op.idesc, 1u);
                            umma_bf16(d, da + 6, db + 3 * bs, op.idesc, 1u);
                        }
                        if (ring) {
                            if (cs == 1) umma_commit(&bars->empty[g % kBars]); else umma_commit_multicast(&bars->empty[g % kBars], cmask);
                        }
                    }
                    __syncwarp();
                    a_lo += L::kChunk >> 4;
                    b_lo += 8192 >> 4;
                    acc = 1u;
                }
                if (PAIR && (pfl & (kTfPairFreeK | kTfPairFreeV))) {      // last product of the unit that reads K (V): tell the partner
                    if (elect_one()) {
                        if (pfl & kTfPairFreeK) umma_commit_multicast(&bars->kfree, static_cast<uint16_t>(1u << (crank ^ 1u)));
                        if (pfl & kTfPairFreeV) umma_commit_multicast(&bars->vfree, static_cast<uint16_t>(1u << (crank ^ 1u)));
                    }
                    __syncwarp();
                }
                if (sig) {
                    if (elect_one()) {
                        umma_commit(&bars->done[sig - 1u]);
#if MMF_TILE_TRACE
                        if (a.trace && blockIdx.x == 0 && step == 1 && i < 128) a.trace[1024 + i] = clock64();
#endif
                    }
                    __syncwarp();
                }
            }
        }
        // the partner's last kvfree commit lands in THIS CTA's shared memory: wait for it before anybody may leave
        if (PAIR && attn_issuer && nK > 0) mbar_wait_cluster(&bars->kfree, (nK - 1) & 1);
        if (PAIR && attn_issuer && nV > 0) mbar_wait_cluster(&bars->vfree, (nV - 1) & 1);
        __syncwarp();
    } else {
        // ---------------------------------------------------- epilogue warps ----------------------------------------
        Epi e;
        e.arena = arena; e.pbuf = pbuf; e.misc = misc; e.bars = bars; e.P = pbuf;
        e.r = (warp & 3) * 32 + lane; e.hf = warp >> 2; e.tid = tid;
        e.taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        e.pd0 = 0; e.pd1 = 0; e.pd2 = 0; e.pd3 = 0; e.pc = 0; e.gc = 0; e.ac = 0;
        const int r = e.r, hf = e.hf;
        const int nrows = meta->nrows;
        float* s_xs = misc + mXs;
        int* s_ks = reinterpret_cast<int*>(misc + mKs);
        if (tid < 128) {
#pragma unroll
            for (int c = 0; c < 3; ++c) s_xs[tid * 3 + c] = a.xs0[(static_cast<size_t>(tile) * 128 + tid) * 3 + c];
            s_ks[tid] = a.ks0[static_cast<size_t>(tile) * 128 + tid];
        }
        // keys this thread's row attends, in the coordinates of the thread's own key range [hf*NK, hf*NK + NK):
        // plain tile: the rows of its jet; pair tile: the real rows of this CTA (hf 0) / of the partner (hf 1)
        constexpr int NK = PAIR ? 80 : 64;
        const int seg_b = PAIR ? hf * NK : meta->seg_beg[r];
        const int seg_e = PAIR ? hf * NK + (hf ? meta->pad[0] : meta->nrows) : meta->seg_end[r];
        const int att_lo = seg_b - hf * NK;
        const uint32_t att_span = static_cast<uint32_t>(seg_e - seg_b);
        e.kmask = 0; e.kfull = 0; e.kpart = 0;
        e.nomax = a.softmax_nomax ? 1u : 0u;
#pragma unroll
        for (int g = 0; g < NK / 16; ++g) {
            const int k0 = hf * NK + 16 * g;
            if (__ballot_sync(0xffffffffu, seg_b < k0 + 16 && seg_e > k0) != 0u) e.kmask |= 1u << g;
            if (__all_sync(0xffffffffu, seg_b <= k0 && seg_e >= k0 + 16)) e.kfull |= 1u << g;
            const bool in = seg_b <= k0 && seg_e >= k0 + 16, out = seg_e <= k0 || seg_b >= k0 + 16;
            if (!__all_sync(0xffffffffu, in || out)) e.kpart |= 1u << g;
        }

        e.kc = 0; e.vc = 0;
        const int tb_row = a.per_jet_time ? meta->row_tb[r] : 0;
        const long long slot = a.row_slot[static_cast<size_t>(tile) * 128 + r];
        float* skipc = a.skip + (static_cast<size_t>(tile) * 256 + hf * 128) * 128 + r;    // + col * 128
        const bool pf = a.arch == MMF_ARCH_PARTICLEFORMER;
        epi_bar();

        e.trace = (blockIdx.x == 0 && tid == 0) ? a.trace : nullptr;
        e.mark_i = 0; e.step = 0;

        for (int step = 0; step < a.nsteps; ++step) {
            e.mark_i = 0; e.step = step;
            mark(e);
            // time embedding of this step: shared row in the sampler, per-jet rows (global) in the forward API
            const float* tb;
            if (a.per_jet_time) {
                tb = a.temb + static_cast<size_t>(tb_row) * 512;
            } else {
                float* s_t = misc + mTemb;
                s_t[tid] = a.temb[static_cast<size_t>(step) * 512 + tid];
                s_t[256 + tid] = a.temb[static_cast<size_t>(step) * 512 + 256 + tid];
                epi_bar();
                tb = s_t;
            }
            const float* tb1 = tb + hf * 128;                 // stream-level embedding, this thread's columns
            const float* tb2 = tb + (pf ? 256 : 0) + hf * 128;  // embedding added inside the main (256-wide) blocks

            // ================= embedding stage =================
            const float* PA = param_acquire(e, 0);            // continuous branch
            {
                const float x0 = s_xs[r * 3], x1 = s_xs[r * 3 + 1], x2 = s_xs[r * 3 + 2];
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int chunk = hf * 2 + cc;
                    float v[64];
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        const float4 w = ldf4(PA + tfp::EA_W0 + (chunk * 64 + i) * 4);
                        v[i] = gelu_tile(fmaf(w.z, x2, fmaf(w.y, x1, fmaf(w.x, x0, w.w))));
                    }
                    if (row_ok<PAIR>(r)) stage_row_bf16(arena + L::oA + chunk * L::kChunk, r, v);
                }
                go(e);
            }
            const float* PB = param_acquire(e, 1);            // discrete branch + first LayerNorm
            wait_done(e, 0);
            {
                // x half: LN_ln1x(wxe.2 output + bias) + temb ; y half: Ytab[k] + temb   -> residual + skip streams
                RowStat sum{0.f, 0.f, 0.f};
                if (hf == 0) {
                    RowStat st1{0.f, 0.f, 0.f};
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        float v[32];
                        tmem_ld32(e.taddr + kScr + cc * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 bx = ldf4(PA + tfp::EA_BXE2 + cc * 32 + i);
                            v[i] += bx.x; v[i + 1] += bx.y; v[i + 2] += bx.z; v[i + 3] += bx.w;
                        }
                        if (cc == 0) st1.c = v[0];
                        stat_regs<32>(v, st1);
                    }
                    const float dm = st1.s1 * (1.0f / 128.0f), mean = st1.c + dm;
                    const float rstd = rsqrtf(fmaxf(fmaf(-st1.s1, dm, st1.s2), 0.f) * (1.0f / 128.0f) + 1e-5f);
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        float v[32];
                        tmem_ld32(e.taddr + kScr + cc * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const int c = cc * 32 + i;
                            const float4 bx = ldf4(PA + tfp::EA_BXE2 + c), gg = ldf4(PA + tfp::EA_LN1X_G + c), bb = ldf4(PA + tfp::EA_LN1X_B + c), tt = ldf4(tb1 + c);
                            v[i] = fmaf((v[i] + bx.x - mean) * rstd, gg.x, bb.x) + tt.x;
                            v[i + 1] = fmaf((v[i + 1] + bx.y - mean) * rstd, gg.y, bb.y) + tt.y;
                            v[i + 2] = fmaf((v[i + 2] + bx.z - mean) * rstd, gg.z, bb.z) + tt.z;
                            v[i + 3] = fmaf((v[i + 3] + bx.w - mean) * rstd, gg.w, bb.w) + tt.w;
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) skipc[(cc * 32 + i) * 128] = v[i];
                        if (cc == 0) sum.c = v[0];
                        stat_regs<32>(v, sum);
                        tmem_st32(e.taddr + cc * 32, v);
                    }
                } else {
                    const float* yt = PB + tfp::EB_YTAB + s_ks[r] * 128;
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 yy = ldf4(yt + cc * 32 + i), tt = ldf4(tb1 + cc * 32 + i);
                            v[i] = yy.x + tt.x; v[i + 1] = yy.y + tt.y; v[i + 2] = yy.z + tt.z; v[i + 3] = yy.w + tt.w;
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) skipc[(cc * 32 + i) * 128] = v[i];
                        if (cc == 0) sum.c = v[0];
                        stat_regs<32>(v, sum);
                        tmem_st32(e.taddr + 128 + cc * 32, v);
                    }
                }
                tmem_st_wait();
                float mean, rstd;
                if (pf) ln_stats<false>(e, sum, 0, mean, rstd); else ln_stats<true>(e, sum, 0, mean, rstd);
                ln_to_abuf<PAIR>(e, mean, rstd);
                go(e);
            }
            param_release(e);
            param_release(e);
            mark(e);

            // ================= stream blocks (ParticleFormer): two independent 128-wide groups =================
            for (int blk = 0; blk < a.n_stream; ++blk) {
                param_acquire(e);                             // attention stage
                const bool last = blk + 1 == a.n_stream;
                for (int g = 0; g < 2; ++g) {
                    const float* G = e.P + g * tfp::SA_GROUP;
                    for (int u = 0; u < 2; ++u)
                    {
                        const int ng = u == 1 ? 1 : g, nu = u ^ 1;          // the unit after (g, u): (g, 1) or (1, 0)
                        const float* NG = e.P + ng * tfp::SA_GROUP;
                        attention_unit<32, PAIR>(e, g == 0 && u == 0, !(g == 1 && u == 1), G + tfp::SA_BQKV + u * 64, G + tfp::SA_BQKV + 128 + u * 64,
                                                 G + tfp::SA_QG, G + tfp::SA_QB, G + tfp::SA_KG, G + tfp::SA_KB, att_lo, att_span,
                                                 NG + tfp::SA_BQKV + nu * 64, NG + tfp::SA_BQKV + 128 + nu * 64, NG + tfp::SA_QG, NG + tfp::SA_QB, NG + tfp::SA_KG, NG + tfp::SA_KB);
                    }
                }
                wait_done(e, 0);                              // last projection of group 1 has landed
                {
                    const float* G = e.P + hf * tfp::SA_GROUP;
                    const RowStat sum = resid_update(e, G + tfp::SA_BPROJ, nullptr, nullptr);
                    float mean, rstd;
                    ln_stats<false>(e, sum, 0, mean, rstd);
                    ln_to_abuf<PAIR>(e, mean, rstd);
                    go(e);
                }
                param_release(e);
                param_acquire(e);                             // MLP stage
                for (int g = 0; g < 2; ++g) {
                    const float* G = e.P + g * tfp::SM_GROUP;
                    for (int q = 0; q < 4; ++q) {
                        // up-projection halves land on done[0]; before H1 is rewritten (q = 3) the down-projection of
                        // quarter 1 must have read it (done[1]) - see emit_mlp
                        if (q == 0 || q == 2) wait_done(e, 0);
                        if (q == 3) wait_done(e, 1);
                        fc_epilogue<PAIR>(e, q, G + tfp::SM_BFC + q * 128);
                        go(e);
                    }
                }
                wait_done(e, 0);                              // last down-projection of group 1
                {
                    const float* G = e.P + hf * tfp::SM_GROUP;
                    if (!last) {
                        const RowStat sum = resid_update(e, G + tfp::SM_BP2, tb1, nullptr);
                        float mean, rstd;
                        ln_stats<false>(e, sum, 0, mean, rstd);
                        ln_to_abuf<PAIR>(e, mean, rstd);
                    } else {
                        // stream junction: x = ln2_x(x + x_skip) | y = ln2_y(y + y_skip); z = cat(x, y) + time_expand(temb)
                        const RowStat sum = resid_update(e, G + tfp::SM_BP2, tb1, skipc);
                        float mean, rstd;
                        ln_stats<false>(e, sum, 0, mean, rstd);
                        const RowStat sum2 = ln_to_resid(e, mean, rstd, e.P + tfp::SM_LNN_G + hf * 128, e.P + tfp::SM_LNN_B + hf * 128, tb2);
                        ln_stats<true>(e, sum2, 1, mean, rstd);
                        ln_to_abuf<PAIR>(e, mean, rstd);
                    }
                    go(e);
                }
                param_release(e);
                mark(e);
            }

            // ================= main blocks: one 256-wide stream =================
            for (int blk = 0; blk < a.n_main; ++blk) {
                param_acquire(e);                             // attention stage
                const bool last = blk + 1 == a.n_main;
                for (int u = 0; u < 4; ++u)
                    attention_unit<64, PAIR>(e, u == 0, u < 3, e.P + tfp::BA_BQKV + u * 64, e.P + tfp::BA_BQKV + 256 + u * 64,
                                       e.P + tfp::BA_QG, e.P + tfp::BA_QB, e.P + tfp::BA_KG, e.P + tfp::BA_KB, att_lo, att_span,
                                       e.P + tfp::BA_BQKV + (u + 1) * 64, e.P + tfp::BA_BQKV + 256 + (u + 1) * 64, e.P + tfp::BA_QG, e.P + tfp::BA_QB,
                                       e.P + tfp::BA_KG, e.P + tfp::BA_KB);
                wait_done(e, 0);
                {
                    const RowStat sum = resid_update(e, e.P + tfp::BA_BPROJ + hf * 128, nullptr, nullptr);
                    float mean, rstd;
                    ln_stats<true>(e, sum, 0, mean, rstd);
                    ln_to_abuf<PAIR>(e, mean, rstd);
                    go(e);
                }
                param_release(e);
                param_acquire(e);                             // MLP stage
                for (int q = 0; q < 4; ++q) {
                    if (q == 0 || q == 2) wait_done(e, 0);
                    if (q == 3) wait_done(e, 1);
                    fc_epilogue<PAIR>(e, q, e.P + tfp::BM_BFC + q * 128);
                    go(e);
                }
                wait_done(e, 0);
                {
                    // not last: z += bias + temb, LayerNorm ln1 of the next block.
                    // last: ParticleFormer  x = ln3_x(x + x_skip) | y = ln3_y(y + y_skip);  Fused  z = ln2(z + z_skip)
                    const RowStat sum = resid_update(e, e.P + tfp::BM_BP2 + hf * 128, tb2, last ? skipc : nullptr);
                    float mean, rstd;
                    if (last && pf) ln_stats<false>(e, sum, 0, mean, rstd); else ln_stats<true>(e, sum, 0, mean, rstd);
                    ln_to_abuf<PAIR>(e, mean, rstd);
                    go(e);
                }
                param_release(e);
                mark(e);
            }

            // ================= heads + the hybrid step =================
            float outx[3] = {0.f, 0.f, 0.f};
            float outy[V];
#pragma unroll
            for (int v = 0; v < V; ++v) outy[v] = 0.f;
            param_acquire(e);                                 // head_x
            for (int q = 0; q < 4; ++q) {
                wait_done(e, q & 1);
                head_epilogue<3>(e, q, e.P + tfp::HX_BIAS + q * 128, e.P + tfp::HX_W2 + q * 128, 512, outx);
                go(e);
            }
            if (hf == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) outx[c] += e.P[tfp::HX_B2 + c];
            }
            param_release(e);
            for (int q = 4; q < 8; ++q) {
                param_acquire(e);                             // head_y, hidden units [(q-4)*128, +128)
                wait_done(e, q & 1);
                head_epilogue<V>(e, q, e.P + tfp::HY_BIAS, e.P + tfp::HY_W2, 128, outy);
                if (q < 6) go(e);
                if (q == 4 && hf == 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) outy[v] += e.P[tfp::HY_B2 + v];
                }
                param_release(e);
            }
            float* s_out = misc + mOut;
            if (hf == 1) {
#pragma unroll
                for (int c = 0; c < 3; ++c) s_out[r * 12 + c] = outx[c];
#pragma unroll
                for (int v = 0; v < V; ++v) s_out[r * 12 + 3 + v] = outy[v];
            }
            epi_bar();
            if (hf == 0 && r < nrows) {
#pragma unroll
                for (int c = 0; c < 3; ++c) outx[c] += s_out[r * 12 + c];
#pragma unroll
                for (int v = 0; v < V; ++v) outy[v] += s_out[r * 12 + 3 + v];
                if (a.vt_out) {                               // forward API: velocity and logits at the padded slots
#pragma unroll
                    for (int c = 0; c < 3; ++c) a.vt_out[slot * 3 + c] = outx[c];
#pragma unroll
                    for (int v = 0; v < V; ++v) a.logits_out[slot * V + v] = outy[v];
                } else {
                    float uu[V], rates[V];
                    if (a.st.u) {
#pragma unroll
                        for (int v = 0; v < V; ++v) uu[v] = __ldg(a.st.u + (static_cast<size_t>(step) * a.st.slots + slot) * V + v);
                    } else {
                        philox_uniforms(a.st.seed, a.st.slot0 + static_cast<uint64_t>(slot), static_cast<uint32_t>(step), V, uu);
                    }
                    const bool lastst = step + 1 == a.nsteps;
                    const bool want_rates = lastst && (a.st.rates_out != nullptr || a.st.argmax_last);
                    const float w = __ldg(a.st.thermo + step * 2), coef = __ldg(a.st.thermo + step * 2 + 1);
                    int kn = step_particle<V>(outy, s_ks[r], w, coef, a.st.sp, uu, want_rates ? rates : nullptr);
                    if (a.st.forced) kn = a.st.forced[static_cast<size_t>(step) * a.st.slots + slot];
                    if (lastst && a.st.rates_out) {
#pragma unroll
                        for (int v = 0; v < V; ++v) a.st.rates_out[slot * V + v] = rates[v];
                    }
                    if (lastst && a.st.argmax_last) {         // use_final_max_rates (reference model/MMF.py:193-196)
                        int best = 0;
#pragma unroll
                        for (int v = 1; v < V; ++v) best = rates[v] > rates[best] ? v : best;
                        kn = best;
                    }
                    s_ks[r] = kn;
#pragma unroll
                    for (int c = 0; c < 3; ++c) s_xs[r * 3 + c] = euler_update(s_xs[r * 3 + c], outx[c], a.st.sp.dt);
                }
            }
            epi_bar();
            mark(e);
        }
        if (a.x_out && tid < nrows) {
            const long long sl = a.row_slot[static_cast<size_t>(tile) * 128 + tid];
#pragma unroll
            for (int c = 0; c < 3; ++c) a.x_out[sl * 3 + c] = s_xs[tid * 3 + c];
            a.k_out[sl] = s_ks[tid];
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // no CTA leaves while a peer may still multicast into it
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

#if MMF_TILE_TRACE
static unsigned long long* g_dbg_host = nullptr;
void tf_tiles_dump_timeouts() {     // after a failed launch: which barrier waits of CTA 0 timed out (see mbar_wait)
    if (!g_dbg_host) return;
    const unsigned n = static_cast<unsigned>(g_dbg_host[0]);
    fprintf(stderr, "tile kernel: %u timed-out barrier waits (first 62; tag: epilogue = stamps passed, issuer = op index, 0x8000 | op = weight tile)\n", n);
    for (unsigned i = 0; i < n && i < 62; ++i) {
        const unsigned long long v = g_dbg_host[1 + i];
        fprintf(stderr, "  cta %llu barrier smem+0x%llx parity %llu warp %llu tag %llu\n", (v >> 12) & 0xffff, (v >> 32) & 0xffff, (v >> 28) & 1, (v & 0xfff) >> 5, v >> 48);
    }
}
int launch_tf_tiles_trace(const TfLaunch& a, int n_tiles, int cluster, bool pair, cudaStream_t stream) {
    if (!g_dbg_host) {
        unsigned long long* dptr = nullptr;
        MMF_CUDA_OK(cudaHostAlloc(&g_dbg_host, 64 * 8, cudaHostAllocMapped));
        memset(g_dbg_host, 0, 64 * 8);
        MMF_CUDA_OK(cudaHostGetDevicePointer(&dptr, g_dbg_host, 0));
        MMF_CUDA_OK(cudaMemcpyToSymbol(mmf_dbg_sink, &dptr, sizeof(dptr)));
    }
#else
int tf_tile_smem_bytes() { return smem_bytes<false>(); }

#if MMF_PROD_DIAG
static unsigned long long* g_diag_host = nullptr;
static void diag_dump() {
    if (!g_diag_host || !g_diag_host[0]) return;
    const unsigned n = static_cast<unsigned>(g_diag_host[0]);
    fprintf(stderr, "tile kernel (production build): %u timed-out barrier waits\n", n);
    for (unsigned i = 0; i < n && i < 62; ++i) {
        const unsigned long long v = g_diag_host[1 + i];
        fprintf(stderr, "  cta %llu barrier smem+0x%llx parity %llu warp %llu lane0 tag %llu\n", (v >> 12) & 0xffff, (v >> 32) & 0xffff, (v >> 28) & 1, (v & 0xfff) >> 5, v >> 48);
    }
}
#endif

int launch_tf_tiles(const TfLaunch& a, int n_tiles, int cluster, bool pair, cudaStream_t stream) {
#if MMF_PROD_DIAG
    if (!g_diag_host) {
        unsigned long long* dptr = nullptr;
        MMF_CUDA_OK(cudaHostAlloc(&g_diag_host, 64 * 8, cudaHostAllocMapped));
        memset(g_diag_host, 0, 64 * 8);
        MMF_CUDA_OK(cudaHostGetDevicePointer(&dptr, g_diag_host, 0));
        MMF_CUDA_OK(cudaMemcpyToSymbol(mmf_dbg_sink, &dptr, sizeof(dptr)));
        atexit(diag_dump);
    }
#endif
#endif
    if (n_tiles == 0) return 0;
    MMF_REQUIRE(a.vocab == 9, "the tile kernel is instantiated for vocab_size 9");
    MMF_REQUIRE((cluster == 1 || cluster == 2 || cluster == 4) && n_tiles % cluster == 0, "tile launch: bad cluster size");
    MMF_REQUIRE(!pair || cluster == 2, "pair tiles run as 2-CTA clusters");
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        MMF_CUDA_OK(cudaFuncSetAttribute(tf_tile_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<false>()));
        MMF_CUDA_OK(cudaFuncSetAttribute(tf_tile_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<true>()));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pair ? smem_bytes<true>() : smem_bytes<false>();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (pair) MMF_CUDA_OK(cudaLaunchKernelEx(&cfg, tf_tile_kernel<9, true>, a, *a.optab, *a.prodtab));
    else MMF_CUDA_OK(cudaLaunchKernelEx(&cfg, tf_tile_kernel<9, false>, a, *a.optab, *a.prodtab));
    return 0;
}

}  // namespace mmf
