// CUDA-core kernels of the hot path: the fused hybrid step, packing, the K=3 input embedding, row
// LayerNorms that sit between streams, and the tiny output projections fused with the step.
// All are memory-bound: one coalesced, vectorised pass over the data.
#include <cstdlib>
#include "mmf_internal.h"
#include "mmf_simt.h"
#include "mmf_ptx.cuh"

namespace mmf {

namespace {

__device__ __forceinline__ float gelu_erf_simt(float x) {
    const float z = fabsf(x) * 0.70710678f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = 1.0f - p * t * exp2f(-1.44269504f * z * z);
    return 0.5f * x * (1.0f + copysignf(e, x));
}

// ---------------------------------------------------------------------------------------------
// 1. fused hybrid step on the padded (B, D) layout  -- north_star kernel (2)
//    reads vt(12) logits(4V) x(12) k(8) [u(4V)]  writes x(12) k(8) [rates(4V)]  bytes per particle
// ---------------------------------------------------------------------------------------------
constexpr int kStepThreads = 256;

// cooperative, coalesced copy of `count` floats between global and shared memory (float4 when aligned)
__device__ __forceinline__ void block_load(float* s, const float* g, int count, int valid) {
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0 && valid == count) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* s4 = reinterpret_cast<float4*>(s);
        for (int i = threadIdx.x; i < count / 4; i += blockDim.x) s4[i] = __ldcs(g4 + i);
        for (int i = (count / 4) * 4 + threadIdx.x; i < count; i += blockDim.x) s[i] = __ldcs(g + i);
    } else {
        for (int i = threadIdx.x; i < valid; i += blockDim.x) s[i] = __ldcs(g + i);
    }
}
__device__ __forceinline__ void block_store(float* g, const float* s, int count, int valid) {
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0 && valid == count) {
        float4* g4 = reinterpret_cast<float4*>(g);
        const float4* s4 = reinterpret_cast<const float4*>(s);
        for (int i = threadIdx.x; i < count / 4; i += blockDim.x) __stcs(g4 + i, s4[i]);
        for (int i = (count / 4) * 4 + threadIdx.x; i < count; i += blockDim.x) __stcs(g + i, s[i]);
    } else {
        for (int i = threadIdx.x; i < valid; i += blockDim.x) __stcs(g + i, s[i]);
    }
}

// FAST: production mode (in-kernel Philox draws, no rates returned) - MUFU exp / division and the two-uniform form of the
// jump law (StepMath / step_particle_2u in step_math.cuh); the parity modes keep the reproducible per-channel arithmetic
template <int V, bool FAST>
__global__ void __launch_bounds__(kStepThreads)
hybrid_step_kernel(const float* __restrict__ vt, const float* __restrict__ logits, float* __restrict__ x,
                   long long* __restrict__ k, const float* __restrict__ t, long long n_particles, int D,
                   const StepLaunch sl, float* __restrict__ rates_out) {
    __shared__ __align__(16) float s_lg[kStepThreads * V];      // logits in, rates out
    __shared__ __align__(16) float s_u[kStepThreads * V];
    __shared__ __align__(16) float s_x[kStepThreads * 3];
    __shared__ __align__(16) float s_v[kStepThreads * 3];

    const long long base = static_cast<long long>(blockIdx.x) * kStepThreads;
    const long long remain = n_particles - base;
    const int nval = remain < kStepThreads ? static_cast<int>(remain) : kStepThreads;
    block_load(s_lg, logits + base * V, kStepThreads * V, nval * V);
    if (sl.u) block_load(s_u, sl.u + base * V, kStepThreads * V, nval * V);
    block_load(s_x, x + base * 3, kStepThreads * 3, nval * 3);
    block_load(s_v, vt + base * 3, kStepThreads * 3, nval * 3);
    __syncthreads();

    const int tid = threadIdx.x;
    const long long i = base + tid;
    if (tid < nval) {
        float lg[V], u[V], rates[V];
#pragma unroll
        for (int v = 0; v < V; ++v) lg[v] = s_lg[tid * V + v];
        long long kc = k[i];
        if (kc < 0 || kc >= V) { atomicOr(sl.err_flag, 2); kc = 0; }
        // jet of this particle: 32-bit division whenever the slot index fits (a 64-bit one costs ~100 instructions)
        const long long jet = n_particles < 0x7fffffffLL ? static_cast<long long>(static_cast<unsigned>(i) / static_cast<unsigned>(D)) : i / D;
        float w, coef;
        if (FAST) {
            const float a = static_cast<float>(-static_cast<double>(V) * static_cast<double>(sl.sp.beta));
            w = StepMath<true>::exp(a * (1.0f - __ldg(t + jet)));
            coef = __fdividef(w * static_cast<float>(V), 1.0f - w);
        } else {
            det_thermostat(__ldg(t + jet), sl.sp.beta, V, &w, &coef);
        }
        int kn;
        if (FAST) {                                   // production mode: one Philox block, two uniforms (step_particle_2u)
            const uint64_t slot = sl.slot0 + static_cast<uint64_t>(i);
            const Philox4 r = philox4x32_10(Philox4{static_cast<uint32_t>(slot), static_cast<uint32_t>(slot >> 32), sl.step, 0x32u},
                                            static_cast<uint32_t>(sl.seed), static_cast<uint32_t>(sl.seed >> 32));
            kn = step_particle_2u<V>(lg, static_cast<int>(kc), w, coef, sl.sp, u01_from_bits(r.x), u01_from_bits(r.y));
        } else {
            if (sl.u) {
#pragma unroll
                for (int v = 0; v < V; ++v) u[v] = s_u[tid * V + v];
            } else {
                philox_uniforms(sl.seed, sl.slot0 + static_cast<uint64_t>(i), sl.step, V, u);
            }
            kn = sl.sp.method == 1 ? step_particle_euler<V>(lg, static_cast<int>(kc), w, coef, sl.sp, u[0], rates_out ? rates : nullptr)
                                   : step_particle<V, false>(lg, static_cast<int>(kc), w, coef, sl.sp, u, rates_out ? rates : nullptr);
        }
        k[i] = kn;
#pragma unroll
        for (int c = 0; c < 3; ++c) s_x[tid * 3 + c] = euler_update(s_x[tid * 3 + c], s_v[tid * 3 + c], sl.sp.dt);
        if (rates_out) {
#pragma unroll
            for (int v = 0; v < V; ++v) s_lg[tid * V + v] = rates[v];
        }
    }
    __syncthreads();
    block_store(x + base * 3, s_x, kStepThreads * 3, nval * 3);
    if (rates_out) block_store(rates_out + base * V, s_lg, kStepThreads * V, nval * V);
}

// ---------------------------------------------------------------------------------------------
// 1b. production mode of the step (in-kernel Philox draws, no rates returned): the HBM-bound form.
//     Persistent CTAs walk 512-particle chunks; a chunk's logits / x / vt / k arrive in shared memory by four 1-D bulk
//     copies (cp.async.bulk + mbarrier) issued one chunk ahead into the other of two stages, and x / k leave by two bulk
//     stores - the SM issues no global load or store instruction on this path, so the ~170 instructions per particle
//     that remain (softmax, thermostat, jump law, Philox) fit under the HBM time of 88 bytes per particle.
//     A thread owns the two ADJACENT particles (2 tid, 2 tid + 1) of the chunk: their 18 logits are nine conflict-free
//     8-byte shared loads (V odd), and one Philox4x32-10 block - keyed on (seed, global slot >> 1, step) - yields the two
//     uniforms of both (x, y for the even slot, z, w for the odd one), independent of launch geometry and sharding.
//     Arithmetic: StepMath<true> (MUFU ex2 / rcp) and the two-uniform jump law of step_particle_2u.
// ---------------------------------------------------------------------------------------------
constexpr int kProdThreads = 256, kProdChunk = 2 * kProdThreads, kProdStages = 2;
template <int V> constexpr int prod_stage_bytes() { return kProdChunk * (8 + 4 * V + 12 + 12); }
template <int V> constexpr int prod_smem_bytes() { return kProdStages * prod_stage_bytes<V>() + 64; }

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct ProdConst {
    float sc;            // log2(e) / temperature: p_v = 2^(sc (l_v - max l))
    float a2;            // -V beta log2(e): w = 2^(a2 (1 - t))
    int filters;         // top-k or top-p active
    unsigned long long div_magic;   // floor(2^64 / D) + 1: jet = umul64hi(slot, magic), exact below 2^64 / D
};

// one particle: logits in registers (lg, also at s_lg for the indexed read of l_k) -> new token
template <int V>
__device__ __forceinline__ int prod_particle(const float* lg, const float* s_lg, int k, float t, const StepParams& sp, const ProdConst& pc,
                                             float u1, float u2) {
    float m = lg[0];
#pragma unroll
    for (int v = 1; v < V; ++v) m = fmaxf(m, lg[v]);
    const float nm = -m * pc.sc;
    float p[V];
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        p[v] = ex2_approx(fmaf(lg[v], pc.sc, nm));
        s += p[v];
    }
    const float inv = rcp_approx(s);
    float qk;
    if (pc.filters) {
#pragma unroll
        for (int v = 0; v < V; ++v) p[v] *= inv;
        step_filters<V, true>(p, sp.top_k, sp.top_p);
        qk = p[0];
#pragma unroll
        for (int v = 1; v < V; ++v) qk = (k == v) ? p[v] : qk;
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) p[v] *= inv;
        qk = ex2_approx(fmaf(s_lg[k], pc.sc, nm)) * inv;
    }
    const float w = ex2_approx(pc.a2 * (1.0f - t));
    const float coef = w * static_cast<float>(V) * rcp_approx(1.0f - w);
    // lam_v = dt (1 + coef p_v + w q_k); sum_v p_v = 1, so L = V dt (1 + w q_k) + dt coef
    const float bdt = fmaf(w, qk, 1.0f) * sp.dt, cdt = coef * sp.dt;
    const float L = fmaf(static_cast<float>(V), bdt, cdt);
    if (!(u1 < L * ex2_approx(-1.44269504f * L))) return k;     // zero or >= 2 events: the token stays (model/solvers.py:49-54)
    const float target = u2 * L;
    float c = 0.0f;
    int j = 0;
#pragma unroll
    for (int v = 0; v < V - 1; ++v) {
        c += fmaf(cdt, p[v], bdt);
        j += (target >= c) ? 1 : 0;
    }
    return j;
}

template <int V>
__global__ void __launch_bounds__(kProdThreads)
hybrid_step_prod_kernel(const float* __restrict__ vt, const float* __restrict__ logits, float* __restrict__ x, long long* __restrict__ k,
                        const float* __restrict__ t, long long n_particles, int D, const StepLaunch sl, const ProdConst pc, int bulk_ok) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int kStage = prod_stage_bytes<V>();
    constexpr uint32_t kBytesK = kProdChunk * 8, kBytesL = kProdChunk * 4 * V, kBytesX = kProdChunk * 12;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kProdStages * kStage);
    const int tid = threadIdx.x;
    const long long n_chunks = (n_particles + kProdChunk - 1) / kProdChunk;
    if (tid == 0) {
        for (int i = 0; i < kProdStages; ++i) mbar_init(&full[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto stage_k = [&](int st) { return reinterpret_cast<long long*>(smem_raw + st * kStage); };
    auto stage_lg = [&](int st) { return reinterpret_cast<float*>(smem_raw + st * kStage + kBytesK); };
    auto stage_x = [&](int st) { return reinterpret_cast<float*>(smem_raw + st * kStage + kBytesK + kBytesL); };
    auto stage_v = [&](int st) { return reinterpret_cast<float*>(smem_raw + st * kStage + kBytesK + kBytesL + kBytesX); };
    // a chunk travels by bulk copies when it is complete and every array is 16-byte aligned (else plain loads / stores)
    auto is_bulk = [&](long long c) { return bulk_ok && c < n_chunks && (c + 1) * kProdChunk <= n_particles; };
    auto issue_load = [&](long long c, int st) {          // one thread
        const long long b = c * kProdChunk;
        mbar_expect_tx(&full[st], kBytesK + kBytesL + 2 * kBytesX);
        bulk_load_1d(stage_k(st), k + b, kBytesK, &full[st]);
        bulk_load_1d(stage_lg(st), logits + b * V, kBytesL, &full[st]);
        bulk_load_1d(stage_x(st), x + b * 3, kBytesX, &full[st]);
        bulk_load_1d(stage_v(st), vt + b * 3, kBytesX, &full[st]);
    };

    uint32_t phase0 = 0, phase1 = 0;
    long long chunk = blockIdx.x;
    if (tid == 0 && is_bulk(chunk)) issue_load(chunk, 0);
    for (int it = 0; chunk < n_chunks; ++it, chunk += gridDim.x) {
        const int st = it & 1;
        const long long next = chunk + gridDim.x;
        if (tid == 0) {
            tma_store_wait_read<0>();                     // the stores of the previous chunk have read the other stage
            if (is_bulk(next)) issue_load(next, st ^ 1);
        }
        const long long base = chunk * kProdChunk;
        const long long remain = n_particles - base;
        const int nval = remain < kProdChunk ? static_cast<int>(remain) : kProdChunk;
        long long* s_k = stage_k(st);
        float* s_lg = stage_lg(st);
        float* s_x = stage_x(st);
        float* s_v = stage_v(st);
        const bool bulk = is_bulk(chunk);
        // jets of this thread's pair and their times: fetched before the wait, so the L2 latency hides under the bulk copies
        const int p0 = 2 * tid;
        const unsigned long long ia = static_cast<unsigned long long>(base + p0);
        float ta = 0.0f, tb = 0.0f;
        if (p0 < nval) {
            const unsigned long long ja = D == 1 ? ia : __umul64hi(ia, pc.div_magic);
            const unsigned long long jb = ja + ((ia - ja * static_cast<unsigned long long>(D) + 1ull == static_cast<unsigned long long>(D)) ? 1ull : 0ull);
            ta = __ldg(t + ja);
            tb = (p0 + 1 < nval && jb != ja) ? __ldg(t + jb) : ta;
        }
        if (bulk) {
            mbar_wait(&full[st], st ? phase1 : phase0);
            if (st) phase1 ^= 1; else phase0 ^= 1;
        } else {
            for (int i = tid; i < nval; i += kProdThreads) s_k[i] = k[base + i];
            for (int i = tid; i < nval * V; i += kProdThreads) s_lg[i] = logits[base * V + i];
            for (int i = tid; i < nval * 3; i += kProdThreads) { s_x[i] = x[base * 3 + i]; s_v[i] = vt[base * 3 + i]; }
            __syncthreads();
        }

        if (p0 < nval) {
            float lg[2 * V], xs[6], vs[6];
            const float2* lg2 = reinterpret_cast<const float2*>(s_lg + 2 * V * tid);
#pragma unroll
            for (int v = 0; v < V; ++v) { const float2 q = lg2[v]; lg[2 * v] = q.x; lg[2 * v + 1] = q.y; }
            const longlong2 kk = *reinterpret_cast<const longlong2*>(s_k + p0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 q = reinterpret_cast<const float2*>(s_x + 6 * tid)[c], r = reinterpret_cast<const float2*>(s_v + 6 * tid)[c];
                xs[2 * c] = q.x; xs[2 * c + 1] = q.y; vs[2 * c] = r.x; vs[2 * c + 1] = r.y;
            }
            long long ka = kk.x, kb = kk.y;
            const bool has_b = p0 + 1 < nval;
            if (ka < 0 || ka >= V) { atomicOr(sl.err_flag, 2); ka = 0; }
            if (has_b && (kb < 0 || kb >= V)) { atomicOr(sl.err_flag, 2); kb = 0; }
            if (!has_b) kb = 0;
            const uint64_t ga = sl.slot0 + ia;
            const uint32_t k0 = static_cast<uint32_t>(sl.seed), k1 = static_cast<uint32_t>(sl.seed >> 32);
            float ua1, ua2, ub1, ub2;
            {
                const uint64_t pa = ga >> 1;
                const Philox4 r = philox4x32_10(Philox4{static_cast<uint32_t>(pa), static_cast<uint32_t>(pa >> 32), sl.step, 0x32u}, k0, k1);
                if ((ga & 1ull) == 0) {
                    ua1 = u01_from_bits(r.x); ua2 = u01_from_bits(r.y); ub1 = u01_from_bits(r.z); ub2 = u01_from_bits(r.w);
                } else {                                  // odd first slot of the launch: the pair straddles two blocks
                    ua1 = u01_from_bits(r.z); ua2 = u01_from_bits(r.w);
                    const uint64_t pb = pa + 1;
                    const Philox4 q = philox4x32_10(Philox4{static_cast<uint32_t>(pb), static_cast<uint32_t>(pb >> 32), sl.step, 0x32u}, k0, k1);
                    ub1 = u01_from_bits(q.x); ub2 = u01_from_bits(q.y);
                }
            }
            const int na = prod_particle<V>(lg, s_lg + 2 * V * tid, static_cast<int>(ka), ta, sl.sp, pc, ua1, ua2);
            const int nb = prod_particle<V>(lg + V, s_lg + 2 * V * tid + V, static_cast<int>(kb), tb, sl.sp, pc, ub1, ub2);
            *reinterpret_cast<longlong2*>(s_k + p0) = make_longlong2(na, has_b ? nb : kk.y);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                reinterpret_cast<float2*>(s_x + 6 * tid)[c] = make_float2(euler_update(xs[2 * c], vs[2 * c], sl.sp.dt), euler_update(xs[2 * c + 1], vs[2 * c + 1], sl.sp.dt));
        }
        if (bulk) {
            fence_proxy_async();                          // the generic-proxy writes above, before the async-proxy reads below
            __syncthreads();
            if (tid == 0) {
                bulk_store_1d(k + base, s_k, kBytesK);
                bulk_store_1d(x + base * 3, s_x, kBytesX);
                tma_store_commit();
            }
        } else {
            __syncthreads();
            for (int i = tid; i < nval; i += kProdThreads) k[base + i] = s_k[i];
            for (int i = tid; i < nval * 3; i += kProdThreads) x[base * 3 + i] = s_x[i];
            __syncthreads();                              // (a later chunk of this CTA may reuse the stage)
        }
    }
    if (tid == 0) tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------
// 1c. jet observables of a sample, one fused pass: de-standardise (utils/callbacks.py:52-56), particle
//     four-momenta and charges (utils/aoj.py:333-368), per-jet sums -> pt, m, eta, phi, charge, jet charge (aoj.py:452-471,
//     514-521), token counts (utils/metrics.py:10-33).  Half a warp per jet, lanes stride over the D slots; padded slots cost
//     their 8 mask bytes only.  Sums are carried in fp64 (the mass is a difference of squares of the sums).
//     HBM-bound: 8 B per slot + 20 B per real particle in, 48 + 4 V bytes per jet out.
// ---------------------------------------------------------------------------------------------
constexpr int kObsLanes = 16;                                 // lanes per jet: a warp works on two jets at once
__device__ __forceinline__ double group_sum(double v) {      // sum over the kObsLanes lanes of this thread's jet
#pragma unroll
    for (int o = kObsLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256, 4)
jet_observables_kernel(const float* __restrict__ x, const long long* __restrict__ k, const long long* __restrict__ mask, const ObsArgs a) {
    // half a warp per jet: the per-jet tail (reductions, token counts, the transcendental finish) is half of the instruction
    // stream at the typical 55 particles per jet, and two jets share every one of those instructions
    const int lane = threadIdx.x & (kObsLanes - 1);
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);     // the lanes of this jet
    const long long jet = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / kObsLanes;
    const bool live = jet < a.B;                               // (an odd last jet leaves half a warp idle, but in step)
    const long long base = (live ? jet : 0) * a.D;
    double px = 0.0, py = 0.0, pz = 0.0, E = 0.0, qpt = 0.0;
    int n = 0, q = 0;
    unsigned long long c_lo = 0ull, c_hi = 0ull;              // 8-bit counters of tokens 0..7 / 8..15 (a lane sees <= 255 slots)
    // kObsUnroll slots per lane and pass (80 slots of the jet): all mask loads of a pass are issued together, then the x / k
    // loads of the unmasked slots, then the arithmetic - two memory round trips per pass.  (Measured alternatives: ten per
    // pass - one pass at D = 150 - spills at 64 registers, 2x slower; masks folded into a bit word and 32-bit tokens, 20 % slower.)
    constexpr int kObsUnroll = 5;
    const int D = live ? a.D : 0;
    for (int d0 = lane; d0 < D; d0 += kObsLanes * kObsUnroll) {
        long long mk[kObsUnroll], tk[kObsUnroll];
        float xv[kObsUnroll][3];
#pragma unroll
        for (int u = 0; u < kObsUnroll; ++u) {
            const int d = d0 + kObsLanes * u;
            mk[u] = d < D ? __ldcs(mask + base + d) : 0;      // mask_bool = mask > 0 (aoj.py:336)
        }
#pragma unroll
        for (int u = 0; u < kObsUnroll; ++u) {
            const long long s = base + d0 + kObsLanes * u;
            xv[u][0] = xv[u][1] = xv[u][2] = 0.0f;
            tk[u] = -1;
            if (mk[u] > 0) {
                xv[u][0] = __ldcs(x + s * 3); xv[u][1] = __ldcs(x + s * 3 + 1); xv[u][2] = __ldcs(x + s * 3 + 2);
                if (k) tk[u] = __ldcs(k + s);
            }
        }
#pragma unroll
        for (int u = 0; u < kObsUnroll; ++u) {
            if (mk[u] <= 0) continue;
            const float pt = det_add(det_mul(xv[u][0], a.std[0]), a.mean[0]);          // (x * sig) + mu, rounded like the reference
            const float eta = det_add(det_mul(xv[u][1], a.std[1]), a.mean[1]);
            const float phi = det_add(det_mul(xv[u][2], a.std[2]), a.mean[2]);
            float sn, cs;
            sincosf(phi, &sn, &cs);
            // sinh / cosh from one exponential: absolute error ~1e-7 cosh(eta), i.e. ~1e-7 of this particle's energy
            const float ex = expf(eta), exi = __frcp_rn(ex);
            px += static_cast<double>(pt * cs);
            py += static_cast<double>(pt * sn);
            pz += static_cast<double>(pt * (0.5f * (ex - exi)));
            E += static_cast<double>(pt * (0.5f * (ex + exi)));
            ++n;
            const long long tok = tk[u];
            if (tok >= 0 && tok < 16) {
                const unsigned sh = 8u * (static_cast<unsigned>(tok) & 7u);
                if (tok < 8) c_lo += 1ull << sh; else c_hi += 1ull << sh;
                // charge (aoj.py:358-368): two bits per token, code 0 -> -1 (tokens 3, 5, 7), 1 -> 0, 2 -> +1 (tokens 4, 6, 8)
                const int ch = static_cast<int>((0x55562215u >> (2u * static_cast<unsigned>(tok))) & 3u) - 1;
                q += ch;
                qpt += static_cast<double>(static_cast<float>(ch) * pt);
            }
        }
    }
    px = group_sum(px); py = group_sum(py); pz = group_sum(pz); E = group_sum(E); qpt = group_sum(qpt);
    n = __reduce_add_sync(gmask, n);
    q = __reduce_add_sync(gmask, q);
    if (a.counts) {
        int mine_tot = 0;
        for (int v = 0; v < a.V; ++v) {
            const int mine = static_cast<int>(((v < 8 ? c_lo : c_hi) >> (8 * (v & 7))) & 0xffull);
            const int tot = __reduce_add_sync(gmask, mine);
            mine_tot = (lane == v) ? tot : mine_tot;
        }
        if (live && lane < a.V) a.counts[jet * a.V + lane] = mine_tot;    // one coalesced store per jet (V <= 16 = kObsLanes)
    }
    if (live && lane == 0) {
        // sums and the mass-squared difference in fp64; the transcendental tail in fp32 (the reference's own precision)
        const double pt2 = px * px + py * py;
        const double m2 = E * E - pt2 - pz * pz;
        const float ptf = sqrtf(static_cast<float>(pt2));
        const double ptj = static_cast<double>(ptf);
        float* o = a.kin + jet * kObsKin;
        o[0] = static_cast<float>(px); o[1] = static_cast<float>(py); o[2] = static_cast<float>(pz); o[3] = static_cast<float>(E);
        o[4] = ptf;
        o[5] = sqrtf(static_cast<float>(m2));
        o[6] = 0.5f * logf(static_cast<float>(ptj + pz) / static_cast<float>(ptj - pz));
        o[7] = atan2f(static_cast<float>(py), static_cast<float>(px));
        o[8] = static_cast<float>(q);
        o[9] = static_cast<float>(qpt) / ptf;
        o[10] = static_cast<float>(n);
        o[11] = static_cast<float>(m2);
    }
}

// ---------------------------------------------------------------------------------------------
// 1d. source state of the sampler, built on the device: multiplicity ~ Categorical(empirical histogram)
//     with prefix masks (utils/aoj.py:875-890), x0 = N(0,1) * mask, k0 = U{1..V-1} * mask (scripts/sample_mmf.py:82-84).
//     Counter-based: every draw is a function of (seed, GLOBAL jet, slot) only, so the sample does not depend on batch size
//     or on how jets are sharded over GPUs, and oracle/source_oracle.py reproduces masks and tokens bit for bit.
//       jet J:   Philox4x32-10(ctr = (J lo, J hi, 0, 'MULT'), key = seed) -> u = (x >> 8) 2^-24 -> n = #{m : cdf[m] <= u}
//       slot S = J D + d:  Philox(ctr = (S lo, S hi, 0, 'SRCE')) -> u_i = ((w_i >> 9) + 0.5) 2^-23;
//                z0, z1 = sqrt(-2 ln u_x) (cos, sin)(2 pi u_y), z2 = sqrt(-2 ln u_z) cos(2 pi u_w);
//                token = 1 + mulhi(low bytes of the four words, V - 1)
//     Write-bound: 28 B per slot (x 12, k 8, mask 8).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kSrcTagMult = 0x4d554c54u, kSrcTagSlot = 0x53524345u;

__global__ void source_mult_kernel(const SourceArgs a, int* __restrict__ n_out) {
    const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const uint64_t J = a.first_jet + static_cast<uint64_t>(b);
    const Philox4 r = philox4x32_10(Philox4{static_cast<uint32_t>(J), static_cast<uint32_t>(J >> 32), 0u, kSrcTagMult},
                                    static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32));
    const float u = u01_from_bits(r.x);
    int n = 0;
    for (int m = 0; m < a.D; ++m) n += (a.cdf[m] <= u) ? 1 : 0;      // cdf[m] = P(multiplicity <= m); cdf[D] = 1 > u
    n_out[b] = n;
}

__global__ void __launch_bounds__(256)
source_fill_kernel(const SourceArgs a, const int* __restrict__ n_in, float* __restrict__ x0, long long* __restrict__ k0,
                   long long* __restrict__ mask) {
    __shared__ __align__(16) float s_x[256 * 3];
    const long long total = a.B * a.D;
    const long long i0 = static_cast<long long>(blockIdx.x) * 256, i = i0 + threadIdx.x;
    const bool valid = i < total;
    float z0 = 0.0f, z1 = 0.0f, z2 = 0.0f;
    long long tok = 0;
    bool real = false;
    if (valid) {
        const unsigned long long b = a.D == 1 ? static_cast<unsigned long long>(i) : __umul64hi(static_cast<unsigned long long>(i), a.div_magic);
        const int d = static_cast<int>(i - static_cast<long long>(b) * a.D);
        real = d < __ldg(n_in + b);
        if (real) {
            const uint64_t S = (a.first_jet + b) * static_cast<uint64_t>(a.D) + static_cast<uint64_t>(d);
            const Philox4 r = philox4x32_10(Philox4{static_cast<uint32_t>(S), static_cast<uint32_t>(S >> 32), 0u, kSrcTagSlot},
                                            static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32));
            // 23-bit uniforms strictly inside (0,1): (w >> 9) + 0.5 is exact in binary32
            const float ux = (static_cast<float>(r.x >> 9) + 0.5f) * 1.1920929e-07f, uy = (static_cast<float>(r.y >> 9) + 0.5f) * 1.1920929e-07f;
            const float uz = (static_cast<float>(r.z >> 9) + 0.5f) * 1.1920929e-07f, uw = (static_cast<float>(r.w >> 9) + 0.5f) * 1.1920929e-07f;
            const float ra = sqrtf(-2.0f * logf(ux)), rb = sqrtf(-2.0f * logf(uz));
            float sn, cs, sn2, cs2;
            sincospif(2.0f * uy, &sn, &cs);
            sincospif(2.0f * uw, &sn2, &cs2);
            z0 = ra * cs; z1 = ra * sn; z2 = rb * cs2;
            const uint32_t bits = (r.x & 0xffu) | ((r.y & 0xffu) << 8) | ((r.z & 0xffu) << 16) | ((r.w & 0xffu) << 24);
            tok = 1 + static_cast<long long>(__umulhi(bits, static_cast<uint32_t>(a.V - 1)));
        }
        if (k0) __stcs(k0 + i, tok);
        __stcs(mask + i, real ? 1ll : 0ll);
    }
    // x0: the block's 768 floats leave as 192 16-byte stores (full blocks of an aligned array), else element-wise
    const bool vec = i0 + 256 <= total && (reinterpret_cast<uintptr_t>(x0) & 15u) == 0;
    if (vec) {
        s_x[threadIdx.x * 3] = z0; s_x[threadIdx.x * 3 + 1] = z1; s_x[threadIdx.x * 3 + 2] = z2;
        __syncthreads();
        if (threadIdx.x < 192) __stcs(reinterpret_cast<float4*>(x0 + i0 * 3) + threadIdx.x, reinterpret_cast<const float4*>(s_x)[threadIdx.x]);
    } else if (valid) {
        __stcs(x0 + i * 3, z0); __stcs(x0 + i * 3 + 1, z1); __stcs(x0 + i * 3 + 2, z2);
    }
}

__global__ void euler_kernel(const float* __restrict__ vt, float* __restrict__ x, float dt, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) x[i] = euler_update(x[i], vt[i], dt);
}

// ---------------------------------------------------------------------------------------------
// 2. packing: (B, D, .) with prefix-or-arbitrary masks  <->  [rows, .] over real particles only
// ---------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ x0, const long long* __restrict__ k0,
                            const int* __restrict__ row_slot, int rows, int V, float* __restrict__ xs,
                            int* __restrict__ ks, int* err_flag) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const long long s = row_slot[r];
    if (s < 0) {                       // padding row of a tile-aligned layout
        xs[r * 3 + 0] = 0.f; xs[r * 3 + 1] = 0.f; xs[r * 3 + 2] = 0.f;
        if (ks) ks[r] = 0;
        return;
    }
    xs[r * 3 + 0] = x0[s * 3 + 0];
    xs[r * 3 + 1] = x0[s * 3 + 1];
    xs[r * 3 + 2] = x0[s * 3 + 2];
    int kk = 0;
    if (k0) {
        const long long kv = k0[s];
        if (kv < 0 || kv >= V) atomicOr(err_flag, 2); else kk = static_cast<int>(kv);
    }
    if (ks) ks[r] = kk;
}

__global__ void unpack_kernel(const float* __restrict__ xs, const int* __restrict__ ks,
                              const int* __restrict__ row_slot, int rows, float* __restrict__ x_out,
                              long long* __restrict__ k_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const long long s = row_slot[r];
    x_out[s * 3 + 0] = xs[r * 3 + 0];
    x_out[s * 3 + 1] = xs[r * 3 + 1];
    x_out[s * 3 + 2] = xs[r * 3 + 2];
    if (k_out) k_out[s] = ks[r];
}

// ---------------------------------------------------------------------------------------------
// 2b. the generated sample as ONE narrow record per jet: what FlowGeneratorCallback does on the host after the run
//     (reference utils/callbacks.py:52-57: continuous * std + mean, then apply_mask) fused with the narrowing of the int64
//     token / mask tensors to one byte per slot.  Record of jet b (R = round_up(13 D, 16) bytes):
//         [D][3] fp32 de-standardised kinematics, zero at padded slots | [D] uint8 token | mask << 7 | zero padding
//     The records of a shard are what the single end-of-run collective moves (13 B instead of 28 B per slot) and what the
//     writer lays out as generated_sample.h5.  HBM: 28 B read + 13 B written per slot.
// ---------------------------------------------------------------------------------------------
struct SampleNorm { float mean[3], std[3]; };

__global__ void sample_pack_kernel(const float* __restrict__ x, const long long* __restrict__ k, const long long* __restrict__ mask,
                                   SampleNorm nm, long long slots, int D, int rec_bytes, unsigned char* __restrict__ rec) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const long long b = i / D;
    const int d = static_cast<int>(i - b * D);
    const bool real = mask[i] != 0;
    unsigned char* r = rec + b * rec_bytes;
    float* xo = reinterpret_cast<float*>(r) + d * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) xo[c] = real ? det_add(det_mul(x[i * 3 + c], nm.std[c]), nm.mean[c]) : 0.0f;
    const int tok = (k != nullptr && real) ? static_cast<int>(k[i]) & 0x7f : 0;
    r[D * 12 + d] = static_cast<unsigned char>(tok | (real ? 0x80 : 0));
    if (d == 0) {                                         // the record's tail padding
        for (int j = D * 13; j < rec_bytes; ++j) r[j] = 0;
    }
}

__global__ void sample_unpack_kernel(const unsigned char* __restrict__ rec, long long slots, int D, int rec_bytes,
                                     float* __restrict__ x, long long* __restrict__ k, long long* __restrict__ mask) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const long long b = i / D;
    const int d = static_cast<int>(i - b * D);
    const unsigned char* r = rec + b * rec_bytes;
    const float* xi = reinterpret_cast<const float*>(r) + d * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) x[i * 3 + c] = xi[c];
    const unsigned char kb = r[D * 12 + d];
    if (k) k[i] = kb & 0x7f;
    if (mask) mask[i] = kb >> 7;
}

__global__ void force_tokens_kernel(const unsigned char* __restrict__ forced, const int* __restrict__ row_slot,
                                    int rows, int* __restrict__ ks) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) ks[r] = forced[row_slot[r]];
}

// ---------------------------------------------------------------------------------------------
// 3. input embedding, first layer: h = GELU(W0 x + b0), K = 3 -> CUDA cores   (wxe.0 / epic.wxe)
//    one thread produces 8 consecutive features of one row (one 16-byte bf16 store)
// ---------------------------------------------------------------------------------------------
__global__ void embed_x_kernel(const float* __restrict__ xs, int rows, const float* __restrict__ w0 /*[E][3]*/,
                               const float* __restrict__ b0, int E, int apply_gelu, bf16* __restrict__ out, int ld_out) {
    const int per_row = E / 8;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int r = static_cast<int>(idx / per_row), c0 = static_cast<int>(idx % per_row) * 8;
    if (r >= rows) return;
    const float a = xs[r * 3 + 0], b = xs[r * 3 + 1], c = xs[r * 3 + 2];
    uint32_t packed[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float h[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int f = c0 + j * 2 + q;
            float v = fmaf(__ldg(w0 + f * 3 + 2), c, fmaf(__ldg(w0 + f * 3 + 1), b, fmaf(__ldg(w0 + f * 3), a, __ldg(b0 + f))));
            h[q] = apply_gelu ? gelu_erf_simt(v) : v;
        }
        __nv_bfloat162 p = __floats2bfloat162_rn(h[0], h[1]);
        packed[j] = *reinterpret_cast<uint32_t*>(&p);
    }
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * ld_out + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

// ---------------------------------------------------------------------------------------------
// 4. row kernels on the 256-wide residual stream, one warp per row, lane owns 8 consecutive columns.
//    LayerNorm groups are either two halves of 128 (ParticleFormer's x | y streams) or the full 256.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float group_sum(float v, int group_width) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    if (group_width == 256) v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}
// v[8] = this lane's columns [lane*8, lane*8+8).  Normalises in place over the lane's group.
__device__ __forceinline__ void warp_layernorm(float* v, int lane, int group_width, const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = group_sum(s, group_width) / group_width;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(group_sum(q, group_width) / group_width + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf((v[i] - mean) * rstd, __ldg(g + lane * 8 + i), b ? __ldg(b + lane * 8 + i) : 0.f);
}
__device__ __forceinline__ void load8(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8_bf16(bf16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// after the wxe.2 GEMM:  x-half = LN_ln1x(raw) + temb ;  y-half = Ytab[k] + temb ;  keep a copy as the skip
// stream; emit the first block's LayerNorm as the bf16 GEMM operand.
__global__ void embed_finish_kernel(EmbedFinishArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.rows) return;
    const int row = warp;
    float v[8], yv[8];
    const float* tb = a.temb + static_cast<size_t>(a.row_jet ? a.row_jet[row] : 0) * a.temb_ld + lane * 8;
    if (lane < 16) {
        load8(a.resid + static_cast<size_t>(row) * 256 + lane * 8, v);
    } else {
        load8(a.ytab + a.ks[row] * 128 + (lane - 16) * 8, yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = yv[i];
    }
    // all 32 lanes take part in the shuffles; the upper half-warp's result is discarded
    warp_layernorm(v, lane & 15, 128, a.ln1x_g, a.ln1x_b);
    if (lane >= 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = yv[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += __ldg(tb + i);
    store8(a.resid + static_cast<size_t>(row) * 256 + lane * 8, v);
    store8(a.skip + static_cast<size_t>(row) * 256 + lane * 8, v);
    warp_layernorm(v, lane, a.next_ln_width, a.next_g, a.next_b);
    store8_bf16(a.act + static_cast<size_t>(row) * 256 + lane * 8, v);
}

// stream junctions: v = resid + skip ; LN_1 (groups) ; + temb2 ; [write resid] ; [LN_2] ; bf16 operand out
__global__ void add_ln_kernel(AddLnArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.rows) return;
    const int row = warp;
    float v[8], s[8];
    load8(a.resid + static_cast<size_t>(row) * 256 + lane * 8, v);
    load8(a.skip + static_cast<size_t>(row) * 256 + lane * 8, s);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += s[i];
    warp_layernorm(v, lane, a.ln1_width, a.ln1_g, a.ln1_b);
    if (a.temb) {
        const float* tb = a.temb + static_cast<size_t>(a.row_jet ? a.row_jet[row] : 0) * a.temb_ld + lane * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += __ldg(tb + i);
    }
    if (a.write_resid) store8(a.resid + static_cast<size_t>(row) * 256 + lane * 8, v);
    if (a.ln2_g) warp_layernorm(v, lane, a.ln2_width, a.ln2_g, a.ln2_b);
    store8_bf16(a.act + static_cast<size_t>(row) * 256 + lane * 8, v);
}

// ---------------------------------------------------------------------------------------------
// 5. output projections (512 -> 3 and 512 -> V) fused with the hybrid step        (head_x.2, head_y.2)
//    one warp per row: lane owns 16 hidden units of each head, 3 + V dot products, xor-reduce, then lane 0
//    takes the Euler + telegraph step for the particle (or writes vt / logits for the forward-only API).
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256)
head_out_kernel(HeadOutArgs a) {
    extern __shared__ float s_w[];            // [3 + V][512] fp32 weights, then [3 + V] biases
    constexpr int NO = 3 + V;
    for (int i = threadIdx.x; i < NO * 512; i += blockDim.x) s_w[i] = i < 3 * 512 ? a.wx[i] : a.wy[i - 3 * 512];
    for (int i = threadIdx.x; i < NO; i += blockDim.x) s_w[NO * 512 + i] = i < 3 ? a.bx[i] : a.by[i - 3];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < a.rows; row += warps_per_grid) {
        float acc[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {            // 0: head_x hidden, 1: head_y hidden
            const uint4* src = reinterpret_cast<const uint4*>(a.hidden + static_cast<size_t>(row) * a.ld_hidden + half * 512 + lane * 16);
            float h[16];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint4 raw = __ldg(src + q);
                const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&w4[j]);
                    h[q * 8 + j * 2] = __low2float(p);
                    h[q * 8 + j * 2 + 1] = __high2float(p);
                }
            }
            const int o0 = half == 0 ? 0 : 3, o1 = half == 0 ? 3 : NO;
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                if (o >= o0 && o < o1) {
                    const float* w = s_w + o * 512 + lane * 16;
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[o] = fmaf(h[j], w[j], acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], m);
            acc[o] += s_w[NO * 512 + o];
        }
        if (lane == 0) {
            const long long slot = a.row_slot[row];
            if (a.vt_out) {                       // forward-only API: padded (B, D, .) outputs
                for (int c = 0; c < 3; ++c) a.vt_out[slot * 3 + c] = acc[c];
                for (int v = 0; v < V; ++v) a.logits_out[slot * V + v] = acc[3 + v];
            }
            if (a.do_step) {
                float u[V], rates[V];
                if (a.sl.u) {
#pragma unroll
                    for (int v = 0; v < V; ++v) u[v] = __ldg(a.sl.u + slot * V + v);
                } else {
                    philox_uniforms(a.sl.seed, a.sl.slot0 + static_cast<uint64_t>(slot), a.sl.step, V, u);
                }
                const int kn = step_particle<V>(acc + 3, a.ks[row], a.w, a.coef, a.sl.sp, u, (a.rates_out || a.argmax_out) ? rates : nullptr);
                a.ks[row] = a.forced ? static_cast<int>(a.forced[slot]) : kn;
#pragma unroll
                for (int c = 0; c < 3; ++c) a.xs[row * 3 + c] = euler_update(a.xs[row * 3 + c], acc[c], a.sl.sp.dt);
                if (a.rates_out) {
#pragma unroll
                    for (int v = 0; v < V; ++v) a.rates_out[slot * V + v] = rates[v];
                }
                if (a.argmax_out) {               // use_final_max_rates (reference model/MMF.py:193-196)
                    int best = 0;
#pragma unroll
                    for (int v = 1; v < V; ++v) best = rates[v] > rates[best] ? v : best;
                    a.ks[row] = best;
                }
            }
        }
    }
}

template <int V>
int launch_head_out_t(const HeadOutArgs& a, cudaStream_t stream) {
    const int smem = ((3 + V) * 512 + (3 + V)) * sizeof(float);
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        MMF_CUDA_OK(cudaFuncSetAttribute(head_out_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    int blocks = (a.rows + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    head_out_kernel<V><<<blocks, 256, smem, stream>>>(a);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int V>
int launch_step_prod(const float* vt, const float* logits, float* x, long long* k, const float* t, long long n, int D,
                     const StepLaunch& sl, cudaStream_t stream) {
    static int resident[64] = {0};                        // CTAs the device holds at once, per device ordinal
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    constexpr int smem = prod_smem_bytes<V>();
    if (dev < 0 || dev >= 64 || resident[dev] == 0) {
        MMF_CUDA_OK(cudaFuncSetAttribute(hybrid_step_prod_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int sms = 0, per_sm = 0;
        MMF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MMF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hybrid_step_prod_kernel<V>, kProdThreads, smem));
        MMF_REQUIRE(per_sm > 0, "step kernel does not fit on this device");
        if (dev >= 0 && dev < 64) resident[dev] = sms * per_sm;
        else return 2;
    }
    ProdConst pc;
    pc.sc = 1.44269504f / sl.sp.temperature;
    pc.a2 = static_cast<float>(-static_cast<double>(V) * static_cast<double>(sl.sp.beta) * 1.4426950408889634);
    pc.filters = (sl.sp.top_k > 0 && sl.sp.top_k != V) || sl.sp.top_p > 0.0f;
    pc.div_magic = D > 1 ? ~0ull / static_cast<unsigned long long>(D) + 1ull : 0ull;
    const auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const int bulk_ok = aligned(vt) && aligned(logits) && aligned(x) && aligned(k);
    const long long chunks = (n + kProdChunk - 1) / kProdChunk;
    const unsigned grid = static_cast<unsigned>(chunks < resident[dev] ? chunks : resident[dev]);
    hybrid_step_prod_kernel<V><<<grid, kProdThreads, smem, stream>>>(vt, logits, x, k, t, n, D, sl, pc, bulk_ok);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

#define MMF_DISPATCH_V(V_, CALL)                                        \
    switch (V_) {                                                       \
        case 2: { constexpr int VV = 2; CALL; } break;                  \
        case 3: { constexpr int VV = 3; CALL; } break;                  \
        case 4: { constexpr int VV = 4; CALL; } break;                  \
        case 5: { constexpr int VV = 5; CALL; } break;                  \
        case 6: { constexpr int VV = 6; CALL; } break;                  \
        case 7: { constexpr int VV = 7; CALL; } break;                  \
        case 8: { constexpr int VV = 8; CALL; } break;                  \
        case 9: { constexpr int VV = 9; CALL; } break;                  \
        case 10: { constexpr int VV = 10; CALL; } break;                \
        case 12: { constexpr int VV = 12; CALL; } break;                \
        case 16: { constexpr int VV = 16; CALL; } break;                \
        default:                                                        \
            set_last_error("vocab_size must be one of 2..10, 12, 16");  \
            return 2;                                                   \
    }

int launch_hybrid_step(const float* vt, const float* logits, float* x, long long* k, const float* t, int B, int D,
                       const StepLaunch& sl, float* rates_out, cudaStream_t stream) {
    const long long n = static_cast<long long>(B) * D;
    if (n == 0) return 0;
    const unsigned blocks = static_cast<unsigned>((n + kStepThreads - 1) / kStepThreads);
    // production mode (Philox draws, no rates) runs the MUFU arithmetic; MMF_STEP_EXACT=1 forces the reproducible one
    const char* fe = getenv("MMF_STEP_EXACT");
    const bool force_exact = fe != nullptr && atoi(fe) != 0;
    if (sl.u == nullptr && rates_out == nullptr && !force_exact && sl.sp.method == 0) {
        MMF_DISPATCH_V(sl.sp.vocab, return launch_step_prod<VV>(vt, logits, x, k, t, n, D, sl, stream));
    } else {
        MMF_DISPATCH_V(sl.sp.vocab, (hybrid_step_kernel<VV, false><<<blocks, kStepThreads, 0, stream>>>(vt, logits, x, k, t, n, D, sl, rates_out)));
    }
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_euler(const float* vt, float* x, float dt, long long n, cudaStream_t stream) {
    if (n == 0) return 0;
    euler_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(vt, x, dt, n);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_jet_observables(const float* x, const long long* k, const long long* mask, const ObsArgs& a, cudaStream_t stream) {
    if (a.B == 0) return 0;
    const long long threads = static_cast<long long>(a.B) * kObsLanes;
    jet_observables_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(x, k, mask, a);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_make_source(const SourceArgs& a, float* x0, long long* k0, long long* mask, int* n_out, cudaStream_t stream) {
    if (a.B == 0) return 0;
    source_mult_kernel<<<static_cast<unsigned>((a.B + 255) / 256), 256, 0, stream>>>(a, n_out);
    MMF_CUDA_OK(cudaGetLastError());
    const long long slots = a.B * a.D;
    source_fill_kernel<<<static_cast<unsigned>((slots + 255) / 256), 256, 0, stream>>>(a, n_out, x0, k0, mask);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_pack(const float* x0, const long long* k0, const int* row_slot, int rows, int V, float* xs, int* ks,
                int* err_flag, cudaStream_t stream) {
    if (rows == 0) return 0;
    pack_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(x0, k0, row_slot, rows, V, xs, ks, err_flag);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_unpack(const float* xs, const int* ks, const int* row_slot, int rows, float* x_out, long long* k_out,
                  cudaStream_t stream) {
    if (rows == 0) return 0;
    unpack_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(xs, ks, row_slot, rows, x_out, k_out);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int sample_record_bytes(int D) { return (D * 13 + 15) / 16 * 16; }

int launch_sample_pack(const float* x, const long long* k, const long long* mask, const float* mean, const float* std_, long long B,
                       int D, unsigned char* rec, cudaStream_t stream) {
    const long long slots = B * D;
    if (slots == 0) return 0;
    SampleNorm nm;
    for (int c = 0; c < 3; ++c) { nm.mean[c] = mean ? mean[c] : 0.0f; nm.std[c] = std_ ? std_[c] : 1.0f; }
    sample_pack_kernel<<<static_cast<unsigned>((slots + 255) / 256), 256, 0, stream>>>(x, k, mask, nm, slots, D, sample_record_bytes(D), rec);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_sample_unpack(const unsigned char* rec, long long B, int D, float* x, long long* k, long long* mask, cudaStream_t stream) {
    const long long slots = B * D;
    if (slots == 0) return 0;
    sample_unpack_kernel<<<static_cast<unsigned>((slots + 255) / 256), 256, 0, stream>>>(rec, slots, D, sample_record_bytes(D), x, k, mask);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_force_tokens(const unsigned char* forced, const int* row_slot, int rows, int* ks, cudaStream_t stream) {
    if (rows == 0) return 0;
    force_tokens_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(forced, row_slot, rows, ks);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_embed_x(const float* xs, int rows, const float* w0, const float* b0, int E, int apply_gelu, bf16* out,
                   int ld_out, cudaStream_t stream) {
    if (rows == 0) return 0;
    const long long threads = static_cast<long long>(rows) * (E / 8);
    embed_x_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(xs, rows, w0, b0, E, apply_gelu, out, ld_out);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_embed_finish(const EmbedFinishArgs& a, cudaStream_t stream) {
    if (a.rows == 0) return 0;
    embed_finish_kernel<<<(a.rows + 7) / 8, 256, 0, stream>>>(a);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_add_ln(const AddLnArgs& a, cudaStream_t stream) {
    if (a.rows == 0) return 0;
    add_ln_kernel<<<(a.rows + 7) / 8, 256, 0, stream>>>(a);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_head_out(const HeadOutArgs& a, int V, cudaStream_t stream) {
    if (a.rows == 0) return 0;
    MMF_DISPATCH_V(V, return launch_head_out_t<VV>(a, stream));
    return 0;
}

}  // namespace mmf
