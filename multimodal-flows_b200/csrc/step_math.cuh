// The hybrid step for ONE particle, in registers.  Mirrors oracle/step_oracle.c operation for operation.
//   softmax(logits / T) -> [top-k] -> [top-p] -> telegraph rate -> lam = rate dt ->
//   per-channel Poisson count in {0,1,>=2} from one uniform -> at most one jump
// reference: model/solvers.py:22-60 (tauleap_step), :101-119 (filters), model/MJB.py:163-195 (rate).
#pragma once
#include "mmf_common.cuh"

namespace mmf {

struct StepParams {
    float temperature;   // logits / T when T != 1
    float dt;
    float beta;
    float top_p;         // <= 0: off
    int top_k;           // <= 0: off
    int vocab;
    int method;          // 0: tau-leap (reference solvers.py:22-60, the one the sampler uses); 1: categorical Euler (:62-91)
};

// Arithmetic policy.  EXACT (default): individually rounded IEEE ops and the polynomial det_expf, reproduced bit for bit by
// oracle/step_oracle.c - used whenever uniforms are supplied, rates are returned or tokens are forced (the parity modes).
// FAST: MUFU ex2 / rcp (relative error ~2^-22) for the production mode with in-kernel Philox draws, where no draw-level
// comparison exists and the step has to stay memory-bound; a decision differs from EXACT only when a uniform falls within
// ~1e-6 of a threshold (tests/test_gpu_step.py::test_fast_step_arithmetic_agrees_with_exact).
template <bool FAST> struct StepMath;
template <> struct StepMath<false> {
    static __device__ __forceinline__ float exp(float x) { return det_expf(x); }
    static __device__ __forceinline__ float div(float a, float b) { return det_div(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return det_mul(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return det_add(a, b); }
};
template <> struct StepMath<true> {
    static __device__ __forceinline__ float exp(float x) {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.44269504f));
        return y;
    }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
};

template <int V, bool FAST = false>
__device__ __forceinline__ void step_softmax(const float* l, float T, float* p) {
    using M = StepMath<FAST>;
    float z[V];
#pragma unroll
    for (int v = 0; v < V; ++v) z[v] = (T != 1.0f) ? M::div(l[v], T) : l[v];
    float m = z[0];
#pragma unroll
    for (int v = 1; v < V; ++v) m = z[v] > m ? z[v] : m;
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        p[v] = M::exp(M::add(z[v], -m));
        s = (v == 0) ? p[0] : M::add(s, p[v]);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) p[v] = M::div(p[v], s);
}

template <int V>
__device__ __forceinline__ void step_ranks(const float* p, int* rank) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
        int r = 0;
#pragma unroll
        for (int w = 0; w < V; ++w) r += (p[w] > p[v]) || (p[w] == p[v] && w < v);
        rank[v] = r;
    }
}

template <int V, bool FAST = false>
__device__ __forceinline__ void step_renorm(float* p, const bool* keep) {
    using M = StepMath<FAST>;
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        p[v] = keep[v] ? p[v] : 0.0f;
        s = (v == 0) ? p[0] : M::add(s, p[v]);
    }
    const float d = M::add(s, 1e-8f);
#pragma unroll
    for (int v = 0; v < V; ++v) p[v] = M::div(p[v], d);
}

template <int V, bool FAST = false>
__device__ __forceinline__ void step_filters(float* p, int top_k, float top_p) {
    using M = StepMath<FAST>;
    if (top_k > 0 && top_k != V) {
        int rank[V];
        bool keep[V];
        step_ranks<V>(p, rank);
#pragma unroll
        for (int v = 0; v < V; ++v) keep[v] = rank[v] < top_k;
        step_renorm<V, FAST>(p, keep);
    }
    if (top_p > 0.0f) {
        int rank[V];
        bool keep[V];
        step_ranks<V>(p, rank);
        bool keep_sorted[V];
        float cum = 0.0f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float pj = 0.0f;
#pragma unroll
            for (int v = 0; v < V; ++v) pj = (rank[v] == j) ? p[v] : pj;
            cum = (j == 0) ? pj : M::add(cum, pj);
            keep_sorted[j] = (j == 0) || (cum <= top_p);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            bool kv = false;
#pragma unroll
            for (int j = 0; j < V; ++j) kv = (rank[v] == j) ? keep_sorted[j] : kv;
            keep[v] = kv;
        }
        step_renorm<V, FAST>(p, keep);
    }
}

// returns the new token; `rates` (V floats) is written when non-null.  `k` must be in [0,V).
template <int V, bool FAST = false>
__device__ __forceinline__ int step_particle(const float* logits, int k, float w, float coef, const StepParams& sp,
                                             const float* u, float* rates) {
    using M = StepMath<FAST>;
    float p[V];
    step_softmax<V, FAST>(logits, sp.temperature, p);
    step_filters<V, FAST>(p, sp.top_k, sp.top_p);
    float qk = p[0];
#pragma unroll
    for (int v = 1; v < V; ++v) qk = (k == v) ? p[v] : qk;
    const float wq = M::mul(w, qk);
    int total = 0, single = k;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float rate = M::add(M::add(1.0f, M::mul(coef, p[v])), wq);
        if (rates) rates[v] = rate;
        const float lam = M::mul(rate, sp.dt);
        const float e = M::exp(-lam);
        const int c = (u[v] >= e) + (u[v] >= M::mul(e, M::add(1.0f, lam)));
        total += c;
        single = (c == 1) ? v : single;
    }
    return (total == 1) ? single : k;
}

// HybridSolver.euler_step (reference model/solvers.py:62-91, T = 1): the categorical jump.  Off-diagonal transition
// probabilities dp_v = min(rate_v dt, 1), diagonal dp_k = max(1 - sum_{v != k} dp_v, 0); top-k / top-p act on dp (not on the
// softmax); k' ~ Categorical(dp), which normalises dp.  One uniform: k' = min{v : u sum(dp) < cum_v} (sequential sums,
// falling back to the last channel with dp > 0).  `rates` (unfiltered softmax, as in the reference) is written when non-null.
template <int V>
__device__ __forceinline__ int step_particle_euler(const float* logits, int k, float w, float coef, const StepParams& sp,
                                                   float u0, float* rates) {
    using M = StepMath<false>;
    float p[V], dp[V];
    step_softmax<V, false>(logits, 1.0f, p);
    float qk = p[0];
#pragma unroll
    for (int v = 1; v < V; ++v) qk = (k == v) ? p[v] : qk;
    const float wq = M::mul(w, qk);
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float rate = M::add(M::add(1.0f, M::mul(coef, p[v])), wq);
        if (rates) rates[v] = rate;
        const float d = fminf(M::mul(rate, sp.dt), 1.0f);
        dp[v] = (v == k) ? 0.0f : d;
        s = (v == 0) ? dp[0] : M::add(s, dp[v]);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) dp[v] = (v == k) ? fmaxf(M::add(1.0f, -s), 0.0f) : dp[v];
    step_filters<V, false>(dp, sp.top_k, sp.top_p);
    float cum[V];
    float tot = 0.0f;
    int last = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        tot = (v == 0) ? dp[0] : M::add(tot, dp[v]);
        cum[v] = tot;
        last = dp[v] > 0.0f ? v : last;
    }
    const float target = M::mul(u0, tot);
    int j = last;
#pragma unroll
    for (int v = V - 1; v >= 0; --v) j = (target < cum[v] && dp[v] > 0.0f) ? v : j;
    return j;
}

// The same transition law from TWO uniforms: the V channel counts are independent Poisson(lam_v), so their
// total is Poisson(L), L = sum lam_v; the token changes iff the total is exactly one - probability L e^-L - and then to
// channel j with probability lam_j / L.  Distributionally identical to the per-channel form, not draw-identical: used only
// by the production mode of the standalone step kernel (in-kernel draws), where it cuts the RNG work by three.
template <int V>
__device__ __forceinline__ int step_particle_2u(const float* logits, int k, float w, float coef, const StepParams& sp, float u1, float u2) {
    using M = StepMath<true>;
    float p[V];
    step_softmax<V, true>(logits, sp.temperature, p);
    step_filters<V, true>(p, sp.top_k, sp.top_p);
    float qk = p[0];
#pragma unroll
    for (int v = 1; v < V; ++v) qk = (k == v) ? p[v] : qk;
    const float base = fmaf(w, qk, 1.0f);
    float cum[V];
    float L = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        L += fmaf(coef, p[v], base) * sp.dt;
        cum[v] = L;
    }
    if (!(u1 < L * M::exp(-L))) return k;            // zero or >= 2 events: the token stays (reference model/solvers.py:49-54)
    const float target = u2 * L;
    int j = V - 1;
#pragma unroll
    for (int v = V - 2; v >= 0; --v) j = (target < cum[v]) ? v : j;
    return j;
}

__device__ __forceinline__ float euler_update(float x, float vt, float dt) { return det_add(x, det_mul(vt, dt)); }

}  // namespace mmf
