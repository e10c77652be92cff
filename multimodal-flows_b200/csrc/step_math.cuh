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
};

template <int V>
__device__ __forceinline__ void step_softmax(const float* l, float T, float* p) {
    float z[V];
#pragma unroll
    for (int v = 0; v < V; ++v) z[v] = (T != 1.0f) ? det_div(l[v], T) : l[v];
    float m = z[0];
#pragma unroll
    for (int v = 1; v < V; ++v) m = z[v] > m ? z[v] : m;
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        p[v] = det_expf(det_add(z[v], -m));
        s = (v == 0) ? p[0] : det_add(s, p[v]);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) p[v] = det_div(p[v], s);
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

template <int V>
__device__ __forceinline__ void step_renorm(float* p, const bool* keep) {
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        p[v] = keep[v] ? p[v] : 0.0f;
        s = (v == 0) ? p[0] : det_add(s, p[v]);
    }
    const float d = det_add(s, 1e-8f);
#pragma unroll
    for (int v = 0; v < V; ++v) p[v] = det_div(p[v], d);
}

template <int V>
__device__ __forceinline__ void step_filters(float* p, int top_k, float top_p) {
    if (top_k > 0 && top_k != V) {
        int rank[V];
        bool keep[V];
        step_ranks<V>(p, rank);
#pragma unroll
        for (int v = 0; v < V; ++v) keep[v] = rank[v] < top_k;
        step_renorm<V>(p, keep);
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
            cum = (j == 0) ? pj : det_add(cum, pj);
            keep_sorted[j] = (j == 0) || (cum <= top_p);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            bool kv = false;
#pragma unroll
            for (int j = 0; j < V; ++j) kv = (rank[v] == j) ? keep_sorted[j] : kv;
            keep[v] = kv;
        }
        step_renorm<V>(p, keep);
    }
}

// returns the new token; `rates` (V floats) is written when non-null.  `k` must be in [0,V).
template <int V>
__device__ __forceinline__ int step_particle(const float* logits, int k, float w, float coef, const StepParams& sp,
                                             const float* u, float* rates) {
    float p[V];
    step_softmax<V>(logits, sp.temperature, p);
    step_filters<V>(p, sp.top_k, sp.top_p);
    float qk = p[0];
#pragma unroll
    for (int v = 1; v < V; ++v) qk = (k == v) ? p[v] : qk;
    const float wq = det_mul(w, qk);
    int total = 0, single = k;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float rate = det_add(det_add(1.0f, det_mul(coef, p[v])), wq);
        if (rates) rates[v] = rate;
        const float lam = det_mul(rate, sp.dt);
        const float e = det_expf(-lam);
        const int c = (u[v] >= e) + (u[v] >= det_mul(e, det_add(1.0f, lam)));
        total += c;
        single = (c == 1) ? v : single;
    }
    return (total == 1) ? single : k;
}

__device__ __forceinline__ float euler_update(float x, float vt, float dt) { return det_add(x, det_mul(vt, dt)); }

}  // namespace mmf
