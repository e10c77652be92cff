// Forward half of the MMF training step: everything around the encoder call of
// MultiModalFlowBridge.loss (reference model/MMF.py:138-170) as three fused, HBM-bound kernels.
//   bridge_sample_kernel   UniformFlow.sample (model/CFM.py:171-184) + RandomTelegraphBridge.sample / transition_probability /
//                          conditional_probability (model/MJB.py:197-257): xt = t x1 + (1 - t) x0 + sigma z and
//                          kt ~ Categorical(P), P(k) = p(k -> k1; t, 1) p(k0 -> k; 0, t) / p(k0 -> k1; 0, 1),
//                          p(a -> b; s, t) = 1/V + w (delta_ab - 1/V), w = exp(-V beta (t - s))    -- one thread per slot
//   multitask_loss_kernel  masked MSE against the conditional drift x1 - x0 (CFM.py:186-193) and cross entropy with
//                          ignore_index = 0, both normalised per jet by clamp_min(#real, 1) (MMF.py:156-165) -- one warp per jet
//   loss_combine_kernel    MultiTaskLoss (MMF.py:203-233): "sum", or "time-weighted" with the uncertainty net
//                          MLP(n_embd -> n_embd -> 2) on the sin/cos time features                   -- one CTA per jet
// The backward pass is not part of this library (the encoder has no backward kernels yet).
#include "mmf_common.cuh"
#include "mmf_simt.h"

namespace mmf {
namespace {

__device__ __forceinline__ float telegraph_p(float w, bool same, float inv_v) { return inv_v + w * ((same ? 1.0f : 0.0f) - inv_v); }

template <int V>
__global__ void bridge_sample_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const long long* __restrict__ k0,
                                     const long long* __restrict__ k1, const float* __restrict__ t, float sigma, float beta,
                                     const float* __restrict__ z, const float* __restrict__ u, unsigned long long seed,
                                     unsigned long long slot0, long long slots, int D, float* __restrict__ xt,
                                     long long* __restrict__ kt, int* __restrict__ err) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const float tj = __ldg(t + i / D);
    // draws: supplied, or Philox4x32-10 keyed on (seed, global slot): words 0-1 -> one normal pair, 2 -> third normal (with 3), 3 -> uniform
    float zz[3], uu;
    if (z) { zz[0] = z[i * 3]; zz[1] = z[i * 3 + 1]; zz[2] = z[i * 3 + 2]; }
    if (!z || !u) {
        const unsigned long long s = slot0 + static_cast<unsigned long long>(i);
        const Philox4 r = philox4x32_10(Philox4{static_cast<uint32_t>(s), static_cast<uint32_t>(s >> 32), 0u, 0x54524e31u},
                                        static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
        const Philox4 r2 = philox4x32_10(Philox4{static_cast<uint32_t>(s), static_cast<uint32_t>(s >> 32), 1u, 0x54524e31u},
                                         static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
        if (!z) {
            const float a0 = (static_cast<float>(r.x >> 9) + 0.5f) * 1.1920929e-07f, a1 = (static_cast<float>(r.y >> 9) + 0.5f) * 1.1920929e-07f;
            const float a2 = (static_cast<float>(r.z >> 9) + 0.5f) * 1.1920929e-07f, a3 = (static_cast<float>(r.w >> 9) + 0.5f) * 1.1920929e-07f;
            const float m0 = sqrtf(-2.0f * logf(a0)), m2 = sqrtf(-2.0f * logf(a2));
            float s0, c0, c2;
            sincospif(2.0f * a1, &s0, &c0);
            c2 = cospif(2.0f * a3);
            zz[0] = m0 * c0; zz[1] = m0 * s0; zz[2] = m2 * c2;
        }
        uu = u01_from_bits(r2.x);
    }
    if (u) uu = u[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // t * x1 + (1 - t) * x0, then + sigma * z: the reference's operation order, every op individually rounded
        const float lin = det_add(det_mul(tj, x1[i * 3 + c]), det_mul(det_add(1.0f, -tj), x0[i * 3 + c]));
        xt[i * 3 + c] = det_add(lin, det_mul(sigma, zz[c]));
    }
    long long a = k0[i], b = k1[i];
    if (a < 0 || a >= V || b < 0 || b >= V) { atomicOr(err, 2); a = 0; b = 0; }
    const float inv_v = 1.0f / static_cast<float>(V);
    const float vb = static_cast<float>(V) * beta;
    const float w_t1 = expf(-vb * (1.0f - tj)), w_0t = expf(-vb * (tj - 0.0f)), w_01 = expf(-vb * (1.0f - 0.0f));
    const float den = telegraph_p(w_01, a == b, inv_v);
    float cum[V];
    float tot = 0.0f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const float p = telegraph_p(w_t1, k == b, inv_v) * telegraph_p(w_0t, a == k, inv_v) / den;
        tot += p;
        cum[k] = tot;
    }
    const float target = uu * tot;                      // Categorical normalises its probabilities
    int pick = V - 1;
#pragma unroll
    for (int k = V - 2; k >= 0; --k) pick = (target < cum[k]) ? k : pick;
    kt[i] = pick;
}

template <int V>
__global__ void multitask_loss_kernel(const float* __restrict__ vt, const float* __restrict__ logits, const float* __restrict__ x0,
                                      const float* __restrict__ x1, const long long* __restrict__ k1,
                                      const long long* __restrict__ mask, int B, int D, float* __restrict__ loss_mse,
                                      float* __restrict__ loss_ce) {
    const int jet = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (jet >= B) return;
    float mse = 0.0f, ce = 0.0f, n = 0.0f;
    for (int d = lane; d < D; d += 32) {
        const long long i = static_cast<long long>(jet) * D + d;
        if (mask[i] == 0) continue;
        n += 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float e = vt[i * 3 + c] - (x1[i * 3 + c] - x0[i * 3 + c]);
            mse += e * e;
        }
        const long long tgt = k1[i];
        if (tgt != 0) {                                 // ignore_index = 0
            float l[V], m = -INFINITY;
#pragma unroll
            for (int v = 0; v < V; ++v) { l[v] = logits[i * V + v]; m = fmaxf(m, l[v]); }
            float s = 0.0f, lt = l[0];
#pragma unroll
            for (int v = 0; v < V; ++v) { s += expf(l[v] - m); lt = (v == tgt) ? l[v] : lt; }
            ce += (m + logf(s)) - lt;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mse += __shfl_xor_sync(0xffffffffu, mse, o);
        ce += __shfl_xor_sync(0xffffffffu, ce, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if (lane == 0) {
        const float dn = fmaxf(n, 1.0f);
        loss_mse[jet] = mse / dn;
        loss_ce[jet] = ce / dn;
    }
}

// out[0..4] += (loss, loss_mse, loss_ce, w_mse, w_ce) / B    (atomics on five floats; out is zeroed by the launcher)
__global__ void __launch_bounds__(256) loss_combine_kernel(const float* __restrict__ t, const float* __restrict__ loss_mse,
                                                            const float* __restrict__ loss_ce, const float* __restrict__ w_fc /*[E][E]*/,
                                                            const float* __restrict__ b_fc, const float* __restrict__ w_pr /*[2][E]*/,
                                                            const float* __restrict__ b_pr, int E, int mode, int B, float* __restrict__ out) {
    __shared__ float s_emb[256], s_red[2][8];
    const int jet = blockIdx.x, tid = threadIdx.x;
    const float l1 = loss_mse[jet], l2 = loss_ce[jet];
    float u1 = 0.0f, u2 = 0.0f;
    if (mode == 1) {                                    // time-weighted: uncertainty_net(transformer_timestep_embedding(t, E))
        const int half = E / 2;
        const float scale = logf(10000.0f) / static_cast<float>(half - 1);
        if (tid < E) {
            const int j = tid < half ? tid : tid - half;
            const float a = t[jet] * expf(static_cast<float>(j) * -scale);
            s_emb[tid] = tid < half ? sinf(a) : cosf(a);
        }
        __syncthreads();
        float h = 0.0f;
        if (tid < E) {
            const float* w = w_fc + static_cast<size_t>(tid) * E;
            float acc = b_fc[tid];
            for (int i = 0; i < E; ++i) acc = fmaf(w[i], s_emb[i], acc);
            h = 0.5f * acc * (1.0f + erff(acc * 0.70710678118654752f));
        }
        float p1 = tid < E ? h * w_pr[tid] : 0.0f, p2 = tid < E ? h * w_pr[E + tid] : 0.0f;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o); }
        if ((tid & 31) == 0) { s_red[0][tid >> 5] = p1; s_red[1][tid >> 5] = p2; }
        __syncthreads();
        if (tid == 0) {
            for (int w8 = 0; w8 < 8; ++w8) { u1 += s_red[0][w8]; u2 += s_red[1][w8]; }
            u1 += b_pr[0]; u2 += b_pr[1];
        }
    }
    if (tid == 0) {
        const float inv = 1.0f / static_cast<float>(B);
        float loss, w1 = 1.0f, w2 = 1.0f;
        if (mode == 0) loss = l1 + l2;
        else { w1 = expf(-u1); w2 = expf(-u2); loss = 0.5f * (u1 + w1 * l1) + 0.5f * (u2 + w2 * l2); }
        atomicAdd(out + 0, loss * inv); atomicAdd(out + 1, l1 * inv); atomicAdd(out + 2, l2 * inv);
        atomicAdd(out + 3, w1 * inv); atomicAdd(out + 4, w2 * inv);
    }
}

// EMA of the weights as timm's ModelEmaV2.update does it for every state_dict entry (reference utils/callbacks.py:152-226,
// on_train_batch_end): ema <- decay * ema + (1 - decay) * p, each operation rounded on its own like the torch expression
__global__ void ema_update_kernel(float* __restrict__ ema, const float* __restrict__ p, float decay, float one_minus, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) ema[i] = det_add(det_mul(decay, ema[i]), det_mul(one_minus, p[i]));
}

}  // namespace

int launch_ema_update(float* ema, const float* p, double decay, long long n, cudaStream_t s) {
    if (n == 0) return 0;
    // decay and (1. - decay) are Python doubles in the reference, each rounded to fp32 when it multiplies an fp32 tensor
    const float one_minus = static_cast<float>(1.0 - decay);
    ema_update_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(ema, p, static_cast<float>(decay), one_minus, n);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_bridge_sample(const float* x0, const float* x1, const long long* k0, const long long* k1, const float* t, float sigma,
                         float beta, int V, const float* z, const float* u, unsigned long long seed, unsigned long long slot0,
                         long long B, int D, float* xt, long long* kt, int* err, cudaStream_t s) {
    const long long slots = B * D;
    if (slots == 0) return 0;
    MMF_REQUIRE(V == 9, "bridge sampling is instantiated for vocab_size 9");
    bridge_sample_kernel<9><<<static_cast<unsigned>((slots + 255) / 256), 256, 0, s>>>(x0, x1, k0, k1, t, sigma, beta, z, u, seed, slot0, slots, D, xt, kt, err);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_multitask_loss(const float* vt, const float* logits, const float* x0, const float* x1, const long long* k1,
                          const long long* mask, int B, int D, int V, float* loss_mse, float* loss_ce, cudaStream_t s) {
    if (B == 0) return 0;
    MMF_REQUIRE(V == 9, "the loss kernel is instantiated for vocab_size 9");
    multitask_loss_kernel<9><<<(B * 32 + 255) / 256, 256, 0, s>>>(vt, logits, x0, x1, k1, mask, B, D, loss_mse, loss_ce);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_loss_combine(const float* t, const float* loss_mse, const float* loss_ce, const float* w_fc, const float* b_fc,
                        const float* w_pr, const float* b_pr, int E, int mode, int B, float* out5, cudaStream_t s) {
    MMF_CUDA_OK(cudaMemsetAsync(out5, 0, 5 * sizeof(float), s));
    if (B == 0) return 0;
    MMF_REQUIRE(mode == 0 || (E >= 4 && E <= 256 && w_fc && b_fc && w_pr && b_pr), "loss_combine: time-weighted needs the uncertainty net (n_embd <= 256)");
    loss_combine_kernel<<<B, 256, 0, s>>>(t, loss_mse, loss_ce, w_fc, b_fc, w_pr, b_pr, E, mode, B, out5);
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace mmf
