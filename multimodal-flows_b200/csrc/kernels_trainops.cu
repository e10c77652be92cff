// CUDA-core kernels of the training step: everything between the tensor-core GEMMs of the encoder's
// forward and backward passes, on the packed row layout (mmf_train.h).  Each kernel is one HBM pass over its operands.
// Reference semantics: networks/ParticleTransformers.py:62-122, 177-210 (encoder), networks/attention.py:23-26, 53-74 (block),
// utils/models.py:8-37, 62-75 (MLP, LayerNorm, time embedding), model/MMF.py:138-170, 203-233 (loss), :77-78 (Adam) - the
// backward formulas are the derivatives torch autograd applies to those lines.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "mmf_ptx.cuh"
#include "mmf_train.h"

namespace mmf {
namespace {

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float bflo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bfhi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// ------------------------------------------------------------------------------------------------ small fp32 GEMM
__global__ void __launch_bounds__(256) tr_sgemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B,
                                                       long long sbk, long long sbn, float* __restrict__ C, long long ldc, int M, int N,
                                                       int K, const float* __restrict__ bias, int accumulate) {
    grid_dep_wait();
    grid_dep_launch();
    __shared__ float As[16][17], Bs[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
        const int ka = k0 + tx, kb = k0 + ty;
        As[ty][tx] = (m < M && ka < K) ? A[m * sam + ka * sak] : 0.f;
        Bs[ty][tx] = (kb < K && n < N) ? B[kb * sbk + n * sbn] : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(As[ty][k], Bs[k][tx], acc);
        __syncthreads();
    }
    if (m < M && n < N) {
        float r = acc + (bias ? bias[n] : 0.f);
        if (accumulate) r += C[m * ldc + n];
        C[m * ldc + n] = r;
    }
}

// ------------------------------------------------------------------------------------------------ cast / transpose / column sums
template <bool F32>
__global__ void __launch_bounds__(256) tr_cast_transpose_kernel(const void* __restrict__ in_, long long ld_in, int rows, int cols,
                                                                bf16* __restrict__ out, long long ld_out, bf16* __restrict__ outT,
                                                                long long ldT, float* __restrict__ colsum) {
    grid_dep_wait();
    grid_dep_launch();
    __shared__ float tile[32][33];
    __shared__ float part[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int c = c0 + tx;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + ty + 8 * j;
        float v = 0.f;
        if (r < rows && c < cols) {
            if (F32) v = static_cast<const float*>(in_)[r * ld_in + c];
            else v = bf2f(static_cast<const bf16*>(in_)[r * ld_in + c]);
            if (out) out[r * ld_out + c] = __float2bfloat16_rn(v);
        }
        tile[ty + 8 * j][tx] = v;
        s += v;
    }
    part[ty][tx] = s;
    __syncthreads();
    if (outT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cc = c0 + ty + 8 * j, rr = r0 + tx;
            if (cc < cols && rr < rows) outT[cc * ldT + rr] = __float2bfloat16_rn(tile[tx][ty + 8 * j]);
        }
    }
    if (colsum && ty == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][tx];
        atomicAdd(colsum + c, t);
    }
}

// all two-dimensional weights of the flat parameter buffer: bf16 transposed copies for the data-gradient GEMMs, one launch
__global__ void __launch_bounds__(256) tr_weights_transpose_kernel(const float* __restrict__ p, bf16* __restrict__ pT,
                                                                   const TrTransposeJob* __restrict__ jobs, int n_jobs) {
    grid_dep_wait();
    grid_dep_launch();
    __shared__ float tile[32][33];
    int j = 0;
    while (j + 1 < n_jobs && jobs[j + 1].tile0 <= static_cast<int>(blockIdx.x)) ++j;
    const TrTransposeJob job = jobs[j];
    const int t = blockIdx.x - job.tile0;
    const int tiles_c = (job.cols + 31) / 32;
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int r = r0 + ty + 8 * jj, c = c0 + tx;
        tile[ty + 8 * jj][tx] = (r < job.rows && c < job.cols) ? p[job.src + static_cast<long long>(r) * job.cols + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int c = c0 + ty + 8 * jj, r = r0 + tx;
        if (c < job.cols && r < job.rows) pT[job.dst + static_cast<long long>(c) * job.rows + r] = __float2bfloat16_rn(tile[tx][ty + 8 * jj]);
    }
}

// ------------------------------------------------------------------------------------------------ inputs
// gathers the real particles: xs = xt, ks = kt, tgt = x1 - x0 (conditional drift, reference model/CFM.py:186-193), k1
__global__ void tr_pack_kernel(const float* __restrict__ xt, const long long* __restrict__ kt, const float* __restrict__ x0,
                               const float* __restrict__ x1, const long long* __restrict__ k1, const int* __restrict__ row_slot, int M,
                               int V, float* __restrict__ xs, int* __restrict__ ks, float* __restrict__ tgt, int* __restrict__ k1p,
                               int* __restrict__ err) {
    grid_dep_wait();
    grid_dep_launch();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const long long s = row_slot[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        xs[r * 3 + c] = xt[s * 3 + c];
        tgt[r * 3 + c] = x1[s * 3 + c] - x0[s * 3 + c];
    }
    long long a = kt[s], b = k1[s];
    if (a < 0 || a >= V || b < 0 || b >= V) { atomicOr(err, 2); a = 0; b = 0; }
    ks[r] = static_cast<int>(a);
    k1p[r] = static_cast<int>(b);
}

// transformer_timestep_embedding (reference utils/models.py:62-75); dup = 1 writes the row twice (x | y halves)
__global__ void tr_time_embed_kernel(const float* __restrict__ t, const int* __restrict__ perm, int B, int dim, int dup, float* __restrict__ out,
                                     long long ld) {
    grid_dep_wait();
    grid_dep_launch();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * dim) return;
    const int b = i / dim, c = i % dim, half = dim / 2;
    const float scale = logf(10000.0f) / static_cast<float>(half - 1);
    const int j = c < half ? c : c - half;
    const float a = t[perm ? perm[b] : b] * expf(static_cast<float>(j) * -scale);     // row b = packed jet b = jet perm[b] of the batch
    const float v = c < half ? sinf(a) : cosf(a);
    out[b * ld + c] = v;
    if (dup) out[b * ld + dim + c] = v;
}

__global__ void tr_embed_x_fwd_kernel(const float* __restrict__ xs, int M, const float* __restrict__ w0, const float* __restrict__ b0,
                                      int E, bf16* __restrict__ h, long long ld) {
    grid_dep_wait();
    grid_dep_launch();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(M) * E) return;
    const int r = static_cast<int>(i / E), j = static_cast<int>(i % E);
    const float z = fmaf(w0[j * 3 + 2], xs[r * 3 + 2], fmaf(w0[j * 3 + 1], xs[r * 3 + 1], fmaf(w0[j * 3], xs[r * 3], b0[j])));
    h[r * ld + j] = __float2bfloat16_rn(gelu_f(z));
}

// thread = hidden unit j, CTA = a run of rows: dz = dh GELU'(z); dW0[j,:] += dz xs, db0[j] += dz
__global__ void tr_embed_x_bwd_kernel(const bf16* __restrict__ dh, long long ld, const float* __restrict__ xs, int M,
                                      const float* __restrict__ w0, const float* __restrict__ b0, int E, int rows_per_cta,
                                      float* __restrict__ dw0, float* __restrict__ db0) {
    grid_dep_wait();
    grid_dep_launch();
    const int j = threadIdx.x;
    if (j >= E) return;
    const int r_beg = blockIdx.x * rows_per_cta, r_end = min(M, r_beg + rows_per_cta);
    const float wa = w0[j * 3], wb = w0[j * 3 + 1], wc = w0[j * 3 + 2], bb = b0[j];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
    for (int r = r_beg; r < r_end; ++r) {
        const float x0 = xs[r * 3], x1 = xs[r * 3 + 1], x2 = xs[r * 3 + 2];
        const float z = fmaf(wc, x2, fmaf(wb, x1, fmaf(wa, x0, bb)));
        const float dz = bf2f(dh[r * ld + j]) * gelu_grad(z);
        a0 = fmaf(dz, x0, a0); a1 = fmaf(dz, x1, a1); a2 = fmaf(dz, x2, a2); ab += dz;
    }
    atomicAdd(dw0 + j * 3, a0); atomicAdd(dw0 + j * 3 + 1, a1); atomicAdd(dw0 + j * 3 + 2, a2); atomicAdd(db0 + j, ab);
}

__global__ void tr_embed_y_fwd_kernel(const int* __restrict__ ks, int M, const float* __restrict__ emb, int E, bf16* __restrict__ g,
                                      long long ld) {
    grid_dep_wait();
    grid_dep_launch();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(M) * E) return;
    const int r = static_cast<int>(i / E), j = static_cast<int>(i % E);
    g[r * ld + j] = __float2bfloat16_rn(gelu_f(emb[ks[r] * E + j]));
}

template <int V>
__global__ void tr_embed_y_bwd_kernel(const bf16* __restrict__ dg, long long ld, const int* __restrict__ ks, int M,
                                      const float* __restrict__ emb, int E, int rows_per_cta, float* __restrict__ demb) {
    grid_dep_wait();
    grid_dep_launch();
    const int j = threadIdx.x;
    if (j >= E) return;
    const int r_beg = blockIdx.x * rows_per_cta, r_end = min(M, r_beg + rows_per_cta);
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    for (int r = r_beg; r < r_end; ++r) {
        const int k = ks[r];
        const float d = bf2f(dg[r * ld + j]);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += (k == v) ? d : 0.f;
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (acc[v] != 0.f) atomicAdd(demb + v * E + j, acc[v] * gelu_grad(emb[v * E + j]));
}

// ------------------------------------------------------------------------------------------------ LayerNorm (warp per row)
template <int C>
__global__ void __launch_bounds__(256) tr_ln_fwd_kernel(const TrLnArgs a) {
    grid_dep_wait();
    grid_dep_launch();
    constexpr int NV = C / 128;                     // float4 per lane
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= a.M) return;
    float v[NV * 4];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = q * 128 + lane * 4;
        float4 x = *reinterpret_cast<const float4*>(a.x + r * a.ldx + c);
        if (a.add) {
            const float4 y = *reinterpret_cast<const float4*>(a.add + r * a.lda + c);
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        }
        v[q * 4] = x.x; v[q * 4 + 1] = x.y; v[q * 4 + 2] = x.z; v[q * 4 + 3] = x.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / C);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) { const float d = v[i] - mean; q2 = fmaf(d, d, q2); }
    const float rstd = rsqrtf(warp_sum(q2) * (1.0f / C) + 1e-5f);
    if (lane == 0 && a.mean) { a.mean[r] = mean; a.rstd[r] = rstd; }
    const float* tadd = a.tadd ? a.tadd + static_cast<long long>(a.row_jet ? a.row_jet[r] : 0) * a.ldt : nullptr;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = q * 128 + lane * 4;
        const float4 g = *reinterpret_cast<const float4*>(a.g + c);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f), t = b;
        if (a.b) b = *reinterpret_cast<const float4*>(a.b + c);
        if (tadd) t = *reinterpret_cast<const float4*>(tadd + c);
        float4 y;
        y.x = fmaf((v[q * 4] - mean) * rstd, g.x, b.x) + t.x;
        y.y = fmaf((v[q * 4 + 1] - mean) * rstd, g.y, b.y) + t.y;
        y.z = fmaf((v[q * 4 + 2] - mean) * rstd, g.z, b.z) + t.z;
        y.w = fmaf((v[q * 4 + 3] - mean) * rstd, g.w, b.w) + t.w;
        if (a.out32) *reinterpret_cast<float4*>(a.out32 + r * a.ld32 + c) = y;
        if (a.out16) *reinterpret_cast<uint2*>(a.out16 + r * a.ld16 + c) = make_uint2(pack2(y.x, y.y), pack2(y.z, y.w));
    }
}

template <int C>
__global__ void __launch_bounds__(256) tr_ln_bwd_kernel(const TrLnBwdArgs a) {
    grid_dep_wait();
    grid_dep_launch();
    constexpr int NV = C / 128;
    __shared__ float red[3][8][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float accg[NV * 4], accb[NV * 4], accx[NV * 4];
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) { accg[i] = 0.f; accb[i] = 0.f; accx[i] = 0.f; }
    for (int r = blockIdx.x * 8 + warp; r < a.M; r += gridDim.x * 8) {
        const float mean = a.mean[r], rstd = a.rstd[r];
        float xh[NV * 4], dxh[NV * 4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int c = q * 128 + lane * 4;
            float4 x = *reinterpret_cast<const float4*>(a.x + r * a.ldx + c);
            if (a.add) {
                const float4 y = *reinterpret_cast<const float4*>(a.add + r * a.lda + c);
                x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
            }
            const float4 dy = *reinterpret_cast<const float4*>(a.dy + r * a.lddy + c);
            const float4 g = *reinterpret_cast<const float4*>(a.g + c);
            const float xv[4] = {x.x, x.y, x.z, x.w}, dv[4] = {dy.x, dy.y, dy.z, dy.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float h = (xv[e] - mean) * rstd;
                xh[q * 4 + e] = h;
                dxh[q * 4 + e] = dv[e] * gv[e];
                accg[q * 4 + e] = fmaf(dv[e], h, accg[q * 4 + e]);
                accb[q * 4 + e] += dv[e];
                s1 += dxh[q * 4 + e];
                s2 = fmaf(dxh[q * 4 + e], h, s2);
            }
        }
        const float c1 = warp_sum(s1) * (1.0f / C), c2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int c = q * 128 + lane * 4;
            float4 d;
            d.x = rstd * (dxh[q * 4] - c1 - xh[q * 4] * c2);
            d.y = rstd * (dxh[q * 4 + 1] - c1 - xh[q * 4 + 1] * c2);
            d.z = rstd * (dxh[q * 4 + 2] - c1 - xh[q * 4 + 2] * c2);
            d.w = rstd * (dxh[q * 4 + 3] - c1 - xh[q * 4 + 3] * c2);
            float4* dst = reinterpret_cast<float4*>(a.dx + r * a.lddx + c);
            if (a.accumulate) { const float4 o = *dst; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
            *dst = d;
            // the new gradient as the bf16 operand of the next linear's products, and its column sums (that linear's bias gradient)
            if (a.dx16) *reinterpret_cast<uint2*>(a.dx16 + r * a.ld16 + c) = make_uint2(pack2(d.x, d.y), pack2(d.z, d.w));
            accx[q * 4] += d.x; accx[q * 4 + 1] += d.y; accx[q * 4 + 2] += d.z; accx[q * 4 + 3] += d.w;
        }
    }
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            red[0][warp][q * 128 + lane * 4 + e] = accg[q * 4 + e];
            red[1][warp][q * 128 + lane * 4 + e] = accb[q * 4 + e];
            red[2][warp][q * 128 + lane * 4 + e] = accx[q * 4 + e];
        }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float sg = 0.f, sb = 0.f, sx = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { sg += red[0][w][c]; sb += red[1][w][c]; sx += red[2][w][c]; }
        atomicAdd(a.dg + c, sg);
        if (a.db) atomicAdd(a.db + c, sb);
        if (a.dxsum) atomicAdd(a.dxsum + c, sx);
    }
}

// ------------------------------------------------------------------------------------------------ per-head LayerNorm of q, k
template <int HS>
__device__ __forceinline__ void load_head(const bf16* src, float* v) {
#pragma unroll
    for (int u = 0; u < HS / 8; ++u) {
        const uint4 w = *reinterpret_cast<const uint4*>(src + u * 8);
        v[u * 8] = bflo(w.x); v[u * 8 + 1] = bfhi(w.x); v[u * 8 + 2] = bflo(w.y); v[u * 8 + 3] = bfhi(w.y);
        v[u * 8 + 4] = bflo(w.z); v[u * 8 + 5] = bfhi(w.z); v[u * 8 + 6] = bflo(w.w); v[u * 8 + 7] = bfhi(w.w);
    }
}
template <int HS>
__device__ __forceinline__ void store_head(bf16* dst, const float* v) {
#pragma unroll
    for (int u = 0; u < HS / 8; ++u)
        *reinterpret_cast<uint4*>(dst + u * 8) = make_uint4(pack2(v[u * 8], v[u * 8 + 1]), pack2(v[u * 8 + 2], v[u * 8 + 3]),
                                                            pack2(v[u * 8 + 4], v[u * 8 + 5]), pack2(v[u * 8 + 6], v[u * 8 + 7]));
}
template <int HS>
__device__ __forceinline__ void head_stats(const float* v, float* mean, float* rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HS; ++i) s += v[i];
    const float m = s * (1.0f / HS);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < HS; ++i) { const float d = v[i] - m; q = fmaf(d, d, q); }
    *mean = m;
    *rstd = rsqrtf(q * (1.0f / HS) + 1e-5f);
}

template <int HS>
__global__ void __launch_bounds__(128) tr_qkln_fwd_kernel(const bf16* __restrict__ qkv, long long ld, int M, int C, int H,
                                                          const float* __restrict__ qg, const float* __restrict__ qb,
                                                          const float* __restrict__ kg, const float* __restrict__ kb,
                                                          bf16* __restrict__ qn, bf16* __restrict__ kn, long long ldn) {
    grid_dep_wait();
    grid_dep_launch();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(M) * 2 * H) return;
    const int head = static_cast<int>(idx % H), which = static_cast<int>((idx / H) % 2);
    const long long row = idx / (2 * H);
    float v[HS];
    load_head<HS>(qkv + row * ld + which * C + head * HS, v);
    float mean, rstd;
    head_stats<HS>(v, &mean, &rstd);
    const float* g = which ? kg : qg;
    const float* b = which ? kb : qb;
#pragma unroll
    for (int i = 0; i < HS; ++i) v[i] = fmaf((v[i] - mean) * rstd, __ldg(g + i), b ? __ldg(b + i) : 0.f);
    store_head<HS>((which ? kn : qn) + row * ldn + head * HS, v);
}

// in place: dqkv[:, which*C + head*HS ..] holds d(normalised q / k) on entry, d(q / k) on exit; blockIdx.y = which
template <int HS>
__global__ void __launch_bounds__(128) tr_qkln_bwd_kernel(bf16* __restrict__ dqkv, long long ldd, const bf16* __restrict__ qkv, long long ld,
                                                          int M, int C, int H, const float* __restrict__ qg, const float* __restrict__ kg,
                                                          float* __restrict__ dqg, float* __restrict__ dqb, float* __restrict__ dkg,
                                                          float* __restrict__ dkb) {
    grid_dep_wait();
    grid_dep_launch();
    // the affine gradients are sums over all (row, head) items: each warp transposes its 32 items through shared memory
    // (conflict-free pitch HS + 1) so that lane j adds up feature j - a quarter of the instructions of 2 HS warp reductions
    __shared__ float tile[4][32][HS + 1];
    const int which = blockIdx.y, lane = threadIdx.x & 31;
    float (*tw)[HS + 1] = tile[threadIdx.x >> 5];
    const float* g = which ? kg : qg;
    float* dgam = which ? dkg : dqg;
    float* dbet = which ? dkb : dqb;
    float accg[HS / 32], accb[HS / 32];
#pragma unroll
    for (int i = 0; i < HS / 32; ++i) { accg[i] = 0.f; accb[i] = 0.f; }
    const long long items = static_cast<long long>(M) * H;
    const long long chunks = (items + 31) / 32;
    const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long ch = warp0; ch < chunks; ch += nwarps) {
        const long long item = ch * 32 + lane;
        const bool valid = item < items;
        float x[HS], dy[HS];
        float mean = 0.f, rstd = 0.f;
        bf16* dptr = nullptr;
        if (valid) {
            const long long row = item / H;
            const int head = static_cast<int>(item % H);
            load_head<HS>(qkv + row * ld + which * C + head * HS, x);
            dptr = dqkv + row * ldd + which * C + head * HS;
            load_head<HS>(dptr, dy);
            head_stats<HS>(x, &mean, &rstd);
        } else {
#pragma unroll
            for (int i = 0; i < HS; ++i) { x[i] = 0.f; dy[i] = 0.f; }
        }
#pragma unroll
        for (int i = 0; i < HS; ++i) {
            x[i] = (x[i] - mean) * rstd;
            tw[lane][i] = dy[i] * x[i];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < HS / 32; ++q) {
            float a = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) a += tw[r][q * 32 + lane];
            accg[q] += a;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < HS; ++i) tw[lane][i] = dy[i];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < HS / 32; ++q) {
            float a = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) a += tw[r][q * 32 + lane];
            accb[q] += a;
        }
        __syncwarp();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < HS; ++i) {
            dy[i] *= __ldg(g + i);
            s1 += dy[i];
            s2 = fmaf(dy[i], x[i], s2);
        }
        if (valid) {
            const float c1 = s1 * (1.0f / HS), c2 = s2 * (1.0f / HS);
#pragma unroll
            for (int i = 0; i < HS; ++i) dy[i] = rstd * (dy[i] - c1 - x[i] * c2);
            store_head<HS>(dptr, dy);
        }
    }
#pragma unroll
    for (int i = 0; i < HS / 32; ++i) {
        atomicAdd(dgam + i * 32 + lane, accg[i]);
        if (dbet) atomicAdd(dbet + i * 32 + lane, accb[i]);
    }
}

// ------------------------------------------------------------------------------------------------ attention, one CTA per (jet, head)
template <int HS>
__device__ __forceinline__ void load_rows(uint32_t* dst, const bf16* src, long long ld, int n, int tid) {
    constexpr int W = HS / 2, PW = W + 1;           // words per row, padded pitch (odd: conflict-free column walks)
    for (int idx = tid; idx < n * W; idx += 256) {
        const int i = idx / W, w = idx % W;
        dst[i * PW + w] = *reinterpret_cast<const uint32_t*>(src + i * ld + 2 * w);
    }
}
template <int HS>
__device__ __forceinline__ float dot_rows(const uint32_t* a, const uint32_t* b) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < HS / 2; ++w) {
        const uint32_t x = a[w], y = b[w];
        acc = fmaf(bflo(x), bflo(y), acc);
        acc = fmaf(bfhi(x), bfhi(y), acc);
    }
    return acc;
}
__device__ __forceinline__ float bf_elem(const uint32_t* row, int d) {
    const uint32_t w = row[d >> 1];
    return (d & 1) ? bfhi(w) : bflo(w);
}

template <int HS>
__global__ void __launch_bounds__(256) tr_attn_fwd_kernel(const bf16* __restrict__ qn, long long ldq, const bf16* __restrict__ kn, long long ldk,
                                                          const bf16* __restrict__ v, long long ldv, const int* __restrict__ jet_off,
                                                          const long long* __restrict__ p_off, int H, float scale, int min_n,
                                                          bf16* __restrict__ o, long long ldo, bf16* __restrict__ P) {
    grid_dep_wait();
    grid_dep_launch();
    extern __shared__ uint32_t sm[];
    constexpr int PW = HS / 2 + 1;
    const int jet = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    const int r0 = jet_off[jet], n = jet_off[jet + 1] - r0;
    if (n <= min_n) return;
    uint32_t* Qs = sm;
    uint32_t* Ks = Qs + n * PW;
    uint32_t* Vs = Ks + n * PW;
    float* S = reinterpret_cast<float*>(Vs + n * PW);
    const int ps = n + 1;
    load_rows<HS>(Qs, qn + static_cast<long long>(r0) * ldq + h * HS, ldq, n, tid);
    load_rows<HS>(Ks, kn + static_cast<long long>(r0) * ldk + h * HS, ldk, n, tid);
    load_rows<HS>(Vs, v + static_cast<long long>(r0) * ldv + h * HS, ldv, n, tid);
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += 256) {
        const int i = idx / n, j = idx % n;
        S[i * ps + j] = dot_rows<HS>(Qs + i * PW, Ks + j * PW) * scale;
    }
    __syncthreads();
    bf16* Pg = P + (p_off[jet] * H + static_cast<long long>(h) * n * n);
    const int lane = tid & 31;
    for (int i = tid >> 5; i < n; i += 8) {
        float m = -INFINITY;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, S[i * ps + j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < n; j += 32) { const float e = __expf(S[i * ps + j] - m); S[i * ps + j] = e; s += e; }
        const float inv = 1.0f / warp_sum(s);
        for (int j = lane; j < n; j += 32) {
            const float p = S[i * ps + j] * inv;
            S[i * ps + j] = p;
            Pg[i * n + j] = __float2bfloat16_rn(p);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < n * HS; idx += 256) {
        const int i = idx / HS, d = idx % HS;
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc = fmaf(S[i * ps + j], bf_elem(Vs + j * PW, d), acc);
        o[static_cast<long long>(r0 + i) * ldo + h * HS + d] = __float2bfloat16_rn(acc);
    }
}

template <int HS>
__global__ void __launch_bounds__(256) tr_attn_bwd_kernel(const bf16* __restrict__ dO, long long lddo, const bf16* __restrict__ o, long long ldo,
                                                          const bf16* __restrict__ P, const bf16* __restrict__ qn, long long ldq,
                                                          const bf16* __restrict__ kn, long long ldk, const bf16* __restrict__ v, long long ldv,
                                                          const int* __restrict__ jet_off, const long long* __restrict__ p_off, int H,
                                                          float scale, int min_n, bf16* __restrict__ dqkv, long long ldd, int C) {
    grid_dep_wait();
    grid_dep_launch();
    extern __shared__ uint32_t sm[];
    constexpr int PW = HS / 2 + 1;
    const int jet = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int r0 = jet_off[jet], n = jet_off[jet + 1] - r0;
    if (n <= min_n) return;
    uint32_t* Qs = sm;
    uint32_t* Ks = Qs + n * PW;
    uint32_t* Vs = Ks + n * PW;
    uint32_t* Ds = Vs + n * PW;
    float* S = reinterpret_cast<float*>(Ds + n * PW);
    const int ps = n + 1;
    float* delta = S + n * ps;
    load_rows<HS>(Qs, qn + static_cast<long long>(r0) * ldq + h * HS, ldq, n, tid);
    load_rows<HS>(Ks, kn + static_cast<long long>(r0) * ldk + h * HS, ldk, n, tid);
    load_rows<HS>(Vs, v + static_cast<long long>(r0) * ldv + h * HS, ldv, n, tid);
    load_rows<HS>(Ds, dO + static_cast<long long>(r0) * lddo + h * HS, lddo, n, tid);
    const bf16* Pg = P + (p_off[jet] * H + static_cast<long long>(h) * n * n);
    for (int idx = tid; idx < n * n; idx += 256) S[(idx / n) * ps + (idx % n)] = bf2f(Pg[idx]);
    // delta_i = sum_d dO_id O_id  (= sum_j P_ij dP_ij)
    for (int i = tid >> 5; i < n; i += 8) {
        float s = 0.f;
        for (int d = lane; d < HS; d += 32) s = fmaf(bf2f(dO[static_cast<long long>(r0 + i) * lddo + h * HS + d]), bf2f(o[static_cast<long long>(r0 + i) * ldo + h * HS + d]), s);
        s = warp_sum(s);
        if (lane == 0) delta[i] = s;
    }
    __syncthreads();
    // dV = P^T dO
    for (int idx = tid; idx < n * HS; idx += 256) {
        const int j = idx / HS, d = idx % HS;
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc = fmaf(S[i * ps + j], bf_elem(Ds + i * PW, d), acc);
        dqkv[static_cast<long long>(r0 + j) * ldd + 2 * C + h * HS + d] = __float2bfloat16_rn(acc);
    }
    __syncthreads();
    // dS = P (dP - delta) scale, dP = dO V^T   (in place over P)
    for (int idx = tid; idx < n * n; idx += 256) {
        const int i = idx / n, j = idx % n;
        const float dp = dot_rows<HS>(Ds + i * PW, Vs + j * PW);
        S[i * ps + j] = S[i * ps + j] * (dp - delta[i]) * scale;
    }
    __syncthreads();
    // dQ = dS K, dK = dS^T Q
    for (int idx = tid; idx < n * HS; idx += 256) {
        const int i = idx / HS, d = idx % HS;
        float aq = 0.f, ak = 0.f;
        for (int j = 0; j < n; ++j) {
            aq = fmaf(S[i * ps + j], bf_elem(Ks + j * PW, d), aq);
            ak = fmaf(S[j * ps + i], bf_elem(Qs + j * PW, d), ak);
        }
        dqkv[static_cast<long long>(r0 + i) * ldd + h * HS + d] = __float2bfloat16_rn(aq);
        dqkv[static_cast<long long>(r0 + i) * ldd + C + h * HS + d] = __float2bfloat16_rn(ak);
    }
}

// ------------------------------------------------------------------------------------------------ element-wise
template <bool F32>
__global__ void tr_gelu_fwd_kernel(const void* __restrict__ z_, void* __restrict__ h_, long long n) {
    grid_dep_wait();
    grid_dep_launch();
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
    if (i >= n) return;
    if (F32) {
        const float* z = static_cast<const float*>(z_);
        float* h = static_cast<float*>(h_);
        h[i] = gelu_f(z[i]);
        if (i + 1 < n) h[i + 1] = gelu_f(z[i + 1]);
    } else {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const bf16*>(z_) + i);
        *reinterpret_cast<uint32_t*>(static_cast<bf16*>(h_) + i) = pack2(gelu_f(bflo(w)), gelu_f(bfhi(w)));
    }
}
template <bool F32>
__global__ void tr_gelu_bwd_kernel(const void* __restrict__ dh_, const void* __restrict__ z_, void* __restrict__ dz_, long long n) {
    grid_dep_wait();
    grid_dep_launch();
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
    if (i >= n) return;
    if (F32) {
        const float *dh = static_cast<const float*>(dh_), *z = static_cast<const float*>(z_);
        float* dz = static_cast<float*>(dz_);
        dz[i] = dh[i] * gelu_grad(z[i]);
        if (i + 1 < n) dz[i + 1] = dh[i + 1] * gelu_grad(z[i + 1]);
    } else {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const bf16*>(z_) + i);
        const uint32_t d = *reinterpret_cast<const uint32_t*>(static_cast<const bf16*>(dh_) + i);
        *reinterpret_cast<uint32_t*>(static_cast<bf16*>(dz_) + i) = pack2(bflo(d) * gelu_grad(bflo(w)), bfhi(d) * gelu_grad(bfhi(w)));
    }
}

__global__ void tr_add_kernel(float* __restrict__ out, long long ldo, const float* __restrict__ a, long long lda, const float* __restrict__ y,
                              long long ldy, const float* __restrict__ tadd, long long ldt, const int* __restrict__ row_jet, int M, int C) {
    grid_dep_wait();
    grid_dep_launch();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int cq = C / 4;
    if (i >= static_cast<long long>(M) * cq) return;
    const long long r = i / cq;
    const int c = static_cast<int>(i % cq) * 4;
    float4 s = *reinterpret_cast<const float4*>(a + r * lda + c);
    if (y) { const float4 t = *reinterpret_cast<const float4*>(y + r * ldy + c); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
    if (tadd) {
        const float4 t = *reinterpret_cast<const float4*>(tadd + static_cast<long long>(row_jet ? row_jet[r] : 0) * ldt + c);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(out + r * ldo + c) = s;
}

__global__ void tr_jet_sum_kernel(const float* __restrict__ g, long long ld, const int* __restrict__ jet_off, int C, float* __restrict__ out,
                                  long long ldo, int accumulate) {
    grid_dep_wait();
    grid_dep_launch();
    const int b = blockIdx.x, c = threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int r = jet_off[b]; r < jet_off[b + 1]; ++r) s += g[r * ld + c];
    if (accumulate) s += out[b * ldo + c];
    out[b * ldo + c] = s;
}

// ------------------------------------------------------------------------------------------------ output projections
// vt = hx Wx^T + bx, logits = hy Wy^T + by; h = [hx | hy] bf16 [M, 2 I].  The (3 + V) x I weights sit in shared memory; a warp
// walks rows (grid-stride), lane l owns features 128 j + 4 l .. + 3 (8-byte loads of h, conflict-free float4 loads of W).
template <int V>
__global__ void __launch_bounds__(256) tr_head_fwd_kernel(const bf16* __restrict__ h, long long ldh, int I, const float* __restrict__ wx,
                                                          const float* __restrict__ bx, const float* __restrict__ wy,
                                                          const float* __restrict__ by, int M, float* __restrict__ vt, float* __restrict__ logits) {
    grid_dep_wait();
    grid_dep_launch();
    extern __shared__ float4 head_w[];                 // [(3 + V)][I / 4]
    const int lane = threadIdx.x & 31, I4 = I / 4;
    for (int i = threadIdx.x; i < (3 + V) * I4; i += 256)
        head_w[i] = i < 3 * I4 ? reinterpret_cast<const float4*>(wx)[i] : reinterpret_cast<const float4*>(wy)[i - 3 * I4];
    __syncthreads();
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < M; r += gridDim.x * 8) {
        float ax[3] = {0.f, 0.f, 0.f}, ay[V];
#pragma unroll
        for (int c = 0; c < V; ++c) ay[c] = 0.f;
        for (int j = 0; j < I / 128; ++j) {
            const int i4 = j * 32 + lane;
            const uint2 ux = *reinterpret_cast<const uint2*>(h + r * ldh + 4 * i4), uy = *reinterpret_cast<const uint2*>(h + r * ldh + I + 4 * i4);
            const float hx[4] = {bflo(ux.x), bfhi(ux.x), bflo(ux.y), bfhi(ux.y)}, hy[4] = {bflo(uy.x), bfhi(uy.x), bflo(uy.y), bfhi(uy.y)};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 w = head_w[c * I4 + i4];
                ax[c] = fmaf(hx[0], w.x, fmaf(hx[1], w.y, fmaf(hx[2], w.z, fmaf(hx[3], w.w, ax[c]))));
            }
#pragma unroll
            for (int c = 0; c < V; ++c) {
                const float4 w = head_w[(3 + c) * I4 + i4];
                ay[c] = fmaf(hy[0], w.x, fmaf(hy[1], w.y, fmaf(hy[2], w.z, fmaf(hy[3], w.w, ay[c]))));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float s = warp_sum(ax[c]); if (lane == 0) vt[r * 3 + c] = s + bx[c]; }
#pragma unroll
        for (int c = 0; c < V; ++c) { const float s = warp_sum(ay[c]); if (lane == 0) logits[r * V + c] = s + by[c]; }
    }
}

// thread = hidden unit i of both heads (I / 256 each), CTA = a run of rows:
//   dz[r, i] = (sum_c dout[r, c] W[c, i]) GELU'(z[r, i]);  dW[c, i] += dout[r, c] h[r, i];  db[c] += dout[r, c]
template <int V, int IPT>
__global__ void __launch_bounds__(256) tr_head_bwd_kernel(const float* __restrict__ dvt, const float* __restrict__ dlog, const bf16* __restrict__ h,
                                                          const bf16* __restrict__ z, long long ldh, int I, const float* __restrict__ wx,
                                                          const float* __restrict__ wy, int M, int rows_per_cta, bf16* __restrict__ dz,
                                                          float* __restrict__ dwx, float* __restrict__ dbx, float* __restrict__ dwy,
                                                          float* __restrict__ dby) {
    grid_dep_wait();
    grid_dep_launch();
    __shared__ float sd[3 + V];
    const int tid = threadIdx.x;
    const int r_beg = blockIdx.x * rows_per_cta, r_end = min(M, r_beg + rows_per_cta);
    float wxr[IPT][3], wyr[IPT][V], gx[IPT][3], gy[IPT][V];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int i = tid + q * 256;
#pragma unroll
        for (int c = 0; c < 3; ++c) { wxr[q][c] = wx[c * I + i]; gx[q][c] = 0.f; }
#pragma unroll
        for (int c = 0; c < V; ++c) { wyr[q][c] = wy[c * I + i]; gy[q][c] = 0.f; }
    }
    float gb = 0.f;
    for (int r = r_beg; r < r_end; ++r) {
        __syncthreads();
        if (tid < 3) sd[tid] = dvt[r * 3 + tid];
        else if (tid < 3 + V) sd[tid] = dlog[r * V + tid - 3];
        __syncthreads();
        if (tid < 3 + V) gb += sd[tid];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int i = tid + q * 256;
            const long long ox = r * ldh + i, oy = r * ldh + I + i;
            const float hx = bf2f(h[ox]), hy = bf2f(h[oy]);
            float dx = 0.f, dy = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { dx = fmaf(sd[c], wxr[q][c], dx); gx[q][c] = fmaf(sd[c], hx, gx[q][c]); }
#pragma unroll
            for (int c = 0; c < V; ++c) { dy = fmaf(sd[3 + c], wyr[q][c], dy); gy[q][c] = fmaf(sd[3 + c], hy, gy[q][c]); }
            dz[ox] = __float2bfloat16_rn(dx * gelu_grad(bf2f(z[ox])));
            dz[oy] = __float2bfloat16_rn(dy * gelu_grad(bf2f(z[oy])));
        }
    }
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int i = tid + q * 256;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(dwx + c * I + i, gx[q][c]);
#pragma unroll
        for (int c = 0; c < V; ++c) atomicAdd(dwy + c * I + i, gy[q][c]);
    }
    if (tid < 3) atomicAdd(dbx + tid, gb);
    else if (tid < 3 + V) atomicAdd(dby + tid - 3, gb);
}

// ------------------------------------------------------------------------------------------------ loss
template <int V>
__global__ void tr_loss_fwd_kernel(const float* __restrict__ vt, const float* __restrict__ logits, const float* __restrict__ tgt,
                                   const int* __restrict__ k1, const int* __restrict__ jet_off, int B, float* __restrict__ loss_mse,
                                   float* __restrict__ loss_ce) {
    grid_dep_wait();
    grid_dep_launch();
    const int jet = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (jet >= B) return;
    float mse = 0.f, ce = 0.f;
    const int r0 = jet_off[jet], r1 = jet_off[jet + 1];
    for (int r = r0 + lane; r < r1; r += 32) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float e = vt[r * 3 + c] - tgt[r * 3 + c]; mse += e * e; }
        const int t = k1[r];
        if (t != 0) {                                   // ignore_index = 0
            float l[V], m = -INFINITY;
#pragma unroll
            for (int v = 0; v < V; ++v) { l[v] = logits[r * V + v]; m = fmaxf(m, l[v]); }
            float s = 0.f, lt = l[0];
#pragma unroll
            for (int v = 0; v < V; ++v) { s += expf(l[v] - m); lt = (v == t) ? l[v] : lt; }
            ce += (m + logf(s)) - lt;
        }
    }
    mse = warp_sum(mse);
    ce = warp_sum(ce);
    if (lane == 0) {
        const float dn = fmaxf(static_cast<float>(r1 - r0), 1.0f);
        loss_mse[jet] = mse / dn;
        loss_ce[jet] = ce / dn;
    }
}

__global__ void tr_loss_combine_kernel(const float* __restrict__ loss_mse, const float* __restrict__ loss_ce, const float* __restrict__ u,
                                       int B, float* __restrict__ out5, float* __restrict__ gl1, float* __restrict__ gl2,
                                       float* __restrict__ du) {
    grid_dep_wait();
    grid_dep_launch();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const float inv = 1.0f / static_cast<float>(B);
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (b < B) {
        const float l1 = loss_mse[b], l2 = loss_ce[b];
        float w1 = 1.f, w2 = 1.f, loss;
        if (u) {
            const float u1 = u[b * 2], u2 = u[b * 2 + 1];
            w1 = expf(-u1); w2 = expf(-u2);
            loss = 0.5f * (u1 + w1 * l1) + 0.5f * (u2 + w2 * l2);
            gl1[b] = 0.5f * w1 * inv; gl2[b] = 0.5f * w2 * inv;
            du[b * 2] = 0.5f * (1.0f - w1 * l1) * inv; du[b * 2 + 1] = 0.5f * (1.0f - w2 * l2) * inv;
        } else {
            loss = l1 + l2;
            gl1[b] = inv; gl2[b] = inv;
        }
        v[0] = loss * inv; v[1] = l1 * inv; v[2] = l2 * inv; v[3] = w1 * inv; v[4] = w2 * inv;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float s = warp_sum(v[i]);
        if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(out5 + i, s);
    }
}

template <int V>
__global__ void tr_loss_bwd_kernel(const float* __restrict__ vt, const float* __restrict__ logits, const float* __restrict__ tgt,
                                   const int* __restrict__ k1, const int* __restrict__ row_jet, const int* __restrict__ jet_off,
                                   const float* __restrict__ gl1, const float* __restrict__ gl2, int M, int B, float* __restrict__ dvt,
                                   float* __restrict__ dlog) {
    grid_dep_wait();
    grid_dep_launch();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    if (r >= jet_off[B]) {                              // rows beyond the last jet (a batch padded to a fixed row capacity)
#pragma unroll
        for (int c = 0; c < 3; ++c) dvt[r * 3 + c] = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) dlog[r * V + v] = 0.f;
        return;
    }
    const int b = row_jet[r];
    const float dn = fmaxf(static_cast<float>(jet_off[b + 1] - jet_off[b]), 1.0f);
    const float a1 = 2.0f * gl1[b] / dn, a2 = gl2[b] / dn;
#pragma unroll
    for (int c = 0; c < 3; ++c) dvt[r * 3 + c] = a1 * (vt[r * 3 + c] - tgt[r * 3 + c]);
    const int t = k1[r];
    float l[V], m = -INFINITY, s = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) { l[v] = logits[r * V + v]; m = fmaxf(m, l[v]); }
#pragma unroll
    for (int v = 0; v < V; ++v) { l[v] = expf(l[v] - m); s += l[v]; }
    const float inv = 1.0f / s;
#pragma unroll
    for (int v = 0; v < V; ++v) dlog[r * V + v] = t != 0 ? a2 * (l[v] * inv - (v == t ? 1.0f : 0.0f)) : 0.0f;
}

// ------------------------------------------------------------------------------------------------ optimiser
__global__ void __launch_bounds__(256) tr_sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
    grid_dep_wait();
    grid_dep_launch();
    __shared__ float red[8];
    float s = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        s = fmaf(g[i], g[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}
// second stage, one block, fixed order: the result is bit-reproducible (replicas of a data-parallel run compute the same clip
// coefficient from the same all-reduced gradient and stay bit-identical)
__global__ void __launch_bounds__(256) tr_sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    grid_dep_wait();
    grid_dep_launch();
    __shared__ float red[256];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

__global__ void tr_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                               float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, const float* __restrict__ sumsq,
                               float max_norm, float grad_scale, bf16* __restrict__ p16) {
    grid_dep_wait();
    grid_dep_launch();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float coef = grad_scale;
    if (sumsq && max_norm > 0.f) {                    // torch.nn.utils.clip_grad_norm_: min(1, max_norm / (||g|| + 1e-6))
        const float total = sqrtf(*sumsq) * grad_scale;
        coef *= fminf(1.0f, max_norm / (total + 1e-6f));
    }
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (p16) p16[i] = __float2bfloat16_rn(pi);
}

inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

}  // namespace

// ================================================================================================== launchers
bool tr_pdl_enabled() {
    static const bool on = [] { const char* e = getenv("MMF_TRAIN_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

int launch_tr_sgemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C, long long ldc,
                    int M, int N, int K, const float* bias, int accumulate, cudaStream_t s) {
    if (M <= 0 || N <= 0) return 0;
    MMF_CUDA_OK(tr_launch(tr_sgemm_kernel, dim3(dim3((N + 15) / 16, (M + 15) / 16)), dim3(256), 0, s, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, accumulate));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_cast_transpose(const void* in, long long ld_in, int in_f32, int rows, int cols, bf16* out, long long ld_out, bf16* outT,
                             long long ldT, float* colsum, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return 0;
    const dim3 grid((rows + 31) / 32, (cols + 31) / 32);
    if (in_f32) MMF_CUDA_OK(tr_launch(tr_cast_transpose_kernel<true>, dim3(grid), dim3(256), 0, s, in, ld_in, rows, cols, out, ld_out, outT, ldT, colsum));
    else MMF_CUDA_OK(tr_launch(tr_cast_transpose_kernel<false>, dim3(grid), dim3(256), 0, s, in, ld_in, rows, cols, out, ld_out, outT, ldT, colsum));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_weights_transpose(const float* p, bf16* pT, const TrTransposeJob* jobs_dev, int n_jobs, int n_tiles, cudaStream_t s) {
    if (n_jobs <= 0 || n_tiles <= 0) return 0;
    MMF_CUDA_OK(tr_launch(tr_weights_transpose_kernel, dim3(n_tiles), dim3(256), 0, s, p, pT, jobs_dev, n_jobs));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_pack(const float* xt, const long long* kt, const float* x0, const float* x1, const long long* k1, const int* row_slot, int M,
                   int V, float* xs, int* ks, float* tgt, int* k1p, int* err, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_CUDA_OK(tr_launch(tr_pack_kernel, dim3(blocks_for(M, 256)), dim3(256), 0, s, xt, kt, x0, x1, k1, row_slot, M, V, xs, ks, tgt, k1p, err));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_time_embed(const float* t, const int* perm, int B, int dim, int dup, float* out, long long ld, cudaStream_t s) {
    if (B <= 0) return 0;
    MMF_CUDA_OK(tr_launch(tr_time_embed_kernel, dim3(blocks_for(static_cast<long long>(B) * dim, 256)), dim3(256), 0, s, t, perm, B, dim, dup, out, ld));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_embed_x_fwd(const float* xs, int M, const float* w0, const float* b0, int E, bf16* h, long long ld, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_CUDA_OK(tr_launch(tr_embed_x_fwd_kernel, dim3(blocks_for(static_cast<long long>(M) * E, 256)), dim3(256), 0, s, xs, M, w0, b0, E, h, ld));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_embed_x_bwd(const bf16* dh, long long ld, const float* xs, int M, const float* w0, const float* b0, int E, float* dw0,
                          float* db0, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_REQUIRE(E <= 1024, "embed_x_bwd: n_embd up to 1024");
    const int rpc = 64;
    MMF_CUDA_OK(tr_launch(tr_embed_x_bwd_kernel, dim3(blocks_for(M, rpc)), dim3(E), 0, s, dh, ld, xs, M, w0, b0, E, rpc, dw0, db0));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_embed_y_fwd(const int* ks, int M, const float* emb, int E, int V, bf16* g, long long ld, cudaStream_t s) {
    if (M <= 0) return 0;
    (void)V;
    MMF_CUDA_OK(tr_launch(tr_embed_y_fwd_kernel, dim3(blocks_for(static_cast<long long>(M) * E, 256)), dim3(256), 0, s, ks, M, emb, E, g, ld));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_embed_y_bwd(const bf16* dg, long long ld, const int* ks, int M, const float* emb, int E, int V, float* demb, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_REQUIRE(V == 9 && E <= 1024, "embed_y_bwd is instantiated for vocab_size 9");
    const int rpc = 64;
    MMF_CUDA_OK(tr_launch(tr_embed_y_bwd_kernel<9>, dim3(blocks_for(M, rpc)), dim3(E), 0, s, dg, ld, ks, M, emb, E, rpc, demb));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_ln_fwd(const TrLnArgs& a, cudaStream_t s) {
    if (a.M <= 0) return 0;
    MMF_REQUIRE(a.C == 128 || a.C == 256, "layernorm: width 128 or 256");
    if (a.C == 128) MMF_CUDA_OK(tr_launch(tr_ln_fwd_kernel<128>, dim3(blocks_for(a.M, 8)), dim3(256), 0, s, a));
    else MMF_CUDA_OK(tr_launch(tr_ln_fwd_kernel<256>, dim3(blocks_for(a.M, 8)), dim3(256), 0, s, a));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_ln_bwd(const TrLnBwdArgs& a, cudaStream_t s) {
    if (a.M <= 0) return 0;
    MMF_REQUIRE(a.C == 128 || a.C == 256, "layernorm: width 128 or 256");
    const unsigned grid = std::min<unsigned>(blocks_for(a.M, 8), 148 * 4);
    if (a.C == 128) MMF_CUDA_OK(tr_launch(tr_ln_bwd_kernel<128>, dim3(grid), dim3(256), 0, s, a));
    else MMF_CUDA_OK(tr_launch(tr_ln_bwd_kernel<256>, dim3(grid), dim3(256), 0, s, a));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_qkln_fwd(const bf16* qkv, long long ld, int M, int C, int H, const float* qg, const float* qb, const float* kg, const float* kb,
                       bf16* qn, bf16* kn, long long ldn, cudaStream_t s) {
    if (M <= 0) return 0;
    const int hs = C / H;
    MMF_REQUIRE(hs == 32 || hs == 64, "q/k LayerNorm: head size 32 or 64");
    const unsigned grid = blocks_for(static_cast<long long>(M) * 2 * H, 128);
    if (hs == 32) MMF_CUDA_OK(tr_launch(tr_qkln_fwd_kernel<32>, dim3(grid), dim3(128), 0, s, qkv, ld, M, C, H, qg, qb, kg, kb, qn, kn, ldn));
    else MMF_CUDA_OK(tr_launch(tr_qkln_fwd_kernel<64>, dim3(grid), dim3(128), 0, s, qkv, ld, M, C, H, qg, qb, kg, kb, qn, kn, ldn));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_qkln_bwd(bf16* dqkv, long long ldd, const bf16* qkv, long long ld, int M, int C, int H, const float* qg, const float* kg,
                       float* dqg, float* dqb, float* dkg, float* dkb, cudaStream_t s) {
    if (M <= 0) return 0;
    const int hs = C / H;
    MMF_REQUIRE(hs == 32 || hs == 64, "q/k LayerNorm: head size 32 or 64");
    const long long chunks = (static_cast<long long>(M) * H + 31) / 32;
    const dim3 grid(static_cast<unsigned>(std::min<long long>((chunks + 3) / 4, 148 * 8)), 2);
    if (hs == 32) MMF_CUDA_OK(tr_launch(tr_qkln_bwd_kernel<32>, dim3(grid), dim3(128), 0, s, dqkv, ldd, qkv, ld, M, C, H, qg, kg, dqg, dqb, dkg, dkb));
    else MMF_CUDA_OK(tr_launch(tr_qkln_bwd_kernel<64>, dim3(grid), dim3(128), 0, s, dqkv, ldd, qkv, ld, M, C, H, qg, kg, dqg, dqb, dkg, dkb));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

static int attn_smem_bytes(int hs, int n, int n_operands, bool with_delta) {
    return n_operands * n * (hs / 2 + 1) * 4 + n * (n + 1) * 4 + (with_delta ? n * 4 : 0);
}
template <int TAG, typename K>                            // TAG: one static table per kernel instantiation (the function TYPES coincide)
static int attn_configure(K kernel, int bytes) {
    static int configured[64] = {0};                       // per device: largest size opted in so far
    int dev = 0;
    MMF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || configured[dev] < bytes) {
        MMF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (dev >= 0 && dev < 64) configured[dev] = bytes;
    }
    return 0;
}

int launch_tr_attn_fwd(const bf16* qn, long long ldq, const bf16* kn, long long ldk, const bf16* v, long long ldv, const int* jet_off,
                       const long long* p_off, int B, int H, int hs, int nmax, int min_n, bf16* o, long long ldo, bf16* P, cudaStream_t s) {
    if (B <= 0 || nmax <= 0) return 0;
    MMF_REQUIRE(hs == 32 || hs == 64, "attention: head size 32 or 64");
    MMF_REQUIRE(nmax <= 176, "attention: jets of up to 176 particles");
    const int bytes = attn_smem_bytes(hs, nmax, 3, false);
    const float scale = 1.0f / sqrtf(static_cast<float>(hs));
    if (hs == 32) {
        if (attn_configure<0>(tr_attn_fwd_kernel<32>, bytes)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_fwd_kernel<32>, dim3(dim3(B, H)), dim3(256), bytes, s, qn, ldq, kn, ldk, v, ldv, jet_off, p_off, H, scale, min_n, o, ldo, P));
    } else {
        if (attn_configure<1>(tr_attn_fwd_kernel<64>, bytes)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_fwd_kernel<64>, dim3(dim3(B, H)), dim3(256), bytes, s, qn, ldq, kn, ldk, v, ldv, jet_off, p_off, H, scale, min_n, o, ldo, P));
    }
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_attn_bwd(const bf16* dO, long long lddo, const bf16* o, long long ldo, const bf16* P, const bf16* qn, long long ldq,
                       const bf16* kn, long long ldk, const bf16* v, long long ldv, const int* jet_off, const long long* p_off, int B, int H,
                       int hs, int nmax, int min_n, bf16* dqkv, long long ldd, int C, cudaStream_t s) {
    if (B <= 0 || nmax <= 0) return 0;
    MMF_REQUIRE(hs == 32 || hs == 64, "attention: head size 32 or 64");
    MMF_REQUIRE(nmax <= 176, "attention: jets of up to 176 particles");
    const int bytes = attn_smem_bytes(hs, nmax, 4, true);
    const float scale = 1.0f / sqrtf(static_cast<float>(hs));
    if (hs == 32) {
        if (attn_configure<2>(tr_attn_bwd_kernel<32>, bytes)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_bwd_kernel<32>, dim3(dim3(B, H)), dim3(256), bytes, s, dO, lddo, o, ldo, P, qn, ldq, kn, ldk, v, ldv, jet_off, p_off, H, scale, min_n, dqkv, ldd, C));
    } else {
        if (attn_configure<3>(tr_attn_bwd_kernel<64>, bytes)) return 1;
        MMF_CUDA_OK(tr_launch(tr_attn_bwd_kernel<64>, dim3(dim3(B, H)), dim3(256), bytes, s, dO, lddo, o, ldo, P, qn, ldq, kn, ldk, v, ldv, jet_off, p_off, H, scale, min_n, dqkv, ldd, C));
    }
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_gelu_fwd(const void* z, void* h, long long n, int f32, cudaStream_t s) {
    if (n <= 0) return 0;
    MMF_REQUIRE(f32 || n % 2 == 0, "gelu: bf16 arrays hold an even number of elements");
    if (f32) MMF_CUDA_OK(tr_launch(tr_gelu_fwd_kernel<true>, dim3(blocks_for((n + 1) / 2, 256)), dim3(256), 0, s, z, h, n));
    else MMF_CUDA_OK(tr_launch(tr_gelu_fwd_kernel<false>, dim3(blocks_for(n / 2, 256)), dim3(256), 0, s, z, h, n));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_gelu_bwd(const void* dh, const void* z, void* dz, long long n, int f32, cudaStream_t s) {
    if (n <= 0) return 0;
    MMF_REQUIRE(f32 || n % 2 == 0, "gelu: bf16 arrays hold an even number of elements");
    if (f32) MMF_CUDA_OK(tr_launch(tr_gelu_bwd_kernel<true>, dim3(blocks_for((n + 1) / 2, 256)), dim3(256), 0, s, dh, z, dz, n));
    else MMF_CUDA_OK(tr_launch(tr_gelu_bwd_kernel<false>, dim3(blocks_for(n / 2, 256)), dim3(256), 0, s, dh, z, dz, n));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_add(float* out, long long ldo, const float* a, long long lda, const float* y, long long ldy, const float* tadd, long long ldt,
                  const int* row_jet, int M, int C, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_REQUIRE(C % 4 == 0, "add: width must be a multiple of 4");
    MMF_CUDA_OK(tr_launch(tr_add_kernel, dim3(blocks_for(static_cast<long long>(M) * (C / 4), 256)), dim3(256), 0, s, out, ldo, a, lda, y, ldy, tadd, ldt, row_jet, M, C));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_jet_sum(const float* g, long long ld, const int* jet_off, int B, int C, float* out, long long ldo, int accumulate, cudaStream_t s) {
    if (B <= 0) return 0;
    MMF_REQUIRE(C <= 1024, "jet_sum: width up to 1024");
    MMF_CUDA_OK(tr_launch(tr_jet_sum_kernel, dim3(B), dim3(C), 0, s, g, ld, jet_off, C, out, ldo, accumulate));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_head_fwd(const bf16* h, long long ldh, int I, const float* wx, const float* bx, const float* wy, const float* by, int V, int M,
                       float* vt, float* logits, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_REQUIRE(V == 9 && I % 128 == 0 && I <= 768, "head kernels are instantiated for vocab_size 9 and n_inner a multiple of 128 up to 768");
    const unsigned grid = std::min<unsigned>(blocks_for(M, 8), 148 * 2);
    MMF_CUDA_OK(tr_launch(tr_head_fwd_kernel<9>, dim3(grid), dim3(256), static_cast<size_t>(12) * I * 4, s, h, ldh, I, wx, bx, wy, by, M, vt, logits));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_head_bwd(const float* dvt, const float* dlog, const bf16* h, const bf16* z, long long ldh, int I, const float* wx,
                       const float* wy, int V, int M, bf16* dz, float* dwx, float* dbx, float* dwy, float* dby, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_REQUIRE(V == 9 && I == 512, "head kernels are instantiated for vocab_size 9 and n_inner 512");
    const int rpc = 32;
    MMF_CUDA_OK(tr_launch(tr_head_bwd_kernel<9, 2>, dim3(blocks_for(M, rpc)), dim3(256), 0, s, dvt, dlog, h, z, ldh, I, wx, wy, M, rpc, dz, dwx, dbx, dwy, dby));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_loss_fwd(const float* vt, const float* logits, const float* tgt, const int* k1, const int* jet_off, int B, int V,
                       float* loss_mse, float* loss_ce, cudaStream_t s) {
    if (B <= 0) return 0;
    MMF_REQUIRE(V == 9, "the loss kernels are instantiated for vocab_size 9");
    MMF_CUDA_OK(tr_launch(tr_loss_fwd_kernel<9>, dim3(blocks_for(static_cast<long long>(B) * 32, 256)), dim3(256), 0, s, vt, logits, tgt, k1, jet_off, B, loss_mse, loss_ce));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_loss_combine(const float* loss_mse, const float* loss_ce, const float* u, int B, float* out5, float* gl1, float* gl2, float* du,
                           cudaStream_t s) {
    MMF_CUDA_OK(cudaMemsetAsync(out5, 0, 5 * sizeof(float), s));
    if (B <= 0) return 0;
    MMF_CUDA_OK(tr_launch(tr_loss_combine_kernel, dim3(blocks_for(B, 256)), dim3(256), 0, s, loss_mse, loss_ce, u, B, out5, gl1, gl2, du));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_loss_bwd(const float* vt, const float* logits, const float* tgt, const int* k1, const int* row_jet, const int* jet_off,
                       const float* gl1, const float* gl2, int M, int B, int V, float* dvt, float* dlog, cudaStream_t s) {
    if (M <= 0) return 0;
    MMF_REQUIRE(V == 9, "the loss kernels are instantiated for vocab_size 9");
    MMF_CUDA_OK(tr_launch(tr_loss_bwd_kernel<9>, dim3(blocks_for(M, 256)), dim3(256), 0, s, vt, logits, tgt, k1, row_jet, jet_off, gl1, gl2, M, B, dvt, dlog));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tr_sumsq(const float* g, long long n, float* out, cudaStream_t s) {
    // out[0] = result, out[1 .. kSumsqScratch) = per-block partial sums
    const int blocks = n <= 0 ? 0 : static_cast<int>(std::min<long long>((n + 255) / 256, kSumsqScratch - 1));
    if (blocks > 0) MMF_CUDA_OK(tr_launch(tr_sumsq_kernel, dim3(blocks), dim3(256), 0, s, g, n, out + 1));
    MMF_CUDA_OK(tr_launch(tr_sumsq_final_kernel, dim3(1), dim3(256), 0, s, static_cast<const float*>(out + 1), blocks, out));
    return 0;
}

int launch_tr_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, int step,
                   const float* sumsq, float max_norm, float grad_scale, bf16* p16, cudaStream_t s) {
    if (n <= 0) return 0;
    MMF_REQUIRE(step >= 1, "adam: steps count from 1");
    const float bc1 = 1.0f - static_cast<float>(pow(static_cast<double>(beta1), step));
    const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
    MMF_CUDA_OK(tr_launch(tr_adam_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, s, p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, sumsq, max_norm, grad_scale, p16));
    MMF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace mmf
