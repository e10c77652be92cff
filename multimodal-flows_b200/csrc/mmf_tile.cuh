// Device helpers shared by the tensor-core kernels: exact-erf GELU and swizzled shared-memory staging rows.
#pragma once
#include "mmf_ptx.cuh"

namespace mmf {

// MUFU wrappers: one instruction each, no slow-path branch (arguments here are never denormal / special)
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float gelu_erf(float x) {
    // 0.5 x (1 + erf(x / sqrt 2)); erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7), two MUFU ops, branch-free
    const float z = fabsf(x) * 0.70710678f;
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = fmaf(-p * t, ex2_approx(-1.44269504f * z * z), 1.0f);
    return 0.5f * x * (1.0f + copysignf(e, x));
}

// GELU through its tanh form with one MUFU.TANH: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))).  It differs from the
// erf form by at most 4.7e-4 absolute (plus 2^-11 relative from tanh.approx), an order of magnitude below the bf16 rounding
// of the value it produces; used where the SIMT epilogue is the bottleneck (MMF_TILE_GELU_EXACT restores the erf form).
__device__ __forceinline__ float gelu_tanh(float x) {
    const float u = x * fmaf(x * x, 0.0356774081f, 0.7978845608f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
// Packed fp32 arithmetic (sm_100 add / mul / fma.rn.f32x2 on an aligned register pair): one instruction for two elements,
// each lane rounded exactly like the scalar instruction.  The row-per-thread epilogues are bound by instruction issue and
// dependent-issue latency (two warps per scheduler), so halving the instruction count of their element-wise passes pays.
__device__ __forceinline__ float2 f2dup(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#define MMF_V2(v, i) make_float2((v)[(i)], (v)[(i) + 1])
#define MMF_SET2(v, i, expr) do { const float2 _r2 = (expr); (v)[(i)] = _r2.x; (v)[(i) + 1] = _r2.y; } while (0)

__device__ __forceinline__ float2 gelu_tanh2(float2 x) {     // gelu_tanh on a pair: five packed instructions + two MUFU.TANH
    const float2 u = f2mul(x, f2fma(f2mul(x, x), f2dup(0.0356774081f), f2dup(0.7978845608f)));
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
    const float2 hx = f2mul(x, f2dup(0.5f));
    return f2fma(hx, t, hx);
}
#ifdef MMF_TILE_GELU_EXACT
__device__ __forceinline__ float2 gelu_tile2(float2 x) { return make_float2(gelu_erf(x.x), gelu_erf(x.y)); }
#else
__device__ __forceinline__ float2 gelu_tile2(float2 x) { return gelu_tanh2(x); }
#endif
// 2 GELU(x) = x + x tanh(..): four packed instructions + two MUFU.TANH.  For activations whose consumer is a linear layer: the
// factor 0.5 moves into that layer's weights on the host, exactly (a power of two).
__device__ __forceinline__ float2 gelu2x_tanh2(float2 x) {
    const float2 u = f2mul(x, f2fma(f2mul(x, x), f2dup(0.0356774081f), f2dup(0.7978845608f)));
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
    return f2fma(x, t, x);
}
#ifdef MMF_TILE_GELU_EXACT
__device__ __forceinline__ float2 gelu2x_tile2(float2 x) { return make_float2(2.0f * gelu_erf(x.x), 2.0f * gelu_erf(x.y)); }
#else
__device__ __forceinline__ float2 gelu2x_tile2(float2 x) { return gelu2x_tanh2(x); }
#endif
#ifdef MMF_TILE_GELU_EXACT
__device__ __forceinline__ float gelu_tile(float x) { return gelu_erf(x); }
#else
__device__ __forceinline__ float gelu_tile(float x) { return gelu_tanh(x); }
#endif

__device__ __forceinline__ void st_shared_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(const void* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(smem_u32(p)) : "memory");
    return v;
}

// write 64 fp32 values of row r as bf16 into a [128][128 B] swizzled staging chunk
__device__ __forceinline__ void stage_row_bf16(uint8_t* chunk, int r, const float* v) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        st_shared_v4(chunk + sw128_offset(r, u), pack_bf16x2(v[u * 8 + 0], v[u * 8 + 1]),
                     pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]), pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]),
                     pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]));
    }
}
// write 32 fp32 values of row r into a [128][128 B] swizzled staging chunk
__device__ __forceinline__ void stage_row_f32(uint8_t* chunk, int r, const float* v) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        st_shared_v4(chunk + sw128_offset(r, u), __float_as_uint(v[u * 4 + 0]), __float_as_uint(v[u * 4 + 1]),
                     __float_as_uint(v[u * 4 + 2]), __float_as_uint(v[u * 4 + 3]));
    }
}

__device__ __forceinline__ float leaky_relu(float x) { return x > 0.f ? x : 0.01f * x; }   // F.leaky_relu default slope
// the same on a pair: max(x, 0.01 x) equals the select for every finite x (one packed multiply + two FMNMX)
__device__ __forceinline__ float2 leaky_relu2(float2 x) {
    const float2 y = f2mul(x, f2dup(0.01f));
    return make_float2(fmaxf(x.x, y.x), fmaxf(x.y, y.y));
}

}  // namespace mmf
