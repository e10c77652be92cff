// Shared host/device helpers: deterministic fp32 math for the step, Philox counter RNG, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <string>

namespace mmf {

constexpr int kMaxV = 16;        // largest vocabulary the step kernels hold in registers
constexpr int kDC = 3;           // continuous features per particle (pT, eta_rel, phi_rel)

// ---------------------------------------------------------------------------------------------
// errors: every extern "C" entry returns int, message kept thread-local (include/mmf_b200.h)
// ---------------------------------------------------------------------------------------------
void set_last_error(const std::string& msg);
#define MMF_CUDA_OK(expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) {                                                                            \
            ::mmf::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +        \
                                  __FILE__ + ":" + std::to_string(__LINE__) + ")");                         \
            return 1;                                                                                       \
        }                                                                                                   \
    } while (0)
#define MMF_REQUIRE(cond, msg)                                                                              \
    do {                                                                                                    \
        if (!(cond)) {                                                                                      \
            ::mmf::set_last_error(std::string(msg) + " [" #cond "]");                                       \
            return 2;                                                                                       \
        }                                                                                                   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Deterministic exp: individually rounded IEEE binary32 ops in a fixed order (no MUFU), so that a
// CPU can reproduce every jump threshold bit for bit.  The specification is
//   n = rint(x log2e); r = fma(n,-ln2_hi,x); r = fma(n,-ln2_lo,r); Taylor-7 Horner in fma; scale by 2^n
// (subnormal results are scaled in two exact steps); x <= -104 -> 0.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float det_bits_to_float(uint32_t b) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ float det_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);          // never contracted into an fma
#else
    volatile float r = a * b;        // host build uses -ffp-contract=off as well; volatile is belt and braces
    return r;
#endif
}
__host__ __device__ __forceinline__ float det_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
__host__ __device__ __forceinline__ float det_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b;
    return r;
#endif
}
__host__ __device__ __forceinline__ float det_expf(float x) {
    if (!(x > -104.0f)) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float n = rintf(det_mul(x, 1.44269502e+00f));
    float r = fmaf(n, -6.93145752e-01f, x);
    r = fmaf(n, -1.42860677e-06f, r);
    float p = 1.98412701e-04f;
    p = fmaf(p, r, 1.38888892e-03f);
    p = fmaf(p, r, 8.33333377e-03f);
    p = fmaf(p, r, 4.16666679e-02f);
    p = fmaf(p, r, 1.66666672e-01f);
    p = fmaf(p, r, 5.00000000e-01f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    const int ni = static_cast<int>(n);
    if (ni >= -126) return det_mul(p, det_bits_to_float(static_cast<uint32_t>(ni + 127) << 23));
    return det_mul(det_mul(p, det_bits_to_float(static_cast<uint32_t>(ni + 127 + 64) << 23)), 5.42101086e-20f);
}

// thermostat weight w = exp(-V beta (1-t)) and coefficient w V / (1-w)
// (reference utils/thermostats.py:20-27, model/MJB.py:189-193)
__host__ __device__ __forceinline__ void det_thermostat(float t, float beta, int V, float* w, float* coef) {
    const float a = static_cast<float>(-static_cast<double>(V) * static_cast<double>(beta));
    const float ww = det_expf(det_mul(a, det_add(1.0f, -t)));
    *w = ww;
    *coef = det_div(det_mul(ww, static_cast<float>(V)), det_add(1.0f, -ww));
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (particle slot, step, lane block, 0),
// key = seed.  Output is independent of launch geometry and of how jets are sharded over GPUs.
// ---------------------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}
__host__ __device__ __forceinline__ Philox4 philox4x32_10(Philox4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = Philox4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
// 24-bit uniform in [0,1), the same construction torch.rand uses for float32
__host__ __device__ __forceinline__ float u01_from_bits(uint32_t b) { return static_cast<float>(b >> 8) * 5.96046448e-08f; }

// V uniforms for (global particle slot, step): blocks of 4 from successive counters
__host__ __device__ __forceinline__ void philox_uniforms(uint64_t seed, uint64_t slot, uint32_t step, int V, float* u) {
    const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    for (int b = 0; b * 4 < V; ++b) {
        Philox4 r = philox4x32_10(Philox4{static_cast<uint32_t>(slot), static_cast<uint32_t>(slot >> 32), step,
                                          static_cast<uint32_t>(b)}, k0, k1);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        for (int j = 0; j < 4 && b * 4 + j < V; ++j) u[b * 4 + j] = u01_from_bits(w[j]);
    }
}

}  // namespace mmf
