// Host-side helpers shared by the translation units that pack checkpoints (model.cu, epic_model.cu).
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mmf_b200.h"
#include "mmf_common.cuh"

namespace mmf {

#define MMF_TRY_RC(expr)          \
    do {                         \
        int _rc = (expr);        \
        if (_rc != 0) return _rc; \
    } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

inline uint16_t f32_to_bf16_bits(float f) {          // round to nearest even, as __float2bfloat16_rn
    uint32_t x;
    memcpy(&x, &f, 4);
    if ((x & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((x >> 16) | 0x40);
    const uint32_t lsb = (x >> 16) & 1u;
    x += 0x7fffu + lsb;
    return static_cast<uint16_t>(x >> 16);
}

// ---------------------------------------------------------------------------------------------
// device memory helpers
// ---------------------------------------------------------------------------------------------
struct DeviceArena {            // one allocation, bump sub-allocation, 256-byte aligned
    uint8_t* base = nullptr;
    size_t cap = 0, used = 0;
    std::vector<uint8_t> staging;
    size_t reserve(size_t bytes) {
        const size_t off = (staging.size() + 255) / 256 * 256;
        staging.resize(off + bytes, 0);
        return off;
    }
    size_t put_f32(const std::vector<float>& v) {
        const size_t off = reserve(v.size() * 4);
        memcpy(staging.data() + off, v.data(), v.size() * 4);
        return off;
    }
    size_t put_bf16(const std::vector<float>& v) {
        const size_t off = reserve(v.size() * 2);
        uint16_t* d = reinterpret_cast<uint16_t*>(staging.data() + off);
        for (size_t i = 0; i < v.size(); ++i) d[i] = f32_to_bf16_bits(v[i]);
        return off;
    }
    int upload() {
        cap = staging.size();
        MMF_CUDA_OK(cudaMalloc(&base, cap ? cap : 256));
        MMF_CUDA_OK(cudaMemcpy(base, staging.data(), cap, cudaMemcpyHostToDevice));
        staging.clear();
        staging.shrink_to_fit();
        return 0;
    }
    template <typename T> T* at(size_t off) const { return reinterpret_cast<T*>(base + off); }
    void release() { if (base) cudaFree(base); base = nullptr; }
};

// Pinned staging for the small per-call tables (tile plans, time tables): the H2D copies are truly asynchronous and the
// host vectors they came from may go out of scope at once.  One buffer per owner; a call first waits (host side) until the
// copies of the previous call have left the buffer - they are queued ahead of that call's kernel, so in steady state the
// wait returns immediately and the entry points never synchronise with running kernels.
struct PinnedStage {
    uint8_t* base = nullptr;
    size_t cap = 0, used = 0;
    cudaEvent_t ev = nullptr;
    bool pending = false;
    int begin(size_t bytes) {
        if (pending) { MMF_CUDA_OK(cudaEventSynchronize(ev)); pending = false; }
        if (!ev) MMF_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        if (bytes > cap) {
            if (base) MMF_CUDA_OK(cudaFreeHost(base));
            base = nullptr;
            cap = bytes + bytes / 4 + 4096;
            MMF_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&base), cap, cudaHostAllocDefault));
        }
        used = 0;
        return 0;
    }
    // copies `bytes` from src into the stage and queues the H2D copy to dst on `s`
    int push(void* dst, const void* src, size_t bytes, cudaStream_t s) {
        if (bytes == 0) return 0;
        const size_t off = (used + 15) / 16 * 16;
        if (off + bytes > cap) { set_last_error("pinned stage overflow"); return 2; }
        memcpy(base + off, src, bytes);
        used = off + bytes;
        MMF_CUDA_OK(cudaMemcpyAsync(dst, base + off, bytes, cudaMemcpyHostToDevice, s));
        return 0;
    }
    int end(cudaStream_t s) {
        MMF_CUDA_OK(cudaEventRecord(ev, s));
        pending = true;
        return 0;
    }
    void release() {
        if (ev) { cudaEventSynchronize(ev); cudaEventDestroy(ev); ev = nullptr; }
        if (base) cudaFreeHost(base);
        base = nullptr;
    }
};

struct WeightMap {
    std::unordered_map<std::string, const MmfWeightRef*> m;
    std::string missing;
    const MmfWeightRef* find(const std::string& name) const {
        auto it = m.find(name);
        return it == m.end() ? nullptr : it->second;
    }
    // returns a copy; records the first missing / mis-shaped parameter
    std::vector<float> get(const std::string& name, int64_t d0, int64_t d1 = -1, bool optional = false) {
        const MmfWeightRef* w = find(name);
        const int64_t n = d0 * (d1 < 0 ? 1 : d1);
        if (!w) {
            if (!optional && missing.empty()) missing = "missing parameter " + name;
            return std::vector<float>(static_cast<size_t>(n), 0.f);
        }
        int64_t have = 1;
        for (int i = 0; i < w->ndim; ++i) have *= w->shape[i];
        const bool ok = have == n && w->shape[0] == d0 && (d1 < 0 || w->ndim < 2 || w->shape[1] == d1);
        if (!ok) {
            if (missing.empty()) missing = "parameter " + name + " has an unexpected shape";
            return std::vector<float>(static_cast<size_t>(n), 0.f);
        }
        return std::vector<float>(w->data, w->data + n);
    }
    bool has(const std::string& name) const { return find(name) != nullptr; }
};

inline void append(std::vector<float>& dst, const std::vector<float>& src) { dst.insert(dst.end(), src.begin(), src.end()); }

// sin/cos time features (reference utils/models.py:62-75), fp32 like the reference
inline void sincos_row(float t, int dim, float* out) {
    const int half = dim / 2;
    const float scale = std::log(10000.0f) / static_cast<float>(half - 1);
    for (int i = 0; i < half; ++i) {
        const float f = std::exp(static_cast<float>(i) * -scale);
        const float a = t * f;
        out[i] = std::sin(a);
        out[half + i] = std::cos(a);
    }
}

}  // namespace mmf
