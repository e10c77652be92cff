// Micro-benchmark (measurement aid): tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps.
#include <cstdio>
#include "../../multimodal-flows_b200/csrc/mmf_ptx.cuh"
using namespace mmf;

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float* v) { tmem_ld32(taddr, v); tmem_ld32(taddr + 32, v + 32); }

// mode 0: ld x32 + wait per iteration; 1: two ld x32 then wait; 2: st x32 + wait; 3: ld x32, wait, st x32, wait
__global__ void __launch_bounds__(512, 1) tmem_kernel(int mode, int iters, long long* out) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = threadIdx.x + i;
    tmem_st32(taddr, v); tmem_st32(taddr + 32, v + 32); tmem_st_wait();
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) { tmem_ld32(taddr + (it & 1) * 32, v); tmem_ld_wait(); acc += v[it & 31]; }
        else if (mode == 1) { tmem_ld64(taddr, v); tmem_ld_wait(); acc += v[it & 63]; }
        else if (mode == 2) { v[0] = acc; tmem_st32(taddr + (it & 1) * 32, v); tmem_st_wait(); acc += 1.f; }
        else { tmem_ld32(taddr, v); tmem_ld_wait(); v[1] += 1.f; tmem_st32(taddr, v); tmem_st_wait(); acc += v[0]; }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 12345.f) out[1] = 1;
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* out; cudaMalloc(&out, 16);
    const int iters = 2000;
    printf("mode warps | cycles/iter/warp | B/clk/SM\n");
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            tmem_kernel<<<1, warps * 32>>>(mode, iters, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("fail %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            const double bytes_per_iter = (mode == 1 ? 8192.0 : mode == 3 ? 8192.0 : 4096.0) * warps;
            printf("%d %2d | %8.1f | %8.1f\n", mode, warps, double(h) / iters, bytes_per_iter * iters / double(h));
        }
    return 0;
}
