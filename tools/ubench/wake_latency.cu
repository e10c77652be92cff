// Micro-benchmark (measurement aid): wake-up latency of an mbarrier wait, spinning try_wait vs try_wait with a
// suspend-time hint, when the arrival comes from another warp of the same CTA or from a tcgen05.commit.
#include <cstdio>
#include "../../multimodal-flows_b200/csrc/mmf_ptx.cuh"
using namespace mmf;

__device__ __forceinline__ bool try_wait_nohint(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// mode 0: spin on try_wait (no hint); 1: try_wait with hint (parked); 2: spin on test_wait
__global__ void wake_kernel(int mode, int delay, long long* out) {
    __shared__ uint64_t bar[2];
    __shared__ long long t_arrive;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncthreads();
    long long acc = 0;
    for (int it = 0; it < 64; ++it) {
        const uint32_t par = it & 1;
        if (warp == 0) {
            // waiter
            if (mode == 0) { while (!try_wait_nohint(&bar[0], par)) {} }
            else if (mode == 1) { while (!mbar_try_wait(&bar[0], par)) {} }
            else { while (!test_wait(&bar[0], par)) {} }
            const long long t = clock64();
            __syncwarp();
            if (lane == 0) acc += t - *reinterpret_cast<volatile long long*>(&t_arrive);
            if (lane == 0) mbar_arrive(&bar[1]);
        } else if (warp == 1) {
            const long long t0 = clock64();
            while (clock64() - t0 < delay) {}
            if (lane == 0) { *reinterpret_cast<volatile long long*>(&t_arrive) = clock64(); __threadfence_block(); mbar_arrive(&bar[0]); }
            while (!try_wait_nohint(&bar[1], par)) {}
        }
    }
    if (threadIdx.x == 0) out[0] = acc / 64;
}

int main() {
    long long* out; cudaMalloc(&out, 8);
    for (int mode = 0; mode < 3; ++mode)
        for (int delay : {200, 2000, 20000}) {
            wake_kernel<<<1, 64>>>(mode, delay, out);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("mode %d (%s) delay %6d: wake latency %lld cycles\n", mode, mode == 0 ? "try_wait spin" : mode == 1 ? "try_wait+hint" : "test_wait spin", delay, h);
        }
    return 0;
}
