// Micro-benchmark (measurement aid, not product code): tcgen05.mma k-tile rate of ONE CTA per SM with both operands in
// shared memory, the B operand streamed through a bulk-copy ring (as in the tile kernels), optionally with extra
// shared-memory traffic from "epilogue" warps.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../multimodal-flows_b200/csrc/mmf_ptx.cuh"
using namespace mmf;

struct Args { const uint8_t* src; size_t src_bytes; int n, stages, ntiles, producers, use_ring, epi_traffic, issuers; unsigned long long* cyc; };
constexpr int kTile = 32768;   // stage stride

__global__ void __launch_bounds__(384, 1) mma_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    uint64_t* done = full + 16;
    uint32_t* tbase = reinterpret_cast<uint32_t*>(smem + 256);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 512);
    uint8_t* abuf = smem + 1024;                 // 16 KB A chunk
    uint8_t* scratch = abuf + 16384;             // 32 KB epilogue scratch
    uint8_t* ring = scratch + 32768;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1); mbar_init(done + 1, 1);
        *stop = 0;
        fence_mbar_init();
    }
    if (warp == 9) { tmem_alloc(tbase, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tbase;
    const int tile_bytes = a.n * 128;
    const long long t0 = clock64();
    if (warp >= 10 && warp < 10 + a.producers && a.use_ring) {
        for (uint32_t it = warp - 10; it < (uint32_t)a.ntiles; it += a.producers) {
            const uint32_t s = it & 3;
            if (it >= 4u) mbar_wait(&empty[s], ((it >> 2) - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(&full[s], tile_bytes);
                const size_t off = (static_cast<size_t>(it) * 32768u) & ((8u << 20) - 1);
                bulk_load_1d(ring + s * kTile, a.src + off, tile_bytes, &full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 9 || (warp == 8 && a.issuers == 2)) {
        const uint32_t wsel = warp == 9 ? 0u : 1u;
        const uint32_t idesc = umma_idesc_bf16(128, a.n);
        const uint64_t da = umma_desc_sw128(smem_u32(abuf));
        for (uint32_t it = 0; it < (uint32_t)a.ntiles; ++it) {
            const uint32_t s = a.use_ring ? (it & 3) : 0;
            if (a.use_ring) mbar_wait(&full[s], (it >> 2) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t db = umma_desc_sw128(smem_u32(ring + s * kTile));
                umma_bf16(tmem + (a.issuers == 2 ? wsel * 256 : (it & 1) * 256), da, db, idesc, 0);
                umma_bf16(tmem + (a.issuers == 2 ? wsel * 256 : (it & 1) * 256), da + 2, db + 2, idesc, 1);
                umma_bf16(tmem + (a.issuers == 2 ? wsel * 256 : (it & 1) * 256), da + 4, db + 4, idesc, 1);
                umma_bf16(tmem + (a.issuers == 2 ? wsel * 256 : (it & 1) * 256), da + 6, db + 6, idesc, 1);
                if (a.use_ring) umma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(done + wsel);
        __syncwarp();
        mbar_wait(done + wsel, 0);
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) a.cyc[wsel] = clock64() - t0;
        if (wsel == 0) *stop = 1;
    } else if (warp < 8 && a.epi_traffic && !(warp == 8)) {
        // epilogue-like traffic: 16-byte swizzled stores + broadcast parameter loads, until the MMA warp is done
        const int r = threadIdx.x & 127, hf = threadIdx.x >> 7;
        float acc = 0.f;
        uint32_t k = 0;
        while (!*stop) {
            for (int u = 0; u < 8; ++u) {
                const float4 p = *reinterpret_cast<const float4*>(abuf + ((k + u) & 255) * 16);      // broadcast read
                acc += p.x;
                if (a.epi_traffic > 1) {
                    const uint32_t v = __float_as_uint(acc);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(scratch + hf * 16384 + sw128_offset(r, u))), "r"(v) : "memory");
                }
            }
            ++k;
        }
        if (acc == 123.f) a.cyc[2] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, 512);
}

int main() {
    const size_t src_bytes = 10u << 20;
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    unsigned long long* cyc; cudaMalloc(&cyc, 24);
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    printf("grid  N producers ring epi issuers | cycles/k-tile (issuer 0, issuer 1)  (tensor floor)\n");
    struct C { int n, ring, producers, epi, issuers; };
    std::vector<C> cs;
    for (int n : {32, 64, 128, 192, 256}) { cs.push_back({n, 0, 0, 0, 1}); cs.push_back({n, 0, 0, 0, 2}); }
    for (int n : {64, 128, 192, 256}) for (int p : {1, 2}) for (int epi : {0, 2}) cs.push_back({n, 1, p, epi, 1});
    for (const C& c : cs) {
        Args a{src, src_bytes, c.n, 4, 4000, c.producers, c.ring, c.epi, c.issuers, cyc};
        const int smem = 1024 + 16384 + 32768 + kTile * 4;
        cudaMemset(cyc, 0, 24);
        mma_kernel<<<110, 384, smem>>>(a);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        unsigned long long h[2]; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("%4d %3d %d %d %d %d | %8.1f %8.1f (%d)\n", 110, c.n, c.producers, c.ring, c.epi, c.issuers, double(h[0]) / a.ntiles, double(h[1]) / a.ntiles, c.n * 2);
    }
    return 0;
}
