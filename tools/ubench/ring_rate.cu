// Micro-benchmark (measurement aid, not product code): how fast can a CTA stream L2-resident weight tiles into a
// shared-memory ring with 1-D bulk copies?  One consumer warp releases every stage as soon as it is full.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ring_rate ring_rate.cu && ./ring_rate
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../multimodal-flows_b200/csrc/mmf_ptx.cuh"
using namespace mmf;

struct Args { const uint8_t* src; size_t src_bytes; int tile_bytes, stages, lg, ntiles, producers, split, lgsplit, mc; unsigned long long* cyc; };

__global__ void __launch_bounds__(256, 1) ring_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 16;
    uint8_t* ring = smem + 1024;
    const int warp = threadIdx.x >> 5;
    const uint32_t cs = cluster_nctarank(), crank = cluster_ctarank();
    const uint16_t cmask = static_cast<uint16_t>((1u << cs) - 1u);
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], cs); }
        fence_mbar_init();
    }
    __syncthreads();
    cluster_sync_all();
    const long long t0 = clock64();
    if (warp < a.producers) {
        for (uint32_t it = warp; it < (uint32_t)a.ntiles; it += a.producers) {
            const uint32_t s = it & (a.stages - 1);
            if (it >= (uint32_t)a.stages) mbar_wait(&empty[s], ((it >> a.lg) - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(&full[s], a.tile_bytes);
                const uint32_t slice = a.tile_bytes >> (cs >> 1), part = slice >> a.lgsplit;
                const size_t off = (static_cast<size_t>(it) * a.tile_bytes) & ((8u << 20) - 1);
                for (int c = 0; c < a.split; ++c) {
                    if (cs == 1 || !a.mc) bulk_load_1d(ring + s * a.tile_bytes + c * part, a.src + off + c * part, part, &full[s]);
                    else bulk_load_1d_multicast(ring + s * a.tile_bytes + crank * slice + c * part, a.src + off + crank * slice + c * part, part, &full[s], cmask);
                }
                if (cs > 1 && !a.mc) { /* every CTA loads the whole tile itself */
                    for (int c = 0; c < a.split; ++c) if (c * part + part > (int)slice) {}
                }
            }
            __syncwarp();
        }
    } else if (warp == 7) {
        for (uint32_t it = 0; it < (uint32_t)a.ntiles; ++it) {
            const uint32_t s = it & (a.stages - 1);
            mbar_wait(&full[s], (it >> a.lg) & 1);
            if (elect_one()) {
                if (cs == 1 || !a.mc) mbar_arrive(&empty[s]);
                else for (uint32_t r = 0; r < cs; ++r) mbar_arrive_remote(dsmem_addr(&empty[s], r));
            }
            __syncwarp();
        }
        if (blockIdx.x == 0 && threadIdx.x == 224) a.cyc[0] = clock64() - t0;
    }
    __syncthreads();
    cluster_sync_all();
}

int main() {
    const size_t src_bytes = 10u << 20;
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    unsigned long long* cyc; cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    struct Cfg { int grid, tile, stages, producers, split, cluster, mc; };
    std::vector<Cfg> cfgs;
    for (int grid : {1, 110})
        for (int tile : {4096, 8192, 16384, 32768})
            for (int stages : {1, 2, 4, 8})
                for (int producers : {1, 2})
                    for (int split : {1, 2, 4}) { if (producers == 2 && stages < 2) continue; cfgs.push_back({grid, tile, stages, producers, split, 1, 0}); }
    for (int tile : {16384, 32768}) for (int cl : {2, 4}) for (int stages : {1, 4}) cfgs.push_back({cl == 4 ? 108 : 110, tile, stages, 2, 1, cl, 1});
    printf("grid tile stages producers split cluster | cycles/tile  B/clk/SM  | kernel ms  aggregate TB/s\n");
    for (const Cfg& c : cfgs) {
        if (c.tile * c.stages > 190 * 1024) continue;
        int lg = 0; while ((1 << lg) < c.stages) ++lg; int lgs = 0; while ((1 << lgs) < c.split) ++lgs;
        Args a{src, src_bytes, c.tile, c.stages, lg, 2000, c.producers, c.split, lgs, c.mc, cyc};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c.grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 1024 + c.tile * c.stages;
        cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = c.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            cudaError_t err = cudaLaunchKernelEx(&cfg, ring_kernel, a);
            cudaEventRecord(e1);
            if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaEventElapsedTime(&ms, e0, e1);
        }
        unsigned long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double cpt = double(h) / a.ntiles;
        printf("%4d %6d %2d %d %d %d | %8.1f %7.1f | %7.3f %6.2f\n", c.grid, c.tile, c.stages, c.producers, c.split, c.cluster, cpt, c.tile / cpt, ms,
               double(c.grid) * a.ntiles * c.tile / (ms * 1e-3) / 1e12);
    }
    return 0;
}
