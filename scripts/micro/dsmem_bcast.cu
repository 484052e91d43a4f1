// How fast can the 8 CTAs of a cluster hand a freshly written bf16 operand slice to each other through
// distributed shared memory?  (the "trajectory on chip" design for c2 needs every CTA to
// broadcast its 32 KB slice [128 chains x 128 dims x bf16] to the 7 other CTAs of its cluster after EVERY
// leapfrog step.)  Each CTA: 32 KB source tile in its own shared memory, 8 x 32 KB landing buffers; per
// iteration it stores its tile into slot `rank` of every peer (st.shared::cluster, 16 B per thread) and the
// cluster synchronises.  Prints microseconds per exchange and the DSMEM store rate per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/dsmem_bcast scripts/micro/dsmem_bcast.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;

constexpr int CL = 8, TILE = 32 * 1024, THREADS = 256;

__global__ void __launch_bounds__(THREADS, 1) k_bcast(int iters, int peers, float* sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint4* src = reinterpret_cast<uint4*>(smem);                     // 32 KB
    uint4* land = reinterpret_cast<uint4*>(smem + TILE);             // 8 x 32 KB landing slots
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < TILE / 16; i += THREADS) src[i] = make_uint4(i, rank, 0, 0);
    cluster.sync();
    for (int it = 0; it < iters; ++it) {
        for (int p = 1; p <= peers; ++p) {
            uint4* dst = cluster.map_shared_rank(land, (rank + p) % CL) + (p - 1) * (TILE / 16);   // receiver's slot p-1
            for (int i = threadIdx.x; i < TILE / 16; i += THREADS) dst[i] = src[i];
        }
        cluster.sync();
    }
    if (threadIdx.x == 0 && land[0].x == 12345u) *sink = 1.f;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = TILE + (size_t)CL * TILE;    // 288 KB?  no: 32 + 8 * 32 = 288 KB exceeds 227 KB -> use 6 landing slots
    float* sink;
    cudaMalloc(&sink, 4);
    for (int peers : {1, 3, 6}) {
        size_t bytes = TILE + (size_t)peers * TILE;      // landing slots only for the peers that write
        if (bytes < 120 * 1024) bytes = 120 * 1024;        // one CTA per SM, like the real kernel
        (void)smem;
        cudaFuncSetAttribute(k_bcast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        cudaFuncSetAttribute(k_bcast, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = bytes;
        int max_clusters = 0;
        cfg.gridDim = dim3(CL * 18);
        cudaOccupancyMaxActiveClusters(&max_clusters, k_bcast, &cfg);
        cfg.gridDim = dim3(CL * (max_clusters > 0 ? max_clusters : 1));
        const int iters = 200;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaLaunchKernelEx(&cfg, k_bcast, 10, peers, sink);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, k_bcast, iters, peers, sink);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / iters;
        printf("cluster of %d, %d co-resident clusters (%d of %d SMs): each CTA stores its 32 KB tile to %d peers: %.2f us per exchange "
               "= %.1f GB/s of DSMEM stores per SM (%s); 7 peers would take %.2f us\n",
               CL, max_clusters, CL * max_clusters, sms, peers, us, peers * TILE / us / 1e3, cudaGetErrorString(err), us * 7 / peers);
    }
    return 0;
}
