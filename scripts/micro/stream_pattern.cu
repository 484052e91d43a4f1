// Memory-pattern microbenchmark of the STEP epilogue (k_dense_tc) without the tcgen05 machinery:
// 148 persistent CTAs x 8 warps stream 256-chain x 128-dim patches: read q_prev, q_cur (fp32), write
// q_next over q_prev (fp32) and the bf16 operand.  Flags isolate what caps the achieved bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 stream_pattern.cu -o stream_pattern
//   ./stream_pattern            (runs every variant, prints us per launch and GB/s on 14 B / element)
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int BM = 128, BN = 256, BK = 64, CWID = 16;
enum { F_BLOCK32 = 1, F_NO_W32 = 2, F_NO_W16 = 4, F_NO_READ = 8, F_W16_VIA_SMEM = 16, F_ROWCOOP = 32 };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// WARPS epilogue warps; warp -> (quarter = w & 3, part = w >> 2), part splits the 256 chains
template <int FLAGS, int EDEPTH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_stream(float* __restrict__ q_prev, const float* __restrict__ q_cur, __nv_bfloat16* __restrict__ hi, int n_patches,
         int kblocks) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int ECHUNK = 2 * CWID * 32 * 4;
    constexpr int PARTS = WARPS / 4;
    constexpr int NC = BN / PARTS / CWID;             // chunks per warp per patch
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3, part = warp >> 2;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem) + warp * EDEPTH * ECHUNK;
    const int row_in = lane >> 3, seg = lane & 7;
    // fp32 layout inside a patch: default [256 chains][128 dims]; BLOCK32: [4 quarters][256 chains][32 dims]
    const int64_t row_stride = (FLAGS & F_BLOCK32) ? 32 : BM;
    const int64_t warp_off = (FLAGS & F_BLOCK32) ? (int64_t)quarter * BN * 32 + (int64_t)part * (BN / PARTS) * 32
                                                 : (int64_t)part * (BN / PARTS) * BM + quarter * 32;
    int pf_patch = blockIdx.x, pf_ch = 0, pf_slot = 0;
    uint32_t pf_dst = ring + row_in * 128 + seg * 16;
    int64_t pf_e = (int64_t)pf_patch * BN * BM + warp_off + row_in * row_stride + seg * 4;
    auto issue = [&]() {
        if (pf_patch < n_patches && !(FLAGS & F_NO_READ)) {
            const float* sp = q_prev + pf_e;
            const float* sc = q_cur + pf_e;
#pragma unroll
            for (int it = 0; it < CWID / 4; ++it) {
                cp_async16(pf_dst + it * 512, sp + it * 4 * row_stride);
                cp_async16(pf_dst + CWID * 128 + it * 512, sc + it * 4 * row_stride);
            }
        }
        if (pf_patch < n_patches) {
            pf_e += CWID * row_stride;
            if (++pf_ch == NC) {
                pf_ch = 0;
                pf_patch += gridDim.x;
                pf_e = (int64_t)pf_patch * BN * BM + warp_off + row_in * row_stride + seg * 4;
            }
        }
        cp_async_commit();
        pf_dst += ECHUNK;
        if (++pf_slot == EDEPTH) { pf_slot = 0; pf_dst -= EDEPTH * ECHUNK; }
    };
#pragma unroll
    for (int p = 0; p < EDEPTH - 1; ++p) issue();
    uint32_t src = ring + lane * 4;
    int slot = 0;
    for (int patch = blockIdx.x; patch < n_patches; patch += gridDim.x) {
        float* qw = q_prev + (int64_t)patch * BN * BM + warp_off + lane;
        // bf16 box-blocked: element (chain row, k) at ((row/BN)*kblocks + k/64)*BN*64 + (row%BN)*64 + k%64
        const int n_tile = patch / (kblocks / 2), m_tile = patch % (kblocks / 2);
        const int d = m_tile * BM + quarter * 32 + lane;
        __nv_bfloat16* hw = hi + ((int64_t)n_tile * kblocks + d / BK) * BN * BK + (int64_t)part * (BN / PARTS) * BK + d % BK;
        for (int ch = 0; ch < NC; ++ch) {
            issue();
            cp_async_wait<EDEPTH - 1>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < CWID; ++j) {
                float pj = 1.f, qj = 2.f;
                if (!(FLAGS & F_NO_READ)) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pj) : "r"(src + j * 128));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qj) : "r"(src + CWID * 128 + j * 128));
                }
                const float qn = qj + fmaf(0.01f, pj, qj - pj);
                if (!(FLAGS & F_NO_W32)) __stcs(qw + j * row_stride, qn);
                if (!(FLAGS & F_NO_W16)) hw[j * BK] = __float2bfloat16_rn(qn);
            }
            qw += CWID * row_stride;
            hw += CWID * BK;
            src += ECHUNK;
            if (++slot == EDEPTH) { slot = 0; src -= EDEPTH * ECHUNK; }
            __syncwarp();
        }
    }
    cp_async_wait<0>();
}

// reference point: plain grid-stride float4 streaming with the same byte counts (read 8 B, write 6 B per element)
__global__ void k_plain(float4* __restrict__ q_prev, const float4* __restrict__ q_cur, uint2* __restrict__ hi, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = q_prev[i], q = __ldcs(q_cur + i);
        float4 r;
        r.x = q.x + fmaf(0.01f, p.x, q.x - p.x); r.y = q.y + fmaf(0.01f, p.y, q.y - p.y);
        r.z = q.z + fmaf(0.01f, p.z, q.z - p.z); r.w = q.w + fmaf(0.01f, p.w, q.w - p.w);
        __stcs(q_prev + i, r);
        __nv_bfloat162 a = __floats2bfloat162_rn(r.x, r.y), b = __floats2bfloat162_rn(r.z, r.w);
        uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        hi[i] = o;
    }
}

template <int FLAGS, int EDEPTH, int WARPS>
void run(const char* name, float* qp, float* qc, __nv_bfloat16* hi, int n_patches, int kblocks, double bytes_per_elem) {
    const int smem = WARPS * EDEPTH * 2 * CWID * 32 * 4;
    cudaFuncSetAttribute(k_stream<FLAGS, EDEPTH, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k_stream<FLAGS, EDEPTH, WARPS><<<148, WARPS * 32, smem>>>(qp, qc, hi, n_patches, kblocks);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k_stream<FLAGS, EDEPTH, WARPS><<<148, WARPS * 32, smem>>>(i & 1 ? qc : qp, i & 1 ? qp : qc, hi, n_patches, kblocks);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / reps, elems = (double)n_patches * BN * BM;
    printf("%-58s depth %d warps %2d: %7.1f us  %6.0f GB/s (%s)\n", name, EDEPTH, WARPS, us, elems * bytes_per_elem / us / 1e3,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int C = 65536, Dp = 1024, kblocks = Dp / BK, n_patches = (C / BN) * (Dp / BM);
    float *qp, *qc; __nv_bfloat16* hi;
    cudaMalloc(&qp, (size_t)C * Dp * 4); cudaMalloc(&qc, (size_t)C * Dp * 4); cudaMalloc(&hi, (size_t)C * Dp * 2);
    cudaMemset(qp, 0, (size_t)C * Dp * 4); cudaMemset(qc, 0, (size_t)C * Dp * 4);
    run<0, 3, 8>("current pattern", qp, qc, hi, n_patches, kblocks, 14);
    run<0, 4, 8>("current pattern", qp, qc, hi, n_patches, kblocks, 14);
    run<0, 3, 16>("current pattern", qp, qc, hi, n_patches, kblocks, 14);
    run<F_NO_W16, 3, 8>("no bf16 writes", qp, qc, hi, n_patches, kblocks, 12);
    run<F_NO_W32 | F_NO_W16, 3, 8>("reads only", qp, qc, hi, n_patches, kblocks, 8);
    run<F_NO_READ, 3, 8>("writes only", qp, qc, hi, n_patches, kblocks, 6);
    run<F_NO_READ | F_NO_W16, 3, 8>("fp32 writes only", qp, qc, hi, n_patches, kblocks, 4);
    run<F_NO_READ | F_NO_W32, 3, 8>("bf16 writes only", qp, qc, hi, n_patches, kblocks, 2);
    run<F_BLOCK32, 3, 8>("fp32 arrays blocked [quarter][chain][32 dims]", qp, qc, hi, n_patches, kblocks, 14);
    run<F_BLOCK32, 3, 16>("fp32 arrays blocked [quarter][chain][32 dims]", qp, qc, hi, n_patches, kblocks, 14);
    run<F_BLOCK32 | F_NO_W16, 3, 8>("blocked32, no bf16 writes", qp, qc, hi, n_patches, kblocks, 12);
    run<F_BLOCK32 | F_NO_W32 | F_NO_W16, 3, 8>("blocked32, reads only", qp, qc, hi, n_patches, kblocks, 8);
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int64_t n4 = (int64_t)C * Dp / 4;
        for (int i = 0; i < 3; ++i) k_plain<<<148 * 8, 256>>>((float4*)qp, (const float4*)qc, (uint2*)hi, n4);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) k_plain<<<148 * 8, 256>>>((float4*)(i & 1 ? qc : qp), (const float4*)(i & 1 ? qp : qc), (uint2*)hi, n4);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-58s                  : %7.1f us  %6.0f GB/s\n", "plain grid-stride float4 kernel, same bytes", ms * 1e3 / 20,
               (double)C * Dp * 14 / (ms * 1e3 / 20) / 1e3);
    }
    return 0;
}
