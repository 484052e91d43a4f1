// fp64 / fp32 FMA issue-rate microbenchmark: nvcc -arch=sm_100a -O3 dfma.cu -o dfma
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void k(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T>
void run(const char* name, int threads, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    T* out; cudaMalloc(&out, sizeof(T) * sms * blocks_per_sm * threads);
    const int iters = 4096; constexpr int ILP = 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<T, ILP><<<sms * blocks_per_sm, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e0);
    k<T, ILP><<<sms * blocks_per_sm, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * blocks_per_sm * threads * iters * ILP;
    printf("%s threads=%d blocks/SM=%d: %.2f TFLOP/s (%.1f FMA/clk/SM at 1.9 GHz)\n", name, threads, blocks_per_sm,
           2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.9e9);
    cudaFree(out);
}
int main() {
    run<double>("fp64", 256, 2); run<double>("fp64", 512, 4); run<double>("fp64", 128, 1);
    run<float>("fp32", 512, 4);
    return 0;
}
