// SHFL throughput per SM: WARPS warps per CTA, one CTA per SM, each thread ILP independent shuffle+fma chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int ILP>
__global__ void k(float* out, int iters, int width) {
    float v[ILP], acc[ILP];
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x * 0.001f + i; acc[i] = 0.f; }
    extern __shared__ float sm[];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float t;
            if (MODE == 0) t = __shfl_down_sync(0xffffffffu, v[i], 1, width);
            else if (MODE == 1) t = __shfl_xor_sync(0xffffffffu, v[i], 1, width);
            else if (MODE == 2) t = __shfl_sync(0xffffffffu, v[i], (threadIdx.x + 1) & 31, width);
            else { sm[threadIdx.x] = v[i]; __syncwarp(); t = sm[(threadIdx.x & ~31) + ((threadIdx.x + 1) & 31)]; __syncwarp(); }
            acc[i] = fmaf(t, 1.0001f, acc[i]);
            v[i] = acc[i];
        }
    }
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int ILP>
void run(const char* name, int warps) {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, ILP><<<148, warps * 32, 4096>>>(d, 10, 32);
    cudaEventRecord(e0);
    k<MODE, ILP><<<148, warps * 32, 4096>>>(d, iters, 32);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    double shfl = (double)iters * ILP * warps;
    printf("%-10s warps=%2d ILP=%2d  %.3f ms  %.3f shfl/clk/SM (at %d MHz nominal)\n", name, warps, ILP, ms, shfl / cycles, clk / 1000);
    cudaFree(d);
}
int main() {
    run<0, 8>("down", 4); run<0, 8>("down", 8); run<0, 8>("down", 16); run<0, 16>("down", 16); run<0, 8>("down", 32);
    run<1, 8>("xor", 16); run<2, 8>("idx", 16); run<3, 8>("smem", 16);
    return 0;
}
