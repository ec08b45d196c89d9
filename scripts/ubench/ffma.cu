// Microbenchmark: FP32 FMA throughput per SM with scalar FFMA vs packed FFMA2 (fma.rn.f32x2),
// and MUFU.EX2 throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma.cu -o ffma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        }
    } else if (MODE == 1) {
        unsigned long long A, Bc;
        float2 av = make_float2(a, a), bv = make_float2(b, b);
        A = *reinterpret_cast<unsigned long long*>(&av);
        Bc = *reinterpret_cast<unsigned long long*>(&bv);
        unsigned long long* X = reinterpret_cast<unsigned long long*>(x);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[i]) : "l"(A), "l"(Bc));
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
    int sms = 148;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * 256);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<sms * ctas_per_sm, 256>>>(out, 100, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<MODE><<<sms * ctas_per_sm, 256>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * ctas_per_sm * 256 * 16.0 * iters;  // scalar op count
    printf("%-8s ctas/sm=%d  %.3f ms  %.1f Gop/s  (%.1f ops/clk/SM at 1.9 GHz)\n", name, ctas_per_sm, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main() {
    for (int c : {1, 2, 4}) {
        run<0>("FFMA", c);
        run<1>("FFMA2", c);
        run<2>("MUFU", c);
    }
    return 0;
}
