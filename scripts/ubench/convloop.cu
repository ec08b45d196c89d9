// Microbenchmark: register-only model of the dwconv inner loop (TH=16 rows, 7 taps, channel pair per lane),
// packed FFMA2 vs scalar FFMA, to see which the register file can feed.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int TH = 16;

template <int MODE>
__global__ void __launch_bounds__(256, 1) bench(float* out, int iters, float seed) {
    float2 acc[TH], col[TH + 6], w[7];
#pragma unroll
    for (int i = 0; i < TH; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < TH + 6; ++i) col[i] = make_float2(seed * (i + threadIdx.x), seed * (i + 1));
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = make_float2(seed + i, seed - i);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // scalar, tap-major (as written in the kernel)
#pragma unroll
            for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                for (int i = 0; i < TH; ++i) {
                    acc[i].x = fmaf(col[i + ky].x, w[ky].x, acc[i].x);
                    acc[i].y = fmaf(col[i + ky].y, w[ky].y, acc[i].y);
                }
        } else if (MODE == 1) {  // packed, tap-major
#pragma unroll
            for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                for (int i = 0; i < TH; ++i) {
                    uint64_t a = fma2(pk2(col[i + ky].x, col[i + ky].y), pk2(w[ky].x, w[ky].y), pk2(acc[i].x, acc[i].y));
                    acc[i] = *reinterpret_cast<float2*>(&a);
                }
        } else {  // packed, halo-value-major (one input row feeds up to 7 output rows)
#pragma unroll
            for (int r = 0; r < TH + 6; ++r)
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    const int i = r - ky;
                    if (i >= 0 && i < TH) {
                        uint64_t a = fma2(pk2(col[r].x, col[r].y), pk2(w[ky].x, w[ky].y), pk2(acc[i].x, acc[i].y));
                        acc[i] = *reinterpret_cast<float2*>(&a);
                    }
                }
        }
        // keep the loop honest: rotate the column values a little
#pragma unroll
        for (int i = 0; i < TH + 6; i += 5) col[i].x += 1.0f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TH; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
    const int sms = 148, iters = 4000;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<sms, 256>>>(out, 10, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<MODE><<<sms, 256>>>(out, iters, 1e-3f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * 256 * iters * 7.0 * TH * 2.0;
    printf("%-28s %.3f ms  %.1f FMA/clk/SM at 1.9 GHz (8 warps/SM)\n", name, ms, fma / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main() {
    run<0>("scalar FFMA tap-major");
    run<1>("FFMA2 tap-major");
    run<2>("FFMA2 halo-major");
    return 0;
}
