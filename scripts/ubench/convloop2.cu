// Microbenchmark: the dwconv inner loop with its shared-memory traffic (halo bf16 pairs + fp32 tap pairs),
// isolating what limits it.  Variants: TH (rows per thread), bf16-unpack vs fp32 halo, FFMA2 vs scalar FFMA,
// warps per SM.  No TMA / LN / stores: pure loop throughput.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#ifdef UNPACK_PRMT
// both halves on the ALU pipe: the compiler turns `u << 16` into IMAD.U32 (FMA pipe), which competes with the FFMA2s
__device__ __forceinline__ uint64_t unpack_bf16x2(uint32_t u) {
    uint32_t lo, hi;
    asm volatile("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(u));
    asm volatile("lop3.b32 %0, %1, 0xffff0000, 0, 0xc0;" : "=r"(hi) : "r"(u));
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
#else
__device__ __forceinline__ uint64_t unpack_bf16x2(uint32_t u) { return (static_cast<uint64_t>(u & 0xFFFF0000u) << 32) | static_cast<uint64_t>(u << 16); }
#endif

// MODE 0: bf16 halo + FFMA2 (kernel as is)   1: fp32 halo (LDS.64, no unpack) + FFMA2   2: bf16 halo + scalar FFMA
// MODE 3: bf16 halo, kx loop fully unrolled + FFMA2
// MODE 4: TWO output columns per warp (8 input columns per chunk, every halo value feeds both, taps kept for the neighbour)
template <int TH, int MODE>
__global__ void __launch_bounds__(256, 2) bench(float* out, int chunks) {
    constexpr int HALO_W = (MODE == 4 || MODE == 5) ? 22 : 14;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* halo = reinterpret_cast<uint32_t*>(smem);                       // [(TH+6)][14][32] bf16 pairs (or [..][64] fp32 for MODE 1)
    uint64_t* taps = reinterpret_cast<uint64_t*>(smem + (TH + 6) * HALO_W * 32 * 8);  // [49][32] fp32 pairs
    for (int i = threadIdx.x; i < (TH + 6) * HALO_W * 32 * 2; i += blockDim.x) halo[i] = 0x3f803f80u + (i & 7);
    for (int i = threadIdx.x; i < 63 * 32; i += blockDim.x) taps[i] = pk2(0.01f * (i & 3), 0.02f);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t accs = pk2(0.f, 0.f);
    for (int c = 0; c < chunks; ++c) {
        uint64_t acc[TH];
#pragma unroll
        for (int i = 0; i < TH; ++i) acc[i] = pk2(0.f, (float)c);
        const uint32_t* hp = halo + (MODE == 1 ? 2 : 1) * (wid * 32 + lane);
        const uint64_t* tp = taps + lane;
        if (MODE == 4 || MODE == 5) {
            uint64_t acc1[TH];
#pragma unroll
            for (int i = 0; i < TH; ++i) acc1[i] = pk2(1.f, (float)c);
            const uint32_t* hp2 = halo + (2 * wid) * 32 + lane;
            if (MODE == 4) {
                // uniform rolled loop over the 8 input columns; the tap table is padded with a zero column at both ends
                // ([ky][9]: entry j = tap kx = j - 1), so j = 0 and j = 7 run one useless half each (12.5 % extra FMAs)
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    uint64_t col[TH + 6], w0[7], w1[7];
#pragma unroll
                    for (int r = 0; r < TH + 6; ++r) col[r] = unpack_bf16x2(hp2[(r * HALO_W + j) * 32]);
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky) { w0[ky] = tp[(ky * 9 + j + 1) * 32]; w1[ky] = tp[(ky * 9 + j) * 32]; }
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                        for (int i = 0; i < TH; ++i) { acc[i] = fma2(col[i + ky], w0[ky], acc[i]); acc1[i] = fma2(col[i + ky], w1[ky], acc1[i]); }
                }
            } else {
                // peeled: j = 0 feeds column 0 only, j = 7 column 1 only, j = 1..6 (rolled) both
                {
                    uint64_t col[TH + 6], w0[7];
#pragma unroll
                    for (int r = 0; r < TH + 6; ++r) col[r] = unpack_bf16x2(hp2[(r * HALO_W + 0) * 32]);
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky) w0[ky] = tp[(ky * 9 + 1) * 32];
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                        for (int i = 0; i < TH; ++i) acc[i] = fma2(col[i + ky], w0[ky], acc[i]);
                }
#pragma unroll 1
                for (int j = 1; j < 7; ++j) {
                    uint64_t col[TH + 6], w0[7], w1[7];
#pragma unroll
                    for (int r = 0; r < TH + 6; ++r) col[r] = unpack_bf16x2(hp2[(r * HALO_W + j) * 32]);
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky) { w0[ky] = tp[(ky * 9 + j + 1) * 32]; w1[ky] = tp[(ky * 9 + j) * 32]; }
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                        for (int i = 0; i < TH; ++i) { acc[i] = fma2(col[i + ky], w0[ky], acc[i]); acc1[i] = fma2(col[i + ky], w1[ky], acc1[i]); }
                }
                {
                    uint64_t col[TH + 6], w1[7];
#pragma unroll
                    for (int r = 0; r < TH + 6; ++r) col[r] = unpack_bf16x2(hp2[(r * HALO_W + 7) * 32]);
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky) w1[ky] = tp[(ky * 9 + 7) * 32];
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                        for (int i = 0; i < TH; ++i) acc1[i] = fma2(col[i + ky], w1[ky], acc1[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < TH; ++i) accs = fma2(acc1[i], pk2(1e-3f, 1e-3f), accs);
        } else if (MODE == 3) {
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                uint64_t col[TH + 6], wv[7];
#pragma unroll
                for (int r = 0; r < TH + 6; ++r) col[r] = unpack_bf16x2(hp[(r * HALO_W + kx) * 32]);
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) wv[ky] = tp[(ky * 7 + kx) * 32];
#pragma unroll
                for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                    for (int i = 0; i < TH; ++i) acc[i] = fma2(col[i + ky], wv[ky], acc[i]);
            }
        } else {
#pragma unroll 1
            for (int kx = 0; kx < 7; ++kx) {
                uint64_t col[TH + 6], wv[7];
#pragma unroll
                for (int r = 0; r < TH + 6; ++r) {
                    if (MODE == 1) col[r] = *reinterpret_cast<const uint64_t*>(hp + (r * HALO_W + kx) * 64);
                    else col[r] = unpack_bf16x2(hp[(r * HALO_W + kx) * 32]);
                }
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) wv[ky] = tp[(ky * 7 + kx) * 32];
                if (MODE == 2) {
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                        for (int i = 0; i < TH; ++i) {
                            float a0, a1, c0, c1, w0, w1;
                            upk2(acc[i], a0, a1); upk2(col[i + ky], c0, c1); upk2(wv[ky], w0, w1);
                            acc[i] = pk2(fmaf(c0, w0, a0), fmaf(c1, w1, a1));
                        }
                } else {
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                        for (int i = 0; i < TH; ++i) acc[i] = fma2(col[i + ky], wv[ky], acc[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TH; ++i) accs = fma2(acc[i], pk2(1e-3f, 1e-3f), accs);
    }
    float lo, hi;
    upk2(accs, lo, hi);
    out[blockIdx.x * blockDim.x + threadIdx.x] = lo + hi;
}

template <int TH, int MODE>
void run(const char* name, int ctas_per_sm) {
    const int sms = 148, chunks = 400;
    const size_t smem = (size_t)(TH + 6) * ((MODE == 4 || MODE == 5) ? 22 : 14) * 32 * 8 + 63 * 32 * 8;
    cudaFuncSetAttribute(bench<TH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<TH, MODE><<<sms * ctas_per_sm, 256, smem>>>(out, 4);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<TH, MODE><<<sms * ctas_per_sm, 256, smem>>>(out, chunks);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * ctas_per_sm * 256 * chunks * 49.0 * TH * 2.0 * ((MODE == 4 || MODE == 5) ? 2.0 : 1.0);
    printf("%-34s TH=%2d ctas/sm=%d  %.3f ms  %.1f FMA/clk/SM at 1.9 GHz  (%s)\n", name, TH, ctas_per_sm, ms,
           fma / (ms * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<8, 0>("bf16 halo, FFMA2, rolled kx", 1);
    run<8, 0>("bf16 halo, FFMA2, rolled kx", 2);
    run<16, 0>("bf16 halo, FFMA2, rolled kx", 1);
    run<16, 0>("bf16 halo, FFMA2, rolled kx", 2);
    run<8, 4>("2 cols/warp, padded taps (useful FMA)", 1);
    run<8, 4>("2 cols/warp, padded taps (useful FMA)", 2);
    run<4, 4>("2 cols/warp, padded taps (useful FMA)", 2);
    run<8, 5>("2 cols/warp, peeled ends", 1);
    run<8, 5>("2 cols/warp, peeled ends", 2);
    run<4, 5>("2 cols/warp, peeled ends", 2);
    run<8, 1>("fp32 halo, FFMA2, rolled kx", 2);
    run<16, 1>("fp32 halo, FFMA2, rolled kx", 2);
    run<8, 2>("bf16 halo, scalar FFMA, rolled kx", 2);
    run<16, 2>("bf16 halo, scalar FFMA, rolled kx", 2);
    return 0;
}
