// Experiment: can a tcgen05.mma A operand (K-major, SWIZZLE_128B, 128-byte rows = pixels x 64 channels) start at an
// arbitrary ROW offset inside a larger swizzled tile (a "shifted view"), and does the descriptor need base_offset?
// D[m][n] = sum_k A[m + sh][kb*16 + k] * B[n][k], B = diag(w) (16x16, no swizzle).  Compared on the host.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../../spine_vision_b200/csrc/svb_common.cuh"
using namespace svb;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)swz << 61;
    return d;
}

// rows: number of 128-byte pixel rows in the A tile (multiple of 8)
__global__ void __launch_bounds__(128) k(const __nv_bfloat16* __restrict__ a /*[rows][64]*/, const float* __restrict__ w /*[16]*/,
                                         float* __restrict__ out /*[128][16]*/, int rows, int sh, int kb, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                       // rows * 128 B, SWIZZLE_128B pattern
    uint8_t* sB = smem + rows * 128;          // 512 B: 16x16 bf16, core matrices (n_hi, k_hi) at n_hi*256 + k_hi*128
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 512);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    // software 128B swizzle: 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
    for (int i = tid; i < rows * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a + (size_t)r * 64 + c * 8);
    }
    for (int i = tid; i < 256; i += 128) reinterpret_cast<uint16_t*>(sB)[i] = 0;
    __syncthreads();
    if (tid < 16) {
        const int n = tid;
        __nv_bfloat16 v = __float2bfloat16_rn(w[n]);
        *reinterpret_cast<__nv_bfloat16*>(sB + (n >> 3) * 256 + (n >> 3) * 128 + (n & 7) * 16 + (n & 7) * 2) = v;
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (tid < 32) tmem_alloc<1>(tptr, 32);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(sA) + sh * 128 + kb * 32;
        const uint32_t boff = mode == 1 ? ((a_addr >> 7) & 7) : 0;
        const uint64_t adesc = make_desc(a_addr, 16, 1024, 2, boff);
        const uint64_t bdesc = make_desc(smem_u32(sB), 128, 256, 0, 0);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        tc_mma_f16<1>(tmem, adesc, bdesc, idesc, 0u);
        tc_commit<1>(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const int warp = tid >> 5, lane = tid & 31;
    float v[16];
    {
        uint32_t* r = reinterpret_cast<uint32_t*>(v);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
        tmem_ld_wait();
    }
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = v[j];
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc<1>(tmem, 32);
}

int main() {
    const int rows = 384;
    std::vector<__nv_bfloat16> ha((size_t)rows * 64);
    std::vector<float> haf((size_t)rows * 64), hw(16);
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) { float f = (float)((rand() % 255) - 127) / 16.0f; ha[i] = __float2bfloat16_rn(f); haf[i] = __bfloat162float(ha[i]); }
    for (int i = 0; i < 16; ++i) hw[i] = __bfloat162float(__float2bfloat16_rn(0.25f * (i + 1)));
    __nv_bfloat16* da; float *dw, *dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&dw, 64); cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), 64, cudaMemcpyHostToDevice);
    const int smem = rows * 128 + 512 + 64 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hout(128 * 16);
    for (int mode = 0; mode < 2; ++mode)
        for (int sh : {0, 8, 1, 3, 7, 38, 41, 117, 200}) {
            for (int kb : {0, 3}) {
                cudaMemset(dout, 0, 128 * 16 * 4);
                k<<<1, 128, smem>>>(da, dw, dout, rows, sh, kb, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d sh %d kb %d: CUDA error %s\n", mode, sh, kb, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0; double maxerr = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 16; ++n) {
                        const float want = haf[(size_t)(m + sh) * 64 + kb * 16 + n] * hw[n];
                        const double err = fabs(hout[m * 16 + n] - want);
                        if (err > 1e-3) ++bad;
                        if (err > maxerr) maxerr = err;
                    }
                printf("base_offset %s  shift %3d rows  kblock %d : %4d / 2048 wrong (max err %.4g)\n", mode ? "=(addr>>7)&7" : "=0          ", sh, kb, bad, maxerr);
            }
        }
    return 0;
}
