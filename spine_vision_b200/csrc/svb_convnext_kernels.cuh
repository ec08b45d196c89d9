// svb_convnext_kernels.cuh -- sm_100a device kernels of the CoordinateRegressor forward
// (spine_vision/training/models/generic.py:389-391 over a timm ConvNeXt backbone).
//
// Activations are NHWC ("token-major": [B*H*W, C]) 16-bit (bf16 or fp16), so every pointwise
// layer is a K-major GEMM operand as stored.  Kernels:
//   stem_ln_kernel      conv4x4 s4 (1 folded input channel) + LayerNorm2d         CUDA cores, HBM-bound
//   dwconv_ln_kernel    depthwise 7x7 + bias + LayerNorm(C)                       TMA halo tiles, warp-shuffle LN
//   gemm_kernel         D = epilogue(A * W^T): fc1+GELU / fc2*gamma+residual / downsample conv
//                       tcgen05.mma (kind::f16) with TMEM accumulators, TMA operand pipeline
//   ln_patchify_kernel  LayerNorm2d + 2x2/s2 patch gather (A operand of the downsample GEMM)
//   head_kernel         avg-pool + LayerNorm2d + LayerNorm + Linear + GELU + Linear + sigmoid
#pragma once
#include "svb_common.cuh"

namespace svb {

constexpr float LN_EPS_BACKBONE = 1e-6f;  // timm ConvNeXt LayerNorm / LayerNorm2d
constexpr float LN_EPS_HEAD = 1e-5f;      // nn.LayerNorm default (generic.py:344)

// --------------------------------------------------------------------------------------------
// GELU(x) = x * Phi(x) with Phi through erfc(|x|/sqrt2) ~= exp2(poly5(|x|)), |x| clamped at 4*sqrt2.
// Max |erfc error| 6.8e-7 (fit in scripts/fit_gelu.py), i.e. far below 16-bit output rounding.
// 10 FP32 ops + 1 MUFU per element: the fc1 epilogue has a budget of ~16 issue slots per element
// at C=128 before it, not the tensor pipe, bounds the kernel.
__device__ __forceinline__ float gelu_fast(float x) {
    const float u = fminf(fabsf(x), 5.65685425f);
    float r = -5.20460508e-04f;            // c5 / 2^(5/2)
    r = fmaf(r, u, 7.39751849e-03f);       // c4 / 4
    r = fmaf(r, u, -5.25612477e-02f);      // c3 / 2^(3/2)
    r = fmaf(r, u, -4.59254682e-01f);      // c2 / 2
    r = fmaf(r, u, -1.15109138e+00f);      // c1 / sqrt2
    r *= u;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(r));
    const float h = (0.5f * x) * e;
    return fmaxf(x, 0.0f) - fabsf(h);
}
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// ============================================================================ stem
// in  u8 [B,H,W] (K1 output, one plane); /255, ImageNet mean/std and the RGB replication of
// cropping.py:463-472 are folded into wf/bf.  out [B,H/4,W/4,C0].
template <typename T, int CPL>
__global__ void __launch_bounds__(256) stem_ln_kernel(const uint8_t* __restrict__ in, const float* __restrict__ wf /*[C0][16]*/,
                                                      const float* __restrict__ bf, const float* __restrict__ lnw,
                                                      const float* __restrict__ lnb, T* __restrict__ out, int B, int H,
                                                      int W) {
    constexpr int C0 = 32 * CPL;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int Ho = H >> 2, Wo = W >> 2;
    const long long tokens = (long long)B * Ho * Wo;

    float w[CPL][16], bias[CPL], g[CPL], be[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int c = lane * CPL + j;
#pragma unroll
        for (int p = 0; p < 16; ++p) w[j][p] = wf[c * 16 + p];
        bias[j] = bf[c];
        g[j] = lnw[c];
        be[j] = lnb[c];
    }
    for (long long t = warp; t < tokens; t += nwarps) {
        const int b = (int)(t / (Ho * Wo));
        const int rem = (int)(t - (long long)b * Ho * Wo);
        const int ty = rem / Wo, tx = rem - ty * Wo;
        const uint8_t* p = in + ((size_t)b * H + (size_t)ty * 4) * W + (size_t)tx * 4;
        float px[16];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p + (size_t)r * W));
            px[r * 4 + 0] = (float)(v & 0xFF);
            px[r * 4 + 1] = (float)((v >> 8) & 0xFF);
            px[r * 4 + 2] = (float)((v >> 16) & 0xFF);
            px[r * 4 + 3] = (float)(v >> 24);
        }
        float acc[CPL];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float a = bias[j];
#pragma unroll
            for (int q = 0; q < 16; ++q) a = fmaf(px[q], w[j][q], a);
            acc[j] = a;
            s += a;
        }
        const float mean = warp_sum(s) * (1.0f / C0);
        float v2 = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) { const float d = acc[j] - mean; v2 = fmaf(d, d, v2); }
        const float rstd = 1.0f / sqrtf(warp_sum(v2) * (1.0f / C0) + LN_EPS_BACKBONE);
        static_assert(CPL % 2 == 0, "stem packs channel pairs");
        uint32_t* o = reinterpret_cast<uint32_t*>(out + (size_t)t * C0 + lane * CPL);
#pragma unroll
        for (int j = 0; j < CPL; j += 2)
            o[j >> 1] = Cvt<T>::pack2(fmaf((acc[j] - mean) * rstd, g[j], be[j]),
                                      fmaf((acc[j + 1] - mean) * rstd, g[j + 1], be[j + 1]));
    }
}

// ============================================================================ depthwise 7x7 + LayerNorm
// One CTA = TH x 8 output pixels x all C channels.  The (TH+6) x 14 x 64-channel halo tile of
// each 64-channel chunk arrives by TMA (4-D tensor map over NHWC, out-of-bounds = zero padding)
// together with the chunk's [49][64] fp32 taps; 8 warps = 8 pixel columns, lane = channel pair,
// each thread slides a TH-row window down its column.  Conv results stay on chip in fp32
// (res[TH*8][C]); LayerNorm over C is a warp-per-pixel shuffle reduction; the 16-bit normalised
// row is the K-major A operand of the fc1 GEMM.
template <int C, int TH>
struct DwCfg {
    static constexpr int TW = 8, CC = 64, STAGES = 2;
    static constexpr int HALO_H = TH + 6, HALO_W = TW + 6;
    static constexpr int HALO_BYTES = HALO_H * HALO_W * CC * 2;
    static constexpr int W_BYTES = 49 * CC * 4;
    static constexpr int STAGE_BYTES = ((HALO_BYTES + W_BYTES + 127) / 128) * 128;
    static constexpr int RES_BYTES = TH * TW * C * 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RES_BYTES + 64 + 1024;  // + barriers + align slack
    static_assert(C % CC == 0, "C must be a multiple of 64");
    static_assert(HALO_BYTES % 128 == 0, "halo stage must keep the tap buffer 128B aligned");
};

template <typename T, int C, int TH>
__global__ void __launch_bounds__(256) dwconv_ln_kernel(const __grid_constant__ CUtensorMap x_map,
                                                        const __grid_constant__ CUtensorMap w_map,
                                                        const float* __restrict__ bdw, const float* __restrict__ lnw,
                                                        const float* __restrict__ lnb, T* __restrict__ out, int H, int W,
                                                        int tiles_x, int tiles_y) {
    using Cfg = DwCfg<C, TH>;
    constexpr int TW = Cfg::TW, CC = Cfg::CC, STAGES = Cfg::STAGES, HALO_W = Cfg::HALO_W;
    constexpr int NCHUNK = C / CC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_stage = smem;
    float* s_res = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::RES_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int x0 = tx * TW, y0 = ty * TH;

    if (tid == 0) {
        tma_prefetch_desc(&x_map);
        tma_prefetch_desc(&w_map);
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int stage, int chunk) {
        uint8_t* dst = s_stage + stage * Cfg::STAGE_BYTES;
        mbar_expect_tx(&s_full[stage], Cfg::HALO_BYTES + Cfg::W_BYTES);
        tma_load_4d(dst, &x_map, &s_full[stage], chunk * CC, x0 - 3, y0 - 3, b);
        tma_load_2d(dst + Cfg::HALO_BYTES, &w_map, &s_full[stage], chunk * CC, 0);
    };
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < NCHUNK; ++s) issue(s, s);
    }

    for (int k = 0; k < NCHUNK; ++k) {
        const int stage = k % STAGES;
        mbar_wait(&s_full[stage], (k / STAGES) & 1);
        const uint32_t* halo = reinterpret_cast<const uint32_t*>(s_stage + stage * Cfg::STAGE_BYTES);  // [HALO_H][HALO_W][32] ch pairs
        const float2* taps = reinterpret_cast<const float2*>(s_stage + stage * Cfg::STAGE_BYTES + Cfg::HALO_BYTES);  // [49][32]
        const int c0 = k * CC + 2 * lane;
        const float2 bias = *reinterpret_cast<const float2*>(bdw + c0);
        float2 acc[TH];
#pragma unroll
        for (int i = 0; i < TH; ++i) acc[i] = bias;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
            float2 col[TH + 6];
#pragma unroll
            for (int r = 0; r < TH + 6; ++r) col[r] = Cvt<T>::unpack2(halo[(r * HALO_W + wid + kx) * 32 + lane]);
#pragma unroll
            for (int ky = 0; ky < 7; ++ky) {
                const float2 wv = taps[(ky * 7 + kx) * 32 + lane];
#pragma unroll
                for (int i = 0; i < TH; ++i) {
                    acc[i].x = fmaf(col[i + ky].x, wv.x, acc[i].x);
                    acc[i].y = fmaf(col[i + ky].y, wv.y, acc[i].y);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TH; ++i) *reinterpret_cast<float2*>(s_res + (size_t)(i * TW + wid) * C + c0) = acc[i];
        __syncthreads();  // every warp is done with this stage (and res of this chunk is visible)
        if (tid == 0 && k + STAGES < NCHUNK) issue(stage, k + STAGES);
    }

    // LayerNorm over C, one warp per pixel
    constexpr int V4 = C / 128;  // float4 groups per lane
    for (int p = wid; p < TH * TW; p += 8) {
        const int oy = p / TW, ox = p - oy * TW;
        const int y = y0 + oy, x = x0 + ox;
        if (y >= H || x >= W) continue;  // warp-uniform
        const float4* rp = reinterpret_cast<const float4*>(s_res + (size_t)p * C);
        float4 v[V4];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < V4; ++j) {
            v[j] = rp[lane + 32 * j];
            s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
        const float mean = warp_sum(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < V4; ++j) {
            const float a = v[j].x - mean, bb = v[j].y - mean, cc = v[j].z - mean, d = v[j].w - mean;
            q += (a * a + bb * bb) + (cc * cc + d * d);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + LN_EPS_BACKBONE);
        uint2* op = reinterpret_cast<uint2*>(out + (((size_t)b * H + y) * W + x) * C);
#pragma unroll
        for (int j = 0; j < V4; ++j) {
            const int c4 = lane + 32 * j;
            const float4 g = __ldg(reinterpret_cast<const float4*>(lnw) + c4);
            const float4 be = __ldg(reinterpret_cast<const float4*>(lnb) + c4);
            uint2 o;
            o.x = Cvt<T>::pack2(fmaf((v[j].x - mean) * rstd, g.x, be.x), fmaf((v[j].y - mean) * rstd, g.y, be.y));
            o.y = Cvt<T>::pack2(fmaf((v[j].z - mean) * rstd, g.z, be.z), fmaf((v[j].w - mean) * rstd, g.w, be.w));
            op[c4] = o;
        }
    }
}

// ============================================================================ tcgen05 GEMM
// D[M,N] = epilogue(A[M,K] * Wt[N,K]^T); A, Wt K-major 16-bit; fp32 accumulation in TMEM.
// Persistent CTAs (one per SM), 128 x BN tiles, BK = 64 (one 128-byte swizzle atom), warp roles:
//   warp 0      TMA producer (one elected lane)           smem ring: STAGES x (A 16 KB + B BN*128 B)
//   warp 1      TMEM allocator + tcgen05.mma issuer (one lane)
//   warps 2..9  epilogue: tcgen05.ld -> bias / GELU / gamma+residual -> 16-bit global stores
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
enum GemmMode { GEMM_GELU = 0, GEMM_RESID = 1, GEMM_BIAS = 2 };

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512: power of two
    static constexpr int NUM_THREADS = 320;
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 256 + 1024;
};

template <typename T> struct UmmaFmt;
template <> struct UmmaFmt<__nv_bfloat16> { static constexpr uint32_t v = 1; };
template <> struct UmmaFmt<__half> { static constexpr uint32_t v = 0; };

template <typename T, int BN, int MODE>
__global__ void __launch_bounds__(320, 1) gemm_kernel(const __grid_constant__ CUtensorMap a_map,
                                                      const __grid_constant__ CUtensorMap w_map, T* out,
                                                      const T* resid, const float* __restrict__ bias,
                                                      const float* __restrict__ gamma, int M, int N, int K) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
    uint64_t* full = bars;                 // [STAGES] TMA -> MMA
    uint64_t* empty = bars + STAGES;       // [STAGES] MMA -> TMA
    uint64_t* tfull = bars + 2 * STAGES;   // [2] MMA -> epilogue
    uint64_t* tempty = tfull + 2;          // [2] epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN;
    const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (K + Cfg::BK - 1) / Cfg::BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&a_map);
        tma_prefetch_desc(&w_map);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
                    tma_load_2d(sA + stage * Cfg::A_BYTES, &a_map, &full[stage], kb * Cfg::BK, m_blk * Cfg::BM);
                    tma_load_2d(sB + stage * Cfg::B_BYTES, &w_map, &full[stage], kb * Cfg::BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=f32, A/B = T, both K-major, M=128, N=BN
            constexpr uint32_t idesc = (1u << 4) | (UmmaFmt<T>::v << 7) | (UmmaFmt<T>::v << 10) |
                                       ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(Cfg::BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(sA + stage * Cfg::A_BYTES));
                    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < Cfg::BK / 16; ++k)  // +32 B per UMMA_K inside the swizzle atom
                        tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[as]);  // accumulator complete -> epilogue
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;             // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;   // which half of the BN columns
        constexpr int CH = BN / 2 / 32;     // 32-column chunks per warp
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const int row = m_blk * Cfg::BM + q * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < CH; ++c) {
                const int col = n_blk * BN + half * (BN / 2) + c * 32;
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2) + c * 32), r);
                tmem_ld_wait();
                if (row < M && col < N) {
                    T* optr = out + (size_t)row * N + col;
                    uint32_t packed[16];
                    if (MODE == GEMM_RESID) {
                        const uint4* rp = reinterpret_cast<const uint4*>(resid + (size_t)row * N + col);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 rv = rp[i];
                            packed[4 * i + 0] = rv.x; packed[4 * i + 1] = rv.y;
                            packed[4 * i + 2] = rv.z; packed[4 * i + 3] = rv.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
                        float v0 = __uint_as_float(r[4 * i + 0]) + bv.x;
                        float v1 = __uint_as_float(r[4 * i + 1]) + bv.y;
                        float v2 = __uint_as_float(r[4 * i + 2]) + bv.z;
                        float v3 = __uint_as_float(r[4 * i + 3]) + bv.w;
                        if (MODE == GEMM_GELU) {
                            v0 = gelu_fast(v0); v1 = gelu_fast(v1); v2 = gelu_fast(v2); v3 = gelu_fast(v3);
                        } else if (MODE == GEMM_RESID) {
                            const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma + col) + i);
                            const float2 x01 = Cvt<T>::unpack2(packed[2 * i]);
                            const float2 x23 = Cvt<T>::unpack2(packed[2 * i + 1]);
                            v0 = fmaf(gv.x, v0, x01.x); v1 = fmaf(gv.y, v1, x01.y);
                            v2 = fmaf(gv.z, v2, x23.x); v3 = fmaf(gv.w, v3, x23.y);
                        }
                        packed[2 * i] = Cvt<T>::pack2(v0, v1);
                        packed[2 * i + 1] = Cvt<T>::pack2(v2, v3);
                    }
                    uint4* op = reinterpret_cast<uint4*>(optr);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        op[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ============================================================================ LayerNorm2d + 2x2/s2 patchify
// x [B,H,W,C] -> a2 [B,H/2,W/2,4C] with k = (ky*2+kx)*C + c, the A operand of the downsample GEMM
// (timm stage.downsample = LayerNorm2d -> Conv2d(k=2,s=2)).
template <typename T, int C>
__global__ void __launch_bounds__(256) ln_patchify_kernel(const T* __restrict__ x, const float* __restrict__ lnw,
                                                          const float* __restrict__ lnb, T* __restrict__ a2, int B, int H,
                                                          int W) {
    constexpr int V4 = C / 128;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long tokens = (long long)B * H * W;
    const int Ho = H >> 1, Wo = W >> 1;
    for (long long t = warp; t < tokens; t += nwarps) {
        const int b = (int)(t / ((long long)H * W));
        const int rem = (int)(t - (long long)b * H * W);
        const int y = rem / W, xx = rem - y * W;
        if ((y >> 1) >= Ho || (xx >> 1) >= Wo) continue;
        const uint2* xp = reinterpret_cast<const uint2*>(x + (size_t)t * C);
        float4 v[V4];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < V4; ++j) {
            const uint2 u = __ldg(xp + lane + 32 * j);
            const float2 a = Cvt<T>::unpack2(u.x), c = Cvt<T>::unpack2(u.y);
            v[j] = make_float4(a.x, a.y, c.x, c.y);
            s += (a.x + a.y) + (c.x + c.y);
        }
        const float mean = warp_sum(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < V4; ++j) {
            const float a = v[j].x - mean, bb = v[j].y - mean, cc = v[j].z - mean, d = v[j].w - mean;
            q += (a * a + bb * bb) + (cc * cc + d * d);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + LN_EPS_BACKBONE);
        const size_t orow = ((size_t)b * Ho + (y >> 1)) * Wo + (xx >> 1);
        uint2* op = reinterpret_cast<uint2*>(a2 + orow * (4 * C) + (size_t)(((y & 1) << 1) | (xx & 1)) * C);
#pragma unroll
        for (int j = 0; j < V4; ++j) {
            const int c4 = lane + 32 * j;
            const float4 g = __ldg(reinterpret_cast<const float4*>(lnw) + c4);
            const float4 be = __ldg(reinterpret_cast<const float4*>(lnb) + c4);
            uint2 o;
            o.x = Cvt<T>::pack2(fmaf((v[j].x - mean) * rstd, g.x, be.x), fmaf((v[j].y - mean) * rstd, g.y, be.y));
            o.y = Cvt<T>::pack2(fmaf((v[j].z - mean) * rstd, g.z, be.z), fmaf((v[j].w - mean) * rstd, g.w, be.w));
            op[c4] = o;
        }
    }
}

// ============================================================================ pool + head
// timm head (num_classes=0): global average pool -> LayerNorm2d(C) -> flatten; then
// generic.py:343-351: LayerNorm(C) -> Linear(C,HID) -> GELU -> Linear(HID,NOUT) -> Sigmoid.  fp32.
__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += s_red[i];
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256) head_kernel(const T* __restrict__ x /*[B,tokens,C]*/, int tokens, int C,
                                                   const float* __restrict__ n0w, const float* __restrict__ n0b,
                                                   const float* __restrict__ n1w, const float* __restrict__ n1b,
                                                   const float* __restrict__ w1 /*[HID][C]*/, const float* __restrict__ b1,
                                                   int HID, const float* __restrict__ w2 /*[NOUT][HID]*/,
                                                   const float* __restrict__ b2, int NOUT, float* __restrict__ coords) {
    extern __shared__ float sm[];
    float* s_feat = sm;          // [C]
    float* s_hid = sm + C;       // [HID]
    float* s_red = s_hid + HID;  // [8]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const T* xb = x + (size_t)b * tokens * C;
    // global average pool (coalesced: consecutive threads = consecutive channel pairs)
    for (int c2 = tid; c2 < C / 2; c2 += 256) {
        float sx = 0.f, sy = 0.f;
        for (int t = 0; t < tokens; ++t) {
            const float2 v = Cvt<T>::unpack2(__ldg(reinterpret_cast<const uint32_t*>(xb + (size_t)t * C) + c2));
            sx += v.x;
            sy += v.y;
        }
        s_feat[2 * c2] = sx / (float)tokens;
        s_feat[2 * c2 + 1] = sy / (float)tokens;
    }
    __syncthreads();
    // two LayerNorms back to back (backbone.head.norm eps 1e-6, head.0 eps 1e-5)
    for (int pass = 0; pass < 2; ++pass) {
        const float* gw = pass == 0 ? n0w : n1w;
        const float* gb = pass == 0 ? n0b : n1b;
        const float eps = pass == 0 ? LN_EPS_BACKBONE : LN_EPS_HEAD;
        float s = 0.f;
        for (int c = tid; c < C; c += 256) s += s_feat[c];
        const float mean = block_sum_256(s, s_red) / (float)C;
        float q = 0.f;
        for (int c = tid; c < C; c += 256) { const float d = s_feat[c] - mean; q = fmaf(d, d, q); }
        const float rstd = 1.0f / sqrtf(block_sum_256(q, s_red) / (float)C + eps);
        for (int c = tid; c < C; c += 256) s_feat[c] = fmaf((s_feat[c] - mean) * rstd, gw[c], gb[c]);
        __syncthreads();
    }
    // Linear(C, HID) + exact GELU: one warp per output, lanes stride the row
    for (int j = wid; j < HID; j += 8) {
        const float* wr = w1 + (size_t)j * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(wr + c), s_feat[c], s);
        s = warp_sum(s);
        if (lane == 0) s_hid[j] = gelu_exact(s + b1[j]);
    }
    __syncthreads();
    for (int j = wid; j < NOUT; j += 8) {
        const float* wr = w2 + (size_t)j * HID;
        float s = 0.f;
        for (int c = lane; c < HID; c += 32) s = fmaf(__ldg(wr + c), s_hid[c], s);
        s = warp_sum(s);
        if (lane == 0) coords[(size_t)b * NOUT + j] = 1.0f / (1.0f + expf(-(s + b2[j])));
    }
}

}  // namespace svb
