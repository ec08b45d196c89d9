// svb_classifier.cu -- K4: classifier-input producer.
//
// Replaces, for a whole batch of (patient, level) samples at once, what ClassificationDataset.__getitem__ does per sample
// on the CPU (training/datasets/classification.py:40-68 construct_3channel, :247-278 _build_transforms without
// augmentation, :283-304 __getitem__):
//     rgb = [T2, T1, T2] (or the one available series three times) -> transforms.Resize(output_size) (Pillow bilinear)
//     -> ToTensor (uint8 -> float32 / 255) -> Normalize(ImageNet mean, std)
// Pillow resizes every band independently, so "stack then resize" equals "resize then stack": the resized planes are
// K3's second output (uint8 [N, H, W], bit-exact to Pillow) and this kernel is the stack + ToTensor + Normalize:
//     out[p, c, y, x] = ((plane[idx_c(p)][y, x] / 255) - mean[c]) / std[c]        fp32, IEEE division, no contraction
// There are only 256 possible inputs per channel, so each CTA builds a [3][256] table with the exact fp32 operation order
// of torch (div, sub, div) and the pixel loop is pure data movement: 2 bytes read and 12 bytes written per pixel
// (HBM-bound; 128-bit stores, four pixel quads in flight per thread).
#include "svb_common.cuh"

namespace svb {

constexpr int K4_THREADS = 256;
constexpr int K4_QUADS_PER_THREAD = 4;

template <typename OutT> struct K4Store;
template <> struct K4Store<float> {
    static __device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    }
};
template <> struct K4Store<__nv_bfloat16> {
    static __device__ __forceinline__ void st4(__nv_bfloat16* p, float a, float b, float c, float d) {
        *reinterpret_cast<uint2*>(p) = make_uint2(Cvt<__nv_bfloat16>::pack2(a, b), Cvt<__nv_bfloat16>::pack2(c, d));
    }
};
template <> struct K4Store<__half> {
    static __device__ __forceinline__ void st4(__half* p, float a, float b, float c, float d) {
        *reinterpret_cast<uint2*>(p) = make_uint2(Cvt<__half>::pack2(a, b), Cvt<__half>::pack2(c, d));
    }
};

struct K4Norm { float mean[3], stdv[3]; int normalize; };

// grid (ceil(quads / (256 * 4)), P); one sample per blockIdx.y
template <typename OutT>
__global__ void __launch_bounds__(K4_THREADS) k4_classifier_input_kernel(const uint8_t* __restrict__ planes,
                                                                          const int32_t* __restrict__ t2_idx,
                                                                          const int32_t* __restrict__ t1_idx, int quads /* H*W/4 */,
                                                                          K4Norm nrm, OutT* __restrict__ out) {
    // channels 0 and 2 always show the same plane: one 64-bit lookup serves both
    __shared__ float2 lut02[256];
    __shared__ float lut1[256];
    for (int v = threadIdx.x; v < 256; v += K4_THREADS) {
        const float f = __fdiv_rn(static_cast<float>(v), 255.0f);  // ToTensor
        float c0 = f, c1 = f, c2 = f;
        if (nrm.normalize) {                                        // Normalize: sub then div, fp32
            c0 = __fdiv_rn(__fsub_rn(f, nrm.mean[0]), nrm.stdv[0]);
            c1 = __fdiv_rn(__fsub_rn(f, nrm.mean[1]), nrm.stdv[1]);
            c2 = __fdiv_rn(__fsub_rn(f, nrm.mean[2]), nrm.stdv[2]);
        }
        lut02[v] = make_float2(c0, c2);
        lut1[v] = c1;
    }
    const int p = blockIdx.y;
    const int i2 = t2_idx[p], i1 = t1_idx[p];
    // construct_3channel: both -> [T2, T1, T2]; one -> that series three times
    const int ia = i2 >= 0 ? i2 : i1;        // channels 0 and 2
    const int ib = (i2 >= 0 && i1 >= 0) ? i1 : ia;  // channel 1
    __syncthreads();
    if (ia < 0) return;  // no series at all: the host mirror raises before launching; leave the sample untouched
    const size_t px = (size_t)quads * 4;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(planes + (size_t)ia * px);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(planes + (size_t)ib * px);
    OutT* o0 = out + (size_t)p * 3 * px;
    OutT* o1 = o0 + px;
    OutT* o2 = o1 + px;
    const int q0 = blockIdx.x * (K4_THREADS * K4_QUADS_PER_THREAD) + threadIdx.x;
    uint32_t a[K4_QUADS_PER_THREAD], b[K4_QUADS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < K4_QUADS_PER_THREAD; ++k) {
        const int q = q0 + k * K4_THREADS;
        a[k] = q < quads ? __ldg(pa + q) : 0u;
        b[k] = q < quads ? __ldg(pb + q) : 0u;
    }
#pragma unroll
    for (int k = 0; k < K4_QUADS_PER_THREAD; ++k) {
        const int q = q0 + k * K4_THREADS;
        if (q >= quads) break;
        const uint32_t va = a[k], vb = b[k];
        const int a0 = va & 255, a1 = (va >> 8) & 255, a2 = (va >> 16) & 255, a3 = va >> 24;
        const int b0 = vb & 255, b1 = (vb >> 8) & 255, b2 = (vb >> 16) & 255, b3 = vb >> 24;
        const float2 e0 = lut02[a0], e1 = lut02[a1], e2 = lut02[a2], e3 = lut02[a3];
        K4Store<OutT>::st4(o0 + (size_t)q * 4, e0.x, e1.x, e2.x, e3.x);
        K4Store<OutT>::st4(o1 + (size_t)q * 4, lut1[b0], lut1[b1], lut1[b2], lut1[b3]);
        K4Store<OutT>::st4(o2 + (size_t)q * 4, e0.y, e1.y, e2.y, e3.y);
    }
}

}  // namespace svb

using namespace svb;

extern "C" int svb_k4_classifier_input(const uint8_t* d_planes, const int32_t* d_t2_idx, const int32_t* d_t1_idx, int P, int H,
                                       int W, const float* h_mean3, const float* h_std3, int normalize, int out_dtype,
                                       void* d_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = check_device_sm100()) return rc;
    SVB_REQUIRE(P >= 0 && H > 0 && W > 0, SVB_ERR_INVALID_ARG, "k4: bad sizes P=%d H=%d W=%d", P, H, W);
    if (P == 0) return SVB_OK;
    SVB_REQUIRE(d_planes && d_t2_idx && d_t1_idx && d_out, SVB_ERR_INVALID_ARG, "k4: null pointer argument");
    SVB_REQUIRE(((size_t)H * W) % 4 == 0, SVB_ERR_INVALID_ARG, "k4: H*W (%d x %d) must be a multiple of 4", H, W);
    SVB_REQUIRE(out_dtype == SVB_BF16 || out_dtype == SVB_FP16 || out_dtype == SVB_F32, SVB_ERR_INVALID_ARG,
                "k4: out_dtype %d (0 = bf16, 1 = fp16, 2 = float32)", out_dtype);
    SVB_REQUIRE(P <= 65535, SVB_ERR_INVALID_ARG, "k4: at most 65535 samples per call (got %d)", P);
    K4Norm nrm;
    const float mean_default[3] = {0.485f, 0.456f, 0.406f}, std_default[3] = {0.229f, 0.224f, 0.225f};  // classification.py:270-273
    for (int c = 0; c < 3; ++c) {
        nrm.mean[c] = h_mean3 ? h_mean3[c] : mean_default[c];
        nrm.stdv[c] = h_std3 ? h_std3[c] : std_default[c];
    }
    nrm.normalize = normalize ? 1 : 0;
    const int quads = (int)(((size_t)H * W) / 4);
    dim3 grid(ceil_div(quads, K4_THREADS * K4_QUADS_PER_THREAD), P);
    if (out_dtype == SVB_F32)
        k4_classifier_input_kernel<float><<<grid, K4_THREADS, 0, stream>>>(d_planes, d_t2_idx, d_t1_idx, quads, nrm,
                                                                           static_cast<float*>(d_out));
    else if (out_dtype == SVB_BF16)
        k4_classifier_input_kernel<__nv_bfloat16><<<grid, K4_THREADS, 0, stream>>>(d_planes, d_t2_idx, d_t1_idx, quads, nrm,
                                                                                   static_cast<__nv_bfloat16*>(d_out));
    else
        k4_classifier_input_kernel<__half><<<grid, K4_THREADS, 0, stream>>>(d_planes, d_t2_idx, d_t1_idx, quads, nrm,
                                                                            static_cast<__half*>(d_out));
    SVB_LAUNCHED();
    return SVB_OK;
}
