// svb_resize.cu -- K1 (min-max normalise + Pillow antialiased bilinear resize) and
// K3 (coordinate-driven crop + per-crop normalise + OpenCV fixed-point letterbox resize
// + Pillow bilinear second output).  Both are HBM-bound integer/byte kernels: coalesced
// 128-bit loads, shared-memory staging, no tensor cores.
//
// Arithmetic follows, step for step, what the reference executes through NumPy, Pillow and
// OpenCV (restated in oracle/fixedpoint.py):
//   normalize_to_uint8          spine_vision/io/__init__.py:15-30
//   predict_ivd_locations       spine_vision/datasets/classification/cropping.py:463-472
//   crop_region_horizontal      cropping.py:316-354
//   resize_with_padding         cropping.py:104-146
//   ClassificationDataset Resize  spine_vision/training/datasets/classification.py:247-278
#include <type_traits>

#include "svb_common.cuh"

namespace svb {

constexpr int PIL_PRECISION_BITS = 22;  // Pillow: 32 - 8 - 2
constexpr int CV_COEF_BITS = 11;        // OpenCV INTER_RESIZE_COEF_BITS

// ============================================================================ Pillow coefficients
// precompute_coeffs + normalize_coeffs_8bpc (Pillow src/libImaging/Resample.c) for the
// BILINEAR (triangle, support 1) filter, in IEEE double with explicit round-to-nearest ops so
// that nvcc cannot contract to FMA: the integer coefficients must equal the CPU's bit for bit.
__host__ __device__ inline int pil_ksize(int in_size, int out_size) {
    double scale = (double)(float)in_size / (double)out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(fs) * 2 + 1;
}

// One output index xx of one axis.  bounds = {xmin, n}; kk[0..ksize_max) zero padded.
__device__ void pil_coeff_entry(int in_size, int out_size, int xx, int ksize_max, int* bounds, int* kk) {
    const double scale = __ddiv_rn((double)(float)in_size, (double)out_size);
    const double fs = scale < 1.0 ? 1.0 : scale;
    const double support = fs;  // filter support 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, fs);
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    int n = xmax - xmin;
    if (n > ksize_max) n = ksize_max;  // cannot happen (ksize_max >= ksize); guards the table
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        ww = __dadd_rn(ww, w);
    }
    for (int x = 0; x < ksize_max; ++x) {
        int k = 0;
        if (x < n) {
            double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
            if (a < 0.0) a = -a;
            double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            // bilinear weights are never negative, the -0.5 branch of normalize_coeffs_8bpc is dead
            k = (int)__dadd_rn(0.5, __dmul_rn(w, (double)(1 << PIL_PRECISION_BITS)));
        }
        kk[x] = k;
    }
    bounds[0] = xmin;
    bounds[1] = n;
}

__device__ __forceinline__ uint32_t pil_clip8(int acc) {
    int v = acc >> PIL_PRECISION_BITS;
    return (uint32_t)min(max(v, 0), 255);
}

// ============================================================================ K1
// Workspace layout (all int32 unless noted):
//   keys   [B][2]                  ordered-uint min / max keys
//   hb     [B][out_w][2]           horizontal bounds
//   hk     [B][out_w][ksw]         horizontal coefficients
//   vb     [B][out_h][2]
//   vk     [B][out_h][ksh]
struct K1Layout {
    size_t keys, hb, hk, vb, vk, total;
    int ksw, ksh;
};
static K1Layout k1_layout(int B, int max_h, int max_w, int out_h, int out_w) {
    K1Layout L;
    L.ksw = pil_ksize(max_w, out_w);
    L.ksh = pil_ksize(max_h, out_h);
    size_t o = 0;
    L.keys = o; o += align_up((size_t)B * 2 * 4, 256);
    L.hb = o;   o += align_up((size_t)B * out_w * 2 * 4, 256);
    L.hk = o;   o += align_up((size_t)B * out_w * L.ksw * 4, 256);
    L.vb = o;   o += align_up((size_t)B * out_h * 2 * 4, 256);
    L.vk = o;   o += align_up((size_t)B * out_h * L.ksh * 4, 256);
    L.total = o;
    return L;
}

// grid (ceil(max(out_w,out_h)/128), 2, B): axis 0 = horizontal, 1 = vertical.  Also resets the keys.
__global__ void k1_coeff_kernel(const int32_t* __restrict__ hw, int out_h, int out_w, int ksh, int ksw,
                                uint32_t* __restrict__ keys, int* __restrict__ hb, int* __restrict__ hk,
                                int* __restrict__ vb, int* __restrict__ vk) {
    const int b = blockIdx.z;
    const int axis = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (axis == 0 && i == 0 && keys != nullptr) {
        keys[2 * b + 0] = 0xFFFFFFFFu;  // min key
        keys[2 * b + 1] = 0u;           // max key
    }
    const int h = hw[2 * b + 0], w = hw[2 * b + 1];
    if (axis == 0) {
        if (i < out_w) pil_coeff_entry(w, out_w, i, ksw, hb + ((size_t)b * out_w + i) * 2, hk + ((size_t)b * out_w + i) * ksw);
    } else {
        if (i < out_h) pil_coeff_entry(h, out_h, i, ksh, vb + ((size_t)b * out_h + i) * 2, vk + ((size_t)b * out_h + i) * ksh);
    }
}

// grid (chunks, nb): 128-bit coalesced min/max over slice (b0 + blockIdx.y)
__global__ void __launch_bounds__(512) k1_minmax_kernel(const float* __restrict__ slices, const int64_t* __restrict__ offs,
                                                        const int32_t* __restrict__ hw, int b0,
                                                        uint32_t* __restrict__ keys) {
    const int b = b0 + blockIdx.y;
    const float* base = slices + offs[b];
    const long long n = (long long)hw[2 * b] * hw[2 * b + 1];
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    long long lo = per * blockIdx.x, hi = min(lo + per, n);
    float mn = INFINITY, mx = -INFINITY;
    if (lo < hi) {
        const float* p = base + lo;
        long long cnt = hi - lo;
        // head up to 16-byte alignment
        int head = (int)(((16 - ((uintptr_t)p & 15)) & 15) >> 2);
        if (head > cnt) head = (int)cnt;
        if ((int)threadIdx.x < head) { float v = p[threadIdx.x]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        const float4* p4 = reinterpret_cast<const float4*>(p + head);
        long long n4 = (cnt - head) >> 2;
        for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 v = __ldg(p4 + i);
            mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
            mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        }
        long long done = head + (n4 << 2);
        if (done + threadIdx.x < cnt) { float v = p[done + threadIdx.x]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    __shared__ float smn[16], smx[16];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smn[wid] = mn; smx[wid] = mx; }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        mn = lane < nw ? smn[lane] : INFINITY;
        mx = lane < nw ? smx[lane] : -INFINITY;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0 && lo < hi) {
            atomicMin(&keys[2 * b + 0], float_key(mn));
            atomicMax(&keys[2 * b + 1], float_key(mx));
        }
    }
}

// normalize_px with the IEEE division taken only when it can matter.  The reference value is
// e = fl(fl((x - mn) / rng) * 255) truncated; y = fl((x - mn) * fl(255 / rng)) differs from e by < 5e-5 for
// 0 <= e <= 255, so trunc(y) == trunc(e) unless y sits within 1e-3 of an integer -- only then (and for
// NaN / inf, which fail the comparison) the exact sequence runs.  Bit-exact by construction.
// The two exact hits of that window that real slices are full of never reach the division either: a background pixel equal
// to the minimum (d == 0 -> 0 / rng * 255 = 0) and the maximum itself (d == rng -> 1 * 255 = 255); without the shortcut every
// warp that touches background took the IEEE-division subroutine (19 % of K1's instructions on the bench slices).
__device__ __forceinline__ uint32_t normalize_px_fast(float x, float mn, float rng, float k) {
    const float d = __fsub_rn(x, mn);
    const float y = __fmul_rn(d, k);
    if (fabsf(y - rintf(y)) > 1e-3f) return static_cast<uint32_t>(__float2int_rz(y)) & 0xFFu;
    if (d == 0.0f) return 0u;
    if (d == rng) return 255u;
    return cast_f32_u8(__fmul_rn(__fdiv_rn(d, rng), 255.0f));
}

// Branch-free first try of normalize_px_fast for a GROUP of pixels: returns the truncated fast product and clears `ok` when
// this pixel would have needed one of the exact paths (fast product within 1e-3 of an integer, NaN / inf).  d == 0 (a pixel
// equal to the minimum -- the background of an MRI slice) needs no exact path: 0 * k is exactly 0.  The caller redoes the
// whole group with normalize_px_fast when `ok` comes back false, so the result is the same function, with one branch per
// group instead of three per pixel (control flow was a third of K3's normalise pass, ncu source view).
__device__ __forceinline__ uint32_t normalize_px_try(float x, float mn, float k, bool& ok) {
    const float d = __fsub_rn(x, mn);
    const float y = __fmul_rn(d, k);
    ok = ok && (fabsf(y - rintf(y)) > 1e-3f || d == 0.0f);
    return static_cast<uint32_t>(__float2int_rz(y)) & 0xFFu;
}

// grid (ceil(out_h / R), nb).  CTA = R output rows of slice b, all out_w columns.
// smem: [src u8: rows_in_max * w_max + 8][tmp u8: rows_in_max * out_w][hk int: out_w * ksw][hb int: out_w*2][vk int: R * ksh][vb int: R*2]
__global__ void __launch_bounds__(512) k1_resize_kernel(const float* __restrict__ slices, const int64_t* __restrict__ offs,
                                                        const int32_t* __restrict__ hw, int b0, int out_h, int out_w,
                                                        int R, int ksh, int ksw, int src_cap, int rows_cap,
                                                        const uint32_t* __restrict__ keys, const int* __restrict__ hb,
                                                        const int* __restrict__ hk, const int* __restrict__ vb,
                                                        const int* __restrict__ vk, uint8_t* __restrict__ out,
                                                        float* __restrict__ minmax) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int b = b0 + blockIdx.y;
    const int H = hw[2 * b], W = hw[2 * b + 1];
    const int r0 = blockIdx.x * R;
    const int r1 = min(r0 + R, out_h);
    const int tid = threadIdx.x, nthr = blockDim.x;

    uint8_t* s_src = smem;                                   // src_cap bytes (multiple of 16)
    uint8_t* s_tmp = s_src + src_cap;                        // rows_cap * out_w (multiple of 16)
    int* s_hk = reinterpret_cast<int*>(s_tmp + (((size_t)rows_cap * out_w + 15) / 16) * 16);
    int* s_hb = s_hk + out_w * ksw;
    int* s_vk = s_hb + out_w * 2;
    int* s_vb = s_vk + R * ksh;

    const float mn = key_float(keys[2 * b + 0]);
    const float mx = key_float(keys[2 * b + 1]);
    const float rng = __fsub_rn(mx, mn);
    if (blockIdx.x == 0 && tid == 0 && minmax != nullptr) { minmax[2 * b] = mn; minmax[2 * b + 1] = mx; }

    const int* vb_b = vb + (size_t)b * out_h * 2;
    const int* vk_b = vk + (size_t)b * out_h * ksh;
    const int y_first = vb_b[2 * r0];
    const int y_last = vb_b[2 * (r1 - 1)] + vb_b[2 * (r1 - 1) + 1];  // bounds are monotone in the row index
    const int rows_in = y_last - y_first;

    // stage the coefficient tables
    {
        const int* hk_b = hk + (size_t)b * out_w * ksw;
        const int* hb_b = hb + (size_t)b * out_w * 2;
        for (int i = tid; i < out_w * ksw; i += nthr) s_hk[i] = hk_b[i];
        for (int i = tid; i < out_w * 2; i += nthr) s_hb[i] = hb_b[i];
        for (int i = tid; i < (r1 - r0) * ksh; i += nthr) s_vk[i] = vk_b[(size_t)r0 * ksh + i];
        for (int i = tid; i < (r1 - r0) * 2; i += nthr) s_vb[i] = vb_b[2 * r0 + i];
    }

    // phase A: contiguous span of rows_in full rows -> normalised u8 in smem (128-bit loads, 4 in flight per thread)
    const float* g = slices + offs[b] + (long long)y_first * W;
    const int mis = (int)(((uintptr_t)g >> 2) & 3);  // element misalignment of the span start
    {
        const long long cnt = (long long)rows_in * W;
        const int head = (4 - mis) & 3;
        uint8_t* dst = s_src + mis;  // element i lives at dst[i]; dst + head is 4-byte aligned
        if (tid < head && tid < cnt) dst[tid] = (uint8_t)normalize_px(g[tid], mn, rng);
        const float4* g4 = reinterpret_cast<const float4*>(g + head);
        const int n4 = cnt > head ? (int)((cnt - head) >> 2) : 0;
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
        if (rng > 0.0f) {
            const float k = __fdiv_rn(255.0f, rng);
            int i = tid;
            for (; i + 3 * nthr < n4; i += 4 * nthr) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(g4 + i + u * nthr);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bool ok = true;
                    uint32_t p = normalize_px_try(v[u].x, mn, k, ok) | (normalize_px_try(v[u].y, mn, k, ok) << 8) |
                                 (normalize_px_try(v[u].z, mn, k, ok) << 16) | (normalize_px_try(v[u].w, mn, k, ok) << 24);
                    if (!ok)
                        p = normalize_px_fast(v[u].x, mn, rng, k) | (normalize_px_fast(v[u].y, mn, rng, k) << 8) |
                            (normalize_px_fast(v[u].z, mn, rng, k) << 16) | (normalize_px_fast(v[u].w, mn, rng, k) << 24);
                    d4[i + u * nthr] = p;
                }
            }
            for (; i < n4; i += nthr) {
                const float4 v = __ldg(g4 + i);
                d4[i] = normalize_px_fast(v.x, mn, rng, k) | (normalize_px_fast(v.y, mn, rng, k) << 8) |
                        (normalize_px_fast(v.z, mn, rng, k) << 16) | (normalize_px_fast(v.w, mn, rng, k) << 24);
            }
        } else {  // constant slice: the reference casts the raw values (io/__init__.py:27-30)
            for (int i = tid; i < n4; i += nthr) {
                const float4 v = __ldg(g4 + i);
                d4[i] = cast_f32_u8(v.x) | (cast_f32_u8(v.y) << 8) | (cast_f32_u8(v.z) << 16) | (cast_f32_u8(v.w) << 24);
            }
        }
        const long long done = head + ((long long)n4 << 2);
        if (done + tid < cnt) dst[done + tid] = (uint8_t)normalize_px(g[done + tid], mn, rng);
    }
    __syncthreads();

    // phase B: horizontal pass, u8 rounding between passes (ImagingResampleHorizontal_8bpc).
    // thread = output column (its coefficients live in registers), loop over the staged rows
    {
        const uint8_t* src = s_src + mis;
        if (out_w == W) {
            for (int i = tid; i < rows_in * out_w; i += nthr) s_tmp[i] = src[i];
        } else if (ksw <= 8) {
            for (int xx = tid; xx < out_w; xx += nthr) {
                const int xmin = s_hb[2 * xx], n = s_hb[2 * xx + 1];
                int kc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) kc[j] = j < n ? s_hk[xx * ksw + j] : 0;
                // one pointer per row, the eight taps at immediate offsets.  A padded tap (zero coefficient) may read up to
                // 7 bytes past the row: the next row, or -- after the last row -- the 16 slack bytes behind the staged span
                const uint8_t* sp = src + xmin;
                uint8_t* dp = s_tmp + xx;
                for (int row = 0; row < rows_in; ++row, sp += W, dp += out_w) {
                    int acc = 1 << (PIL_PRECISION_BITS - 1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc += (int)sp[j] * kc[j];
                    *dp = (uint8_t)pil_clip8(acc);
                }
            }
        } else {
            for (int i = tid; i < rows_in * out_w; i += nthr) {
                const int row = i / out_w, xx = i - row * out_w;
                const int xmin = s_hb[2 * xx], n = s_hb[2 * xx + 1];
                const uint8_t* sp = src + (size_t)row * W + xmin;
                const int* kp = s_hk + xx * ksw;
                int acc = 1 << (PIL_PRECISION_BITS - 1);
                for (int j = 0; j < n; ++j) acc += (int)sp[j] * kp[j];
                s_tmp[i] = (uint8_t)pil_clip8(acc);
            }
        }
    }
    __syncthreads();

    // phase C: vertical pass (ImagingResampleVertical_8bpc), 4 output pixels per thread
    {
        uint8_t* out_b = out + (size_t)b * out_h * out_w;
        const int w4 = out_w >> 2;  // out_w % 4 == 0 is required by the host wrapper
        const int items = (r1 - r0) * w4;
        for (int i = tid; i < items; i += nthr) {
            const int rr = i / w4, x4 = (i - rr * w4) << 2;
            const int r = r0 + rr;
            uint32_t packed;
            if (out_h == H) {
                packed = *reinterpret_cast<const uint32_t*>(s_tmp + (size_t)(r - y_first) * out_w + x4);
            } else {
                const int ymin = s_vb[2 * rr] - y_first, n = s_vb[2 * rr + 1];
                const int* kp = s_vk + rr * ksh;
                int a0 = 1 << (PIL_PRECISION_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
                for (int j = 0; j < n; ++j) {
                    const uint32_t px = *reinterpret_cast<const uint32_t*>(s_tmp + (size_t)(ymin + j) * out_w + x4);
                    const int k = kp[j];
                    a0 += (int)(px & 0xFF) * k;
                    a1 += (int)((px >> 8) & 0xFF) * k;
                    a2 += (int)((px >> 16) & 0xFF) * k;
                    a3 += (int)(px >> 24) * k;
                }
                packed = pil_clip8(a0) | (pil_clip8(a1) << 8) | (pil_clip8(a2) << 16) | (pil_clip8(a3) << 24);
            }
            *reinterpret_cast<uint32_t*>(out_b + (size_t)r * out_w + x4) = packed;
        }
    }
}

// ---- normalise only (no resize): the localization dataset builder's per-image normalize_to_uint8
// (spine_vision/datasets/localization.py:262-267, 147-151).  grid (chunks, nb); the uint8 pool uses the same element
// offsets as the float32 pool (slices start 16-byte aligned, so every 128-bit load pairs with one 32-bit store).
__global__ void k1_keys_reset_kernel(uint32_t* __restrict__ keys, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { keys[2 * b + 0] = 0xFFFFFFFFu; keys[2 * b + 1] = 0u; }
}

__global__ void __launch_bounds__(512) k1_normalize_only_kernel(const float* __restrict__ slices, const int64_t* __restrict__ offs,
                                                                const int32_t* __restrict__ hw, int b0,
                                                                const uint32_t* __restrict__ keys, uint8_t* __restrict__ out,
                                                                float* __restrict__ minmax) {
    const int b = b0 + blockIdx.y;
    const float* base = slices + offs[b];
    uint8_t* dst = out + offs[b];
    const long long n = (long long)hw[2 * b] * hw[2 * b + 1];
    const float mn = key_float(keys[2 * b + 0]), mx = key_float(keys[2 * b + 1]);
    const float rng = __fsub_rn(mx, mn);
    if (blockIdx.x == 0 && threadIdx.x == 0 && minmax != nullptr) { minmax[2 * b] = mn; minmax[2 * b + 1] = mx; }
    const float k = rng > 0.0f ? __fdiv_rn(255.0f, rng) : 0.0f;
    const long long n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(base);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 128-bit loads in flight
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(p4 + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            d4[i + u * stride] = rng > 0.0f ? (normalize_px_fast(v[u].x, mn, rng, k) | (normalize_px_fast(v[u].y, mn, rng, k) << 8) |
                                               (normalize_px_fast(v[u].z, mn, rng, k) << 16) | (normalize_px_fast(v[u].w, mn, rng, k) << 24))
                                            : (cast_f32_u8(v[u].x) | (cast_f32_u8(v[u].y) << 8) | (cast_f32_u8(v[u].z) << 16) | (cast_f32_u8(v[u].w) << 24));
    }
    for (; i < n4; i += stride) {
        const float4 v = __ldg(p4 + i);
        d4[i] = rng > 0.0f ? (normalize_px_fast(v.x, mn, rng, k) | (normalize_px_fast(v.y, mn, rng, k) << 8) |
                              (normalize_px_fast(v.z, mn, rng, k) << 16) | (normalize_px_fast(v.w, mn, rng, k) << 24))
                           : (cast_f32_u8(v.x) | (cast_f32_u8(v.y) << 8) | (cast_f32_u8(v.z) << 16) | (cast_f32_u8(v.w) << 24));
    }
    if (blockIdx.x == 0) {
        const long long t = (n4 << 2) + threadIdx.x;
        if (t < n) dst[t] = (uint8_t)normalize_px(base[t], mn, rng);
    }
}

// rows of input one CTA of R output rows can need, for the largest slice of the batch
static int k1_rows_in(int in_h, int out_h, int R) {
    double scale = (double)in_h / out_h;
    double fs = scale < 1.0 ? 1.0 : scale;
    return (int)(R * scale + 2.0 * fs + 3.0);
}

}  // namespace svb

using namespace svb;

extern "C" size_t svb_k1_workspace_bytes(int B, int max_h, int max_w, int out_h, int out_w) {
    if (B <= 0 || max_h <= 0 || max_w <= 0 || out_h <= 0 || out_w <= 0) return 0;
    return k1_layout(B, max_h, max_w, out_h, out_w).total;
}

namespace svb {

// launch plan of the K1 pair for one batch: validated sizes, workspace pointers, rows per CTA
struct K1Plan {
    K1Layout L;
    int R = 32, src_cap = 0, rows_cap = 0, mm_blocks = 1, group = 1;
    size_t smem_bytes = 0;
    uint32_t* keys = nullptr;
    int *hb = nullptr, *hk = nullptr, *vb = nullptr, *vk = nullptr;
};

// Slices per (min/max, resize) launch pair -- or per (K0, resize) pair of the fused entry.  Default: the whole batch in one
// pair.  The second launch re-reads what the first one read (or wrote) from HBM, but it is bound by instruction issue (Pillow's
// fixed-point taps), so the re-read hides under the arithmetic.  Measured alternative (B200, round 2, SVB_K1_GROUP=8: groups of
// 46 MB that the second launch reads back from the 126 MB L2): DRAM reads fall to ~1.1x the algorithmic bytes but the 32 launch
// pairs of 128 CTAs no longer fill 148 SMs x 2 -- K1 1.05 -> 1.65 ms per 256 slices, end-to-end 3,690 -> 3,448 series/s.
// Time, not traffic, is what the step pays for; SVB_K1_GROUP=n keeps the grouped schedule available for profiling.
static int k1_group_size(int B, int max_h, int max_w) {
    (void)max_h; (void)max_w;
    const char* e = getenv("SVB_K1_GROUP");  // read on every call (tests walk the settings)
    const int forced = (e && e[0]) ? atoi(e) : 0;
    if (forced <= 0) return B;
    return forced < B ? forced : B;
}

// validates, sizes the resize kernel, builds the Pillow coefficient tables (one launch); `reset_keys`: the same launch
// resets the per-slice min / max keys
static int k1_prepare(const int32_t* d_hw, int B, int max_h, int max_w, int out_h, int out_w, void* d_ws, size_t ws_bytes,
                      bool reset_keys, cudaStream_t stream, K1Plan* P) {
    SVB_REQUIRE(out_w % 4 == 0, SVB_ERR_INVALID_ARG, "k1: out_w (%d) must be a multiple of 4", out_w);
    P->L = k1_layout(B, max_h, max_w, out_h, out_w);
    const K1Layout& L = P->L;
    SVB_REQUIRE(ws_bytes >= L.total, SVB_ERR_WORKSPACE_TOO_SMALL, "k1: workspace %zu < %zu bytes", ws_bytes, L.total);
    uint8_t* ws = static_cast<uint8_t*>(d_ws);
    P->keys = reinterpret_cast<uint32_t*>(ws + L.keys);
    P->hb = reinterpret_cast<int*>(ws + L.hb);
    P->hk = reinterpret_cast<int*>(ws + L.hk);
    P->vb = reinterpret_cast<int*>(ws + L.vb);
    P->vk = reinterpret_cast<int*>(ws + L.vk);
    {
        dim3 grid(ceil_div(out_w > out_h ? out_w : out_h, 128), 2, B);
        k1_coeff_kernel<<<grid, 128, 0, stream>>>(d_hw, out_h, out_w, L.ksh, L.ksw, reset_keys ? P->keys : nullptr, P->hb, P->hk, P->vb, P->vk);
        SVB_LAUNCHED();
    }
    // rows per CTA: largest power of two whose staging fits ~100 KB (2 CTAs / SM), else smaller
    int R = 32;
    size_t smem_bytes = 0;
    int src_cap = 0, rows_cap = 0;
    for (;; R >>= 1) {
        rows_cap = k1_rows_in(max_h, out_h, R);
        if (rows_cap > max_h) rows_cap = max_h;
        src_cap = (int)align_up((size_t)rows_cap * max_w + 16, 16);
        const size_t table_bytes = (size_t)out_w * L.ksw * 4 + (size_t)out_w * 2 * 4 + (size_t)R * L.ksh * 4 + (size_t)R * 2 * 4;
        smem_bytes = (size_t)src_cap + align_up((size_t)rows_cap * out_w, 16) + table_bytes;
        const size_t budget = R > 1 ? 100 * 1024 : 227 * 1024;
        if (smem_bytes <= budget || R == 1) break;
    }
    SVB_REQUIRE(smem_bytes <= 227 * 1024, SVB_ERR_INVALID_ARG,
                "k1: a %dx%d slice needs %zu bytes of shared memory per output row (limit 232448)", max_h, max_w,
                smem_bytes);
    SVB_CUDA_OK(cudaFuncSetAttribute(k1_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    P->R = R; P->src_cap = src_cap; P->rows_cap = rows_cap; P->smem_bytes = smem_bytes;
    const size_t slice_bytes = (size_t)max_h * max_w * 4;
    int mm_blocks = (int)ceil_div<size_t>(slice_bytes, (size_t)512 * 16 * 8);  // ~8 float4 per thread
    P->mm_blocks = mm_blocks < 1 ? 1 : (mm_blocks > 1024 ? 1024 : mm_blocks);
    P->group = k1_group_size(B, max_h, max_w);
    if (P->group > 65535) P->group = 65535;
    return SVB_OK;
}
static int k1_minmax_group(const K1Plan& P, const float* d_slices, const int64_t* d_offs, const int32_t* d_hw, int b0, int nb,
                           cudaStream_t stream) {
    k1_minmax_kernel<<<dim3(P.mm_blocks, nb), 512, 0, stream>>>(d_slices, d_offs, d_hw, b0, P.keys);
    SVB_LAUNCHED();
    return SVB_OK;
}
static int k1_resize_group(const K1Plan& P, const float* d_slices, const int64_t* d_offs, const int32_t* d_hw, int b0, int nb,
                           int out_h, int out_w, uint8_t* d_out_u8, float* d_minmax, cudaStream_t stream) {
    k1_resize_kernel<<<dim3(ceil_div(out_h, P.R), nb), 512, P.smem_bytes, stream>>>(
        d_slices, d_offs, d_hw, b0, out_h, out_w, P.R, P.L.ksh, P.L.ksw, P.src_cap, P.rows_cap, P.keys, P.hb, P.hk, P.vb, P.vk,
        d_out_u8, d_minmax);
    SVB_LAUNCHED();
    return SVB_OK;
}

}  // namespace svb

extern "C" int svb_k1_normalize_resize(const float* d_slices, const int64_t* d_offs, const int32_t* d_hw, int B,
                                       int max_h, int max_w, int out_h, int out_w, uint8_t* d_out_u8,
                                       float* d_minmax, void* d_ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = check_device_sm100()) return rc;
    SVB_REQUIRE(B >= 0 && max_h > 0 && max_w > 0 && out_h > 0 && out_w > 0, SVB_ERR_INVALID_ARG,
                "k1: bad sizes B=%d max_hw=(%d,%d) out=(%d,%d)", B, max_h, max_w, out_h, out_w);
    if (B == 0) return SVB_OK;
    SVB_REQUIRE(d_slices && d_offs && d_hw && d_out_u8 && d_ws, SVB_ERR_INVALID_ARG, "k1: null pointer argument");
    K1Plan P;
    if (int rc = k1_prepare(d_hw, B, max_h, max_w, out_h, out_w, d_ws, ws_bytes, true, stream, &P)) return rc;
    // Per group of slices: min/max of every slice (the normalisation needs the global extrema first), then normalise + resize.
    // The second launch re-reads the group while it is still in L2.
    for (int b0 = 0; b0 < B; b0 += P.group) {
        const int nb = (B - b0) < P.group ? (B - b0) : P.group;
        if (int rc = k1_minmax_group(P, d_slices, d_offs, d_hw, b0, nb, stream)) return rc;
        if (int rc = k1_resize_group(P, d_slices, d_offs, d_hw, b0, nb, out_h, out_w, d_out_u8, d_minmax, stream)) return rc;
    }
    return SVB_OK;
}

extern "C" size_t svb_normalize_u8_workspace_bytes(int B) { return B > 0 ? align_up((size_t)B * 2 * 4, 256) + 256 : 0; }

extern "C" int svb_normalize_u8(const float* d_slices, const int64_t* d_offs, const int32_t* d_hw, int B, int max_h, int max_w,
                                uint8_t* d_out_u8, float* d_minmax, void* d_ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = check_device_sm100()) return rc;
    SVB_REQUIRE(B >= 0 && max_h > 0 && max_w > 0, SVB_ERR_INVALID_ARG, "normalize_u8: bad sizes B=%d max_hw=(%d,%d)", B, max_h, max_w);
    if (B == 0) return SVB_OK;
    SVB_REQUIRE(d_slices && d_offs && d_hw && d_out_u8 && d_ws, SVB_ERR_INVALID_ARG, "normalize_u8: null pointer argument");
    SVB_REQUIRE(ws_bytes >= svb_normalize_u8_workspace_bytes(B), SVB_ERR_WORKSPACE_TOO_SMALL, "normalize_u8: workspace %zu < %zu bytes",
                ws_bytes, svb_normalize_u8_workspace_bytes(B));
    uint32_t* keys = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~uintptr_t(255));
    k1_keys_reset_kernel<<<ceil_div(B, 256), 256, 0, stream>>>(keys, B);
    SVB_LAUNCHED();
    const size_t slice_bytes = (size_t)max_h * max_w * 4;
    int blocks = (int)ceil_div<size_t>(slice_bytes, (size_t)512 * 16 * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 1024) blocks = 1024;
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
        k1_minmax_kernel<<<dim3(blocks, nb), 512, 0, stream>>>(d_slices, d_offs, d_hw, b0, keys);
        SVB_LAUNCHED();
    }
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
        k1_normalize_only_kernel<<<dim3(blocks, nb), 512, 0, stream>>>(d_slices, d_offs, d_hw, b0, keys, d_out_u8, d_minmax);
        SVB_LAUNCHED();
    }
    return SVB_OK;
}

// ============================================================================ K3
namespace svb {

struct K3Geom {
    int x1, x2, y1, y2, new_h, new_w, y_off, x_off;
    double scale_x, scale_y;  // cv2.resize: 1.0 / (dst / src) per axis, double
};

// Workspace: Pillow tables for the second output: hb2[ow2][2], hk2[ow2][ks2w], vb2[oh2][2], vk2[oh2][ks2h]
struct K3Layout {
    size_t hb, hk, vb, vk, total;
    int ksw, ksh;
};
static K3Layout k3_layout(int ch, int cw, int oh2, int ow2) {
    K3Layout L{};
    if (oh2 <= 0 || ow2 <= 0) { L.total = 256; L.ksw = L.ksh = 1; return L; }
    L.ksw = pil_ksize(cw, ow2);
    L.ksh = pil_ksize(ch, oh2);
    size_t o = 0;
    L.hb = o; o += align_up((size_t)ow2 * 2 * 4, 256);
    L.hk = o; o += align_up((size_t)ow2 * L.ksw * 4, 256);
    L.vb = o; o += align_up((size_t)oh2 * 2 * 4, 256);
    L.vk = o; o += align_up((size_t)oh2 * L.ksh * 4, 256);
    L.total = o;
    return L;
}

__global__ void k3_coeff_kernel(int ch, int cw, int oh2, int ow2, int ksh, int ksw, int* __restrict__ hb,
                                int* __restrict__ hk, int* __restrict__ vb, int* __restrict__ vk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.y == 0) {
        if (i < ow2) pil_coeff_entry(cw, ow2, i, ksw, hb + 2 * i, hk + (size_t)i * ksw);
    } else {
        if (i < oh2) pil_coeff_entry(ch, oh2, i, ksh, vb + 2 * i, vk + (size_t)i * ksh);
    }
}

// OpenCV resizeGeneric_ linear table entry (modules/imgproc/src/resize.cpp): source index and the
// two 11-bit weights for destination index i.  `horizontal` zeroes the fraction when clamped.
__device__ __forceinline__ void cv_axis_entry(int i, int src, double scale, bool horizontal, int& si, int& a0, int& a1) {
    float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)i, 0.5), scale), 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (horizontal) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), (float)(1 << CV_COEF_BITS)));
    a1 = __float2int_rn(__fmul_rn(f, (float)(1 << CV_COEF_BITS)));
    si = s;
}

// One CTA per crop.
// smem: [box u8: box_cap][canvas u8: ch*cw][tmp2 u8: ch*ow2 (if second output needs a pass)]
//       [xs,xa0,xa1: 3*cw int][ys,yb0,yb1: 3*ch int]
constexpr int K3_THREADS = 512;  // 16 warps per crop, two crops per SM: the kernel is latency-bound, so warps are what it needs

__global__ void __launch_bounds__(K3_THREADS, 2) k3_crop_kernel(const float* __restrict__ slices, const int64_t* __restrict__ offs,
                                                      const int32_t* __restrict__ hw, const int32_t* __restrict__ slice_idx,
                                                      const float* __restrict__ xy, const int32_t* __restrict__ delta_px,
                                                      int ch, int cw, uint8_t* __restrict__ crops, int oh2, int ow2,
                                                      uint8_t* __restrict__ crops2, int32_t* __restrict__ geom_out,
                                                      int flags, int box_cap, int ksh2, int ksw2, const int* __restrict__ hb2,
                                                      const int* __restrict__ hk2, const int* __restrict__ vb2,
                                                      const int* __restrict__ vk2, const double* __restrict__ rot,
                                                      const int32_t* __restrict__ pixel_kind) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nwarps = nthr >> 5;
    const bool second = crops2 != nullptr;
    const bool second_identity = second && oh2 == ch && ow2 == cw;

    uint8_t* s_box = smem;
    uint8_t* s_canvas = s_box + box_cap;  // box_cap is a multiple of 16
    uint8_t* s_tmp2 = s_canvas + (size_t)ch * cw;
    int* s_tab = reinterpret_cast<int*>(s_tmp2 + ((second && !second_identity) ? (size_t)ch * ow2 : 0));
    int* xs = s_tab;
    int* xa0 = xs + cw;
    int* xa1 = xa0 + cw;
    int* ys = xa1 + cw;
    int* yb0 = ys + ch;
    int* yb1 = yb0 + ch;

    __shared__ K3Geom g;
    __shared__ float s_red[2][K3_THREADS / 32];
    __shared__ float s_mm[2];

    const int b = slice_idx[n];
    const int H = hw[2 * b], W = hw[2 * b + 1];
    if (tid == 0) {
        // cropping.py:338-348: cx = int(x*w), cy = int(y*h) in Python float (double) arithmetic
        // (the model's float32 outputs widened to double, as float(output_np[i, 0]) does at cropping.py:481-483; or doubles
        //  as given -- SVB_K3_XY_F64 -- for centres that never were float32: the fallback table, user dictionaries)
        const double xn = (flags & SVB_K3_XY_F64) ? reinterpret_cast<const double*>(xy)[2 * n + 0] : (double)xy[2 * n + 0];
        const double yn = (flags & SVB_K3_XY_F64) ? reinterpret_cast<const double*>(xy)[2 * n + 1] : (double)xy[2 * n + 1];
        const int cx = (int)__dmul_rn(xn, (double)W);
        const int cy = (int)__dmul_rn(yn, (double)H);
        const int left = delta_px[4 * n + 0], right = delta_px[4 * n + 1];
        const int top = delta_px[4 * n + 2], bottom = delta_px[4 * n + 3];
        K3Geom q;
        q.x1 = max(0, cx - left);
        q.x2 = min(W, cx + right);
        q.y1 = max(0, cy - top);
        q.y2 = min(H, cy + bottom);
        const int bh = q.y2 - q.y1, bw = q.x2 - q.x1;
        q.new_h = q.new_w = q.y_off = q.x_off = 0;
        if (bh > 0 && bw > 0 && (long long)bh * bw <= box_cap) {
            // cropping.py:121-141: scale = min(th/h, tw/w); new = int(round(dim*scale)); offsets (t-new)//2
            const double sh = __ddiv_rn((double)ch, (double)bh), sw = __ddiv_rn((double)cw, (double)bw);
            const double scale = sh < sw ? sh : sw;
            q.new_h = (int)rint(__dmul_rn((double)bh, scale));
            q.new_w = (int)rint(__dmul_rn((double)bw, scale));
            q.y_off = (ch - q.new_h) >> 1;  // floor division, also for a negative numerator
            q.x_off = (cw - q.new_w) >> 1;
            if (q.new_h <= 0 || q.new_w <= 0 || q.new_h > ch || q.new_w > cw) q.new_h = q.new_w = 0;
        }
        q.scale_x = q.scale_y = 1.0;
        if (q.new_h > 0 && q.new_w > 0) {  // two IEEE double divisions per axis: once per crop, not once per thread
            q.scale_x = __ddiv_rn(1.0, __ddiv_rn((double)q.new_w, (double)bw));
            q.scale_y = __ddiv_rn(1.0, __ddiv_rn((double)q.new_h, (double)bh));
        }
        g = q;
        if (geom_out != nullptr) {
            int32_t* go = geom_out + 8 * (size_t)n;
            go[0] = q.x1; go[1] = q.x2; go[2] = q.y1; go[3] = q.y2;
            go[4] = q.new_h; go[5] = q.new_w; go[6] = q.y_off; go[7] = q.x_off;
        }
    }
    __syncthreads();
    const int bh = g.y2 - g.y1, bw = g.x2 - g.x1;
    const bool valid = g.new_h > 0 && g.new_w > 0;
    const float* img = slices + offs[b];
    const float* src = img + (long long)g.y1 * W + g.x1;
    // rotated mode (cropping.py:258-313): the box is cut from cv2.warpAffine(image, R, INTER_LINEAR, BORDER_REPLICATE).
    // rot = the INVERTED 2x3 map (host, double).  Per destination pixel OpenCV takes the source position in 1/1024 px
    // fixed point, rounds it to 1/32 px, and blends four taps with exact fp32 bilinear weights, left to right.
    // The reference warps the slice in the FILE's pixel type (extract_middle_slice keeps it, cropping.py:63-79): float32
    // sources blend in fp32 (remapBilinear<Cast<float,float>>); 16-bit integer sources blend the same way and are rounded
    // back to the type (Cast<float,short/ushort> = cvRound, half to even); uint8 sources use OpenCV's 15-bit fixed-point
    // weights ((32-fx)(32-fy)*32, ... exact for bilinear) and (sum + 2^14) >> 15.
    double r00 = 0, r01 = 0, r02 = 0, r10 = 0, r11 = 0, r12 = 0;
    int kind = SVB_PIXEL_FLOAT;
    if (rot != nullptr) {
        const double* rp = rot + 6 * (size_t)n;
        r00 = rp[0]; r01 = rp[1]; r02 = rp[2]; r10 = rp[3]; r11 = rp[4]; r12 = rp[5];
        if (pixel_kind != nullptr) kind = pixel_kind[b];
    }
    auto box_value = [&](int bx, int by) -> float {
        if (rot == nullptr) return __ldg(src + (long long)by * W + bx);
        const int x = g.x1 + bx, y = g.y1 + by;
        const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(r00, (double)x), 1024.0));
        const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(r10, (double)x), 1024.0));
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(r01, (double)y), r02), 1024.0)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(r11, (double)y), r12), 1024.0)) + 16;
        const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
        const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
        const float vx1 = __fmul_rn((float)(X & 31), 0.03125f), vy1 = __fmul_rn((float)(Y & 31), 0.03125f);
        const float vx0 = __fsub_rn(1.0f, vx1), vy0 = __fsub_rn(1.0f, vy1);
        const int xa = min(max(sx, 0), W - 1), xb = min(max(sx + 1, 0), W - 1);
        const int ya = min(max(sy, 0), H - 1), yb = min(max(sy + 1, 0), H - 1);
        const float s00 = __ldg(img + (long long)ya * W + xa), s01 = __ldg(img + (long long)ya * W + xb);
        const float s10 = __ldg(img + (long long)yb * W + xa), s11 = __ldg(img + (long long)yb * W + xb);
        if (kind == SVB_PIXEL_UINT8) {
            const int fx = X & 31, fy = Y & 31;
            const int acc = (int)s00 * ((32 - fy) * (32 - fx) * 32) + (int)s01 * ((32 - fy) * fx * 32) +
                            (int)s10 * (fy * (32 - fx) * 32) + (int)s11 * (fy * fx * 32);
            return (float)min(max((acc + (1 << 14)) >> 15, 0), 255);
        }
        float acc = __fadd_rn(__fmul_rn(s00, __fmul_rn(vy0, vx0)), __fmul_rn(s01, __fmul_rn(vy0, vx1)));
        acc = __fadd_rn(acc, __fmul_rn(s10, __fmul_rn(vy1, vx0)));
        acc = __fadd_rn(acc, __fmul_rn(s11, __fmul_rn(vy1, vx1)));
        if (kind == SVB_PIXEL_INT16) acc = fminf(fmaxf(rintf(acc), -32768.0f), 32767.0f);
        else if (kind == SVB_PIXEL_UINT16) acc = fminf(fmaxf(rintf(acc), 0.0f), 65535.0f);
        return acc;
    };

    if (valid) {
        // pass 1: per-crop min / max (normalize_to_uint8 on the crop, cropping.py:350).  A warp walks whole box rows
        // (lane = column, stride 32), two rows x eight column groups = up to 16 independent loads in flight per thread;
        // the address of a load is one add away from the row pointer (the kernel is issue-bound).
        float mn = INFINITY, mx = -INFINITY;
        if (rot == nullptr) {
            // axis-aligned box: row pointers once per row pair, every load one immediate offset away; a column beyond the
            // box repeats the thread's first pixel and a missing second row re-reads the first (min / max do not care)
            for (int y = wid; y < bh; y += 2 * nwarps) {
                const float* ra = src + (long long)y * W;
                const float* rb = (y + nwarps < bh) ? ra + (long long)nwarps * W : ra;
                for (int x0 = lane; x0 < bw; x0 += 256) {
                    const float* pa = ra + x0;
                    const float* pb = rb + x0;
                    float v[16];
                    v[0] = __ldg(pa);
                    v[8] = __ldg(pb);
#pragma unroll
                    for (int u = 1; u < 8; ++u) {
                        const bool in = x0 + 32 * u < bw;
                        v[u] = in ? __ldg(pa + 32 * u) : v[0];
                        v[8 + u] = in ? __ldg(pb + 32 * u) : v[8];
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        mn = fminf(mn, v[u]);
                        mx = fmaxf(mx, v[u]);
                    }
                }
            }
        } else {
        for (int y = wid; y < bh; y += 2 * nwarps) {
            const int y2 = y + nwarps;
            for (int x0 = lane; x0 < bw; x0 += 256) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u;
                    v[u] = x < bw ? box_value(x, y) : 0.0f;
                    v[8 + u] = (x < bw && y2 < bh) ? box_value(x, y2) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u;
                    if (x < bw) {
                        mn = fminf(mn, v[u]);
                        mx = fmaxf(mx, v[u]);
                        if (y2 < bh) {
                            mn = fminf(mn, v[8 + u]);
                            mx = fmaxf(mx, v[8 + u]);
                        }
                    }
                }
            }
        }
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) { s_red[0][wid] = mn; s_red[1][wid] = mx; }
        __syncthreads();
        if (wid == 0) {
            mn = lane < nwarps ? s_red[0][lane] : INFINITY;
            mx = lane < nwarps ? s_red[1][lane] : -INFINITY;
            mn = warp_min(mn);
            mx = warp_max(mx);
            if (lane == 0) { s_mm[0] = mn; s_mm[1] = mx; }
        }
        // resize tables while the reduction finishes
        {
            const double scale_x = g.scale_x, scale_y = g.scale_y;
            for (int i = tid; i < g.new_w; i += nthr) cv_axis_entry(i, bw, scale_x, true, xs[i], xa0[i], xa1[i]);
            for (int i = tid; i < g.new_h; i += nthr) cv_axis_entry(i, bh, scale_y, false, ys[i], yb0[i], yb1[i]);
        }
        __syncthreads();
        const float mnv = s_mm[0];
        // rng <= 0 makes normalize_px a plain cast: the max == min case, and SVB_K3_NO_NORMALIZE
        const float rng = (flags & SVB_K3_NO_NORMALIZE) ? 0.0f : __fsub_rn(s_mm[1], mnv);
        // pass 2: normalise the box into shared memory (second read is an L1/L2 hit); same row-wise walk
        const float kfast = rng > 0.0f ? __fdiv_rn(255.0f, rng) : 0.0f;
        if (rot == nullptr) {
            // axis-aligned box, same walk as pass 1; the scale-or-cast decision is taken once, outside the loops
            auto rows = [&](auto conv) {
                for (int y = wid; y < bh; y += 2 * nwarps) {
                    const bool two = y + nwarps < bh;
                    const float* ra = src + (long long)y * W;
                    const float* rb = two ? ra + (long long)nwarps * W : ra;
                    uint8_t* da = s_box + y * bw;  // box pixels are stored row-major
                    uint8_t* db = da + nwarps * bw;
                    for (int x0 = lane; x0 < bw; x0 += 256) {
                        const float* pa = ra + x0;
                        const float* pb = rb + x0;
                        float v[16];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const bool in = x0 + 32 * u < bw;
                            v[u] = in ? __ldg(pa + 32 * u) : 0.0f;
                            v[8 + u] = in ? __ldg(pb + 32 * u) : 0.0f;
                        }
                        uint8_t* qa = da + x0;
                        uint8_t* qb = db + x0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (x0 + 32 * u < bw) {
                                qa[32 * u] = conv(v[u]);
                                if (two) qb[32 * u] = conv(v[8 + u]);
                            }
                        }
                    }
                }
            };
            // per-pixel fast / exact decision here: grouping 16 pixels behind one branch (as K1 does with 4) was measured
            // SLOWER for crops (295.9 vs 277.5 us / 1280 crops) -- two warps in three hold at least one near-integer pixel
            if (rng > 0.0f) rows([&](float v) { return (uint8_t)normalize_px_fast(v, mnv, rng, kfast); });
            else rows([&](float v) { return (uint8_t)cast_f32_u8(v); });
        } else {
        for (int y = wid; y < bh; y += 2 * nwarps) {
            const int y2 = y + nwarps;
            for (int x0 = lane; x0 < bw; x0 += 256) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u;
                    v[u] = x < bw ? box_value(x, y) : 0.0f;
                    v[8 + u] = (x < bw && y2 < bh) ? box_value(x, y2) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u;
                    if (x < bw) {  // box pixels are stored row-major
                        s_box[y * bw + x] = (uint8_t)(rng > 0.0f ? normalize_px_fast(v[u], mnv, rng, kfast) : cast_f32_u8(v[u]));
                        if (y2 < bh)
                            s_box[y2 * bw + x] = (uint8_t)(rng > 0.0f ? normalize_px_fast(v[8 + u], mnv, rng, kfast) : cast_f32_u8(v[8 + u]));
                    }
                }
            }
        }
        }
    }
    __syncthreads();

    // letterboxed fixed-point bilinear resize (cv2.resize 8U INTER_LINEAR) onto the zero canvas
    {
        const int cw4 = cw >> 2;
        const bool same = (g.new_h == bh && g.new_w == bw);  // cv2.resize returns a copy
        if (nthr % cw4 == 0) {
            // a thread keeps ONE quad of output columns for all its rows: the four columns' source index and weights
            // are fetched once, and no index has to be divided (the kernel is issue-bound, not bandwidth-bound)
            const int ox4 = (tid % cw4) << 2, oy0 = tid / cw4, oys = nthr / cw4;
            int sx0[4], sx1[4], wa0[4], wa1[4], dxs[4];
            bool inx[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dx = ox4 + k - g.x_off;
                inx[k] = valid && dx >= 0 && dx < g.new_w;
                dxs[k] = inx[k] ? dx : 0;
                sx0[k] = inx[k] ? xs[dx] : 0;
                sx1[k] = min(sx0[k] + 1, bw - 1);
                wa0[k] = inx[k] ? xa0[dx] : 0;
                wa1[k] = inx[k] ? xa1[dx] : 0;
            }
            for (int oy = oy0; oy < ch; oy += oys) {
                uint32_t packed = 0;
                const int dy = oy - g.y_off;
                if (valid && dy >= 0 && dy < g.new_h) {
                    const int sy = ys[dy];
                    const int r0 = min(max(sy, 0), bh - 1), r1 = min(max(sy + 1, 0), bh - 1);
                    const int b0 = yb0[dy], b1 = yb1[dy];
                    const uint8_t* p0 = s_box + (size_t)r0 * bw;
                    const uint8_t* p1 = s_box + (size_t)r1 * bw;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t v = 0;
                        if (inx[k]) {
                            if (same) {
                                v = s_box[(size_t)dy * bw + dxs[k]];
                            } else {
                                const int h0 = (int)p0[sx0[k]] * wa0[k] + (int)p0[sx1[k]] * wa1[k];
                                const int h1 = (int)p1[sx0[k]] * wa0[k] + (int)p1[sx1[k]] * wa1[k];
                                const int t = ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16);
                                v = (uint32_t)min(max((t + 2) >> 2, 0), 255);
                            }
                        }
                        packed |= v << (8 * k);
                    }
                }
                *reinterpret_cast<uint32_t*>(s_canvas + (size_t)oy * cw + ox4) = packed;
            }
        } else {
            for (int i = tid; i < ch * cw4; i += nthr) {
                const int oy = i / cw4, ox4 = (i - oy * cw4) << 2;
                uint32_t packed = 0;
                const int dy = oy - g.y_off;
                if (valid && dy >= 0 && dy < g.new_h) {
                    const int sy = ys[dy];
                    const int r0 = min(max(sy, 0), bh - 1), r1 = min(max(sy + 1, 0), bh - 1);
                    const int b0 = yb0[dy], b1 = yb1[dy];
                    const uint8_t* p0 = s_box + (size_t)r0 * bw;
                    const uint8_t* p1 = s_box + (size_t)r1 * bw;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int dx = ox4 + k - g.x_off;
                        uint32_t v = 0;
                        if (dx >= 0 && dx < g.new_w) {
                            if (same) {
                                v = s_box[(size_t)dy * bw + dx];
                            } else {
                                const int sx = xs[dx];
                                const int sx1 = min(sx + 1, bw - 1);
                                const int a0 = xa0[dx], a1 = xa1[dx];
                                const int h0 = (int)p0[sx] * a0 + (int)p0[sx1] * a1;
                                const int h1 = (int)p1[sx] * a0 + (int)p1[sx1] * a1;
                                const int t = ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16);
                                v = (uint32_t)min(max((t + 2) >> 2, 0), 255);
                            }
                        }
                        packed |= v << (8 * k);
                    }
                }
                *reinterpret_cast<uint32_t*>(s_canvas + (size_t)oy * cw + ox4) = packed;
            }
        }
    }
    __syncthreads();

    // crop out: ch*cw bytes, 32-bit coalesced
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(crops + (size_t)n * ch * cw);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(s_canvas);
        for (int i = tid; i < (ch * cw) >> 2; i += nthr) dst[i] = s[i];
    }
    if (!second) return;

    uint8_t* out2 = crops2 + (size_t)n * oh2 * ow2;
    if (second_identity) {  // Pillow skips both passes when the size is unchanged
        uint32_t* dst = reinterpret_cast<uint32_t*>(out2);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(s_canvas);
        for (int i = tid; i < (ch * cw) >> 2; i += nthr) dst[i] = s[i];
        return;
    }
    // second output: Pillow BILINEAR (ch,cw) -> (oh2,ow2): horizontal pass into tmp2, vertical to global.
    // Fast mapping (taps per output <= 4, i.e. up-sampling or mild down-sampling, and a block size the row length
    // divides): a thread owns one output column (horizontal) / one quad of columns (vertical), its taps sit in
    // registers, and no flat index is divided.  Same sums in the same order as the generic loops below.
    constexpr int KREG = 4;
    if (ow2 == cw) {
        for (int i = tid; i < ch * ow2; i += nthr) s_tmp2[i] = s_canvas[i];
    } else if (ksw2 <= KREG && nthr % ow2 == 0) {
        const int xx = tid % ow2, row0 = tid / ow2, rs = nthr / ow2;
        const int xmin = __ldg(hb2 + 2 * xx), cnt = __ldg(hb2 + 2 * xx + 1);
        // taps at immediate offsets from one pointer per row; a tap beyond cnt has weight 0 and may read up to 3 bytes past the
        // canvas row (the next row, or the first bytes of s_tmp2 behind the canvas).  Up-sampling has 3 taps (ksize = 3).
        auto hpass = [&](auto kt) {
            constexpr int KT = decltype(kt)::value;
            int kk[KT];
#pragma unroll
            for (int j = 0; j < KT; ++j) kk[j] = j < cnt ? __ldg(hk2 + (size_t)xx * ksw2 + j) : 0;
            const uint8_t* sp = s_canvas + (size_t)row0 * cw + xmin;
            uint8_t* dp = s_tmp2 + (size_t)row0 * ow2 + xx;
            for (int row = row0; row < ch; row += rs, sp += rs * cw, dp += rs * ow2) {
                int acc = 1 << (PIL_PRECISION_BITS - 1);
#pragma unroll
                for (int j = 0; j < KT; ++j) acc += (int)sp[j] * kk[j];
                *dp = (uint8_t)pil_clip8(acc);
            }
        };
        if (ksw2 <= 3) hpass(std::integral_constant<int, 3>{});
        else hpass(std::integral_constant<int, KREG>{});
    } else {
        for (int i = tid; i < ch * ow2; i += nthr) {
            const int row = i / ow2, xx = i - row * ow2;
            const int xmin = __ldg(hb2 + 2 * xx), cnt = __ldg(hb2 + 2 * xx + 1);
            const uint8_t* sp = s_canvas + (size_t)row * cw + xmin;
            const int* kp = hk2 + (size_t)xx * ksw2;
            int acc = 1 << (PIL_PRECISION_BITS - 1);
            for (int j = 0; j < cnt; ++j) acc += (int)sp[j] * __ldg(kp + j);
            s_tmp2[i] = (uint8_t)pil_clip8(acc);
        }
    }
    __syncthreads();
    {
        const int w4 = ow2 >> 2;
        if (oh2 != ch && ksh2 <= KREG && nthr % w4 == 0) {
            const int x4 = (tid % w4) << 2, r0 = tid / w4, rs = nthr / w4;
            auto vpass = [&](auto kt) {
                constexpr int KT = decltype(kt)::value;
                for (int r = r0; r < oh2; r += rs) {
                    const int ymin = __ldg(vb2 + 2 * r), cnt = __ldg(vb2 + 2 * r + 1);  // warp-uniform when w4 >= 32
                    int a0 = 1 << (PIL_PRECISION_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
                    for (int j = 0; j < KT; ++j) {
                        const int k = j < cnt ? __ldg(vk2 + (size_t)r * ksh2 + j) : 0;
                        const int yy = j < cnt ? ymin + j : ymin;
                        const uint32_t px = *reinterpret_cast<const uint32_t*>(s_tmp2 + (size_t)yy * ow2 + x4);
                        a0 += (int)(px & 0xFF) * k;
                        a1 += (int)((px >> 8) & 0xFF) * k;
                        a2 += (int)((px >> 16) & 0xFF) * k;
                        a3 += (int)(px >> 24) * k;
                    }
                    *reinterpret_cast<uint32_t*>(out2 + (size_t)r * ow2 + x4) =
                        pil_clip8(a0) | (pil_clip8(a1) << 8) | (pil_clip8(a2) << 16) | (pil_clip8(a3) << 24);
                }
            };
            if (ksh2 <= 3) vpass(std::integral_constant<int, 3>{});
            else vpass(std::integral_constant<int, KREG>{});
        } else {
            for (int i = tid; i < oh2 * w4; i += nthr) {
                const int r = i / w4, x4 = (i - r * w4) << 2;
                uint32_t packed;
                if (oh2 == ch) {
                    packed = *reinterpret_cast<const uint32_t*>(s_tmp2 + (size_t)r * ow2 + x4);
                } else {
                    const int ymin = __ldg(vb2 + 2 * r), cnt = __ldg(vb2 + 2 * r + 1);
                    const int* kp = vk2 + (size_t)r * ksh2;
                    int a0 = 1 << (PIL_PRECISION_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
                    for (int j = 0; j < cnt; ++j) {
                        const uint32_t px = *reinterpret_cast<const uint32_t*>(s_tmp2 + (size_t)(ymin + j) * ow2 + x4);
                        const int k = __ldg(kp + j);
                        a0 += (int)(px & 0xFF) * k;
                        a1 += (int)((px >> 8) & 0xFF) * k;
                        a2 += (int)((px >> 16) & 0xFF) * k;
                        a3 += (int)(px >> 24) * k;
                    }
                    packed = pil_clip8(a0) | (pil_clip8(a1) << 8) | (pil_clip8(a2) << 16) | (pil_clip8(a3) << 24);
                }
                *reinterpret_cast<uint32_t*>(out2 + (size_t)r * ow2 + x4) = packed;
            }
        }
    }
}

}  // namespace svb

extern "C" size_t svb_k3_workspace_bytes(int ch, int cw, int oh2, int ow2) {
    return k3_layout(ch, cw, oh2, ow2).total;
}

extern "C" int svb_k3_crop_resample_rotated(const float* d_slices, const int64_t* d_offs, const int32_t* d_hw,
                                            const int32_t* d_slice_idx, const float* d_xy, const int32_t* d_delta_px,
                                            const double* d_inv_affine, const int32_t* d_pixel_kind, int N, int max_box_h,
                                            int max_box_w, int ch, int cw, uint8_t* d_crops, int oh2, int ow2, uint8_t* d_crops2,
                                            int32_t* d_geom, int flags, void* d_ws, size_t ws_bytes, void* stream_);

extern "C" int svb_k3_crop_resample(const float* d_slices, const int64_t* d_offs, const int32_t* d_hw,
                                    const int32_t* d_slice_idx, const float* d_xy, const int32_t* d_delta_px, int N,
                                    int max_box_h, int max_box_w, int ch, int cw, uint8_t* d_crops, int oh2, int ow2,
                                    uint8_t* d_crops2, int32_t* d_geom, int flags, void* d_ws, size_t ws_bytes,
                                    void* stream_) {
    return svb_k3_crop_resample_rotated(d_slices, d_offs, d_hw, d_slice_idx, d_xy, d_delta_px, nullptr, nullptr, N, max_box_h,
                                        max_box_w, ch, cw, d_crops, oh2, ow2, d_crops2, d_geom, flags, d_ws, ws_bytes, stream_);
}

extern "C" int svb_k3_crop_resample_rotated(const float* d_slices, const int64_t* d_offs, const int32_t* d_hw,
                                            const int32_t* d_slice_idx, const float* d_xy, const int32_t* d_delta_px,
                                            const double* d_inv_affine, const int32_t* d_pixel_kind, int N, int max_box_h,
                                            int max_box_w, int ch, int cw, uint8_t* d_crops, int oh2, int ow2, uint8_t* d_crops2,
                                            int32_t* d_geom, int flags, void* d_ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = check_device_sm100()) return rc;
    SVB_REQUIRE(N >= 0 && ch > 0 && cw > 0 && max_box_h > 0 && max_box_w > 0, SVB_ERR_INVALID_ARG,
                "k3: bad sizes N=%d crop=(%d,%d) max_box=(%d,%d)", N, ch, cw, max_box_h, max_box_w);
    if (N == 0) return SVB_OK;
    SVB_REQUIRE(d_slices && d_offs && d_hw && d_slice_idx && d_xy && d_delta_px && d_crops, SVB_ERR_INVALID_ARG,
                "k3: null pointer argument");
    SVB_REQUIRE(cw % 4 == 0, SVB_ERR_INVALID_ARG, "k3: crop width (%d) must be a multiple of 4", cw);
    const bool second = d_crops2 != nullptr;
    if (second) {
        SVB_REQUIRE(oh2 > 0 && ow2 > 0 && ow2 % 4 == 0, SVB_ERR_INVALID_ARG,
                    "k3: second output size (%d,%d) invalid (width must be a multiple of 4)", oh2, ow2);
        SVB_REQUIRE(d_ws != nullptr, SVB_ERR_INVALID_ARG, "k3: workspace required for the second output");
    }
    const K3Layout L = k3_layout(ch, cw, second ? oh2 : 0, second ? ow2 : 0);
    SVB_REQUIRE(!second || ws_bytes >= L.total, SVB_ERR_WORKSPACE_TOO_SMALL, "k3: workspace %zu < %zu bytes", ws_bytes,
                L.total);
    const bool identity = second && oh2 == ch && ow2 == cw;
    const size_t box_cap = align_up((size_t)max_box_h * max_box_w, 16);
    const size_t smem_bytes = box_cap + (size_t)ch * cw + ((second && !identity) ? (size_t)ch * ow2 : 0) +
                              (size_t)3 * (cw + ch) * 4;
    SVB_REQUIRE(smem_bytes <= 227 * 1024, SVB_ERR_BOX_TOO_LARGE,
                "k3: crop box %dx%d with crop %dx%d (second %dx%d) needs %zu bytes of shared memory (limit 232448)",
                max_box_h, max_box_w, ch, cw, oh2, ow2, smem_bytes);
    int *hb = nullptr, *hk = nullptr, *vb = nullptr, *vk = nullptr;
    if (second && !identity) {
        uint8_t* ws = static_cast<uint8_t*>(d_ws);
        hb = reinterpret_cast<int*>(ws + L.hb);
        hk = reinterpret_cast<int*>(ws + L.hk);
        vb = reinterpret_cast<int*>(ws + L.vb);
        vk = reinterpret_cast<int*>(ws + L.vk);
        dim3 grid(ceil_div(ow2 > oh2 ? ow2 : oh2, 128), 2);
        k3_coeff_kernel<<<grid, 128, 0, stream>>>(ch, cw, oh2, ow2, L.ksh, L.ksw, hb, hk, vb, vk);
        SVB_LAUNCHED();
    }
    SVB_CUDA_OK(cudaFuncSetAttribute(k3_crop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    k3_crop_kernel<<<N, K3_THREADS, smem_bytes, stream>>>(d_slices, d_offs, d_hw, d_slice_idx, d_xy, d_delta_px, ch, cw, d_crops,
                                                   oh2, ow2, d_crops2, d_geom, flags, (int)box_cap, L.ksh, L.ksw, hb, hk, vb, vk, d_inv_affine, d_pixel_kind);
    SVB_LAUNCHED();
    return SVB_OK;
}

// ============================================================================ K0
// Middle sagittal plane of the 0.3 mm isotropic resample, straight from the source volume: replaces
// resample_to_isotropic + extract_middle_slice + get_slice_spacing (cropping.py:37-101; SimpleITK/ITK on the CPU,
// 285 M voxels per series of which one plane is kept).  The host resolves the LPI orientation into "which image axis
// runs down the rows / across the columns / is fixed" and the two source planes around the fixed index; the device
// evaluates ITK's linear interpolation for the surviving plane only (double arithmetic, nested lerp x -> y -> z,
// neighbours clamped, default pixel 0 outside [-0.5, size - 0.5)).  Arithmetic = oracle/itk_resample.py.
namespace svb {

// per-axis sample table entry
struct K0Tap { int lo, hi, inside, pad; double frac; };

// grid (ceil(max(out_h,out_w)/128), 2, B): tables for rows (y = 0) and columns (y = 1)
__global__ void k0_table_kernel(const svb_k0_series* __restrict__ desc, int max_h, int max_w, K0Tap* __restrict__ taps) {
    const svb_k0_series d = desc[blockIdx.z];
    const int which = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_out = which == 0 ? d.out_h : d.out_w;
    if (i >= n_out) return;
    const int ax = which == 0 ? d.ax_row : d.ax_col;
    const int size = ax == 0 ? d.nx : (ax == 1 ? d.ny : d.nz);
    const int flip = which == 0 ? d.flip_row : d.flip_col;
    const double sp = which == 0 ? d.sp_row : d.sp_col, nsp = which == 0 ? d.new_sp_row : d.new_sp_col;
    const int idx = flip ? n_out - 1 - i : i;
    const double u = __ddiv_rn(__dmul_rn((double)idx, nsp), sp);
    K0Tap t;
    t.inside = (u >= -0.5 && u < (double)size - 0.5) ? 1 : 0;
    const double base = floor(u);
    t.frac = __dsub_rn(u, base);
    const int b = (int)base;
    t.lo = min(max(b, 0), size - 1);
    t.hi = min(max(b + 1, 0), size - 1);
    t.pad = 0;
    taps[((size_t)blockIdx.z * 2 + which) * (size_t)max(max_h, max_w) + i] = t;
}

// grid (ceil(max_w / 32), ceil(max_h / K0_ROWS), B), block (32, 8): a warp owns 32 output columns of a band of K0_BAND
// consecutive output rows (a thread: one column of it) -- no index division per pixel, coalesced stores along the row.
//
// The nested lerp runs over the IMAGE axes in the order x, y, z (ITK's LinearInterpolateImageFunction).  For the orientation
// every sagittal acquisition has (SPIDER .mha, Phenikaa DICOM series, BASELINE config 1: image x = columns, y = rows, z = the
// fixed Left-Right axis) the x-lerp of a (plane, source row, output column) is shared by every output row that reads that
// source row -- 0.3 mm output rows over 0.7 mm source rows: each x-lerp is needed by 2.3 output rows twice over.  A thread
// walking DOWN its column keeps the x-lerps of its last two source rows in registers and only computes the rows that are new:
// 3.9 instead of 7 double lerps and 1.7 instead of 8 loads + float->double conversions per pixel, the SAME operations in the
// SAME order for every output pixel (bit-identical to the generic path, which serves every other orientation).
constexpr int K0_BAND = 8;              // consecutive output rows per thread
constexpr int K0_ROWS = 8 * K0_BAND;    // output rows per CTA (8 warps)
__global__ void __launch_bounds__(256) k0_midplane_kernel(const float* __restrict__ vol, const svb_k0_series* __restrict__ desc,
                                                          int max_h, int max_w, const K0Tap* __restrict__ taps,
                                                          float* __restrict__ out, uint32_t* __restrict__ keys) {
    const svb_k0_series d = desc[blockIdx.z];
    const int stride = max(max_h, max_w);
    const K0Tap* rt = taps + ((size_t)blockIdx.z * 2 + 0) * stride;
    const K0Tap* ct = taps + ((size_t)blockIdx.z * 2 + 1) * stride;
    const float* v = vol + d.vol_off;
    float* o = out + d.out_off;
    const long long sx = 1, sy = d.nx, sz = (long long)d.nx * d.ny;
    const long long s_row = d.ax_row == 0 ? sx : (d.ax_row == 1 ? sy : sz);
    const long long s_col = d.ax_col == 0 ? sx : (d.ax_col == 1 ? sy : sz);
    const long long s_fix = d.ax_fix == 0 ? sx : (d.ax_fix == 1 ? sy : sz);
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * K0_ROWS + threadIdx.y * K0_BAND;
    const int r1 = min(r0 + K0_BAND, d.out_h);
    auto lerp = [](double a, double b, double f) -> double { return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), f)); };
    float mn = INFINITY, mx = -INFINITY;
    if (c < d.out_w && r0 < d.out_h) {
        const K0Tap tc = ct[c];
        if (d.ax_col == 0 && d.ax_row == 1) {
            // ---- x = columns, y = rows, z = fixed
            const float* p0 = v + (long long)d.fix_lo * sz;
            const float* p1 = v + (long long)d.fix_hi * sz;
            int ya = -1, yb = -1;            // source rows whose x-lerps are held
            double a0 = 0, a1 = 0, b0 = 0, b1 = 0;  // x-lerp at (plane 0 / 1, source row ya / yb)
            auto xrow = [&](int y, double& q0, double& q1) {
                const long long ro = (long long)y * sy;
                q0 = lerp((double)__ldg(p0 + ro + tc.lo), (double)__ldg(p0 + ro + tc.hi), tc.frac);
                q1 = lerp((double)__ldg(p1 + ro + tc.lo), (double)__ldg(p1 + ro + tc.hi), tc.frac);
            };
            for (int r = r0; r < r1; ++r) {
                const K0Tap tr = rt[r];
                float res = 0.0f;
                if (tr.inside && tc.inside && d.fix_inside) {
                    if (tr.lo != ya) {
                        if (tr.lo == yb) { ya = yb; a0 = b0; a1 = b1; }
                        else { ya = tr.lo; xrow(ya, a0, a1); }
                    }
                    if (tr.hi != yb) {
                        if (tr.hi == ya) { yb = ya; b0 = a0; b1 = a1; }
                        else { yb = tr.hi; xrow(yb, b0, b1); }
                    }
                    const double rd = lerp(lerp(a0, b0, tr.frac), lerp(a1, b1, tr.frac), d.fix_frac);
                    res = (float)(d.integer_pixels ? trunc(rd) : rd);  // integer pixel types: ITK's static_cast<PixelType>
                }
                o[(long long)r * d.out_w + c] = res;
                mn = fminf(mn, res);
                mx = fmaxf(mx, res);
            }
        } else {
            // ---- any other orientation: (lo, hi, frac) per IMAGE axis, selected without indexed local arrays
            const int ar = d.ax_row, ac = d.ax_col;
            for (int r = r0; r < r1; ++r) {
                const K0Tap tr = rt[r];
                float res = 0.0f;
                if (tr.inside && tc.inside && d.fix_inside) {
                    const long long rl = tr.lo * s_row, rh = tr.hi * s_row, cl = tc.lo * s_col, ch = tc.hi * s_col;
                    const long long fl = d.fix_lo * s_fix, fh = d.fix_hi * s_fix;
                    auto pick = [&](int axis, long long vr, long long vc, long long vf) { return ar == axis ? vr : (ac == axis ? vc : vf); };
                    auto pickf = [&](int axis) { return ar == axis ? tr.frac : (ac == axis ? tc.frac : d.fix_frac); };
                    const long long xl = pick(0, rl, cl, fl), xh = pick(0, rh, ch, fh);
                    const long long yl = pick(1, rl, cl, fl), yh = pick(1, rh, ch, fh);
                    const long long zl = pick(2, rl, cl, fl), zh = pick(2, rh, ch, fh);
                    const double fx = pickf(0), fy = pickf(1), fz = pickf(2);
                    auto g = [&](long long z, long long y, long long x) -> double { return (double)__ldg(v + z + y + x); };
                    const double qa = lerp(lerp(g(zl, yl, xl), g(zl, yl, xh), fx), lerp(g(zl, yh, xl), g(zl, yh, xh), fx), fy);
                    const double qb = lerp(lerp(g(zh, yl, xl), g(zh, yl, xh), fx), lerp(g(zh, yh, xl), g(zh, yh, xh), fx), fy);
                    const double rd = lerp(qa, qb, fz);
                    res = (float)(d.integer_pixels ? trunc(rd) : rd);
                }
                o[(long long)r * d.out_w + c] = res;
                mn = fminf(mn, res);
                mx = fmaxf(mx, res);
            }
        }
    }
    if (keys != nullptr) {
        // the plane's min / max while it is being written (same fminf / fmaxf reduction as k1_minmax_kernel, so K1 can skip
        // its own pass over the plane): warp, then CTA, then one ordered-key atomic pair per CTA
        mn = warp_min(mn);
        mx = warp_max(mx);
        __shared__ float smn[8], smx[8];
        const int wid = threadIdx.y, lane = threadIdx.x;
        if (lane == 0) { smn[wid] = mn; smx[wid] = mx; }
        __syncthreads();
        if (wid == 0) {
            mn = lane < 8 ? smn[lane] : INFINITY;
            mx = lane < 8 ? smx[lane] : -INFINITY;
            mn = warp_min(mn);
            mx = warp_max(mx);
            if (lane == 0 && mn <= mx) {  // the CTA wrote at least one pixel
                atomicMin(&keys[2 * blockIdx.z + 0], float_key(mn));
                atomicMax(&keys[2 * blockIdx.z + 1], float_key(mx));
            }
        }
    }
}

}  // namespace svb

extern "C" size_t svb_k0_workspace_bytes(int B, int max_out_h, int max_out_w) {
    if (B <= 0 || max_out_h <= 0 || max_out_w <= 0) return 0;
    return (size_t)B * 2 * (size_t)(max_out_h > max_out_w ? max_out_h : max_out_w) * sizeof(svb::K0Tap) + 256;
}

namespace svb {
static int k0_tables(const svb_k0_series* d_desc, int B, int max_out_h, int max_out_w, K0Tap* taps, cudaStream_t stream) {
    const int m = max_out_h > max_out_w ? max_out_h : max_out_w;
    k0_table_kernel<<<dim3(ceil_div(m, 128), 2, B), 128, 0, stream>>>(d_desc, max_out_h, max_out_w, taps);
    SVB_LAUNCHED();
    return SVB_OK;
}
// series [b0, b0 + nb) of the batch the tables were built for
static int k0_planes(const float* d_volumes, const svb_k0_series* d_desc, int b0, int nb, int max_out_h, int max_out_w,
                     const K0Tap* taps, float* d_out, uint32_t* keys, cudaStream_t stream) {
    const size_t stride = (size_t)(max_out_h > max_out_w ? max_out_h : max_out_w);
    k0_midplane_kernel<<<dim3(ceil_div(max_out_w, 32), ceil_div(max_out_h, K0_ROWS), nb), dim3(32, 8), 0, stream>>>(d_volumes, d_desc + b0, max_out_h, max_out_w,
                                                                      taps + (size_t)b0 * 2 * stride, d_out,
                                                                      keys ? keys + 2 * (size_t)b0 : nullptr);
    SVB_LAUNCHED();
    return SVB_OK;
}
}  // namespace svb

extern "C" int svb_k0_midplane_resample(const float* d_volumes, const svb_k0_series* d_desc, int B, int max_out_h,
                                        int max_out_w, float* d_out, void* d_ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = check_device_sm100()) return rc;
    SVB_REQUIRE(B >= 0 && max_out_h > 0 && max_out_w > 0, SVB_ERR_INVALID_ARG, "k0: bad sizes B=%d out=(%d,%d)", B, max_out_h, max_out_w);
    if (B == 0) return SVB_OK;
    SVB_REQUIRE(d_volumes && d_desc && d_out && d_ws, SVB_ERR_INVALID_ARG, "k0: null pointer argument");
    SVB_REQUIRE(ws_bytes >= svb_k0_workspace_bytes(B, max_out_h, max_out_w), SVB_ERR_WORKSPACE_TOO_SMALL, "k0: workspace too small");
    SVB_REQUIRE(B <= 65535, SVB_ERR_INVALID_ARG, "k0: at most 65535 series per call");
    K0Tap* taps = reinterpret_cast<K0Tap*>((reinterpret_cast<uintptr_t>(d_ws) + 15) & ~uintptr_t(15));
    if (int rc = k0_tables(d_desc, B, max_out_h, max_out_w, taps, stream)) return rc;
    return k0_planes(d_volumes, d_desc, 0, B, max_out_h, max_out_w, taps, d_out, nullptr, stream);
}

// ============================================================================ K0 + K1 in one call
extern "C" size_t svb_k01_workspace_bytes(int B, int max_out_h, int max_out_w, int out_h, int out_w) {
    if (B <= 0 || max_out_h <= 0 || max_out_w <= 0 || out_h <= 0 || out_w <= 0) return 0;
    return align_up(k1_layout(B, max_out_h, max_out_w, out_h, out_w).total, 256) + svb_k0_workspace_bytes(B, max_out_h, max_out_w);
}

extern "C" int svb_k01_midplane_normalize_resize(const float* d_volumes, const svb_k0_series* d_desc, int B, int max_out_h,
                                                 int max_out_w, float* d_slices, const int64_t* d_offs, const int32_t* d_hw,
                                                 int out_h, int out_w, uint8_t* d_out_u8, float* d_minmax, void* d_ws,
                                                 size_t ws_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = check_device_sm100()) return rc;
    SVB_REQUIRE(B >= 0 && max_out_h > 0 && max_out_w > 0 && out_h > 0 && out_w > 0, SVB_ERR_INVALID_ARG,
                "k01: bad sizes B=%d iso=(%d,%d) out=(%d,%d)", B, max_out_h, max_out_w, out_h, out_w);
    if (B == 0) return SVB_OK;
    SVB_REQUIRE(d_volumes && d_desc && d_slices && d_offs && d_hw && d_out_u8 && d_ws, SVB_ERR_INVALID_ARG, "k01: null pointer argument");
    SVB_REQUIRE(B <= 65535, SVB_ERR_INVALID_ARG, "k01: at most 65535 series per call");
    SVB_REQUIRE(ws_bytes >= svb_k01_workspace_bytes(B, max_out_h, max_out_w, out_h, out_w), SVB_ERR_WORKSPACE_TOO_SMALL,
                "k01: workspace %zu < %zu bytes", ws_bytes, svb_k01_workspace_bytes(B, max_out_h, max_out_w, out_h, out_w));
    K1Plan P;
    const size_t k1_bytes = align_up(k1_layout(B, max_out_h, max_out_w, out_h, out_w).total, 256);
    if (int rc = k1_prepare(d_hw, B, max_out_h, max_out_w, out_h, out_w, d_ws, k1_bytes, true, stream, &P)) return rc;
    K0Tap* taps = reinterpret_cast<K0Tap*>((reinterpret_cast<uintptr_t>(static_cast<uint8_t*>(d_ws) + k1_bytes) + 15) & ~uintptr_t(15));
    if (int rc = k0_tables(d_desc, B, max_out_h, max_out_w, taps, stream)) return rc;
    // Per group: K0 writes the group's isotropic planes AND their min / max; the resize launch that follows normalises with
    // those extrema and reads the planes back while they are still in L2.  No min/max pass, no second trip through HBM.
    for (int b0 = 0; b0 < B; b0 += P.group) {
        const int nb = (B - b0) < P.group ? (B - b0) : P.group;
        if (int rc = k0_planes(d_volumes, d_desc, b0, nb, max_out_h, max_out_w, taps, d_slices, P.keys, stream)) return rc;
        if (int rc = k1_resize_group(P, d_slices, d_offs, d_hw, b0, nb, out_h, out_w, d_out_u8, d_minmax, stream)) return rc;
    }
    return SVB_OK;
}
