// Single-degradation dataset generators with the reference's exact (quirky) arithmetic, and the image-quality
// reduction: the callers either side of the hot path.  All are streaming kernels bound by HBM:
// algorithmic bytes = elems x (bytes read + bytes written) per image.
//
//   b2r_lut_u8              out = lut[n][in]            the reference's float64 point operations on u8 images, e.g. fog on
//                                                       image/255.0 (04_gen_fog.py:17-30, 13_pipeline_stress_test.py:50-56):
//                                                       256 possible results per image, evaluated by the host with the
//                                                       reference's own arithmetic
//   b2r_minmax_u8           per-image extrema over all channels (cv2.minMaxIdx inside cv2.normalize, 03_gen_blur.py:29)
//   b2r_normalize_minmax_u8 cv2.normalize(x, x, 0, 255, NORM_MINMAX): scale / shift in double, cast to float,
//                           saturate(rint(fma(src, scale, shift)))  (OpenCV 4.13 cvtScale8u; pinned against cv2 in
//                           tests/test_generators_host.py)
//   b2r_noise02             02_gen_noise.py:12-27: float64 image/255 + noise, lower clip -1 when ANY value of the image is
//                           negative, then np.uint8(out * 255), which wraps negative values modulo 256
//   b2r_sse_u8              per-image sum of squared differences (PSNR = 10 log10(255^2 N / SSE), 08_run_inference.py:118-129)
#include "b2r_internal.h"
#include "philox.cuh"
#include "stream_u8.cuh"

namespace b2r {

__global__ void __launch_bounds__(kGenThreads) lut_u8_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ lut,
                                                              uint8_t* __restrict__ out, long elems) {
    __shared__ uint8_t s_lut[256];
    const int n = blockIdx.y;
    if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(s_lut)[threadIdx.x] = reinterpret_cast<const uint32_t*>(lut + n * 256)[threadIdx.x];
    __syncthreads();
    apply_table(s_lut, in + (long)n * elems, out + (long)n * elems, elems);
}

__global__ void minmax_init_kernel(int32_t* minmax, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        minmax[2 * i] = 255;
        minmax[2 * i + 1] = 0;
    }
}

__global__ void __launch_bounds__(kGenThreads) minmax_u8_kernel(const uint8_t* __restrict__ in, int32_t* __restrict__ minmax,
                                                                 long elems) {
    const int n = blockIdx.y;
    const uint8_t* src = in + (long)n * elems;
    const bool vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;   // four byte lanes each
    const long step = (long)gridDim.x * kGenThreads * kGenBytesPerThread;
    long i = ((long)blockIdx.x * kGenThreads + threadIdx.x) * kGenBytesPerThread;
#if B2R_STREAM_UNROLL > 1
    constexpr int kU = 2 * B2R_STREAM_UNROLL;   // a single read stream: twice the loads in flight of the copy kernels
    if (vec) {
        for (; i + (kU - 1) * step + kGenBytesPerThread <= elems; i += kU * step) {
            uint4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) v[u] = __ldcs(reinterpret_cast<const uint4*>(src + i + u * step));
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                lo = __vminu4(__vminu4(lo, v[u].x), __vminu4(v[u].y, __vminu4(v[u].z, v[u].w)));
                hi = __vmaxu4(__vmaxu4(hi, v[u].x), __vmaxu4(v[u].y, __vmaxu4(v[u].z, v[u].w)));
            }
        }
    }
#endif
    for (; i < elems; i += step) {
        if (vec && i + kGenBytesPerThread <= elems) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + i);
            lo = __vminu4(__vminu4(lo, v.x), __vminu4(v.y, __vminu4(v.z, v.w)));
            hi = __vmaxu4(__vmaxu4(hi, v.x), __vmaxu4(v.y, __vmaxu4(v.z, v.w)));
        } else {
            for (long j = i; j < elems && j < i + kGenBytesPerThread; ++j) {
                const uint32_t b = src[j] * 0x01010101u;
                lo = __vminu4(lo, b);
                hi = __vmaxu4(hi, b);
            }
        }
    }
    int mn = min(min(lo & 0xFF, (lo >> 8) & 0xFF), min((lo >> 16) & 0xFF, lo >> 24));
    int mx = max(max(hi & 0xFF, (hi >> 8) & 0xFF), max((hi >> 16) & 0xFF, hi >> 24));
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[2 * n], mn);
        atomicMax(&minmax[2 * n + 1], mx);
    }
}

__global__ void __launch_bounds__(kGenThreads) normalize_minmax_u8_kernel(const uint8_t* __restrict__ in,
                                                                           const int32_t* __restrict__ minmax,
                                                                           uint8_t* __restrict__ out, long elems) {
    __shared__ uint8_t s_lut[256];
    const int n = blockIdx.y;
    {
        // cv::normalize(NORM_MINMAX, alpha 0, beta 255): scale, shift in double; convertTo 8U -> 8U evaluates
        // saturate_cast<uchar>(cvRound(fmaf(src, (float)scale, (float)shift)))
        const double smin = minmax[2 * n], smax = minmax[2 * n + 1];
        const double scale = 255.0 * ((smax - smin) > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
        const double shift = 0.0 - smin * scale;
        const float a = float(scale), b = float(shift);
        const float r = rintf(fmaf(float(threadIdx.x), a, b));   // rintf: ties to even, as cvRound
        s_lut[threadIdx.x] = uint8_t(fminf(fmaxf(r, 0.f), 255.f));
    }
    __syncthreads();
    apply_table(s_lut, in + (long)n * elems, out + (long)n * elems, elems);
}

// 02_gen_noise.py: out = image / 255 (float64) + noise;  pass 0 records whether any value of the image is negative,
// pass 1 clips to [low, 1] and converts with np.uint8(out * 255): truncation toward zero, then wrap modulo 256.
template <int PASS>
__global__ void __launch_bounds__(kGenThreads) noise02_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                               long elems, const float* __restrict__ sigma,
                                                               const double* __restrict__ noise, uint64_t seed,
                                                               uint64_t image_index0, int32_t* __restrict__ neg_flags) {
    __shared__ double s_unit[256];   // u / 255.0 with a true float64 division, once per block
    s_unit[threadIdx.x] = double(threadIdx.x) / 255.0;
    __syncthreads();
    const int n = blockIdx.y;
    const uint8_t* src = in + (long)n * elems;
    const double* nz = noise ? noise + (long)n * elems : nullptr;
    const double sg = double(sigma[n]);
    const double low = (PASS == 1 && neg_flags[n]) ? -1.0 : 0.0;
    bool any_neg = false;
    const long pixels = elems / 3;
    for (long pix = (long)blockIdx.x * kGenThreads + threadIdx.x; pix < pixels; pix += (long)gridDim.x * kGenThreads) {
        float z[3] = {0.f, 0.f, 0.f};
        if (!nz) pixel_normals(seed, image_index0 + uint64_t(n), uint32_t(pix), z);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long i = pix * 3 + c;
            const double v = s_unit[src[i]] + (nz ? nz[i] : sg * double(z[c]));
            if (PASS == 0) {
                any_neg |= v < 0.0;
            } else {
                const double y = fmin(fmax(v, low), 1.0) * 255.0;
                out[(long)n * elems + i] = uint8_t(int(y) & 0xFF);   // C cast toward zero, then two's-complement wrap
            }
        }
    }
    if (PASS == 0 && __any_sync(0xffffffffu, any_neg) && (threadIdx.x & 31) == 0) atomicOr(&neg_flags[n], 1);
}

// sum of the sixteen squared byte differences of two 16-byte words
__device__ __forceinline__ uint32_t sq_diff16(const uint4 va, const uint4 vb) {
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t d = __vabsdiffu4(wa[k], wb[k]);
        s = __dp4a(d, d, s);
    }
    return s;
}

__global__ void __launch_bounds__(kGenThreads) sse_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                              unsigned long long* __restrict__ sse, long elems) {
    const int n = blockIdx.y;
    const uint8_t* pa = a + (long)n * elems;
    const uint8_t* pb = b + (long)n * elems;
    const bool vec = ((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 15) == 0;
    unsigned long long acc = 0;
    const long step = (long)gridDim.x * kGenThreads * kGenBytesPerThread;
    long i = ((long)blockIdx.x * kGenThreads + threadIdx.x) * kGenBytesPerThread;
#if B2R_STREAM_UNROLL > 1
    if (vec) {
        for (; i + (B2R_STREAM_UNROLL - 1) * step + kGenBytesPerThread <= elems; i += B2R_STREAM_UNROLL * step) {
            uint4 va[B2R_STREAM_UNROLL], vb[B2R_STREAM_UNROLL];
#pragma unroll
            for (int u = 0; u < B2R_STREAM_UNROLL; ++u) {
                va[u] = __ldcs(reinterpret_cast<const uint4*>(pa + i + u * step));
                vb[u] = __ldcs(reinterpret_cast<const uint4*>(pb + i + u * step));
            }
            uint32_t s = 0;   // <= 16 * 4 * 255^2 per trip: no overflow
#pragma unroll
            for (int u = 0; u < B2R_STREAM_UNROLL; ++u) s += sq_diff16(va[u], vb[u]);
            acc += s;
        }
    }
#endif
    for (; i < elems; i += step) {
        if (vec && i + kGenBytesPerThread <= elems) {
            acc += sq_diff16(*reinterpret_cast<const uint4*>(pa + i), *reinterpret_cast<const uint4*>(pb + i));
        } else {
            for (long j = i; j < elems && j < i + kGenBytesPerThread; ++j) {
                const int d = int(pa[j]) - int(pb[j]);
                acc += (unsigned long long)(d * d);
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sse[n], acc);
}

// ------------------------------------------------------------------------------------------------------------
// VGG feature taps: mean over one axis of a bf16 tensor viewed as [outer][reduce][inner].
//   inner == 1: channel mean of an NHWC feature map (11_visualize_hidden_states.py:50 torch.mean(features, dim=1));
//               the reduce axis is contiguous: 8 lanes x 16 bytes per row step, shuffle reduction, fp32 accumulation
//   inner  > 1: global average pooling of [N][H*W][C] (12_generate_umap_pt.py:52 torch.mean(feature, dim=[2, 3]));
//               one thread per 8 inner columns, coalesced across the inner axis
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void acc8_bf16(const uint4 v, float (&a)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[2 * k] += __uint_as_float(w[k] << 16);
        a[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
    }
}

__global__ void __launch_bounds__(256) mean_last_axis_kernel(const uint4* __restrict__ in, float* __restrict__ out,
                                                             long outer, int reduce8) {
    // 8 consecutive lanes own one row
    const long row = (blockIdx.x * 256L + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (row < outer)
        for (int i = sub; i < reduce8; i += 8) acc8_bf16(__ldg(&in[row * reduce8 + i]), a);
    float s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (row < outer && sub == 0) out[row] = s / float(reduce8 * 8);
}

__global__ void __launch_bounds__(256) mean_middle_axis_kernel(const uint4* __restrict__ in, float* __restrict__ out,
                                                               long outer, int reduce, int inner8) {
    const long gid = blockIdx.x * 256L + threadIdx.x;
    if (gid >= outer * inner8) return;
    const long o = gid / inner8;
    const int c = int(gid - o * inner8);
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < reduce; ++r) acc8_bf16(__ldg(&in[(o * reduce + r) * inner8 + c]), a);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[(o * inner8 + c) * 8 + k] = a[k] / float(reduce);
}

// ------------------------------------------------------------------------------------------------------------
// Resize of a ragged batch with Pillow's BILINEAR resampling (what transforms.Resize((224, 224)) does to the PIL images
// of 17_run_unified_inference.py:66,79-82 and 18:28-30): separable triangle filter whose support grows with the
// down-scaling factor, coefficients normalised in double and rounded to 22-bit fixed point, horizontal pass rounded to
// u8, then vertical pass.  The host builds the per-size coefficient tables exactly as Pillow's precompute_coeffs /
// normalize_coeffs_8bpc do (imageio.resample_table); this kernel does the integer accumulation:
//     out = clip8((2^21 + sum_k pixel[min + k] * kk[k]) >> 22)
// One CTA per (image, tile of output rows): the input rows the tile needs go through the horizontal pass into shared
// memory, the vertical pass reads them from there.  Table layout per output index: {min, count, kk[0..K)}.
// ------------------------------------------------------------------------------------------------------------
constexpr int kResizeThreads = 256;

__device__ __forceinline__ int clip8_fixed(int ss) { return min(max(ss >> 22, 0), 255); }

__global__ void __launch_bounds__(kResizeThreads) resize_bilinear_u8_kernel(
    const uint8_t* __restrict__ src, const long long* __restrict__ offsets, const int32_t* __restrict__ hw,
    const int32_t* __restrict__ xtab_index, const int32_t* __restrict__ ytab_index, const int32_t* __restrict__ tabs,
    int K, int S, uint8_t* __restrict__ out, int out_h, int out_w, int tile_rows, int max_rows) {
    extern __shared__ uint8_t s_tmp[];   // [rows][out_w][3]
    const int n = blockIdx.y;
    const int yy0 = blockIdx.x * tile_rows;
    const int yy1 = min(yy0 + tile_rows, out_h);
    const int in_h = hw[2 * n], in_w = hw[2 * n + 1];
    const uint8_t* img = src + offsets[n];
    const int E = 2 + K;
    const int32_t* xt = tabs + (long)xtab_index[n] * S * E;
    const int32_t* yt = tabs + (long)ytab_index[n] * S * E;
    // input rows needed by this tile of output rows
    const int r_lo = yt[yy0 * E];
    int r_hi = r_lo;
    for (int yy = yy0; yy < yy1; ++yy) r_hi = max(r_hi, yt[yy * E] + yt[yy * E + 1]);
    const int rows = min(r_hi, in_h) - r_lo;
    if (rows > max_rows) return;   // cannot happen: the host sized max_rows from the same tables
    // horizontal pass
    for (int i = threadIdx.x; i < rows * out_w; i += kResizeThreads) {
        const int r = i / out_w, xx = i - r * out_w;
        const int32_t* t = xt + xx * E;
        const int xmin = t[0], cnt = t[1];
        const uint8_t* p = img + ((long)(r_lo + r) * in_w + xmin) * 3;
        int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
        for (int k = 0; k < cnt; ++k) {
            const int c = t[2 + k];
            s0 += int(p[3 * k]) * c;
            s1 += int(p[3 * k + 1]) * c;
            s2 += int(p[3 * k + 2]) * c;
        }
        uint8_t* d = s_tmp + (long)i * 3;
        d[0] = uint8_t(clip8_fixed(s0));
        d[1] = uint8_t(clip8_fixed(s1));
        d[2] = uint8_t(clip8_fixed(s2));
    }
    __syncthreads();
    // vertical pass
    const int row_bytes = out_w * 3;
    for (int i = threadIdx.x; i < (yy1 - yy0) * row_bytes; i += kResizeThreads) {
        const int yl = i / row_bytes, b = i - yl * row_bytes;
        const int32_t* t = yt + (yy0 + yl) * E;
        const int ymin = t[0], cnt = t[1];
        int ss = 1 << 21;
        for (int k = 0; k < cnt; ++k) ss += int(s_tmp[(long)(ymin - r_lo + k) * row_bytes + b]) * t[2 + k];
        out[((long)n * out_h + yy0 + yl) * row_bytes + b] = uint8_t(clip8_fixed(ss));
    }
}


// cv2.resize(img, (out_w, out_h)) with the default INTER_LINEAR on u8 images, bit for bit (08_run_inference.py:119 resizes
// the clean image this way before PSNR / SSIM).  OpenCV's fixed-point scheme: 11-bit coefficients (cvRound(c * 2048),
// from the host: imageio.cv_linear_table), horizontal pass kept as int, vertical pass
//     dst = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
// tabs int32 [T][S][3] = {first source index, c0, c1}; rows are clamped to the image, columns are clamped by the host
// (OpenCV zeroes the fraction there).  One thread per output pixel: 12 source bytes in, 3 bytes out.
__global__ void __launch_bounds__(kResizeThreads) resize_cv_linear_u8_kernel(
    const uint8_t* __restrict__ src, const long long* __restrict__ offsets, const int32_t* __restrict__ hw,
    const int32_t* __restrict__ xtab_index, const int32_t* __restrict__ ytab_index, const int32_t* __restrict__ tabs, int S,
    uint8_t* __restrict__ out, int out_h, int out_w) {
    const int n = blockIdx.y;
    const int in_h = hw[2 * n], in_w = hw[2 * n + 1];
    const uint8_t* img = src + offsets[n];
    const int32_t* xt = tabs + (long)xtab_index[n] * S * 3;
    const int32_t* yt = tabs + (long)ytab_index[n] * S * 3;
    for (int i = blockIdx.x * kResizeThreads + threadIdx.x; i < out_h * out_w; i += gridDim.x * kResizeThreads) {
        const int yy = i / out_w, xx = i - yy * out_w;
        const int sx = xt[3 * xx], a0 = xt[3 * xx + 1], a1 = xt[3 * xx + 2];
        const int sy = yt[3 * yy], b0 = yt[3 * yy + 1], b1 = yt[3 * yy + 2];
        const int x1 = min(sx + 1, in_w - 1);
        const uint8_t* r0 = img + (long)min(max(sy, 0), in_h - 1) * in_w * 3;
        const uint8_t* r1 = img + (long)min(max(sy + 1, 0), in_h - 1) * in_w * 3;
        uint8_t* d = out + ((long)n * out_h * out_w + i) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int s0 = int(r0[3 * sx + c]) * a0 + int(r0[3 * x1 + c]) * a1;
            const int s1 = int(r1[3 * sx + c]) * a0 + int(r1[3 * x1 + c]) * a1;
            const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
            d[c] = uint8_t(min(max(v, 0), 255));
        }
    }
}


}  // namespace b2r

extern "C" {

int b2r_lut_u8(const uint8_t* in, const uint8_t* lut, uint8_t* out, int N, int64_t elems_per_image, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && lut && out, "null pointer");
    B2R_REQUIRE(N > 0 && N <= 65535 && elems_per_image > 0, "bad shape N=%d elems=%lld", N, (long long)elems_per_image);
    B2R_REQUIRE((reinterpret_cast<uintptr_t>(lut) & 3) == 0, "lut must be 4-byte aligned");
    dim3 grid;
    int rc = gen_grid(elems_per_image, N, &grid, kGenBytesPerThread);
    if (rc) return rc;
    lut_u8_kernel<<<grid, kGenThreads, 0, stream>>>(in, lut, out, elems_per_image);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_minmax_u8(const uint8_t* in, int32_t* minmax, int N, int64_t elems_per_image, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && minmax, "null pointer");
    B2R_REQUIRE(N > 0 && N <= 65535 && elems_per_image > 0, "bad shape N=%d elems=%lld", N, (long long)elems_per_image);
    minmax_init_kernel<<<(N + 255) / 256, 256, 0, stream>>>(minmax, N);
    B2R_CHECK_LAUNCH();
    dim3 grid;
    int rc = gen_grid(elems_per_image, N, &grid, kGenBytesPerThread);
    if (rc) return rc;
    minmax_u8_kernel<<<grid, kGenThreads, 0, stream>>>(in, minmax, elems_per_image);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_normalize_minmax_u8(const uint8_t* in, const int32_t* minmax, uint8_t* out, int N, int64_t elems_per_image,
                            void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && minmax && out, "null pointer");
    B2R_REQUIRE(N > 0 && N <= 65535 && elems_per_image > 0, "bad shape N=%d elems=%lld", N, (long long)elems_per_image);
    dim3 grid;
    int rc = gen_grid(elems_per_image, N, &grid, kGenBytesPerThread);
    if (rc) return rc;
    normalize_minmax_u8_kernel<<<grid, kGenThreads, 0, stream>>>(in, minmax, out, elems_per_image);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_noise02(const uint8_t* in, uint8_t* out, int N, int64_t elems_per_image, const float* sigma, const double* noise,
                uint64_t seed, uint64_t image_index0, int32_t* neg_flags, int clip_rule, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && out && sigma && neg_flags, "null pointer");
    B2R_REQUIRE(N > 0 && N <= 65535 && elems_per_image > 0 && elems_per_image % 3 == 0 && elems_per_image / 3 < (1LL << 32),
                "bad shape N=%d elems=%lld (3 interleaved channels expected)", N, (long long)elems_per_image);
    B2R_REQUIRE(clip_rule == B2R_NOISE_CLIP_SCRIPT02 || clip_rule == B2R_NOISE_CLIP_UNIT, "clip_rule=%d", clip_rule);
    B2R_CUDA(cudaMemsetAsync(neg_flags, 0, sizeof(int32_t) * N, stream));
    dim3 grid;
    int rc = gen_grid(elems_per_image, N, &grid, 3);
    if (rc) return rc;
    if (clip_rule == B2R_NOISE_CLIP_SCRIPT02) {   // the lower clip depends on the whole image: decide it first
        noise02_kernel<0><<<grid, kGenThreads, 0, stream>>>(in, out, elems_per_image, sigma, noise, seed, image_index0, neg_flags);
        B2R_CHECK_LAUNCH();
    }
    noise02_kernel<1><<<grid, kGenThreads, 0, stream>>>(in, out, elems_per_image, sigma, noise, seed, image_index0, neg_flags);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_resize_bilinear_u8(const uint8_t* src, const int64_t* offsets, const int32_t* hw, const int32_t* xtab_index,
                           const int32_t* ytab_index, const int32_t* tabs, int K, int S, uint8_t* out, int N, int out_h,
                           int out_w, int tile_rows, int max_rows, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(src && offsets && hw && xtab_index && ytab_index && tabs && out, "null pointer");
    B2R_REQUIRE(N > 0 && N <= 65535 && out_h > 0 && out_w > 0 && K > 0 && S >= out_h && S >= out_w, "bad shape");
    B2R_REQUIRE(tile_rows > 0 && max_rows > 0, "tile_rows=%d max_rows=%d", tile_rows, max_rows);
    const size_t smem = (size_t)max_rows * out_w * 3;
    B2R_REQUIRE(smem <= 200 * 1024, "the tile needs %zu bytes of shared memory: use fewer tile_rows", smem);
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        B2R_CUDA(cudaFuncSetAttribute(resize_bilinear_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev < 64) attr_set[dev] = true;
    }
    dim3 grid((unsigned)((out_h + tile_rows - 1) / tile_rows), (unsigned)N, 1);
    resize_bilinear_u8_kernel<<<grid, kResizeThreads, smem, stream>>>(src, reinterpret_cast<const long long*>(offsets), hw,
                                                                      xtab_index, ytab_index, tabs, K, S, out, out_h, out_w,
                                                                      tile_rows, max_rows);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_resize_cv_linear_u8(const uint8_t* src, const int64_t* offsets, const int32_t* hw, const int32_t* xtab_index,
                            const int32_t* ytab_index, const int32_t* tabs, int S, uint8_t* out, int N, int out_h,
                            int out_w, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(src && offsets && hw && xtab_index && ytab_index && tabs && out, "null pointer");
    B2R_REQUIRE(N > 0 && N <= 65535 && out_h > 0 && out_w > 0 && S >= out_h && S >= out_w, "bad shape N=%d out=%dx%d S=%d",
                N, out_h, out_w, S);
    int sms = 0;
    const int rc = device_sm_count(&sms);
    if (rc) return rc;
    long blocks = ((long)out_h * out_w + kResizeThreads - 1) / kResizeThreads;
    const long cap = (8L * sms + N - 1) / N > 1 ? (8L * sms + N - 1) / N : 1;
    if (blocks > cap) blocks = cap;
    resize_cv_linear_u8_kernel<<<dim3((unsigned)blocks, (unsigned)N, 1), kResizeThreads, 0, stream>>>(
        src, reinterpret_cast<const long long*>(offsets), hw, xtab_index, ytab_index, tabs, S, out, out_h, out_w);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_mean_bf16(const void* in, float* out, int64_t outer, int reduce, int inner, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && out, "null pointer");
    B2R_REQUIRE(outer > 0 && reduce > 0 && inner > 0, "bad shape outer=%lld reduce=%d inner=%d", (long long)outer, reduce, inner);
    B2R_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, "input must be 16-byte aligned");
    if (inner == 1) {
        B2R_REQUIRE(reduce % 8 == 0, "reduce=%d must be a multiple of 8 when it is the contiguous axis", reduce);
        const long blocks = (outer * 8 + 255) / 256;
        B2R_REQUIRE(blocks < (1L << 31), "too many rows");
        mean_last_axis_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint4*>(in), out, outer, reduce / 8);
    } else {
        B2R_REQUIRE(inner % 8 == 0, "inner=%d must be a multiple of 8", inner);
        const long blocks = (outer * (inner / 8) + 255) / 256;
        B2R_REQUIRE(blocks < (1L << 31), "too many columns");
        mean_middle_axis_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint4*>(in), out, outer, reduce, inner / 8);
    }
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_sse_u8(const uint8_t* a, const uint8_t* b, uint64_t* sse, int N, int64_t elems_per_image, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(a && b && sse, "null pointer");
    B2R_REQUIRE(N > 0 && N <= 65535 && elems_per_image > 0, "bad shape N=%d elems=%lld", N, (long long)elems_per_image);
    B2R_CUDA(cudaMemsetAsync(sse, 0, sizeof(uint64_t) * N, stream));
    dim3 grid;
    int rc = gen_grid(elems_per_image, N, &grid, kGenBytesPerThread);
    if (rc) return rc;
    sse_u8_kernel<<<grid, kGenThreads, 0, stream>>>(a, b, reinterpret_cast<unsigned long long*>(sse), elems_per_image);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

}  // extern "C"
