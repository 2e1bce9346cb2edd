// Small CUDA-core kernels around the tensor-core convs: the 64 -> 3 output head with the clamp / truncating u8 quantiser, stand-alone pooling,
// the 43-way classifier row and the arg-max + correct-count reduction.  All are HBM- or latency-bound.
#include <cstring>

#include "b2r_internal.h"
#include <cuda_bf16.h>

namespace b2r {

__device__ __forceinline__ float act_fn(float x, int act, float slope) {
    if (act == B2R_ACT_RELU) return fmaxf(x, 0.f);
    if (act == B2R_ACT_PRELU) return x >= 0.f ? x : x * slope;
    return x;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}

// ------------------------------------------------------------------------------------------------------------
// final 1x1 conv 64 -> 3: 8 threads per pixel (one 16-byte load each), shuffle reduce, f32 NCHW and/or u8 NHWC out
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) final_conv1x1_kernel(const uint4* __restrict__ in, const float* __restrict__ w,
                                                            const float* __restrict__ b, float* __restrict__ out_f32,
                                                            uint8_t* __restrict__ out_u8, long npix, int HW) {
    const long gid = blockIdx.x * 256L + threadIdx.x;
    const long pix = gid >> 3;
    const int cg = int(gid & 7);
    float wr[3][8];
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[o][j] = __ldg(&w[o * 64 + cg * 8 + j]);
    float s[3] = {0.f, 0.f, 0.f};
    if (pix < npix) {
        float f[8];
        unpack8(__ldg(&in[pix * 8 + cg]), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[0] = fmaf(f[j], wr[0][j], s[0]);
            s[1] = fmaf(f[j], wr[1][j], s[1]);
            s[2] = fmaf(f[j], wr[2][j], s[2]);
        }
    }
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        s[o] += __shfl_xor_sync(0xffffffffu, s[o], 1);
        s[o] += __shfl_xor_sync(0xffffffffu, s[o], 2);
        s[o] += __shfl_xor_sync(0xffffffffu, s[o], 4);
    }
    if (pix < npix && cg < 3) {
        const float v = (cg == 0 ? s[0] : (cg == 1 ? s[1] : s[2])) + __ldg(&b[cg]);
        if (out_f32) {
            const long n = pix / HW, hw = pix % HW;
            out_f32[(n * 3 + cg) * HW + hw] = v;
        }
        if (out_u8) {
            // torch.clamp(x, 0, 1); (x * 255).astype(np.uint8): truncation (17_run_unified_inference.py:86-92)
            const float c = fminf(fmaxf(v, 0.f), 1.f);
            out_u8[pix * 3 + cg] = static_cast<uint8_t>(c * 255.0f);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// F.interpolate(x, size=(out_h, out_w)) with the default mode 'nearest' on NHWC bf16, 8 channels (16 B) per thread: the
// re-alignment of ResUNet.forward (14_train_unified_advanced.py:169-183).  PyTorch's rule: src = min(int(floorf(dst *
// scale)), in - 1) with scale = float(in) / float(out).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_nearest_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H,
                                                             int W, int Ho, int Wo, int C8, float sh, float sw) {
    const long total = (long)N * Ho * Wo * C8;
    const long gid = blockIdx.x * 256L + threadIdx.x;
    if (gid >= total) return;
    const int c = int(gid % C8);
    long t = gid / C8;
    const int wo = int(t % Wo);
    t /= Wo;
    const int ho = int(t % Ho);
    const int n = int(t / Ho);
    const int hi = min(int(floorf(float(ho) * sh)), H - 1);
    const int wi = min(int(floorf(float(wo) * sw)), W - 1);
    out[gid] = __ldg(&in[(((long)n * H + hi) * W + wi) * C8 + c]);
}

// ------------------------------------------------------------------------------------------------------------
// pooling on NHWC bf16, 8 channels (16 B) per thread
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N,
                                                         int H, int W, int C8) {
    const int Ho = H >> 1, Wo = W >> 1;
    const long total = (long)N * Ho * Wo * C8;
    const long gid = blockIdx.x * 256L + threadIdx.x;
    if (gid >= total) return;
    const int c = int(gid % C8);
    long t = gid / C8;
    const int wo = int(t % Wo);
    t /= Wo;
    const int ho = int(t % Ho);
    const int n = int(t / Ho);
    const long base = (((long)n * H + 2 * ho) * W + 2 * wo) * C8 + c;
    const uint4 q[4] = {__ldg(&in[base]), __ldg(&in[base + C8]), __ldg(&in[base + (long)W * C8]),
                        __ldg(&in[base + (long)W * C8 + C8])};
    uint4 r = q[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        const uint32_t a[4] = {r.x, r.y, r.z, r.w};
        const uint32_t bb[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
        uint32_t m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&a[i]);
            __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&bb[i]);
            __nv_bfloat162 z = __hmax2(x, y);
            m[i] = *reinterpret_cast<uint32_t*>(&z);
        }
        r = make_uint4(m[0], m[1], m[2], m[3]);
    }
    out[gid] = r;
}

__global__ void __launch_bounds__(256) adaptive_avgpool7_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                                int N, int H, int W, int C8) {
    const long total = (long)N * 49 * C8;
    const long gid = blockIdx.x * 256L + threadIdx.x;
    if (gid >= total) return;
    const int c = int(gid % C8);
    long t = gid / C8;
    const int j = int(t % 7);
    t /= 7;
    const int i = int(t % 7);
    const int n = int(t / 7);
    // torch adaptive pooling windows: [floor(i*H/7), ceil((i+1)*H/7))
    const int hs = (i * H) / 7, he = ((i + 1) * H + 6) / 7;
    const int ws = (j * W) / 7, we = ((j + 1) * W + 6) / 7;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
            float f[8];
            unpack8(__ldg(&in[(((long)n * H + h) * W + w) * C8 + c]), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
    const float cnt = float((he - hs) * (we - ws));
    uint4 o;
    o.x = pack2(acc[0] / cnt, acc[1] / cnt);
    o.y = pack2(acc[2] / cnt, acc[3] / cnt);
    o.z = pack2(acc[4] / cnt, acc[5] / cnt);
    o.w = pack2(acc[6] / cnt, acc[7] / cnt);
    out[gid] = o;
}

// ------------------------------------------------------------------------------------------------------------
// small-N linear with f32 output: one block per batch row, row staged in smem, one warp per output neuron
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_f32out_kernel(const uint4* __restrict__ in, const uint4* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            int K8, int O) {
    extern __shared__ uint4 s_row[];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < K8; i += 256) s_row[i] = __ldg(&in[(long)b * K8 + i]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < O; o += 8) {
        float s = 0.f;
        for (int i = lane; i < K8; i += 32) {
            float a[8], ww[8];
            unpack8(s_row[i], a);
            unpack8(__ldg(&w[(long)o * K8 + i]), ww);
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(a[k], ww[k], s);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) out[(long)b * O + o] = s + __ldg(&bias[o]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// arg-max (lowest index wins ties), softmax confidence, correct count: one warp per row, one atomic per block
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) argmax_count_kernel(const float* __restrict__ logits,
                                                           const int64_t* __restrict__ labels,
                                                           int64_t* __restrict__ pred, float* __restrict__ conf,
                                                           unsigned long long* __restrict__ counts, int N, int C) {
    __shared__ int s_correct;
    if (threadIdx.x == 0) s_correct = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row < N) {
        const float* r = logits + (long)row * C;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = r[c];
            if (v > best || (v == best && c < bi) || bi == 0x7fffffff) {
                best = v;
                bi = c;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) {
                best = ov;
                bi = oi;
            }
        }
        if (conf) {
            float s = 0.f;
            for (int c = lane; c < C; c += 32) s += expf(r[c] - best);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (lane == 0) conf[row] = 1.0f / s;
        }
        if (lane == 0) {
            if (pred) pred[row] = bi;
            if (labels && labels[row] == (int64_t)bi) atomicAdd(&s_correct, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && counts) {
        const int rows_here = min(8, N - (int)blockIdx.x * 8);
        if (s_correct) atomicAdd(&counts[0], (unsigned long long)s_correct);
        atomicAdd(&counts[1], (unsigned long long)rows_here);
    }
}

}  // namespace b2r

// ================================================================================================================
extern "C" {

int b2r_final_conv1x1(const void* in, const float* weights, const float* bias, float* out_f32, uint8_t* out_u8, int N,
                      int H, int W, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && weights && bias, "null pointer");
    B2R_REQUIRE(out_f32 || out_u8, "no output requested");
    B2R_REQUIRE(N > 0 && H > 0 && W > 0, "bad shape");
    const long npix = (long)N * H * W;
    const long threads = npix * 8;
    const long blocks = (threads + 255) / 256;
    B2R_REQUIRE(blocks < (1L << 31), "too many pixels");
    final_conv1x1_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint4*>(in), weights, bias, out_f32,
                                                               out_u8, npix, H * W);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_maxpool2x2(const void* in, void* out, int N, int H, int W, int C, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && out, "null pointer");
    B2R_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad shape");
    B2R_REQUIRE(H >= 2 && W >= 2, "maxpool2x2 needs H, W >= 2");   // odd sizes: the last row / column is dropped (floor), like nn.MaxPool2d(2, 2)
    const long total = (long)N * (H / 2) * (W / 2) * (C / 8);
    const long blocks = (total + 255) / 256;
    B2R_REQUIRE(blocks < (1L << 31), "too large");
    maxpool2x2_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), N,
                                                            H, W, C / 8);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_resize_nearest_bf16(const void* in, void* out, int N, int H, int W, int out_h, int out_w, int C, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && out, "null pointer");
    B2R_REQUIRE(N > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && C > 0 && C % 8 == 0, "bad shape");
    const long total = (long)N * out_h * out_w * (C / 8);
    const long blocks = (total + 255) / 256;
    B2R_REQUIRE(blocks < (1L << 31), "too large");
    resize_nearest_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), N, H, W,
                                                                out_h, out_w, C / 8, float(H) / float(out_h), float(W) / float(out_w));
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_adaptive_avgpool7(const void* in, void* out, int N, int H, int W, int C, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && out, "null pointer");
    B2R_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad shape");
    const long total = (long)N * 49 * (C / 8);
    const long blocks = (total + 255) / 256;
    adaptive_avgpool7_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint4*>(in),
                                                                   static_cast<uint4*>(out), N, H, W, C / 8);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_linear_f32out(const void* in, const void* w, const float* bias, float* out, int B, int K, int O,
                      void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && w && bias && out, "null pointer");
    B2R_REQUIRE(B > 0 && K > 0 && O > 0 && K % 8 == 0, "bad shape B=%d K=%d O=%d", B, K, O);
    B2R_REQUIRE(K * 2 <= 48 * 1024, "K=%d too large for the staged row", K);
    linear_f32out_kernel<<<B, 256, K * 2, stream>>>(static_cast<const uint4*>(in), static_cast<const uint4*>(w), bias,
                                                    out, K / 8, O);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

int b2r_argmax_count(const float* logits, const int64_t* labels, int64_t* pred, float* conf, int64_t* counts, int N,
                     int C, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(logits != nullptr, "logits is null");
    B2R_REQUIRE(N > 0 && C > 0, "bad shape N=%d C=%d", N, C);
    B2R_REQUIRE(!(counts && !labels), "counts requested without labels");
    argmax_count_kernel<<<(N + 7) / 8, 256, 0, stream>>>(logits, labels, pred, conf,
                                                         reinterpret_cast<unsigned long long*>(counts), N, C);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

}  // extern "C"
