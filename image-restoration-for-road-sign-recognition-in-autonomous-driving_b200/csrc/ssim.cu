// Mean structural similarity of two u8 image batches: the second image-quality number 08_run_inference.py reports
// (08:123  ssim_metric(clean_img, output_bgr, data_range=255, channel_axis=2), skimage.metrics.structural_similarity).
//
// skimage's defaults on u8 input: 7x7 uniform window, sample covariance (x 49/48), K1 = 0.01, K2 = 0.03, float64, per
// channel, the 3-pixel border cropped before the mean, then the mean over channels.  For u8 images the five window
// statistics are ratios of integers, so the kernel forms the window sums exactly in int32
//     Sx, Sy <= 49*255      Sxx, Syy, Sxy <= 49*255^2      49*Sxx - Sx^2 <= 1.6e8
// and evaluates the SSIM ratio once per window in float64:
//     A1 = 2 Sx Sy / 49^2 + C1               B1 = (Sx^2 + Sy^2) / 49^2 + C1
//     A2 = 2 (49 Sxy - Sx Sy) / (49*48) + C2   B2 = ((49 Sxx - Sx^2) + (49 Syy - Sy^2)) / (49*48) + C2
//     S  = (A1 A2) / (B1 B2)
// (scipy's uniform_filter reaches the same means through two float64 passes; the results agree to ~1e-13.)
//
// One CTA per image.  A thread owns one interleaved column (x, c) of the cropped map and walks it top to bottom: the
// 7-tap horizontal sums of each input row are formed once and kept in a 7-row register ring, the window sums slide
// (+ new row, - row that left).  Per-thread, per-warp and per-CTA sums are taken in a fixed order: the result is
// deterministic.  Not HBM-bound (2 B per pixel-channel against ~100 instructions): a metric, not a hot-path kernel.
#include "b2r_internal.h"

namespace b2r {

constexpr int kSsimWin = 7;
constexpr int kSsimMaxThreads = 256;

__global__ void __launch_bounds__(kSsimMaxThreads) ssim_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                                   double* __restrict__ out, int H, int W, int C, double c1,
                                                                   double c2) {
    __shared__ double s_part[kSsimMaxThreads / 32];
    const int n = blockIdx.x;
    const long pitch = long(W) * C;
    const uint8_t* pa = a + long(n) * H * pitch;
    const uint8_t* pb = b + long(n) * H * pitch;
    const int cols = (W - (kSsimWin - 1)) * C;
    const double k_mean = 1.0 / 2401.0;   // 1 / 49^2
    const double k_cov = 1.0 / 2352.0;    // 1 / (49 * 48): mean of the window times the sample-covariance factor 49/48
    double acc = 0.0;
    for (int j = threadIdx.x; j < cols; j += blockDim.x) {
        int ring[kSsimWin][5];
        int tot[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < kSsimWin; ++k)
#pragma unroll
            for (int q = 0; q < 5; ++q) ring[k][q] = 0;
        for (int r0 = 0; r0 < H; r0 += kSsimWin) {
#pragma unroll
            for (int k = 0; k < kSsimWin; ++k) {   // ring slot == row % 7: static register indices
                const int r = r0 + k;
                if (r < H) {
                    const uint8_t* ra = pa + r * pitch + j;
                    const uint8_t* rb = pb + r * pitch + j;
                    int h[5] = {0, 0, 0, 0, 0};
#pragma unroll
                    for (int dx = 0; dx < kSsimWin; ++dx) {
                        const int x = __ldg(ra + dx * C), y = __ldg(rb + dx * C);
                        h[0] += x;
                        h[1] += y;
                        h[2] += x * x;
                        h[3] += y * y;
                        h[4] += x * y;
                    }
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        tot[q] += h[q] - ring[k][q];
                        ring[k][q] = h[q];
                    }
                    if (r >= kSsimWin - 1) {
                        const int sx = tot[0], sy = tot[1];
                        const int pxy = sx * sy;
                        const double a1 = fma(double(2 * pxy), k_mean, c1);
                        const double b1 = fma(double(sx * sx + sy * sy), k_mean, c1);
                        const double a2 = fma(double(2 * (49 * tot[4] - pxy)), k_cov, c2);
                        const double b2 = fma(double((49 * tot[2] - sx * sx) + (49 * tot[3] - sy * sy)), k_cov, c2);
                        acc += (a1 * a2) / (b1 * b2);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int(blockDim.x) + 31) / 32; ++w) s += s_part[w];
        out[n] = s / (double(cols) * double(H - (kSsimWin - 1)));
    }
}

}  // namespace b2r

extern "C" int b2r_ssim_u8(const uint8_t* a, const uint8_t* b, double* ssim, int N, int H, int W, int C, double data_range,
                           void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(a && b && ssim, "null pointer");
    B2R_REQUIRE(N > 0 && C > 0, "bad shape N=%d C=%d", N, C);
    B2R_REQUIRE(H >= kSsimWin && W >= kSsimWin, "win_size 7 exceeds the image extent %dx%d (skimage raises ValueError)", H, W);
    B2R_REQUIRE((long)H * W * C < (1L << 31), "image too large");
    B2R_REQUIRE(data_range > 0, "data_range must be positive");
    const int cols = (W - (kSsimWin - 1)) * C;
    const int passes = (cols + kSsimMaxThreads - 1) / kSsimMaxThreads;
    int threads = (((cols + passes - 1) / passes) + 31) / 32 * 32;   // equal column shares, whole warps
    if (threads > kSsimMaxThreads) threads = kSsimMaxThreads;
    const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
    ssim_u8_kernel<<<N, threads, 0, stream>>>(a, b, ssim, H, W, C, c1, c2);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}
