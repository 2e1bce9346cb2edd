// Fused compound degradation: motion blur (+) fog (+) AWGN on u8 NHWC images, one launch, HBM in -> HBM out once.
//
// Reference arithmetic being replaced (all per image, NumPy/OpenCV on the CPU):
//   blur  : cv2.filter2D(u8, -1, k)  — correlation, anchor d/2, BORDER_REFLECT_101, f32 accumulate in row-major tap
//           order, round-half-even + saturate                           (16_gen_compound_data.py:23-26, 14:54-61)
//   fog   : img*t + A*(1-t) in float32 (t, A Python floats)              (16:30-31, 14:39-43, 15:99-103)
//   noise : img + np.random.normal(0, sigma, shape) -> float64           (16:34-35, 14:46-49, 15:106-108)
//   quant : np.clip(x*255, 0, 255).astype(np.uint8) — truncation         (16:37, 14:53, 14:64, 15:111)
// chain(v) = quant(noise(fog(v/255)));  script 16 computes chain(blur(in)), scripts 14/15 compute blur(chain(in)).
// (u8 -> f32/255 -> *255 -> truncate is the identity on 0..255, so the reference's intermediate re-quantisations
//  that are not listed here are no-ops; tests/test_oracle_degrade.py checks that claim.)
//
// Structure (v3).  One CTA = 16 image rows x full width of one image.
//   stage 1  the pre-blur image (raw input, or chain(input) for the scripts-14/15 order) is staged in shared memory
//            as three PLANAR, row-major fp32 planes with the halo of this image's kernel (REFLECT_101), so the u8->f32
//            conversion is paid once per staged pixel instead of once per tap.  Image column x sits at staged column
//            x + 8 (kStageShift), which keeps every 4-pixel group 16-byte aligned: one STS.128 per plane and group;
//   stage 2  a work item is 4 consecutive pixels x 3 channels (12 accumulators).  The kernel row is held as a
//            zero-padded tap array T[t] = k[t - (8 - anchor)], so that tap t of pixel x reads staged column x + t:
//            groups of four taps start on 16-byte boundaries for every item.  Per group ONE LDS.128 per plane (the
//            next four staged columns) and one LDS.128 of taps feed 48 FMAs with static register indices.  Only the
//            groups between a row's first and last non-zero tap are walked; the zero taps inside are applied too:
//            fma(0, x, acc) == acc, so the result is bit-identical to OpenCV's "non-zero taps, row-major" engine
//            (degree <= 11; larger kernels go through OpenCV's DFT path, see tests/test_degrade_gpu.py);
//   stage 3  round-half-even to u8, then chain() for the script-16 order, and one aligned 12-byte store per item.
// Noise: one Philox4x32-10 call per pixel keyed by (seed; pixel index, global image index) -> Box-Muller with the
// hardware log2/sin/cos units -> three normals; fp32 chain.  When the caller injects the reference's own noise tensor
// (parity path) the add is done in float64 exactly as NumPy promotes it.
// Point-wise recipes (no image of the batch blurs or gets noise: fog alone, the fog half of BASELINE config 2) take
// degrade_pointwise_kernel instead: chain() then depends on the byte value only, so every CTA evaluates it for the 256
// possible inputs of its image and the image streams through that table at the HBM rate (16-byte loads / stores).
#include "b2r_internal.h"
#include "philox.cuh"
#include "stream_u8.cuh"

namespace b2r {

constexpr int kDegRows = 16;
constexpr int kDegThreads = 448;   // 14 warps, two CTAs per SM (shared memory): 28 warps hide the staging loads

struct DegradeParams {
    const uint8_t* in;
    uint8_t* out;
    const float* taps;
    const int32_t* ksize;
    const float* fog_t;
    const float* fog_add;
    const int32_t* fog_on;
    const float* sigma;
    const double* noise;
    uint64_t seed, image_index0;
    int N, H, W, order, flags;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

struct ImgParams {
    float t, add, sigma;
    int fog_on, noise_on, clip_after;
    uint64_t seed, image;
    const double* noise;  // injected noise for this image or nullptr
};

// chain(v): quant(noise(fog(v / 255))) for the 3 channels of the pixel at linear index `pix`.
// `unit` = shared-memory table of float(u)/255 (IEEE division, as NumPy computes `astype(float32) / 255.0`).
// (item, j) = the pixel's 4-pixel work item and its position in it: they key the Philox noise (stream 2, philox.cuh).
__device__ __forceinline__ void chain3(const ImgParams& ip, const float* unit, uint32_t pix, uint32_t item, int j,
                                       const int (&v)[3], int (&q)[3]) {
    float x[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        x[c] = unit[v[c]];
        if (ip.fog_on) x[c] = __fadd_rn(__fmul_rn(x[c], ip.t), ip.add);  // two roundings, like NumPy (no FMA)
    }
    if (ip.noise_on && ip.noise != nullptr) {
        // parity path: float32 image + float64 noise -> float64, exactly as NumPy promotes
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double y = double(x[c]) + ip.noise[size_t(pix) * 3 + c];
            if (ip.clip_after) y = fmin(fmax(y, 0.0), 1.0);
            y = fmin(fmax(y * 255.0, 0.0), 255.0);
            q[c] = int(y);  // truncation
        }
        return;
    }
    if (ip.noise_on) {
        float z[3];
        pixel_normals_v2(ip.seed, ip.image, item, j, z);
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = fmaf(ip.sigma, z[c], x[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float y = x[c];
        if (ip.clip_after) y = fminf(fmaxf(y, 0.f), 1.f);
        // np.clip(y * 255, 0, 255).astype(np.uint8): the saturating, truncating conversion does both (F2IP.U8.TRUNC)
        uint32_t b;
        asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(b) : "f"(__fmul_rn(y, 255.0f)));
        q[c] = int(b);
    }
}

// chain() for a whole work item (4 consecutive pixels of row y, group gx): the noise of the item comes from three Philox
// calls (noise stream 2, philox.cuh).  `pix0` = linear index of the item's first pixel (injected-noise parity path only).
__device__ __forceinline__ void chain_item(const ImgParams& ip, const float* unit, uint32_t item, uint32_t pix0, int nvalid,
                                           const int (&v)[4][3], int (&q)[4][3]) {
    if (ip.noise_on && ip.noise != nullptr) {   // parity path: per pixel, float64 like NumPy (chain3); never past the row end
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            q[j][0] = q[j][1] = q[j][2] = 0;
            if (j < nvalid) chain3(ip, unit, pix0 + j, item, j, v[j], q[j]);
        }
        return;
    }
    float x[12];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float f = unit[v[j][c]];
            if (ip.fog_on) f = __fadd_rn(__fmul_rn(f, ip.t), ip.add);  // two roundings, like NumPy (no FMA)
            x[3 * j + c] = f;
        }
    if (ip.noise_on) {
        float z[12];
        item_normals(ip.seed, ip.image, item, z);
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = fmaf(ip.sigma, z[i], x[i]);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        float y = x[i];
        if (ip.clip_after) y = fminf(fmaxf(y, 0.f), 1.f);
        uint32_t b;
        asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(b) : "f"(__fmul_rn(y, 255.0f)));
        q[i / 3][i % 3] = int(b);
    }
}

// 12 bytes of an item <-> three 32-bit words (pixel j channel c = byte 3j + c), one PRMT per byte
__device__ __forceinline__ void unpack12(const uint32_t (&w)[3], int (&v)[4][3]) {
#pragma unroll
    for (int b = 0; b < 12; ++b) v[b / 3][b % 3] = int(__byte_perm(w[b >> 2], 0u, 0x4440u + uint32_t(b & 3)));
}
__device__ __forceinline__ void pack12(const int (&q)[4][3], uint32_t (&w)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int b = 4 * i;
        const uint32_t lo = __byte_perm(uint32_t(q[b / 3][b % 3]), uint32_t(q[(b + 1) / 3][(b + 1) % 3]), 0x0040u);
        const uint32_t hi = __byte_perm(uint32_t(q[(b + 2) / 3][(b + 2) % 3]), uint32_t(q[(b + 3) / 3][(b + 3) % 3]), 0x0040u);
        w[i] = __byte_perm(lo, hi, 0x5410u);
    }
}

// No blur for this image: the per-pixel chain on items of 4 pixels (three aligned 32-bit loads / stores) over image rows
// r0 .. r0 + rows - 1.  Shared by the blur kernel (images of a mixed batch that do not blur) and degrade_noblur_kernel.
__device__ __forceinline__ void chain_rows(const ImgParams& ip, const float* s_unit, const uint8_t* __restrict__ img_in,
                                           uint8_t* __restrict__ img_out, int W, int r0, int rows, bool rows_aligned,
                                           int tid, int nthreads) {
    const int groups = (W + 3) >> 2;
    for (int item = tid; item < rows * groups; item += nthreads) {
        const int y = item / groups;
        const int x0 = (item - y * groups) << 2;
        const size_t off = (size_t(r0 + y) * W + x0) * 3;
        const bool fast = rows_aligned && x0 + 4 <= W;
        uint32_t bytes[3] = {0u, 0u, 0u};
        if (fast) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(img_in + off);
            bytes[0] = __ldg(s32);
            bytes[1] = __ldg(s32 + 1);
            bytes[2] = __ldg(s32 + 2);
        } else {
            const int nb = min(4, W - x0) * 3;
            for (int b = 0; b < nb; ++b) bytes[b >> 2] |= uint32_t(img_in[off + b]) << (8 * (b & 3));
        }
        uint32_t outb[3];
        int v[4][3], q[4][3];
        unpack12(bytes, v);
        chain_item(ip, s_unit, uint32_t((r0 + y) * groups + (x0 >> 2)), uint32_t((r0 + y) * W + x0), min(4, W - x0), v, q);
        pack12(q, outb);
        if (fast) {
            uint32_t* d32 = reinterpret_cast<uint32_t*>(img_out + off);
            d32[0] = outb[0];
            d32[1] = outb[1];
            d32[2] = outb[2];
        } else {
            const int nb = min(4, W - x0) * 3;
            for (int b = 0; b < nb; ++b) img_out[off + b] = uint8_t(outb[b >> 2] >> (8 * (b & 3)));
        }
    }
}

// Planar staging layout: plane c, staged row sy, staged column sc at  c * plane + sy * pitch + sc,  image column x at
// sc = x + kStageShift.  Tap t of the shifted tap array (kTapSlots entries, groups of 4) reads column x + t, so the widest
// read of an item at x0 (a multiple of 4) is x0 + kTapSlots + 3: pitch = roundup4(W) + kTapSlots keeps it inside the row.
constexpr int kStageShift = 8;                       // >= the largest anchor (B2R_MAX_BLUR / 2 = 7), multiple of 4
constexpr int kTapSlots = 24;                        // (kStageShift - anchor) + B2R_MAX_BLUR <= 7 + 15, rounded up to groups of 4
constexpr int kTapGroups = kTapSlots / 4;

static __host__ __device__ int degrade_pitch(int W) { return ((W + 3) & ~3) + kTapSlots; }

__device__ __forceinline__ int round_u8(float v) {   // cvRound + saturate_cast<uchar>: round half to even, clamp
    uint32_t b;
    asm("cvt.rni.u8.f32 %0, %1;" : "=r"(b) : "f"(v));
    return int(b);
}

__global__ void __launch_bounds__(kDegThreads, 2) degrade_kernel(const DegradeParams P) {
    extern __shared__ __align__(16) float s_planes[];  // [3][srows][pitch], row-major
    __shared__ __align__(16) float s_taps[B2R_MAX_BLUR][kTapSlots];  // shifted, zero-padded tap rows (see above)
    __shared__ float s_unit[256];
    __shared__ int s_rows[B2R_MAX_BLUR + 1];  // [0] = number of non-empty kernel rows, then ky | first group << 8 | groups << 16
    __shared__ int s_seg[B2R_MAX_BLUR][2];  // per kernel row: first tap group with a non-zero tap / number of groups up to the last one (0: empty row)

    const int n = blockIdx.y;
    const int r0 = blockIdx.x * kDegRows;
    const int H = P.H, W = P.W;
    const int rows = min(kDegRows, H - r0);
    const int tid = threadIdx.x;

    int d = P.ksize ? P.ksize[n] : 0;
    if (d <= 1) d = 0;  // "if degree > 1" (14_train_unified_advanced.py:56)
    if (d > B2R_MAX_BLUR) __trap();  // ksize[] is a device array the host cannot validate: fail the launch, never overrun s_taps / s_rows
    ImgParams ip;
    ip.fog_on = P.fog_on ? P.fog_on[n] : 0;
    ip.t = ip.fog_on ? P.fog_t[n] : 1.f;
    ip.add = ip.fog_on ? P.fog_add[n] : 0.f;
    ip.sigma = P.sigma ? P.sigma[n] : 0.f;
    ip.noise_on = ip.sigma > 0.f;
    ip.clip_after = (P.flags & B2R_DEG_CLIP_AFTER_NOISE) != 0;
    ip.seed = P.seed;
    ip.image = P.image_index0 + uint64_t(n);
    ip.noise = P.noise ? P.noise + size_t(n) * H * W * 3 : nullptr;
    // without fog and noise chain() is the identity on 0..255 (u8 -> /255 -> *255 -> truncate; tests/test_oracle_degrade.py)
    const bool chain_on = ip.fog_on || ip.noise_on;

    const uint8_t* img_in = P.in + size_t(n) * H * W * 3;
    uint8_t* img_out = P.out + size_t(n) * H * W * 3;
    const bool rows_aligned = ((W * 3) & 3) == 0 && ((reinterpret_cast<uintptr_t>(img_in) | reinterpret_cast<uintptr_t>(img_out)) & 3) == 0;
    const int groups = (W + 3) >> 2;

    for (int i = tid; i < 256; i += kDegThreads) s_unit[i] = __fdiv_rn(float(i), 255.0f);

    if (d == 0) {
        __syncthreads();  // s_unit
        chain_rows(ip, s_unit, img_in, img_out, W, r0, rows, rows_aligned, tid, kDegThreads);
        return;
    }

    // halo of THIS kernel: cv2 anchor = d/2, so taps reach a pixels up/left and d-1-a pixels down/right
    const int a = d / 2;
    const int hb = d - 1 - a;
    const int sh = kStageShift - a;   // tap kx of pixel x reads staged column x + kx + sh
    for (int i = tid; i < B2R_MAX_BLUR * kTapSlots; i += kDegThreads) {
        const int ky = i / kTapSlots, kx = i - ky * kTapSlots - sh;
        s_taps[ky][i - ky * kTapSlots] =
            (ky < d && kx >= 0 && kx < d) ? P.taps[size_t(n) * (B2R_MAX_BLUR * B2R_MAX_BLUR) + ky * d + kx] : 0.f;
    }
    if (tid < d) {
        const float* tr = P.taps + size_t(n) * (B2R_MAX_BLUR * B2R_MAX_BLUR) + tid * d;
        int first = d, last = -1;
        for (int k = 0; k < d; ++k)
            if (tr[k] != 0.f) {
                if (first == d) first = k;
                last = k;
            }
        s_seg[tid][0] = (first + sh) >> 2;
        s_seg[tid][1] = last < first ? 0 : ((last + sh) >> 2) - ((first + sh) >> 2) + 1;
    }

    const int srows = rows + a + hb;
    const int pitch = degrade_pitch(W);
    const int plane = srows * pitch;
    const bool chain_first = P.order == B2R_ORDER_FOG_NOISE_BLUR;
    __syncthreads();  // s_unit, s_seg
    if (tid == 0) {
        int m = 0;
        for (int ky = 0; ky < d; ++ky)
            if (s_seg[ky][1]) s_rows[++m] = ky | s_seg[ky][0] << 8 | s_seg[ky][1] << 16;
        s_rows[0] = m;   // published by the barrier that ends the staging
    }

    // (a) image columns: items of 4 pixels = three aligned, coalesced 32-bit loads and one 16-byte store per plane
    //     (the staging loop is where the kernel meets HBM latency, so it is unrolled to keep several loads in flight)
#pragma unroll 2
    for (int it = tid; it < srows * groups; it += kDegThreads) {
        const int sy = it / groups;
        const int g = it - sy * groups;
        const int x0 = g << 2;
        const int h = reflect101(r0 - a + sy, H);
        const size_t off = (size_t(h) * W + x0) * 3;
        const bool full = x0 + 4 <= W;
        uint32_t bytes[3] = {0u, 0u, 0u};
        if (rows_aligned && full) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(img_in + off);
            bytes[0] = __ldg(s32);
            bytes[1] = __ldg(s32 + 1);
            bytes[2] = __ldg(s32 + 2);
        } else {
            const int nb = min(4, W - x0) * 3;
            for (int b2 = 0; b2 < nb; ++b2) bytes[b2 >> 2] |= uint32_t(img_in[off + b2]) << (8 * (b2 & 3));
        }
        float f[3][4];
        {
            int v[4][3];
            unpack12(bytes, v);
            if (chain_first && chain_on) {
                int q[4][3];
                chain_item(ip, s_unit, uint32_t(h * groups + g), uint32_t(h * W + x0), min(4, W - x0), v, q);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[j][c] = q[j][c];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) f[c][j] = float(v[j][c]);
        }
        float* dstp = s_planes + sy * pitch + kStageShift + x0;
        if (full) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                *reinterpret_cast<float4*>(dstp + c * plane) = make_float4(f[c][0], f[c][1], f[c][2], f[c][3]);
        } else {   // the columns right of the image belong to loop (b)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (x0 + j < W) dstp[c * plane + j] = f[c][j];
        }
    }
    // (b) the columns left / right of the image: REFLECT_101 where this kernel's taps can reach, zero elsewhere (the
    //     padded zero taps must never meet a NaN)
    const int hc = pitch - W;
    for (int it = tid; it < srows * hc; it += kDegThreads) {
        const int sy = it / hc;
        const int k = it - sy * hc;
        const int sc = k < kStageShift ? k : W + k;
        float f[3] = {0.f, 0.f, 0.f};
        if (sc >= kStageShift - a && sc < kStageShift + W + hb) {
            const int h = reflect101(r0 - a + sy, H);
            const int w = reflect101(sc - kStageShift, W);
            const uint32_t pix = uint32_t(h * W + w);
            int v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = img_in[size_t(pix) * 3 + c];
            if (chain_first && chain_on) {
                int q[3];
                chain3(ip, s_unit, pix, uint32_t(h * groups + (w >> 2)), w & 3, v, q);
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = q[c];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] = float(v[c]);
        }
        const int o = sy * pitch + sc;
#pragma unroll
        for (int c = 0; c < 3; ++c) s_planes[c * plane + o] = f[c];
    }
    __syncthreads();

    const int plane4 = plane >> 2;   // pitch is a multiple of 4
    const int nrows = s_rows[0];
    for (int item = tid; item < rows * groups; item += kDegThreads) {
        const int y = item / groups;
        const int g = item - y * groups;
        const int x0 = g << 2;
        float acc[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
        for (int r = 1; r <= nrows; ++r) {                   // non-empty kernel rows, top to bottom (uniform per CTA)
            const int e = s_rows[r];
            const int ky = e & 0xFF, g0 = (e >> 8) & 0xFF, ng = e >> 16;
            // staged columns x0 + 4 * (g0 + i) .. + 3 for i = 0 .. ng: 16-byte aligned for every item
            const float4* wp = reinterpret_cast<const float4*>(s_planes + (y + ky) * pitch + x0) + g0;
            const float4* tp = reinterpret_cast<const float4*>(&s_taps[ky][0]) + g0;
            float4 lo[3], hi[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) lo[c] = wp[c * plane4];
#define B2R_DEG_GROUP(LO, HI, I)                                                  \
    {                                                                             \
        const float4 k = tp[I];                                                   \
        _Pragma("unroll") for (int c = 0; c < 3; ++c) {                           \
            HI[c] = wp[c * plane4 + (I) + 1];                                     \
            acc[c][0] = fmaf(k.x, LO[c].x, acc[c][0]);                            \
            acc[c][1] = fmaf(k.x, LO[c].y, acc[c][1]);                            \
            acc[c][2] = fmaf(k.x, LO[c].z, acc[c][2]);                            \
            acc[c][3] = fmaf(k.x, LO[c].w, acc[c][3]);                            \
            acc[c][0] = fmaf(k.y, LO[c].y, acc[c][0]);                            \
            acc[c][1] = fmaf(k.y, LO[c].z, acc[c][1]);                            \
            acc[c][2] = fmaf(k.y, LO[c].w, acc[c][2]);                            \
            acc[c][3] = fmaf(k.y, HI[c].x, acc[c][3]);                            \
            acc[c][0] = fmaf(k.z, LO[c].z, acc[c][0]);                            \
            acc[c][1] = fmaf(k.z, LO[c].w, acc[c][1]);                            \
            acc[c][2] = fmaf(k.z, HI[c].x, acc[c][2]);                            \
            acc[c][3] = fmaf(k.z, HI[c].y, acc[c][3]);                            \
            acc[c][0] = fmaf(k.w, LO[c].w, acc[c][0]);                            \
            acc[c][1] = fmaf(k.w, HI[c].x, acc[c][1]);                            \
            acc[c][2] = fmaf(k.w, HI[c].y, acc[c][2]);                            \
            acc[c][3] = fmaf(k.w, HI[c].z, acc[c][3]);                            \
        }                                                                         \
    }
            // the two window registers swap roles, so the loop is unrolled by two with a uniform early exit
            for (int i = 0;;) {
                B2R_DEG_GROUP(lo, hi, i)
                if (++i == ng) break;
                B2R_DEG_GROUP(hi, lo, i)
                if (++i == ng) break;
            }
#undef B2R_DEG_GROUP
        }
        // OpenCV's row filter (4.13, AVX2 build; measured against cv2 for W = 21..33) runs its vector body, which fuses
        // multiply and add, over 4 * floor(3W / 4) interleaved values of a row and a scalar loop with a separate
        // multiply and add over the last 3W mod 4 values: channels c >= 3 - (3W mod 4) of pixel W - 1.  Those few
        // values are recomputed here the scalar way (exact .5 ties round differently otherwise).  Absent for W = 224.
        const int tail = (3 * W) & 3;
        const bool tail_item = tail != 0 && x0 <= W - 1 && W - 1 < x0 + 4;
        float tacc[3] = {0.f, 0.f, 0.f};
        if (tail_item) {
            for (int ky = 0; ky < d; ++ky)
                for (int kx = 0; kx < d; ++kx) {
                    const float k = s_taps[ky][kx + sh];
                    if (k != 0.f) {
                        const float* vp = s_planes + (y + ky) * pitch + (W - 1) + kx + sh;
#pragma unroll
                        for (int c = 0; c < 3; ++c) tacc[c] = __fadd_rn(tacc[c], __fmul_rn(k, vp[c * plane]));
                    }
                }
        }
        uint32_t bytes[3];  // 12 output bytes: pixel j channel c -> byte 3*j + c
        {
            int v[4][3];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    v[j][c] = round_u8(acc[c][j]);
                    if (tail_item && x0 + j == W - 1 && c >= 3 - tail) v[j][c] = round_u8(tacc[c]);
                }
            if (!chain_first && chain_on) {
                int q[4][3];
                chain_item(ip, s_unit, uint32_t((r0 + y) * groups + g), uint32_t((r0 + y) * W + x0), min(4, W - x0), v, q);
                pack12(q, bytes);
            } else {
                pack12(v, bytes);
            }
        }
        uint8_t* dst = img_out + (size_t(r0 + y) * W + x0) * 3;
        if (rows_aligned && x0 + 4 <= W) {
            uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
            d32[0] = bytes[0];
            d32[1] = bytes[1];
            d32[2] = bytes[2];
        } else {
            const int nb = min(4, W - x0) * 3;
            for (int b = 0; b < nb; ++b) dst[b] = uint8_t(bytes[b >> 2] >> (8 * (b & 3)));
        }
    }
}

// ksize == NULL (no image of the batch blurs) with noise somewhere: the chain alone, without the blur kernel's register
// cap and shared-memory planes.  Same grid as degrade_kernel: one CTA = 16 rows of one image.
constexpr int kNoBlurThreads = 224;
__global__ void __launch_bounds__(kNoBlurThreads) degrade_noblur_kernel(const DegradeParams P) {
    __shared__ float s_unit[256];
    const int n = blockIdx.y;
    const int r0 = blockIdx.x * kDegRows;
    const int H = P.H, W = P.W;
    ImgParams ip;
    ip.fog_on = P.fog_on ? P.fog_on[n] : 0;
    ip.t = ip.fog_on ? P.fog_t[n] : 1.f;
    ip.add = ip.fog_on ? P.fog_add[n] : 0.f;
    ip.sigma = P.sigma ? P.sigma[n] : 0.f;
    ip.noise_on = ip.sigma > 0.f;
    ip.clip_after = (P.flags & B2R_DEG_CLIP_AFTER_NOISE) != 0;
    ip.seed = P.seed;
    ip.image = P.image_index0 + uint64_t(n);
    ip.noise = P.noise ? P.noise + size_t(n) * H * W * 3 : nullptr;
    const uint8_t* img_in = P.in + size_t(n) * H * W * 3;
    uint8_t* img_out = P.out + size_t(n) * H * W * 3;
    const bool rows_aligned = ((W * 3) & 3) == 0 && ((reinterpret_cast<uintptr_t>(img_in) | reinterpret_cast<uintptr_t>(img_out)) & 3) == 0;
    for (int i = threadIdx.x; i < 256; i += kNoBlurThreads) s_unit[i] = __fdiv_rn(float(i), 255.0f);
    __syncthreads();
    chain_rows(ip, s_unit, img_in, img_out, W, r0, min(kDegRows, H - r0), rows_aligned, threadIdx.x, kNoBlurThreads);
}

// ksize == NULL and sigma == NULL: out = table_n[in], table_n[v] = chain(v) with image n's fog parameters.
__global__ void __launch_bounds__(kGenThreads) degrade_pointwise_kernel(const DegradeParams P) {
    __shared__ float s_unit[256];
    __shared__ uint8_t s_lut[256];
    const int n = blockIdx.y;
    const int tid = threadIdx.x;   // kGenThreads == 256: one table entry per thread
    ImgParams ip;
    ip.fog_on = P.fog_on ? P.fog_on[n] : 0;
    ip.t = ip.fog_on ? P.fog_t[n] : 1.f;
    ip.add = ip.fog_on ? P.fog_add[n] : 0.f;
    ip.sigma = 0.f;
    ip.noise_on = 0;
    ip.clip_after = (P.flags & B2R_DEG_CLIP_AFTER_NOISE) != 0;
    ip.seed = 0;
    ip.image = 0;
    ip.noise = nullptr;
    s_unit[tid] = __fdiv_rn(float(tid), 255.0f);
    __syncthreads();
    const int v[3] = {tid, tid, tid};
    int q[3];
    chain3(ip, s_unit, 0u, 0u, 0, v, q);
    s_lut[tid] = uint8_t(q[0]);
    __syncthreads();
    const long elems = long(P.H) * P.W * 3;
    apply_table(s_lut, P.in + long(n) * elems, P.out + long(n) * elems, elems);
}

static size_t degrade_smem_bytes(int W) {
    return size_t(3) * (kDegRows + B2R_MAX_BLUR - 1) * degrade_pitch(W) * sizeof(float);
}

}  // namespace b2r

extern "C" int b2r_degrade(const uint8_t* in, uint8_t* out, int N, int H, int W, const float* taps,
                           const int32_t* ksize, const float* fog_t, const float* fog_add, const int32_t* fog_on,
                           const float* sigma, const double* noise, uint64_t seed, uint64_t image_index0, int order,
                           int flags, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && out, "null image pointer");
    B2R_REQUIRE(N > 0 && H > 0 && W > 0, "bad shape N=%d H=%d W=%d", N, H, W);
    B2R_REQUIRE(N <= 65535, "N=%d exceeds gridDim.y; split the batch", N);
    B2R_REQUIRE((long)H * W < (1L << 31) / 3, "image too large");
    B2R_REQUIRE(order == B2R_ORDER_BLUR_FOG_NOISE || order == B2R_ORDER_FOG_NOISE_BLUR, "order=%d", order);
    B2R_REQUIRE((ksize == nullptr) == (taps == nullptr), "ksize and taps must both be given or both be null");
    B2R_REQUIRE(fog_on == nullptr || (fog_t && fog_add), "fog_on given without fog_t / fog_add");
    B2R_REQUIRE(!(noise && !sigma), "injected noise needs sigma[] as the per-image on/off switch");
    {   // the blur reads halo rows that neighbouring CTAs write when the ranges overlap
        const size_t bytes = size_t(N) * H * W * 3;
        const bool overlap = in < out + bytes && out < in + bytes;
        B2R_REQUIRE(!(ksize && overlap), "in / out ranges overlap and ksize is given: the blur cannot run in place");
    }
    const size_t smem = ksize ? degrade_smem_bytes(W) : 0;
    B2R_REQUIRE(smem <= 200 * 1024, "W=%d too wide for the staged tile", W);
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (smem > 48 * 1024 && (dev >= 64 || !attr_set[dev])) {
        B2R_CUDA(cudaFuncSetAttribute(degrade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev < 64) attr_set[dev] = true;
    }
    DegradeParams P;
    P.in = in;
    P.out = out;
    P.taps = taps;
    P.ksize = ksize;
    P.fog_t = fog_t;
    P.fog_add = fog_add;
    P.fog_on = fog_on;
    P.sigma = sigma;
    P.noise = noise;
    P.seed = seed;
    P.image_index0 = image_index0;
    P.N = N;
    P.H = H;
    P.W = W;
    P.order = order;
    P.flags = flags;
    if (!ksize && !sigma) {
        static_assert(kGenThreads == 256, "one table entry per thread");
        dim3 pgrid;
        const int rc = gen_grid(long(H) * W * 3, N, &pgrid, kGenBytesPerThread);
        if (rc) return rc;
        degrade_pointwise_kernel<<<pgrid, kGenThreads, 0, stream>>>(P);
        B2R_CHECK_LAUNCH();
        return B2R_OK;
    }
    dim3 grid((H + kDegRows - 1) / kDegRows, N);
    if (!ksize) {
        degrade_noblur_kernel<<<grid, kNoBlurThreads, 0, stream>>>(P);
        B2R_CHECK_LAUNCH();
        return B2R_OK;
    }
    degrade_kernel<<<grid, kDegThreads, smem, stream>>>(P);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}
