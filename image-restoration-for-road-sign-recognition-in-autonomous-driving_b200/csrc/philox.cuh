// Counter-based noise shared by degrade.cu and generators.cu: results depend only on (seed, global image index, pixel),
// never on the launch geometry or the world size.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2r {

// Philox4x32-10 (Salmon et al., SC'11); constants and round structure of Random123.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

// sqrt(-2 ln u) for u in [2^-33, 1]: u is never denormal, so the bare SFU forms (MUFU.LG2, MUFU.SQRT) are used without
// the denormal / special-value guards that __log2f and sqrtf carry (~10 instructions per call).
__device__ __forceinline__ float box_muller_radius(float u) {
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l));
    return r;
}

// three standard normals for one pixel (channels 0..2): Box-Muller on the four Philox words
__device__ __forceinline__ void pixel_normals(uint64_t seed, uint64_t image, uint32_t pixel, float (&z)[3]) {
    const uint4 r = philox4x32_10(make_uint4(pixel, uint32_t(image), uint32_t(image >> 32), 0u),
                                  make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
    const float k = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = fmaf(float(r.x), k, 0.5f * k), u1 = fmaf(float(r.y), k, 0.5f * k);
    const float u2 = fmaf(float(r.z), k, 0.5f * k), u3 = fmaf(float(r.w), k, 0.5f * k);
    // sqrt(-2 ln u) = sqrt(-2 ln2 * log2 u); sin/cos(2 pi u) on the SFU (abs error ~4e-7: far below one u8 LSB / sigma)
    const float ra = box_muller_radius(u0), rb = box_muller_radius(u2);
    float s, c;
    __sincosf(6.283185307179586f * u1, &s, &c);
    z[0] = ra * s;
    z[1] = ra * c;
    z[2] = rb * __sinf(6.283185307179586f * u3);
}

// ---- noise stream 2 (b2r_degrade since round 2): FOUR normals per Philox call instead of three.
// A work item is 4 consecutive pixels of one image row (x = 4 gx .. 4 gx + 3) = 12 values; its normals come from three calls
// with counter (item, image_lo, image_hi, 4 + k), item = y * ceil(W / 4) + gx, k = 0..2: call k yields
//   n[4k] = r(u0) sin(2 pi u1), n[4k + 1] = r(u0) cos(2 pi u1), n[4k + 2] = r(u2) sin(2 pi u3), n[4k + 3] = r(u2) cos(2 pi u3)
// and value (pixel j, channel c) of the item takes n[3j + c].  A pure function of (seed, image, y, x, c, W): independent of
// the launch geometry, the micro-batch split and the world size.  25 % fewer Philox rounds and transcendental calls per
// pixel than stream 1 (one call per pixel, one of its four normals unused); oracle/degrade_oracle.py::philox_normals_v2.
__device__ __forceinline__ void box_muller4(const uint4 r, float* n) {
    const float k = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = fmaf(float(r.x), k, 0.5f * k), u1 = fmaf(float(r.y), k, 0.5f * k);
    const float u2 = fmaf(float(r.z), k, 0.5f * k), u3 = fmaf(float(r.w), k, 0.5f * k);
    const float ra = box_muller_radius(u0), rb = box_muller_radius(u2);
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
    n[0] = ra * s0;
    n[1] = ra * c0;
    n[2] = rb * s1;
    n[3] = rb * c1;
}

__device__ __forceinline__ void item_normals(uint64_t seed, uint64_t image, uint32_t item, float (&z)[12]) {
    const uint2 key = make_uint2(uint32_t(seed), uint32_t(seed >> 32));
#pragma unroll
    for (int k = 0; k < 3; ++k)
        box_muller4(philox4x32_10(make_uint4(item, uint32_t(image), uint32_t(image >> 32), 4u + k), key), &z[4 * k]);
}

// the three normals of ONE pixel (j = x & 3 inside its item): one or two of the item's calls (halo columns only)
__device__ __forceinline__ void pixel_normals_v2(uint64_t seed, uint64_t image, uint32_t item, int j, float (&z)[3]) {
    const uint2 key = make_uint2(uint32_t(seed), uint32_t(seed >> 32));
    const int i0 = 3 * j, k0 = i0 >> 2, k1 = (i0 + 2) >> 2;
    float n[8];
    box_muller4(philox4x32_10(make_uint4(item, uint32_t(image), uint32_t(image >> 32), 4u + k0), key), &n[0]);
    if (k1 != k0) box_muller4(philox4x32_10(make_uint4(item, uint32_t(image), uint32_t(image >> 32), 4u + k1), key), &n[4]);
#pragma unroll
    for (int c = 0; c < 3; ++c) z[c] = n[i0 + c - 4 * k0];
}

}  // namespace b2r
