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

}  // namespace b2r
