// Streaming u8 -> u8 kernels (csrc/generators.cu, the point-wise path of csrc/degrade.cu): launch shape and the
// table walk they share.  Bound by HBM: algorithmic bytes = elems x (1 read + 1 written) per image.
#pragma once
#include "b2r_internal.h"

namespace b2r {

constexpr int kGenThreads = 256;
constexpr int kGenBytesPerThread = 16;

#ifndef B2R_STREAM_UNROLL
#define B2R_STREAM_UNROLL 4   // 16-byte loads a thread keeps in flight per loop trip (1 -> 4: 77 % -> 86 % of the HBM copy rate)
#endif

__device__ __forceinline__ uint4 table_lookup16(const uint8_t* s_lut, const uint4 v) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        w[k] = uint32_t(s_lut[w[k] & 0xFF]) | uint32_t(s_lut[(w[k] >> 8) & 0xFF]) << 8 |
               uint32_t(s_lut[(w[k] >> 16) & 0xFF]) << 16 | uint32_t(s_lut[w[k] >> 24]) << 24;
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// dst[i] = s_lut[src[i]] for one image: 16 bytes per thread and step when both pointers are 16-byte aligned
__device__ __forceinline__ void apply_table(const uint8_t* s_lut, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                            long elems) {
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const long step = (long)gridDim.x * kGenThreads * kGenBytesPerThread;
    long i = ((long)blockIdx.x * kGenThreads + threadIdx.x) * kGenBytesPerThread;
#if B2R_STREAM_UNROLL > 1
    if (vec) {
        for (; i + (B2R_STREAM_UNROLL - 1) * step + kGenBytesPerThread <= elems; i += B2R_STREAM_UNROLL * step) {
            uint4 v[B2R_STREAM_UNROLL];
#pragma unroll
            for (int u = 0; u < B2R_STREAM_UNROLL; ++u) v[u] = __ldcs(reinterpret_cast<const uint4*>(src + i + u * step));
#pragma unroll
            for (int u = 0; u < B2R_STREAM_UNROLL; ++u)
                __stcs(reinterpret_cast<uint4*>(dst + i + u * step), table_lookup16(s_lut, v[u]));
        }
    }
#endif
    for (; i < elems; i += step) {
        if (vec && i + kGenBytesPerThread <= elems) {
            *reinterpret_cast<uint4*>(dst + i) = table_lookup16(s_lut, *reinterpret_cast<const uint4*>(src + i));
        } else {
            for (long j = i; j < elems && j < i + kGenBytesPerThread; ++j) dst[j] = s_lut[src[j]];
        }
    }
}

static inline int gen_grid(long elems, int N, dim3* grid, int bytes_per_thread) {
    long blocks = (elems + (long)kGenThreads * bytes_per_thread - 1) / ((long)kGenThreads * bytes_per_thread);
    if (blocks < 1) blocks = 1;
    // enough blocks per image to fill the GPU even for N = 1, without a tail of tiny blocks
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const long cap = (8L * sms + N - 1) / N > 1 ? (8L * sms + N - 1) / N : 1;
    if (blocks > cap) blocks = cap;
    *grid = dim3((unsigned)blocks, (unsigned)N, 1);
    return B2R_OK;
}

}  // namespace b2r
