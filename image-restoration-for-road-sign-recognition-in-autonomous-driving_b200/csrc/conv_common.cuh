// Shared between conv_gemm.cu (generic) and conv_n64.cu (C_out = 64): epilogue helpers and parameter blocks.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/b2r.h"
#include "ptx_sm100.cuh"

namespace b2r {

constexpr int kEpiThreadsC = 128;
constexpr int kN64MaxRing = 8;
constexpr int kN64MaxSlots = 40;
constexpr int kN64MaxSmem = 227 * 1024;

// conv_n64 ring-slot encoding: [0,2) source | [2] centre (1 tap) | [4,6) dw+1 | [8,20) channel chunk | [20,32) first k-block
struct alignas(64) ConvN64Params {
    CUtensorMap a3_map[B2R_MAX_SRC];  // box 64 ch x TW x (TH+2) x 1
    CUtensorMap a1_map[B2R_MAX_SRC];  // box 64 ch x TW x TH x 1
    CUtensorMap b_map;                // weights [64][K], box 64 x 64
    CUtensorMap out_map, pool_map;
    const float* bias;
    float slope;
    int act;
    int num_slots, num_kblocks;
    int ring_slots, slot_bytes;
    int tiles_w, tiles_h, n_img;
    int tile_w, tile_h;
    int store_full, store_pool;
    uint32_t slot[kN64MaxSlots];
};

int launch_conv_n64(const ConvN64Params& p, int grid, size_t smem_bytes, cudaStream_t stream);

// conv_w3 group encoding: [0,2) source | [2] centre (one k-step, kernel row 1) | [8,20) channel chunk | [20,32) first k-step
constexpr int kW3MaxGroups = 24;
constexpr int kW3MaxStageBufs = 4;
struct alignas(64) ConvW3Params {
    CUtensorMap a_map[B2R_MAX_SRC];  // box 64 ch x 16 x 10 x 1
    CUtensorMap b_map;               // wide weights [192][64 * num_ksteps], box 64 x 192
    CUtensorMap out_map, pool_map;   // boxes 64 x 14 x 8 x 1 and 64 x 7 x 4 x 1
    const float* bias;
    float slope;
    int act;
    int num_groups, num_ksteps, ring_slots;
    int b_bytes;                     // shared memory of the weights region (resident blocks, or b_slots x 24 KB)
    int stage_stride;                // bytes per staging buffer
    int prefetch;                    // 1: L2-prefetch the boxes of the tile four steps ahead (only pays when the ring holds < 2 tiles)
    int stage_bufs;                  // 2..kW3MaxStageBufs staging buffers, used round-robin (decouples the epilogue from slow TMA stores)
    CUtensorMap b_map_c;             // box 64 x 64: the kw = 1 rows of a 1x1 k-step (compact resident layout)
    uint32_t group_boff[kW3MaxGroups];   // resident mode: shared-memory offset of each group's weights, in 16-byte units
    int b_slots;                     // 0: weights resident (num_ksteps x 24 KB); > 0: weights streamed through this many slots
    int tiles_w, tiles_h, n_img;
    int store_full, store_pool;
    long long* dbg;                  // optional [B2R_DBG_TILES][8] clock64 stamps written by CTA 0
    const float* head_w;             // fused 64 -> 3 head (nullptr: none)
    const float* head_b;
    float* head_f32;
    uint8_t* head_u8;
    int H, W;
    uint32_t group[kW3MaxGroups];
};

size_t conv_w3_smem_bytes(size_t b_bytes, int ring_slots, size_t stage_stride, int stage_bufs);
int launch_conv_w3(const ConvW3Params& p, int grid, cudaStream_t stream, bool pair);

#ifdef __CUDACC__
// Walks tiles first, first + stride, ... as (image, tile row, tile column) without per-tile integer divisions: the two
// divisions per tile were ~350 cycles of dependent latency, and the compiler sinks them into the single thread that
// issues the TMA store, i.e. onto the epilogue's critical path (profiles/r01_c3_timeline.md).
struct TileWalk {
    int n, th, tw;
    int dn, dth, dtw;
    __device__ __forceinline__ void init(long first, long stride, int tiles_w, int tiles_h) {
        const long per_img = long(tiles_w) * tiles_h;
        n = int(first / per_img);
        int t = int(first - n * per_img);
        th = t / tiles_w;
        tw = t - th * tiles_w;
        dn = int(stride / per_img);
        t = int(stride - dn * per_img);
        dth = t / tiles_w;
        dtw = t - dth * tiles_w;
    }
    __device__ __forceinline__ void next(int tiles_w, int tiles_h) {
        tw += dtw;
        th += dth;
        n += dn;
        if (tw >= tiles_w) {
            tw -= tiles_w;
            ++th;
        }
        if (th >= tiles_h) {
            th -= tiles_h;
            ++n;
        }
    }
};

// All three activations are y = max(x, 0) + ns * min(x, 0) with ns = 0 (ReLU), slope (PReLU), 1 (none): branch-free and
// exact in each case (one of the two terms is always a signed zero).
__device__ __forceinline__ float act_neg_slope(int act, float slope) {
    return act == B2R_ACT_RELU ? 0.f : (act == B2R_ACT_PRELU ? slope : 1.f);
}
__device__ __forceinline__ float apply_act_ns(float x, float ns) { return fmaf(ns, fminf(x, 0.f), fmaxf(x, 0.f)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162 r = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&r);
}

// 32 fp32 accumulator columns of this thread's pixel row -> +bias, activation, bf16 -> four 16-byte chunks of the
// 128-byte staging row, written with the 128B TMA swizzle (chunk index XOR row & 7).
// 32 bias values from shared memory into registers with explicit ld.shared (a `const float*` into shared memory is a
// GENERIC pointer: the compiler emits LD.E through the L1TEX path, ~10x the latency of LDS; profiles/r01_c3_epilogue.md)
__device__ __forceinline__ void lds_bias32(const float* bias_smem, float (&b)[32]) {
    const uint32_t a = smem_u32(bias_smem);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(b[4 * i]), "=f"(b[4 * i + 1]), "=f"(b[4 * i + 2]), "=f"(b[4 * i + 3])
                     : "r"(a + 16 * i));
}

// Activation forms, chosen once per launch (warp-uniform).  PReLU with 0 <= slope <= 1 (every trained slope we have seen; nn.PReLU
// starts at 0.25) is max(x, slope * x): two ops instead of three and the same single rounding of slope * x as the general form;
// no activation is a plain pack.  (Only an exact -0 accumulator can differ from the general form, as +0 vs -0.)
enum : int { kActNone = 0, kActRelu = 1, kActPrelu01 = 2, kActGeneral = 3 };
__device__ __forceinline__ int act_form(int act, float slope) {
    if (act == B2R_ACT_RELU) return kActRelu;
    if (act == B2R_ACT_PRELU) return (slope >= 0.f && slope <= 1.f) ? kActPrelu01 : kActGeneral;
    return kActNone;
}

template <int FORM>
__device__ __forceinline__ void epilogue_store_half_form(const uint32_t (&v)[32], const float (&bias32)[32], float ns,
                                                         uint8_t* sfull, int row, int half) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = q * 8 + e * 2;
            const float y0 = __uint_as_float(v[j]) + bias32[j], y1 = __uint_as_float(v[j + 1]) + bias32[j + 1];
            if (FORM == kActRelu)           // rounding is monotonic and max(+0, -0) = +0, so this is
                o[e] = bf16x2_max(pack_bf16x2(y0, y1), 0u);   // bit-identical to rounding max(y, 0)
            else if (FORM == kActNone)
                o[e] = pack_bf16x2(y0, y1);
            else if (FORM == kActPrelu01)
                o[e] = pack_bf16x2(fmaxf(y0, ns * y0), fmaxf(y1, ns * y1));
            else
                o[e] = pack_bf16x2(apply_act_ns(y0, ns), apply_act_ns(y1, ns));
        }
        const int jj = half * 4 + q;
        const uint32_t addr = smem_u32(sfull) + uint32_t(row * 128 + ((jj ^ (row & 7)) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                     : "memory");
    }
}

__device__ __forceinline__ void epilogue_store_half(const uint32_t (&v)[32], const float (&bias32)[32], int act,
                                                    float slope, uint8_t* sfull, int row, int half) {
    const float ns = act_neg_slope(act, slope);
    switch (act_form(act, slope)) {   // warp-uniform
        case kActRelu: epilogue_store_half_form<kActRelu>(v, bias32, ns, sfull, row, half); break;
        case kActNone: epilogue_store_half_form<kActNone>(v, bias32, ns, sfull, row, half); break;
        case kActPrelu01: epilogue_store_half_form<kActPrelu01>(v, bias32, ns, sfull, row, half); break;
        default: epilogue_store_half_form<kActGeneral>(v, bias32, ns, sfull, row, half); break;
    }
}

// 2x2 max-pool of a staged 128-pixel x 64-channel tile (tw x th x tn pixels) into the 32-pixel pooled staging tile.
// Thread -> pooled pixel tid/4, 16 channels (two 16-byte chunks).
__device__ __forceinline__ void epilogue_pool_chunk(const uint8_t* sfull, uint8_t* spool, int epi_tid, int tw, int th) {
    const int rp = epi_tid >> 2;
    const int cg = epi_tid & 3;
    const int pw = tw >> 1, ph = th >> 1;
    const int wp = rp % pw;
    const int hp = (rp / pw) % ph;
    const int nl = rp / (pw * ph);
    const int r00 = (nl * th + 2 * hp) * tw + 2 * wp;
    const int rr[4] = {r00, r00 + 1, r00 + tw, r00 + tw + 1};
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int jj = cg * 2 + cc;
        uint32_t mx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t a = smem_u32(sfull) + uint32_t(rr[k] * 128 + ((jj ^ (rr[k] & 7)) << 4));
            uint32_t t0, t1, t2, t3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3) : "r"(a));
            if (k == 0) {
                mx[0] = t0; mx[1] = t1; mx[2] = t2; mx[3] = t3;
            } else {
                mx[0] = bf16x2_max(mx[0], t0);
                mx[1] = bf16x2_max(mx[1], t1);
                mx[2] = bf16x2_max(mx[2], t2);
                mx[3] = bf16x2_max(mx[3], t3);
            }
        }
        const uint32_t d = smem_u32(spool) + uint32_t(rp * 128 + ((jj ^ (rp & 7)) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(mx[0]), "r"(mx[1]), "r"(mx[2]), "r"(mx[3])
                     : "memory");
    }
}
#endif  // __CUDACC__

}  // namespace b2r
