// First conv layer (C_in = 3 -> 64) on tcgen05: im2col built on the fly by producer warps.
//
// K = 27 is far too small for TMA-staged k-blocks, and on CUDA cores the layer costs 1728 FMA per pixel (it was 11 %
// of the whole pipeline, profiles/r01_launches_v1_summary.md).  Here producer warps (groups of four, one thread per
// output pixel of the 8 x 16 tile) fetch the tile's 10 x 18 x 3 input halo once (<= 5 values per thread, prefetched a
// tile ahead), convert it and park it in shared memory; each thread then reads the 27 values of its pixel from there
// (no per-tap address arithmetic or bounds checks) and writes one 128-byte row of the A operand directly in
// the SWIZZLE_128B K-major layout the MMA reads.  K is laid out as 27 (hi, lo) bf16 pairs: x = hi + lo carries ~16
// mantissa bits of the fp32 activation, and the packed weight matrix repeats each weight for both halves, so the
// activation side of this layer stays near fp32 accuracy at no cost (K = 54 <= 64, four MMAs of N = 64 per tile).
// For u8 input the ToTensor / Normalize arithmetic (u8/255, then (x - mean)/std, IEEE division as PyTorch does on the
// CPU) is evaluated once per block into a 3 x 256 table of packed (hi, lo) pairs, so the hand-off is bit-faithful to
// the reference's preprocessing (17_run_unified_inference.py:66, 18_test_unified_benchmark.py:28-32) and costs one
// shared-memory load per value.  Zero padding is applied after the normalisation, as nn.Conv2d does.
// Warps 0..7 are the producers; warp 8 loads the 8 KB weight tile once, allocates TMEM and issues the MMAs; warps 9..12 run the usual epilogue
// (bias, ReLU / PReLU, bf16, swizzled staging, TMA store).  Persistent, one CTA per SM.
#include <cstring>

#include "b2r_internal.h"
#include "conv_common.cuh"
#include "ptx_sm100.cuh"

namespace b2r {

constexpr int kC3ProducerWarps = 8;   // two groups of 128 threads build alternate tiles (hides the gather latency)
constexpr int kC3EpiWarps = 8;        // two warps per TMEM lane quarter, 32 channels each
constexpr int kC3Threads = (kC3ProducerWarps + 1 + kC3EpiWarps) * 32;
constexpr int kC3Stages = 4;
constexpr int kC3AccStages = 4;       // TMEM accumulator stages (4 x 64 columns)
constexpr int kC3HaloRS = 80;         // words per halo row (54 used): RS % 32 == 16 keeps a warp's two pixel rows on disjoint banks
constexpr int kC3HaloWords = 10 * kC3HaloRS;
constexpr int kC3HaloPerThread = 5;   // ceil(10 * 18 * 3 / 128)
constexpr size_t kC3Smem = 1024 + kC3Stages * 16384 + 8192 + kC3EpiWarps * 4096 + 768 * 4 + 4 * kC3HaloWords * 4 + 256 + 256;

struct alignas(64) ConvC3Params {
    CUtensorMap b_map;    // packed weights bf16 [64][64], box 64 x 64
    CUtensorMap out_map;  // NHWC bf16 [N,H,W,64], box 64 x 16 x 2 x 1
    const void* in;
    const float* bias;
    float mean[3], stdv[3];
    int normalize;
    float slope;
    int act;
    int N, H, W;
    int tiles_w, tiles_h;
    long long* dbg;
};

static long long* g_c3_dbg = nullptr;

__device__ __forceinline__ uint32_t split_hi_lo(float x) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    return uint32_t(__bfloat16_as_ushort(hi)) | (uint32_t(__bfloat16_as_ushort(lo)) << 16);
}

template <int IN_FMT>
__global__ void __launch_bounds__(kC3Threads, 1) conv_c3_kernel(const __grid_constant__ ConvC3Params p) {
    constexpr uint32_t kIdesc = make_idesc_bf16_f32(128, 64);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_st = smem;                               // kC3Stages x 16 KB
    uint8_t* b_s = a_st + kC3Stages * 16384;            // 8 KB
    uint8_t* sfull = b_s + 8192;                        // kC3EpiWarps x 4 KB staging slabs (32 pixels x 64 ch each)
    uint32_t* lut = reinterpret_cast<uint32_t*>(sfull + kC3EpiWarps * 4096);  // [3][256] packed (hi, lo)
    uint32_t* halo = lut + 768;                         // [2 groups][2 buffers][10][kC3HaloRS] packed (hi, lo)
    float* bias_s = reinterpret_cast<float*>(halo + 4 * kC3HaloWords);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 64);
    uint64_t* full_bar = bars;                  // [kC3Stages], 128 producer arrivals
    uint64_t* empty_bar = bars + kC3Stages;     // [kC3Stages]
    uint64_t* tmem_full_bar = bars + 2 * kC3Stages;
    uint64_t* tmem_empty_bar = tmem_full_bar + kC3AccStages;
    uint64_t* b_full_bar = tmem_empty_bar + kC3AccStages;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(b_full_bar + 1);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int total_tiles = tiles_per_img * p.N;
    const int H = p.H, W = p.W;
    // role timeline: debug builds only (-DB2R_TIMELINE, tools/c3_timeline.py); this kernel is issue-bound, so even
    // predicated-off stamps would cost throughput
#ifdef B2R_TIMELINE
#define C3_STAMP(iter, slot) \
    do { if (p.dbg != nullptr && blockIdx.x == 0 && (iter) < B2R_DBG_TILES) p.dbg[(iter) * 8 + (slot)] = clock64(); } while (0)
#else
#define C3_STAMP(iter, slot) do { } while (0)
#endif

    if (warp_idx == kC3ProducerWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&p.b_map);
            tma_prefetch_desc(&p.out_map);
            for (int s = 0; s < kC3Stages; ++s) {
                mbar_init(&full_bar[s], 128);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < kC3AccStages; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);   // the four warps of one epilogue group
            }
            mbar_init(b_full_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<kC3AccStages * 64>(tmem_ptr_s);
    }
    if (IN_FMT == B2R_IN_U8_NHWC) {
        for (int i = threadIdx.x; i < 768; i += kC3Threads) {
            const int c = i >> 8, u = i & 255;
            float v = __fdiv_rn(float(u), 255.0f);                                   // ToTensor
            if (p.normalize) v = __fdiv_rn(v - p.mean[c], p.stdv[c]);                 // Normalize
            lut[i] = split_hi_lo(v);
        }
    }
    if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias[threadIdx.x];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp_idx < kC3ProducerWarps) {
        // ===================================== im2col producers =====================================
        const int group = warp_idx >> 2;      // 0 / 1: this group builds tile iterations group, group + 2, ...
        const int r = threadIdx.x & 127;      // tile row == pixel (h = r / 16, w = r % 16)
        const int ph = r >> 4, pw = r & 15;
        // The 10 x 18 x 3 input halo of a tile (540 values) is fetched ONCE by the group (<= 5 values per thread,
        // consecutive threads on consecutive addresses), converted to packed (hi, lo) words and parked in shared
        // memory; every thread then reads the 27 words of its pixel from there.  The per-thread element assignment
        // does not depend on the tile, so its decomposition is done once, here.
        uint32_t e_meta[kC3HaloPerThread];    // row | px << 8 | c << 16 | valid << 24
        uint32_t e_rel[kC3HaloPerThread];     // element offset from the halo's top-left input element
#pragma unroll
        for (int j = 0; j < kC3HaloPerThread; ++j) {
            const int e = r + 128 * j;
            int row, px, c;
            if (IN_FMT == B2R_IN_U8_NHWC) {
                row = e / 54;
                px = (e % 54) / 3;
                c = e % 3;
                e_rel[j] = uint32_t(row * W * 3 + px * 3 + c);
            } else {
                c = e / 180;
                row = (e % 180) / 18;
                px = e % 18;
                e_rel[j] = uint32_t((c * H + row) * W + px);
            }
            e_meta[j] = uint32_t(row) | uint32_t(px) << 8 | uint32_t(c) << 16 | uint32_t(e < 540) << 24;
        }
        const uint32_t halo_base = smem_u32(halo) + uint32_t(group * 2 * kC3HaloWords * 4);
        const uint32_t lut_addr = smem_u32(lut);
        uint32_t raw[kC3HaloPerThread];

        const long stride = 2L * gridDim.x;
        long tile = (long)blockIdx.x + (long)group * gridDim.x;
        TileWalk ti;   // the tile whose halo is fetched next
        ti.init(tile, stride, p.tiles_w, p.tiles_h);

        auto gather = [&]() {
            const int n0 = ti.n;
            const int w0 = ti.tw * 16 - 1;
            const int h0 = ti.th * 8 - 1;
            const long base = (IN_FMT == B2R_IN_U8_NHWC) ? ((long(n0) * H + h0) * W + w0) * 3
                                                          : (long(n0) * 3 * H + h0) * W + w0;
#pragma unroll
            for (int j = 0; j < kC3HaloPerThread; ++j) {
                const int hh = h0 + int(e_meta[j] & 0xFFu);
                const int ww = w0 + int((e_meta[j] >> 8) & 0xFFu);
                const bool ok = (e_meta[j] >> 24) && hh >= 0 && hh < H && ww >= 0 && ww < W;
                if (IN_FMT == B2R_IN_U8_NHWC) {
                    // raw byte now, table lookup later (0x100 marks padding)
                    raw[j] = ok ? uint32_t(__ldg(static_cast<const uint8_t*>(p.in) + base + e_rel[j])) : 0x100u;
                } else {
                    raw[j] = ok ? __float_as_uint(__ldg(static_cast<const float*>(p.in) + base + e_rel[j])) : 0u;
                }
            }
        };

        int it = group;
        uint32_t par = 0;
        if (tile < total_tiles) gather();
        for (; tile < total_tiles; tile += stride, it += 2, par ^= 1u) {
            // halo words of this tile -> shared memory.  Two buffers per group: a thread can only be here after the
            // group barrier of the previous tile, i.e. after every thread finished reading the tile before that.
            const uint32_t hb = halo_base + par * uint32_t(kC3HaloWords * 4);
#pragma unroll
            for (int j = 0; j < kC3HaloPerThread; ++j) {
                uint32_t word;
                if (IN_FMT == B2R_IN_U8_NHWC) {
                    // explicit ld.shared (a generic pointer into shared memory would compile to LD.E)
                    uint32_t e;
                    asm volatile("ld.shared.b32 %0, [%1];"
                                 : "=r"(e)
                                 : "r"(lut_addr + (((e_meta[j] >> 16) & 0xFFu) * 256u + (raw[j] & 0xFFu)) * 4u));
                    word = (raw[j] & 0x100u) ? 0u : e;
                } else {
                    word = split_hi_lo(__uint_as_float(raw[j]));
                }
                const uint32_t widx = (e_meta[j] & 0xFFu) * uint32_t(kC3HaloRS) + ((e_meta[j] >> 8) & 0xFFu) * 3u +
                                      ((e_meta[j] >> 16) & 0xFFu);
                if (e_meta[j] >> 24) asm volatile("st.shared.b32 [%0], %1;" ::"r"(hb + widx * 4u), "r"(word) : "memory");
            }
            const long next = tile + stride;
            ti.next(p.tiles_w, p.tiles_h);
            if (next < total_tiles) gather();   // these loads complete while the row is built
            named_barrier_sync(2 + group, 128);
#if defined(B2R_EXP_C3_HALF_TAPS)
            uint32_t wrd[32] = {0u};
#else
            uint32_t wrd[32];
#endif
            const uint32_t px_addr = hb + uint32_t((ph * kC3HaloRS + pw * 3) * 4);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
#if defined(B2R_EXP_C3_HALF_TAPS)   // experiment: half of the tap reads (as if two bf16 values came per word).  Timing only.
                for (int i = 0; i < 5; ++i)
#else
                for (int i = 0; i < 9; ++i)   // i = kw * 3 + c: nine consecutive words of halo row ph + kh
#endif
                    asm volatile("ld.shared.b32 %0, [%1];"
                                 : "=r"(wrd[kh * 9 + i])
                                 : "r"(px_addr + uint32_t((kh * kC3HaloRS + i) * 4)));
#pragma unroll
            for (int i = 27; i < 32; ++i) wrd[i] = 0u;
            const int stage = it % kC3Stages;
            const uint32_t phase = uint32_t(it / kC3Stages) & 1u;
            if (r == 0) C3_STAMP(it, 0);
            mbar_wait_warp(&empty_bar[stage], phase ^ 1);
            if (r == 0) C3_STAMP(it, 1);
            const uint32_t row_addr = smem_u32(a_st + stage * 16384) + uint32_t(r * 128);
#if defined(B2R_EXP_C3_HALF_ROW)   // experiment builds only (tools/exp/README.md): what would a 64-byte A row (K = 32) buy?  Timing only.
            constexpr int kRowChunks = 4;
#else
            constexpr int kRowChunks = 8;
#endif
#pragma unroll
            for (int j = 0; j < kRowChunks; ++j) {
                const uint32_t addr = row_addr + uint32_t((j ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wrd[4 * j]), "r"(wrd[4 * j + 1]),
                             "r"(wrd[4 * j + 2]), "r"(wrd[4 * j + 3])
                             : "memory");
            }
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
            mbar_arrive(&full_bar[stage]);
            if (r == 0) C3_STAMP(it, 2);
        }
    } else if (warp_idx == kC3ProducerWarps) {
        // ===================================== weight load + MMA issuer =====================================
        if (lane == 0) {
            mbar_arrive_expect_tx(b_full_bar, 8192);
            tma_load_2d(b_s, &p.b_map, b_full_bar, 0, 0);
            mbar_wait(b_full_bar, 0);
            tc_fence_after();
            const uint64_t bdesc = make_sdesc_sw128(smem_u32(b_s), 1024);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                C3_STAMP(it, 3);
                const uint64_t adesc = make_sdesc_sw128(smem_u32(a_st + stage * 16384), 1024);
                const uint32_t tmem_d = tmem_base + uint32_t(acc * 64);
#if defined(B2R_EXP_C3_HALF_ROW)
                constexpr int kSteps = 2;
#else
                constexpr int kSteps = 4;
#endif
#pragma unroll
                for (int k = 0; k < kSteps; ++k)
                    umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, k > 0 ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                umma_commit(&tmem_full_bar[acc]);
                if (++stage == kC3Stages) {
                    stage = 0;
                    phase ^= 1;
                }
                if (++acc == kC3AccStages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================================== epilogue =====================================
        // Every warp is on its own: it drains the 32 TMEM lanes (= 32 pixels = two tile rows) it may access, all 64
        // channels, stages them in a private 4 KB slab and stores that slab with its own TMA box (64 ch x 16 x 2).
        // No CTA-wide barrier anywhere; the two groups of four warps take alternate tiles (TMEM stages it % 4).
        const int quarter = warp_idx & 3;                              // TMEM lane quarter == tile rows 2q, 2q + 1
        const int group = (warp_idx - (kC3ProducerWarps + 1)) >> 2;
        const bool leader = (warp_idx == kC3ProducerWarps + 1 && lane == 0);
        const uint32_t lane_base = uint32_t(quarter * 32) << 16;
        uint8_t* slab = sfull + (group * 4 + quarter) * 4096;
        const long stride = 2L * gridDim.x;
        long tile = (long)blockIdx.x + (long)group * gridDim.x;
        TileWalk ti;
        ti.init(tile, stride, p.tiles_w, p.tiles_h);
        for (int it = group; tile < total_tiles; tile += stride, it += 2, ti.next(p.tiles_w, p.tiles_h)) {
            const int acc = it & (kC3AccStages - 1);
            const uint32_t acc_phase = uint32_t(it / kC3AccStages) & 1u;
            mbar_wait_warp(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (leader) C3_STAMP(it, 4);
            uint32_t v[32];
            float b32[32];
            tmem_ld_32x32(tmem_base + lane_base + uint32_t(acc * 64), v);
            tmem_ld_wait();
            if (lane == 0) tma_store_wait_read<0>();   // this warp's previous store has finished reading the slab
            __syncwarp();
            if (leader) C3_STAMP(it, 5);
            lds_bias32(bias_s, b32);
            epilogue_store_half(v, b32, p.act, p.slope, slab, lane, 0);
            tmem_ld_32x32(tmem_base + lane_base + uint32_t(acc * 64 + 32), v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            lds_bias32(bias_s + 32, b32);
            epilogue_store_half(v, b32, p.act, p.slope, slab, lane, 1);
            if (leader) C3_STAMP(it, 6);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_4d(&p.out_map, slab, 0, ti.tw * 16, ti.th * 8 + quarter * 2, ti.n);
                tma_store_commit();
            }
            if (leader) C3_STAMP(it, 7);
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == kC3ProducerWarps) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc<kC3AccStages * 64>(tmem_base);
    }
}

}  // namespace b2r

// debug facility of tools/c3_timeline.py (declared in csrc/b2r_debug.h, NOT part of the product header include/b2r.h)
extern "C" void b2r_debug_timeline(int64_t* device_buf) { b2r::g_c3_dbg = reinterpret_cast<long long*>(device_buf); }

extern "C" int b2r_conv3x3_c3(const void* in, int in_fmt, const float* mean_host, const float* std_host,
                              const void* weights_packed, const float* bias, int act, float slope, void* out, int N,
                              int H, int W, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(in && weights_packed && bias && out, "null pointer");
    B2R_REQUIRE(N > 0 && H > 0 && W > 0, "bad shape N=%d H=%d W=%d", N, H, W);
    B2R_REQUIRE(in_fmt == B2R_IN_F32_NCHW || in_fmt == B2R_IN_U8_NHWC, "in_fmt=%d", in_fmt);
    B2R_REQUIRE((mean_host == nullptr) == (std_host == nullptr), "mean/std must both be given or both be null");
    B2R_REQUIRE(!(mean_host && in_fmt != B2R_IN_U8_NHWC), "normalisation is only defined for the u8 hand-off");
    B2R_REQUIRE(act >= B2R_ACT_NONE && act <= B2R_ACT_PRELU, "act=%d", act);
    static thread_local ConvC3Params P;
    memset(&P, 0, sizeof(P));
    {
        const uint64_t dims[2] = {64, 64};
        const uint64_t strides[1] = {128};
        const uint32_t box[2] = {64, 64};
        int rc = encode_tmap_bf16(&P.b_map, weights_packed, 2, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint64_t strides[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
        const uint32_t box[4] = {64, 16, 2, 1};   // one epilogue warp's slab: 32 pixels = two tile rows
        int rc = encode_tmap_bf16(&P.out_map, out, 4, dims, strides, box);
        if (rc) return rc;
    }
    P.in = in;
    P.dbg = g_c3_dbg;
    P.bias = bias;
    P.normalize = mean_host != nullptr;
    for (int c = 0; c < 3; ++c) {
        P.mean[c] = mean_host ? mean_host[c] : 0.f;
        P.stdv[c] = std_host ? std_host[c] : 1.f;
    }
    P.slope = slope;
    P.act = act;
    P.N = N;
    P.H = H;
    P.W = W;
    P.tiles_w = (W + 15) / 16;
    P.tiles_h = (H + 7) / 8;
    const long total_tiles = (long)P.tiles_w * P.tiles_h * N;
    B2R_REQUIRE(total_tiles < (1L << 31), "too many tiles");
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const int grid = (int)(total_tiles < sms ? total_tiles : sms);
    const size_t smem = kC3Smem;
    static bool attr_set[64][2] = {{false}};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (in_fmt == B2R_IN_U8_NHWC) {
        if (dev >= 64 || !attr_set[dev][0]) {
            B2R_CUDA(cudaFuncSetAttribute(conv_c3_kernel<B2R_IN_U8_NHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
            if (dev < 64) attr_set[dev][0] = true;
        }
        conv_c3_kernel<B2R_IN_U8_NHWC><<<grid, kC3Threads, smem, stream>>>(P);
    } else {
        if (dev >= 64 || !attr_set[dev][1]) {
            B2R_CUDA(cudaFuncSetAttribute(conv_c3_kernel<B2R_IN_F32_NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
            if (dev < 64) attr_set[dev][1] = true;
        }
        conv_c3_kernel<B2R_IN_F32_NCHW><<<grid, kC3Threads, smem, stream>>>(P);
    }
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}
