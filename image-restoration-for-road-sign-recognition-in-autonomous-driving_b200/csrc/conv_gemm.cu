// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   M = output pixels (128 per tile: a TW x TH x TN box of the NHWC pixel grid)
//   N = output channels (BLOCK_N = 64 / 128 / 256 per tile)
//   K = list of "k-blocks": 64 input channels of one source tensor at one spatial offset (dh, dw)
//
// A k-block's A operand is ONE 4-D TMA box load (64 ch x TW x TH x TN) at coordinates shifted by (dw, dh):
// out-of-bounds rows/columns are zero-filled by the TMA unit, which is exactly conv padding=1, and the box lands
// in shared memory as 128 rows x 128 B with the 128-byte swizzle — the canonical K-major UMMA operand layout.
// The B operand is a 2-D TMA box (64 x BLOCK_N) of the packed weight matrix.  Accumulators live in TMEM
// (2 stages x BLOCK_N fp32 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/activation -> bf16 -> swizzled smem -> TMA store, plus the
// fused 2x2 max-pool computed from the staged tile).  Persistent: one CTA per SM walks tiles round-robin.
//
// Reference semantics replaced: nn.Conv2d(3x3, pad 1) / ConvTranspose2d(2,2) / BatchNorm2d(eval) / ReLU / PReLU /
// MaxPool2d(2,2) / torch.cat / the ResidualBlock add, as used by SimpleUNet.forward (07_train_restoration.py:99-120),
// ResUNet.forward (14_train_unified_advanced.py:151-186) and torchvision VGG16 (18_test_unified_benchmark.py:46).
#include <cstdlib>
#include <cstring>

#include "b2r_internal.h"
#include "conv_common.cuh"
#include "ptx_sm100.cuh"

namespace b2r {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kStagingFull = 16384;                  // 128 px x 64 ch bf16
constexpr int kStagingPool = 4096;                   // 32 px x 64 ch bf16
constexpr int kNumThreads = 192;
constexpr int kEpiThreads = 128;

struct alignas(64) ConvGemmParams {
    CUtensorMap a_map[B2R_MAX_SRC];
    CUtensorMap b_map;
    CUtensorMap out_map[4];
    CUtensorMap pool_map;
    const float* bias;
    float slope;
    int act;
    int num_kblocks;
    int linear_k;
    int tiles_w, tiles_h, tiles_n;
    int tile_w, tile_h, tile_n;
    int n_tiles;          // cout_total / BLOCK_N
    int cout_per_out;     // channels per output map (CONVT2X2: C_out, four maps; NHWC: cout_total, one map)
    int store_full, store_pool;
    uint32_t kblk[B2R_MAX_KBLOCKS];
    // halo mode (conv_gemm_halo_kernel): one (TH + 2)-row input box per (source, 64-channel chunk, dw) feeds the three
    // kernel rows; slot encoding as in conv_common.cuh (ConvN64Params)
    CUtensorMap a3_map[B2R_MAX_SRC];
    int num_slots, a_slot_bytes, a_slots, b_slots;
    uint32_t slot[kN64MaxSlots];
};

template <int BLOCK_N, bool kDeep = false>
struct GemmCfg {
    static constexpr int kBStageBytes = BLOCK_N * 128;
    static constexpr int kStageBytes = kAStageBytes + kBStageBytes;
    // BLOCK_N = 128: two co-resident CTAs per SM (2 x 256 TMEM columns, 2 x 106 KB of shared memory) with two stages
    // each beat one CTA with five stages by 5-12 % (A/B in one gpurun call): while one CTA's roles are handing off
    // (commit -> epilogue wake -> release -> MMA wake), the other CTA keeps the tensor pipe busy.
    static constexpr int kStages = BLOCK_N == 64 ? 6 : (BLOCK_N == 128 ? 2 : 3);
    static constexpr int kCtasPerSm = BLOCK_N == 128 ? 2 : 1;
    static constexpr int kTmemCols = 2 * BLOCK_N;  // 128 / 256 / 512: powers of two
    // Two epilogue warpgroups (warps 2..5 and 6..9) take alternate 64-column chunks of a tile, each with its own staging
    // buffer(s), bias copy, named barrier and TMA-store queue.  A lone warp per scheduler cannot hide tcgen05.ld / LDS / STS
    // latency, and with short K (ConvTranspose as four 1x1 taps: ONE k-block per 128 x 256 tile) nothing else hides the
    // epilogue either: measured on up1 (64 -> 4 x 64 @112, 256 images) epilogue maths alone 413 us, stores alone 374 us,
    // together 532 us with one group (profiles/r02_convt_epilogue.md).
    // kDeep (short-K layers without a fused pool, i.e. the ConvTranspose layers): the pooled staging is dropped and each
    // group gets TWO 16 KB buffers, so a group converts its next chunk while the store of the previous one drains.
    static constexpr int kEpiGroups = BLOCK_N >= 128 ? 2 : 1;
    static constexpr int kGroupBufs = kDeep ? 2 : 1;
    static constexpr int kBufStride = kDeep ? kStagingFull : kStagingFull + kStagingPool;
    static constexpr int kStagingBytes = kDeep ? 4 * kStagingFull : 2 * (kStagingFull + kStagingPool);
    static constexpr int kThreads = 64 + 128 * kEpiGroups;
    static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kStagingBytes +
                                      kEpiGroups * BLOCK_N * 4 /*bias*/ + 256 /*barriers + tmem ptr*/;
    static_assert(!kDeep || kEpiGroups == 2, "deep staging is for the two-group epilogue");
};

// Which tiles a CTA's epilogue walks, and where it hands the accumulator stage back.
struct SchedSingle {   // one CTA per tile, round-robin over the grid
    __device__ static int first() { return blockIdx.x; }
    __device__ static int stride() { return gridDim.x; }
    __device__ static void decode(const ConvGemmParams& p, int unit, int* n_tile, int* m, bool* valid) {
        *n_tile = unit % p.n_tiles;
        *m = unit / p.n_tiles;
        *valid = true;
    }
    __device__ static void release(uint64_t* bar) { mbar_arrive(bar); }
};
struct SchedPair {     // cta_group::2: a cluster walks PAIRS of pixel tiles, rank r takes tile 2 * pair + r of the same n-tile
    __device__ static int first() { return blockIdx.x >> 1; }
    __device__ static int stride() { return gridDim.x >> 1; }
    __device__ static void decode(const ConvGemmParams& p, int unit, int* n_tile, int* m, bool* valid) {
        const int num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
        *n_tile = unit % p.n_tiles;
        *m = 2 * (unit / p.n_tiles) + int(cluster_ctarank());
        *valid = *m < num_m_tiles;   // odd tile count: the last pair's second CTA computes a duplicate and stores nothing
    }
    __device__ static void release(uint64_t* bar) { mbar_arrive_leader(bar); }   // the leader's barrier, from either CTA
};

// Epilogue role shared by the generic, halo and pair kernels: warps 2..5, TMEM -> registers -> bias / activation -> bf16
// -> swizzled staging -> TMA store (+ fused 2x2 max-pool).  BUFS = number of (staging + pool) buffers; `total_units` =
// tiles (SchedSingle) or tile pairs (SchedPair).
// Named barrier of one epilogue group's 128 threads: ids 1 / 2 as immediates (a register id makes ptxas reserve all 16).
template <int GROUPS>
__device__ __forceinline__ void group_barrier(int grp) {
    if (GROUPS == 1 || grp == 0) named_barrier_sync(1, kEpiThreads);
    else named_barrier_sync(2, kEpiThreads);
}

template <int BLOCK_N, int BUFS, class Sched = SchedSingle, int GROUPS = 1, int GBUFS = 1, int BUF_STRIDE = kStagingFull + kStagingPool>
__device__ __forceinline__ void conv_epilogue_role(const ConvGemmParams& p, uint32_t tmem_base, uint8_t* staging, float* bias_s,
                                                   uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, int total_units,
                                                   int warp_idx, int lane) {
        static_assert(GROUPS == 1 || (GROUPS == 2 && BUFS == 2), "two epilogue groups own one staging buffer each");
        constexpr int kChunks = BLOCK_N / 64;
        const int grp = GROUPS == 2 ? ((warp_idx - 2) >> 2) : 0;   // warps 2..5 / 6..9
        const int quarter = warp_idx & 3;          // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;       // tile row (pixel) == TMEM lane
        const int epi_tid = row;                   // 0..127 inside the group
        const uint32_t lane_base = uint32_t(quarter * 32) << 16;
        if (GROUPS == 2) bias_s += grp * BLOCK_N;  // private copy: no cross-group hazard when the groups are a tile apart
        const bool idle = grp >= kChunks;          // (never true for the instantiations in use)
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t chunk_counter = 0;
        const int tw = p.tile_w, th = p.tile_h;
        for (int unit = Sched::first(); unit < total_units; unit += Sched::stride()) {
            int n_tile, m;
            bool valid_tile;
            Sched::decode(p, unit, &n_tile, &m, &valid_tile);
            const int w0 = (m % p.tiles_w) * p.tile_w;
            const int h0 = ((m / p.tiles_w) % p.tiles_h) * p.tile_h;
            const int n0 = (m / (p.tiles_w * p.tiles_h)) * p.tile_n;
            // first channel of this group's first chunk inside its output map (CONVT2X2: an n-tile may span several of
            // the four maps, e.g. C_out = 64 with BLOCK_N = 256 covers all four taps in one tile)
            int out_idx = (n_tile * BLOCK_N + grp * 64) / p.cout_per_out;
            int ch_in_out = n_tile * BLOCK_N + grp * 64 - out_idx * p.cout_per_out;

            // bias of this n-tile -> smem (previous tile's readers are past their last named barrier)
            for (int i = epi_tid; i < BLOCK_N; i += kEpiThreads) bias_s[i] = p.bias[n_tile * BLOCK_N + i];

            mbar_wait_warp(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();

#pragma unroll 1
            for (int c = grp; c < kChunks; c += GROUPS) {
                const int buf = GROUPS == 2 ? grp * GBUFS + (GBUFS == 2 ? int(chunk_counter & 1) : 0)
                                            : (BUFS == 2 ? int(chunk_counter & 1) : 0);
                ++chunk_counter;
                uint8_t* sfull = staging + buf * BUF_STRIDE;
                uint8_t* spool = sfull + kStagingFull;   // (unused when the stride leaves no pooled staging: store_pool == 0)
                if (epi_tid == 0) {   // the store that last read `buf` has drained (bulk groups are per issuing thread)
                    if ((GROUPS == 1 && BUFS == 2) || (GROUPS == 2 && GBUFS == 2)) tma_store_wait_read<1>();
                    else tma_store_wait_read<0>();
                }
                group_barrier<GROUPS>(grp);

#if !defined(B2R_EXP_EPI_NO_MATH)   // experiment builds only (tools/exp/README.md): what bounds a short-K layer's epilogue?
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + lane_base + uint32_t(acc * BLOCK_N + c * 64 + half * 32), v);
                    tmem_ld_wait();
                    float b32[32];
                    lds_bias32(bias_s + c * 64 + half * 32, b32);
                    epilogue_store_half(v, b32, p.act, p.slope, sfull, row, half);
                }
#endif
                if (c + GROUPS >= kChunks) {
                    // all of this warp's TMEM reads of the accumulator are done -> hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) Sched::release(&tmem_empty_bar[acc]);
                }
                fence_proxy_async_smem();
                group_barrier<GROUPS>(grp);

                if (p.store_pool) {
                    epilogue_pool_chunk(sfull, spool, epi_tid, tw, th);
                    fence_proxy_async_smem();
                    group_barrier<GROUPS>(grp);
                }

                if (epi_tid == 0 && valid_tile) {
#if !defined(B2R_EXP_EPI_NO_STORE)
                    if (p.store_full) tma_store_4d(&p.out_map[out_idx], sfull, ch_in_out, w0, h0, n0);
                    if (p.store_pool) tma_store_4d(&p.pool_map, spool, ch_in_out, w0 >> 1, h0 >> 1, n0);
#endif
                    tma_store_commit();
                }
                ch_in_out += 64 * GROUPS;
                while (ch_in_out >= p.cout_per_out) {
                    ch_in_out -= p.cout_per_out;
                    ++out_idx;
                }
            }
            if (idle) {   // a group without a chunk still owes its share of the accumulator hand-back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) Sched::release(&tmem_empty_bar[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (epi_tid == 0) tma_store_wait_all<0>();
    }

template <int BLOCK_N, bool kDeep = false>
__global__ void __launch_bounds__(GemmCfg<BLOCK_N>::kThreads, GemmCfg<BLOCK_N>::kCtasPerSm) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, kDeep>;
    constexpr int kStages = Cfg::kStages;
    constexpr uint32_t kIdesc = make_idesc_bf16_f32(kBlockM, BLOCK_N);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stages = smem;
    uint8_t* staging = smem + kStages * Cfg::kStageBytes;
    float* bias_s = reinterpret_cast<float*>(staging + Cfg::kStagingBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + Cfg::kEpiGroups * BLOCK_N);
    uint64_t* full_bar = bars;                       // [kStages]
    uint64_t* empty_bar = bars + kStages;            // [kStages]
    uint64_t* tmem_full_bar = bars + 2 * kStages;    // [2]
    uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;  // [2]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int total_tiles = num_m_tiles * p.n_tiles;

    if (warp_idx == 0 && lane == 0) {
        for (int i = 0; i < B2R_MAX_SRC; ++i) tma_prefetch_desc(&p.a_map[i]);
        tma_prefetch_desc(&p.b_map);
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.out_map[i]);
        tma_prefetch_desc(&p.pool_map);
    }
    if (warp_idx == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4 * Cfg::kEpiGroups);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<Cfg::kTmemCols>(tmem_ptr_s);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                const int m = tile / p.n_tiles;
                const int w0 = (m % p.tiles_w) * p.tile_w;
                const int h0 = ((m / p.tiles_w) % p.tiles_h) * p.tile_h;
                const int n0 = (m / (p.tiles_w * p.tiles_h)) * p.tile_n;
                for (int kb = 0; kb < p.num_kblocks; ++kb) {
                    int src = 0, dh = 0, dw = 0, c0 = kb * kBlockK;
                    if (!p.linear_k) {
                        const uint32_t e = p.kblk[kb];
                        src = e & 3;
                        dh = int((e >> 2) & 3) - 1;
                        dw = int((e >> 4) & 3) - 1;
                        c0 = int(e >> 8) * kBlockK;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = stages + stage * Cfg::kStageBytes;
#if defined(B2R_EXP_GEMM_NO_B)      // experiment builds only (tools/exp/README.md): is the layer bound by smem-fill traffic?
                    mbar_arrive_expect_tx(&full_bar[stage], kAStageBytes);
                    tma_load_4d(sa, &p.a_map[src], &full_bar[stage], c0, w0 + dw, h0 + dh, n0);
#elif defined(B2R_EXP_GEMM_NO_A)
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes - kAStageBytes);
                    tma_load_2d(sa + kAStageBytes, &p.b_map, &full_bar[stage], kb * kBlockK, n_tile * BLOCK_N);
#else
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_4d(sa, &p.a_map[src], &full_bar[stage], c0, w0 + dw, h0 + dh, n0);
                    tma_load_2d(sa + kAStageBytes, &p.b_map, &full_bar[stage], kb * kBlockK, n_tile * BLOCK_N);
#endif
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * BLOCK_N);
                for (int kb = 0; kb < p.num_kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stages + stage * Cfg::kStageBytes);
                    const uint64_t adesc = make_sdesc_sw128(sa, 1024);
                    const uint64_t bdesc = make_sdesc_sw128(sa + kAStageBytes, 1024);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 32 B (16 bf16) along K inside the 128-B swizzle row: +2 in the >>4 address field
                        umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc,
                                     (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs have read it
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================================== epilogue =====================================
        conv_epilogue_role<BLOCK_N, 2, SchedSingle, Cfg::kEpiGroups, Cfg::kGroupBufs, Cfg::kBufStride>(
            p, tmem_base, staging, bias_s, tmem_full_bar, tmem_empty_bar, total_tiles, warp_idx, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Halo mode: the same GEMM with 2.4x less A traffic into shared memory.
//
// Measured (B2R_EXP_GEMM_NO_A / NO_B experiment builds, profiles/r01_smem_fill.md): the generic kernel is bound by the
// rate at which TMA can fill shared memory, ~55 B/clk per SM, not by the tensor pipe: at full MMA rate one k-block needs
// 16 KB of A + BLOCK_N x 128 B of B every 4 MMAs = 94 B/clk (N = 256) or 106 B/clk (N = 128).  Every one of the nine
// taps re-loads an A tile that is the previous one shifted by a pixel.  Here the producer loads, per (source, 64-channel
// chunk, dw), ONE box of TH + 2 rows (the conv_n64 trick): the operand of kernel row dh is the 128 rows starting at row
// (dh + 1) * TW of that box, a 1024-byte aligned descriptor start as long as TW is a multiple of 8.  A and B travel through
// separate rings (an A slot lives for three k-blocks, a B slot for one).  Everything downstream of the MMA is the
// generic epilogue.  Requirements (checked by the dispatcher): NHWC output, one image per tile, TW % 8 == 0, k-blocks in
// the (chunk, dw, dh) order packing.py emits; 1x1 centre k-blocks (ResidualBlock shortcut) ride on the same box.
// ------------------------------------------------------------------------------------------------------------
constexpr int kHaloMaxRing = 8;

template <int BLOCK_N>
struct HaloCfg {
    static constexpr int kBBytes = BLOCK_N * 128;
    static constexpr int kCtasPerSm = BLOCK_N == 128 ? 2 : 1;
    // one staging buffer: the shared memory is worth more as ring depth (bytes in flight = fill bandwidth x latency)
    static constexpr int kStagingBufs = 1;
    static constexpr int kTmemCols = 2 * BLOCK_N;
    static constexpr int kFixedBytes = 1024 + kStagingBufs * (kStagingFull + kStagingPool) + BLOCK_N * 4 + 512;
    static constexpr int kMaxSmem = (227 * 1024) / kCtasPerSm - (kCtasPerSm > 1 ? 1024 : 0);
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, HaloCfg<BLOCK_N>::kCtasPerSm) conv_gemm_halo_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = HaloCfg<BLOCK_N>;
    constexpr uint32_t kIdesc = make_idesc_bf16_f32(kBlockM, BLOCK_N);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;                                        // a_slots x a_slot_bytes (multiples of 1 KB)
    uint8_t* b_ring = a_ring + p.a_slots * p.a_slot_bytes;         // b_slots x BLOCK_N x 128 B
    uint8_t* staging = b_ring + p.b_slots * Cfg::kBBytes;
    float* bias_s = reinterpret_cast<float*>(staging + Cfg::kStagingBufs * (kStagingFull + kStagingPool));
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + BLOCK_N);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + kHaloMaxRing;
    uint64_t* b_full = bars + 2 * kHaloMaxRing;
    uint64_t* b_empty = bars + 3 * kHaloMaxRing;
    uint64_t* tmem_full_bar = bars + 4 * kHaloMaxRing;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
    const int AR = p.a_slots, BR = p.b_slots;

    if (warp_idx == 0 && lane == 0) {
        for (int i = 0; i < B2R_MAX_SRC; ++i) tma_prefetch_desc(&p.a3_map[i]);
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.out_map[0]);
        tma_prefetch_desc(&p.pool_map);
    }
    if (warp_idx == 1) {
        if (lane == 0) {
            for (int s = 0; s < kHaloMaxRing; ++s) {
                mbar_init(&a_full[s], 1);
                mbar_init(&a_empty[s], 1);
                mbar_init(&b_full[s], 1);
                mbar_init(&b_empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<Cfg::kTmemCols>(tmem_ptr_s);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                const int m = tile / p.n_tiles;
                const int w0 = (m % p.tiles_w) * p.tile_w;
                const int h0 = ((m / p.tiles_w) % p.tiles_h) * p.tile_h;
                const int n0 = m / (p.tiles_w * p.tiles_h);
                for (int s = 0; s < p.num_slots; ++s) {
                    const uint32_t e = p.slot[s];
                    const int src = e & 3, ntaps = ((e >> 2) & 1) ? 1 : 3, dw = int((e >> 4) & 3) - 1;
                    const int c0 = int((e >> 8) & 0xFFF) * 64, kb0 = int(e >> 20);
                    mbar_wait(&a_empty[as], aph ^ 1);
                    mbar_arrive_expect_tx(&a_full[as], uint32_t(p.a_slot_bytes));
                    tma_load_4d(a_ring + as * p.a_slot_bytes, &p.a3_map[src], &a_full[as], c0, w0 + dw, h0 - 1, n0);
                    if (++as == AR) {
                        as = 0;
                        aph ^= 1;
                    }
                    for (int t = 0; t < ntaps; ++t) {
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        mbar_arrive_expect_tx(&b_full[bs], Cfg::kBBytes);
                        tma_load_2d(b_ring + bs * Cfg::kBBytes, &p.b_map, &b_full[bs], (kb0 + t) * kBlockK, n_tile * BLOCK_N);
                        if (++bs == BR) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t row_stride = uint32_t(p.tile_w) * 128u;   // one kernel row further down the box
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * BLOCK_N);
                uint32_t accum = 0;
                for (int s = 0; s < p.num_slots; ++s) {
                    const uint32_t e = p.slot[s];
                    const bool centre = ((e >> 2) & 1) != 0;
                    const int ntaps = centre ? 1 : 3;
                    mbar_wait(&a_full[as], aph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(a_ring + as * p.a_slot_bytes);
                    for (int t = 0; t < ntaps; ++t) {
                        mbar_wait(&b_full[bs], bph);
                        tc_fence_after();
                        const uint64_t adesc = make_sdesc_sw128(sa + uint32_t(centre ? 1 : t) * row_stride, 1024);
                        const uint64_t bdesc = make_sdesc_sw128(smem_u32(b_ring + bs * Cfg::kBBytes), 1024);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, accum);
                            accum = 1;
                        }
                        umma_commit(&b_empty[bs]);
                        if (++bs == BR) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                    umma_commit(&a_empty[as]);
                    if (++as == AR) {
                        as = 0;
                        aph ^= 1;
                    }
                }
                umma_commit(&tmem_full_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        conv_epilogue_role<BLOCK_N, Cfg::kStagingBufs>(p, tmem_base, staging, bias_s, tmem_full_bar, tmem_empty_bar, total_tiles,
                                                        warp_idx, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pair mode (cta_group::2) for C_out % 256 == 0: two CTAs of a cluster, on the two SMs of a TPC, execute MMAs of
// M = 256 x N = 256.  Each CTA owns one 128-pixel tile (its rows of A, its rows of D in its own TMEM, its own
// epilogue and stores) and loads only HALF of each weight k-block: B is two thirds of what the generic kernel pours
// into shared memory, and shared-memory fill is what bounds it (profiles/r01_smem_fill.md).  32 KB per k-block
// and CTA instead of 48 KB, so the ring is also deeper (5 stages).
// Protocol: the leader (cluster rank 0) issues all MMAs.  Its "full" barriers count the TMA bytes of both CTAs (the
// peer's loads name the leader's barrier); tcgen05.commit multicasts to the "empty" / "accumulator ready" barriers of
// both CTAs; the peer's epilogue releases the accumulator stage on the leader's barrier.
// ------------------------------------------------------------------------------------------------------------
template <int kPairN>
struct PairCfg {
#ifndef B2R_PAIR_STAGES
#define B2R_PAIR_STAGES 5
#endif
#ifndef B2R_PAIR_BUFS
#define B2R_PAIR_BUFS 2
#endif
    // N = 128 (experiment, B2R_PAIR128=1): two co-resident clusters per SM pair, 113 KB and 256 TMEM columns per CTA
    static constexpr int kStages = kPairN == 256 ? B2R_PAIR_STAGES : 3;
    static constexpr int kBufs = kPairN == 256 ? B2R_PAIR_BUFS : 1;
    static constexpr int kCtasPerSm = kPairN == 256 ? 1 : 2;
    static constexpr int kStageBytes = kAStageBytes + (kPairN / 2) * 128;   // 16 KB of A + half a weight k-block per CTA
#ifndef B2R_PAIR_GROUPS
#define B2R_PAIR_GROUPS 1
#endif
    // two epilogue warpgroups on alternate 64-column chunks (see GemmCfg): measured 1-1.5 % SLOWER on every pair layer
    // (conv3_1 337 -> 342 us, conv4_2 621 -> 624 us at 256 images): with K >= 1152 the epilogue is already hidden and the
    // extra warps only take issue slots.  Kept as a variant-build knob (--define B2R_PAIR_GROUPS=2).
    static constexpr int kEpiGroups = (kPairN == 256 && kBufs == 2) ? B2R_PAIR_GROUPS : 1;
    static constexpr int kThreads = 64 + 128 * kEpiGroups;
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kBufs * (kStagingFull + kStagingPool) + kEpiGroups * kPairN * 4 + 256;
};

template <int kPairN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg<kPairN>::kThreads, PairCfg<kPairN>::kCtasPerSm)
conv_gemm_pair_kernel(const __grid_constant__ ConvGemmParams p) {
    constexpr int kPairStages = PairCfg<kPairN>::kStages;
    constexpr int kPairStageBytes = PairCfg<kPairN>::kStageBytes;
    constexpr uint32_t kIdesc = make_idesc_bf16_f32(256, kPairN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stages = smem;
    constexpr int kBufs = PairCfg<kPairN>::kBufs;
    uint8_t* staging = smem + kPairStages * kPairStageBytes;
    float* bias_s = reinterpret_cast<float*>(staging + kBufs * (kStagingFull + kStagingPool));
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + PairCfg<kPairN>::kEpiGroups * kPairN);
    uint64_t* full_bar = bars;                            // [stages]  used in the leader only
    uint64_t* empty_bar = bars + kPairStages;             // [stages]  one multicast commit per phase, in each CTA
    uint64_t* tmem_full_bar = bars + 2 * kPairStages;     // [2]       one multicast commit per phase, in each CTA
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2]       leader only: 4 epilogue warps x 2 CTAs
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int m_pairs = (num_m_tiles + 1) / 2;
    const int total_pairs = m_pairs * p.n_tiles;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp_idx == 0 && lane == 0) {
        for (int i = 0; i < B2R_MAX_SRC; ++i) tma_prefetch_desc(&p.a_map[i]);
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.out_map[0]);
        tma_prefetch_desc(&p.pool_map);
    }
    if (warp_idx == 1) {
        if (lane == 0) {
            for (int s = 0; s < kPairStages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 8 * PairCfg<kPairN>::kEpiGroups);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair<2 * kPairN>(tmem_ptr_s);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // both CTAs' barriers are initialised before any remote arrive / complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp_idx == 0) {
        // ===================================== TMA producer (both CTAs) =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
                const int n_tile = pair % p.n_tiles;
                int m = 2 * (pair / p.n_tiles) + int(rank);
                if (m >= num_m_tiles) m = num_m_tiles - 1;   // odd tile count: the last pair computes one tile twice
                const int w0 = (m % p.tiles_w) * p.tile_w;
                const int h0 = ((m / p.tiles_w) % p.tiles_h) * p.tile_h;
                const int n0 = (m / (p.tiles_w * p.tiles_h)) * p.tile_n;
                for (int kb = 0; kb < p.num_kblocks; ++kb) {
                    int src = 0, dh = 0, dw = 0, c0 = kb * kBlockK;
                    if (!p.linear_k) {
                        const uint32_t e = p.kblk[kb];
                        src = e & 3;
                        dh = int((e >> 2) & 3) - 1;
                        dw = int((e >> 4) & 3) - 1;
                        c0 = int(e >> 8) * kBlockK;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * kPairStageBytes);   // bytes of both CTAs
                    uint8_t* sa = stages + stage * kPairStageBytes;
                    tma_load_4d_pair(sa, &p.a_map[src], &full_bar[stage], c0, w0 + dw, h0 + dh, n0);
                    tma_load_2d_pair(sa + kAStageBytes, &p.b_map, &full_bar[stage], kb * kBlockK,
                                     n_tile * kPairN + int(rank) * (kPairN / 2));
                    if (++stage == kPairStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================================
        if (leader && lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * kPairN);
                for (int kb = 0; kb < p.num_kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stages + stage * kPairStageBytes);
                    const uint64_t adesc = make_sdesc_sw128(sa, 1024);
                    const uint64_t bdesc = make_sdesc_sw128(sa + kAStageBytes, 1024);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_bf16_ss_pair(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc,
                                          (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_pair(&empty_bar[stage]);
                    if (++stage == kPairStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_pair(&tmem_full_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================================== epilogue (both CTAs, own tile) =====================================
        conv_epilogue_role<kPairN, kBufs, SchedPair, PairCfg<kPairN>::kEpiGroups>(p, tmem_base, staging, bias_s, tmem_full_bar,
                                                                                  tmem_empty_bar, total_pairs, warp_idx, lane);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // nobody leaves (or frees TMEM) while the peer may still read its shared memory / barriers
    if (warp_idx == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc_pair<2 * kPairN>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pair + halo mode for C_out = 128 (round 2): cta_group::2 AND halo boxes.  The N = 128 layers are bound by the rate at
// which TMA fills shared memory (profiles/r01_smem_fill.md): per k-block of 4 MMAs (302 cycles at the N = 128 rate) the
// halo kernel pours 16 KB of weights + 6.7 KB of activations into each SM = 75 B/clk against ~55 B/clk available.  Here the
// two CTAs of a cluster execute MMAs of M = 256 x N = 128: each CTA owns one 128-pixel tile (its halo boxes, its rows of D,
// its epilogue) and loads only HALF of every weight k-block (8 KB): 49 B/clk.  Round 1 tried pair mode WITHOUT the halo
// boxes for N = 128 (16 KB + 8 KB per k-block) and only tied halo mode; the two savings are needed together.
// Two clusters are co-resident per SM pair (2 x 256 TMEM columns, 2 x <= 113 KB of shared memory), as in halo mode.
// Protocol = the pair kernel's (leader issues; its "full" barriers count both CTAs' TMA bytes; commits are multicast to
// both CTAs' "empty" / "accumulator ready" barriers) on the halo kernel's two rings.  Same MMAs in the same order as the
// halo kernel: bit-identical results (A/B test in tests/test_conv_gemm_gpu.py).
// ------------------------------------------------------------------------------------------------------------
template <int kPairN>
struct PairHaloCfg {
#ifndef B2R_PAIRHALO_BUFS
#define B2R_PAIRHALO_BUFS 1
#endif
    static constexpr int kBBytes = (kPairN / 2) * 128;   // this CTA's half of a weight k-block
    static constexpr int kCtasPerSm = kPairN == 128 ? 2 : 1;
    static constexpr int kStagingBufs = B2R_PAIRHALO_BUFS;
    static constexpr int kFixedBytes = 1024 + kStagingBufs * (kStagingFull + kStagingPool) + kPairN * 4 + 512;
    static constexpr int kMaxSmem = (227 * 1024) / kCtasPerSm - (kCtasPerSm > 1 ? 1024 : 0);
};

template <int kPairN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, PairHaloCfg<kPairN>::kCtasPerSm)
conv_gemm_pairhalo_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = PairHaloCfg<kPairN>;
    constexpr uint32_t kIdesc = make_idesc_bf16_f32(256, kPairN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + p.a_slots * p.a_slot_bytes;
    uint8_t* staging = b_ring + p.b_slots * Cfg::kBBytes;
    float* bias_s = reinterpret_cast<float*>(staging + Cfg::kStagingBufs * (kStagingFull + kStagingPool));
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + kPairN);
    uint64_t* a_full = bars;                           // leader only: TMA bytes of both CTAs
    uint64_t* a_empty = bars + kHaloMaxRing;           // each CTA: one multicast commit per phase
    uint64_t* b_full = bars + 2 * kHaloMaxRing;
    uint64_t* b_empty = bars + 3 * kHaloMaxRing;
    uint64_t* tmem_full_bar = bars + 4 * kHaloMaxRing;   // each CTA: one multicast commit per phase
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // leader only: 4 epilogue warps x 2 CTAs
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int m_pairs = (num_m_tiles + 1) / 2;
    const int total_pairs = m_pairs * p.n_tiles;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int AR = p.a_slots, BR = p.b_slots;

    if (warp_idx == 0 && lane == 0) {
        for (int i = 0; i < B2R_MAX_SRC; ++i) tma_prefetch_desc(&p.a3_map[i]);
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.out_map[0]);
        tma_prefetch_desc(&p.pool_map);
    }
    if (warp_idx == 1) {
        if (lane == 0) {
            for (int s = 0; s < kHaloMaxRing; ++s) {
                mbar_init(&a_full[s], 1);
                mbar_init(&a_empty[s], 1);
                mbar_init(&b_full[s], 1);
                mbar_init(&b_empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 8);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair<2 * kPairN>(tmem_ptr_s);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // both CTAs' barriers are initialised before any remote arrive / complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp_idx == 0) {
        // ===================================== TMA producer (both CTAs, own tile + own half of B) =====================================
        if (lane == 0) {
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
                const int n_tile = pair % p.n_tiles;
                int m = 2 * (pair / p.n_tiles) + int(rank);
                if (m >= num_m_tiles) m = num_m_tiles - 1;   // odd tile count: the last pair computes one tile twice
                const int w0 = (m % p.tiles_w) * p.tile_w;
                const int h0 = ((m / p.tiles_w) % p.tiles_h) * p.tile_h;
                const int n0 = m / (p.tiles_w * p.tiles_h);
                for (int s = 0; s < p.num_slots; ++s) {
                    const uint32_t e = p.slot[s];
                    const int src = e & 3, ntaps = ((e >> 2) & 1) ? 1 : 3, dw = int((e >> 4) & 3) - 1;
                    const int c0 = int((e >> 8) & 0xFFF) * 64, kb0 = int(e >> 20);
                    mbar_wait(&a_empty[as], aph ^ 1);
                    if (leader) mbar_arrive_expect_tx(&a_full[as], 2u * uint32_t(p.a_slot_bytes));
                    tma_load_4d_pair(a_ring + as * p.a_slot_bytes, &p.a3_map[src], &a_full[as], c0, w0 + dw, h0 - 1, n0);
                    if (++as == AR) {
                        as = 0;
                        aph ^= 1;
                    }
                    for (int t = 0; t < ntaps; ++t) {
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        if (leader) mbar_arrive_expect_tx(&b_full[bs], 2u * Cfg::kBBytes);
                        tma_load_2d_pair(b_ring + bs * Cfg::kBBytes, &p.b_map, &b_full[bs], (kb0 + t) * kBlockK,
                                         n_tile * kPairN + int(rank) * (kPairN / 2));
                        if (++bs == BR) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================================
        if (leader && lane == 0) {
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t row_stride = uint32_t(p.tile_w) * 128u;
            for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * kPairN);
                uint32_t accum = 0;
                for (int s = 0; s < p.num_slots; ++s) {
                    const uint32_t e = p.slot[s];
                    const bool centre = ((e >> 2) & 1) != 0;
                    const int ntaps = centre ? 1 : 3;
                    mbar_wait(&a_full[as], aph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(a_ring + as * p.a_slot_bytes);
                    for (int t = 0; t < ntaps; ++t) {
                        mbar_wait(&b_full[bs], bph);
                        tc_fence_after();
                        const uint64_t adesc = make_sdesc_sw128(sa + uint32_t(centre ? 1 : t) * row_stride, 1024);
                        const uint64_t bdesc = make_sdesc_sw128(smem_u32(b_ring + bs * Cfg::kBBytes), 1024);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            umma_bf16_ss_pair(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, accum);
                            accum = 1;
                        }
                        umma_commit_pair(&b_empty[bs]);
                        if (++bs == BR) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                    umma_commit_pair(&a_empty[as]);
                    if (++as == AR) {
                        as = 0;
                        aph ^= 1;
                    }
                }
                umma_commit_pair(&tmem_full_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        conv_epilogue_role<kPairN, Cfg::kStagingBufs, SchedPair>(p, tmem_base, staging, bias_s, tmem_full_bar, tmem_empty_bar,
                                                                 total_pairs, warp_idx, lane);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // nobody leaves (or frees TMEM) while the peer may still read its shared memory / barriers
    if (warp_idx == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc_pair<2 * kPairN>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// Pick (tw, th, tn), powers of two with product 128, minimising the number of tiles, then the halo area.
// Pass 0 keeps every box dimension within the next power of two of the tensor dimension; pass 1 (tiny tensors,
// e.g. a batch of one row through a Linear layer) drops that restriction and relies on TMA clipping.
static void choose_tile(int N, int H, int W, bool pool, bool spatial_taps, int* tw_o, int* th_o, int* tn_o) {
    for (int pass = 0; pass < 2; ++pass) {
        long best_tiles = -1, best_halo = 0;
        for (int tw = 128; tw >= 1; tw >>= 1) {
            for (int th = 128 / tw; th >= 1; th >>= 1) {
                const int tn = 128 / (tw * th);
                if (pass == 0 && (tw > next_pow2(W) || th > next_pow2(H) || tn > next_pow2(N))) continue;
                if (pool && (tw < 2 || th < 2)) continue;
                const long tiles = (long)ceil_div(W, tw) * ceil_div(H, th) * ceil_div(N, tn);
                const long halo = spatial_taps ? (long)(tw + 2) * (th + 2) * tn : 0;
                if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && halo < best_halo)) {
                    best_tiles = tiles;
                    best_halo = halo;
                    *tw_o = tw;
                    *th_o = th;
                    *tn_o = tn;
                }
            }
        }
        if (best_tiles >= 0) return;
    }
}


// ------------------------------------------------------------------------------------------------------------
// dispatch to the C_out = 64 specialisation (conv_n64.cu) when the layer has its shape
// ------------------------------------------------------------------------------------------------------------
static bool n64_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B2R_DISABLE_N64");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// returns B2R_OK and sets *handled when the layer was launched on the specialised kernel
// cout = 64: resident weights (conv_n64.cu).  (A C_out = 128 variant with streamed weights was tried and was slower than
// the generic kernel with two CTAs per SM; it is not kept.)
static int try_conv_halo(const b2r_conv_gemm_desc* d, cudaStream_t stream, bool* handled, int cout) {
    *handled = false;
    if (n64_disabled() || (d->flags & B2R_CONV_GENERIC_ONLY) || d->out_mode != B2R_OUT_NHWC || d->cout_total != cout || d->kblocks_host == nullptr) return B2R_OK;
    if (d->block_n != 0 && d->block_n != cout) return B2R_OK;
    if (d->tile_n > 1) return B2R_OK;
    const int nk = d->num_kblocks;
    uint32_t slots[kN64MaxSlots];
    int ns = 0;
    for (int i = 0; i < nk;) {
        const uint32_t e = d->kblocks_host[i];
        const int src = e & 3, dh = int((e >> 2) & 3) - 1, dw = int((e >> 4) & 3) - 1, c64 = int(e >> 8);
        bool group = (i + 9 <= nk);
        for (int t = 0; group && t < 9; ++t)  // dw-major, dh-minor: k-block i+t is tap (dh = t%3 - 1, dw = t/3 - 1)
            group = d->kblocks_host[i + t] == B2R_KBLOCK(src, t % 3 - 1, t / 3 - 1, c64);
        if (group) {
            if (ns + 3 > kN64MaxSlots) return B2R_OK;
            for (int j = 0; j < 3; ++j)
                slots[ns++] = uint32_t(src) | (uint32_t(j) << 4) | (uint32_t(c64) << 8) | (uint32_t(i + 3 * j) << 20);
            i += 9;
        } else if (dh == 0 && dw == 0) {
            if (ns + 1 > kN64MaxSlots) return B2R_OK;
            slots[ns++] = uint32_t(src) | (1u << 2) | (1u << 4) | (uint32_t(c64) << 8) | (uint32_t(i) << 20);
            i += 1;
        } else {
            return B2R_OK;  // some other tap order: the generic kernel handles it
        }
    }
    // tile: one image per tile, 8 x 16 or 16 x 8 pixels (W x H)
    int tw = d->tile_w, th = d->tile_h;
    if (tw == 0 && th == 0) {
        const long t_a = (long)ceil_div(d->W, 8) * ceil_div(d->H, 16);
        const long t_b = (long)ceil_div(d->W, 16) * ceil_div(d->H, 8);
        if (t_a <= t_b) { tw = 8; th = 16; } else { tw = 16; th = 8; }
    }
    if (!((tw == 8 && th == 16) || (tw == 16 && th == 8))) return B2R_OK;
    const int slot_bytes = (th + 2) * tw * 128;
    const long fixed = 1024 + (long)nk * 8192 + 16384 + 4096 + 256 + 256;
    int ring = int((kN64MaxSmem - fixed) / slot_bytes);
    if (ring > kN64MaxRing) ring = kN64MaxRing;
    if (ring < 3) return B2R_OK;  // weights too large to keep resident next to a useful ring

    static thread_local ConvN64Params tp;
    ConvN64Params& P = tp;
    memset(&P, 0, sizeof(P));
    const uint64_t N = d->N, H = d->H, W = d->W;
    for (int i = 0; i < B2R_MAX_SRC; ++i) {
        const int s = i < d->num_src ? i : 0;
        const uint64_t C = d->src_C[s];
        const uint64_t dims[4] = {C, W, H, N};
        const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
        const uint32_t box3[4] = {64, (uint32_t)tw, (uint32_t)(th + 2), 1};
        const uint32_t box1[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
        int rc = encode_tmap_bf16(&P.a3_map[i], d->src[s], 4, dims, strides, box3);
        if (rc) return rc;
        rc = encode_tmap_bf16(&P.a1_map[i], d->src[s], 4, dims, strides, box1);
        if (rc) return rc;
    }
    {
        const uint64_t K = (uint64_t)nk * 64;
        const uint64_t dims[2] = {K, (uint64_t)cout};
        const uint64_t strides[1] = {K * 2};
        const uint32_t box[2] = {64, (uint32_t)cout};
        int rc = encode_tmap_bf16(&P.b_map, d->weights, 2, dims, strides, box);
        if (rc) return rc;
    }
    const uint64_t OC = d->out_C;
    const uint64_t pd[4] = {OC, W / 2, H / 2, N};
    const uint64_t ps[3] = {OC * 2, (W / 2) * OC * 2, (H / 2) * (W / 2) * OC * 2};
    const uint32_t pb[4] = {64, (uint32_t)(tw / 2), (uint32_t)(th / 2), 1};
    if (d->out) {
        const uint64_t dims[4] = {OC, W, H, N};
        const uint64_t strides[3] = {OC * 2, W * OC * 2, H * W * OC * 2};
        const uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
        int rc = encode_tmap_bf16(&P.out_map, d->out, 4, dims, strides, box);
        if (rc) return rc;
    }
    if (d->out_pool) {
        int rc = encode_tmap_bf16(&P.pool_map, d->out_pool, 4, pd, ps, pb);
        if (rc) return rc;
    }
    if (!d->out) P.out_map = P.pool_map;
    if (!d->out_pool) P.pool_map = P.out_map;
    P.bias = d->bias;
    P.slope = d->slope;
    P.act = d->act;
    P.num_slots = ns;
    P.num_kblocks = nk;
    P.ring_slots = ring;
    P.slot_bytes = slot_bytes;
    P.tile_w = tw;
    P.tile_h = th;
    P.tiles_w = ceil_div(d->W, tw);
    P.tiles_h = ceil_div(d->H, th);
    P.n_img = d->N;
    P.store_full = d->out != nullptr;
    P.store_pool = d->out_pool != nullptr;
    memcpy(P.slot, slots, sizeof(uint32_t) * ns);
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const long total_tiles = (long)P.tiles_w * P.tiles_h * P.n_img;
    if (total_tiles >= (1L << 31)) return B2R_OK;
    int grid = d->max_ctas > 0 ? d->max_ctas : sms;
    if (grid > total_tiles) grid = (int)total_tiles;
    const size_t smem = size_t(fixed) + size_t(ring) * slot_bytes;
    rc = launch_conv_n64(P, grid, smem, stream);
    if (rc) return rc;
    *handled = true;
    return B2R_OK;
}


// ------------------------------------------------------------------------------------------------------------
// dispatch to the tap-folded C_out = 64 kernel (conv_w3.cu) when the caller supplied the wide weight layout
// ------------------------------------------------------------------------------------------------------------
static int try_conv_w3(const b2r_conv_gemm_desc* d, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (d->weights_w3 == nullptr || (d->flags & (B2R_CONV_GENERIC_ONLY | B2R_CONV_NO_W3))) return B2R_OK;
    if (d->out_mode != B2R_OUT_NHWC || d->cout_total != 64 || d->kblocks_host == nullptr) return B2R_OK;
    if (d->block_n != 0 && d->block_n != 64) return B2R_OK;
    if (d->tile_w != 0 || d->tile_h != 0 || d->tile_n != 0) return B2R_OK;  // explicit tiles belong to the other kernels
    const int nk = d->num_kblocks;
    uint32_t groups[kW3MaxGroups];
    int ng = 0, nsteps = 0;
    for (int i = 0; i < nk;) {
        const uint32_t e = d->kblocks_host[i];
        const int src = e & 3, dh = int((e >> 2) & 3) - 1, dw = int((e >> 4) & 3) - 1, c64 = int(e >> 8);
        bool group = (i + 9 <= nk);
        for (int t = 0; group && t < 9; ++t)
            group = d->kblocks_host[i + t] == B2R_KBLOCK(src, t % 3 - 1, t / 3 - 1, c64);
        if (ng >= kW3MaxGroups) return B2R_OK;
        if (group) {
            groups[ng++] = uint32_t(src) | (uint32_t(c64) << 8) | (uint32_t(nsteps) << 20);
            nsteps += 3;
            i += 9;
        } else if (dh == 0 && dw == 0) {
            groups[ng++] = uint32_t(src) | (1u << 2) | (uint32_t(c64) << 8) | (uint32_t(nsteps) << 20);
            nsteps += 1;
            i += 1;
        } else {
            return B2R_OK;
        }
    }
    // shared-memory plan.  Resident weights: 24 KB per k-step of a 3x3 group, 8 KB (the kw = 1 rows) for a 1x1 group that is
    // not the first group; HALF of that per CTA in pair mode (cta_group::2, two CTAs share every weight block).  The head
    // variant stages only its partial sums (6 KB per buffer), other layers a full + a pooled tile.  What is left goes to the
    // input ring: two tiles' worth of slots lets the two MMA issuers really alternate (with one tile's worth the next
    // tile's boxes cannot even be requested before this tile's MMAs retire).
    const size_t stage_stride = d->head_w ? 6144 : (14336 + 4096);
    int sms_w3 = 0;
    {
        int src_ = device_sm_count(&sms_w3);
        if (src_) return src_;
    }
    static const bool no_w3pair_env = getenv("B2R_NO_W3PAIR") != nullptr;   // A/B switch for tools/layer_bench.py, read once
    const long tiles_total = (long)ceil_div(d->W, 14) * ceil_div(d->H, 8) * d->N;
    bool pair = !no_w3pair_env && !(d->flags & B2R_CONV_NO_PAIR) && sms_w3 >= 2 && tiles_total >= 2 && d->max_ctas != 1;
    uint32_t boff[kW3MaxGroups];
    size_t b_bytes = 0;
    int ring = 0, b_slots = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t div = pair ? 2 : 1;
        b_bytes = 0;
        for (int g = 0; g < ng; ++g) {
            boff[g] = uint32_t(b_bytes >> 4);
            const bool centre = (groups[g] >> 2) & 1;
            b_bytes += (centre ? (g > 0 ? 8192 : 24576) : 3 * 24576) / div;
        }
        ring = 2 * ng > kN64MaxRing ? kN64MaxRing : (2 * ng < 4 ? 4 : 2 * ng);
        if (pair && ring < kN64MaxRing && ring < 6) ring = 6 < kN64MaxRing ? 6 : kN64MaxRing;   // the halved weights leave room: a third tile in flight
        while (ring >= 2 && conv_w3_smem_bytes(b_bytes, ring, stage_stride, 2) > (size_t)kN64MaxSmem) --ring;
        if (ring >= (pair ? ng : 2) || !pair) break;
        pair = false;   // even the halved weights do not leave a ring of one tile: single-CTA kernel (streams the weights)
    }
    if (ring < 2) {
        // weights too large to stay resident (192 -> 64 without pair mode: 216 KB): stream the k-steps through a ring instead
        if (d->head_w) return B2R_OK;
        ring = 3;
        b_slots = kN64MaxRing;
        while (b_slots >= 3 && conv_w3_smem_bytes((size_t)b_slots * 24576, ring, stage_stride, 2) > (size_t)kN64MaxSmem) --b_slots;
        if (b_slots < 3) return B2R_OK;
        b_bytes = (size_t)b_slots * 24576;
    }
    // staging buffers: two are the minimum (one being stored while the next is written); what the ring leaves over buys up
    // to two more, which decouples the epilogue (and through it the TMEM stage and the MMA issuers) from TMA stores that the
    // HBM write path delays.  B2R_W3_STAGE_BUFS / B2R_W3_RING pin either for experiments (tools/layer_bench.py).
    static const int stage_bufs_env = [] { const char* e = getenv("B2R_W3_STAGE_BUFS"); return e ? atoi(e) : 0; }();
    static const int ring_env = [] { const char* e = getenv("B2R_W3_RING"); return e ? atoi(e) : 0; }();
    if (ring_env >= 2 && ring_env <= ring && b_slots == 0) ring = ring_env;
    int stage_bufs = 2;
    if (!d->head_w && b_slots == 0) {
        const int want = stage_bufs_env >= 2 && stage_bufs_env <= kW3MaxStageBufs ? stage_bufs_env : 3;
        while (stage_bufs < want && conv_w3_smem_bytes(b_bytes, ring, stage_stride, stage_bufs + 1) <= (size_t)kN64MaxSmem) ++stage_bufs;
    }

    static thread_local ConvW3Params tp;
    ConvW3Params& P = tp;
    memset(&P, 0, sizeof(P));
    const uint64_t N = d->N, H = d->H, W = d->W;
    for (int i = 0; i < B2R_MAX_SRC; ++i) {
        const int s = i < d->num_src ? i : 0;
        const uint64_t C = d->src_C[s];
        const uint64_t dims[4] = {C, W, H, N};
        const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
        const uint32_t box[4] = {64, 16, 10, 1};
        int rc = encode_tmap_bf16(&P.a_map[i], d->src[s], 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t K = (uint64_t)nsteps * 64;
        const uint64_t dims[2] = {K, 192};
        const uint64_t strides[1] = {K * 2};
        const uint32_t box[2] = {64, pair ? 96u : 192u};       // pair mode: each CTA loads its half of the rows
        int rc = encode_tmap_bf16(&P.b_map, d->weights_w3, 2, dims, strides, box);
        if (rc) return rc;
        const uint32_t box_c[2] = {64, pair ? 32u : 64u};
        rc = encode_tmap_bf16(&P.b_map_c, d->weights_w3, 2, dims, strides, box_c);
        if (rc) return rc;
    }
    const uint64_t OC = d->out_C;
    if (d->out) {
        const uint64_t dims[4] = {OC, W, H, N};
        const uint64_t strides[3] = {OC * 2, W * OC * 2, H * W * OC * 2};
        const uint32_t box[4] = {64, 14, 8, 1};
        int rc = encode_tmap_bf16(&P.out_map, d->out, 4, dims, strides, box);
        if (rc) return rc;
    }
    if (d->out_pool) {
        const uint64_t pd[4] = {OC, W / 2, H / 2, N};
        const uint64_t ps[3] = {OC * 2, (W / 2) * OC * 2, (H / 2) * (W / 2) * OC * 2};
        const uint32_t pb[4] = {64, 7, 4, 1};
        int rc = encode_tmap_bf16(&P.pool_map, d->out_pool, 4, pd, ps, pb);
        if (rc) return rc;
    }
    if (!d->out && !d->out_pool) {
        P.out_map = P.a_map[0];   // never used for a store (store_full = store_pool = 0); keeps prefetch valid
        P.pool_map = P.a_map[0];
    } else {
        if (!d->out) P.out_map = P.pool_map;
        if (!d->out_pool) P.pool_map = P.out_map;
    }
    P.bias = d->bias;
    P.slope = d->slope;
    P.act = d->act;
    P.dbg = reinterpret_cast<long long*>(d->debug_timeline);
    P.head_w = d->head_w;
    P.head_b = d->head_b;
    P.head_f32 = d->head_out_f32;
    P.head_u8 = d->head_out_u8;
    P.H = d->H;
    P.W = d->W;
    P.num_groups = ng;
    P.num_ksteps = nsteps;
    P.ring_slots = ring;
    P.b_slots = b_slots;
    P.b_bytes = (int)b_bytes;
    P.stage_stride = (int)stage_stride;
    P.stage_bufs = stage_bufs;
    // L2 prefetch of the boxes four tiles ahead: needed when the ring cannot hold two tiles' boxes (the loads of the next tile
    // then cannot be issued before this tile's MMAs retire); with a deeper ring it costs 1.5 - 3 % (tools/exp/tma_copy.cu: the
    // same pattern without MMAs runs 250 us without and 306 us with the prefetch)
    P.prefetch = ring < 2 * ng ? 1 : 0;
    memcpy(P.group_boff, boff, sizeof(uint32_t) * ng);
    P.tiles_w = ceil_div(d->W, 14);
    P.tiles_h = ceil_div(d->H, 8);
    P.n_img = d->N;
    P.store_full = d->out != nullptr;
    P.store_pool = d->out_pool != nullptr;
    memcpy(P.group, groups, sizeof(uint32_t) * ng);
    const int sms = sms_w3;
    const long total_tiles = (long)P.tiles_w * P.tiles_h * P.n_img;
    if (total_tiles >= (1L << 31)) return B2R_OK;
    int grid = d->max_ctas > 0 ? d->max_ctas : sms;
    if (pair) {
        const long pairs = (total_tiles + 1) / 2;
        long clusters = grid / 2;
        if (clusters < 1) clusters = 1;
        if (clusters > pairs) clusters = pairs;
        grid = int(2 * clusters);
    } else if (grid > total_tiles) {
        grid = (int)total_tiles;
    }
    int rc = launch_conv_w3(P, grid, stream, pair);
    if (rc) return rc;
    *handled = true;
    return B2R_OK;
}

template <int BLOCK_N, bool kDeep = false>
static int launch(const ConvGemmParams& p, int grid, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, kDeep>;
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        B2R_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, kDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
        if (dev < 64) attr_set[dev] = true;
    }
    note_conv_kernel(BLOCK_N == 64 ? "conv_gemm_kernel<64>" : BLOCK_N == 128 ? "conv_gemm_kernel<128>"
                     : kDeep ? "conv_gemm_kernel<256,deep>" : "conv_gemm_kernel<256>");
    conv_gemm_kernel<BLOCK_N, kDeep><<<grid * Cfg::kCtasPerSm, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(p);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

template <int N>
static int launch_pair(const ConvGemmParams& p, int clusters, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        B2R_CUDA(cudaFuncSetAttribute(conv_gemm_pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PairCfg<N>::kSmemBytes));
        if (dev < 64) attr_set[dev] = true;
    }
    note_conv_kernel(N == 256 ? "conv_gemm_pair_kernel<256>" : "conv_gemm_pair_kernel<128>");
    conv_gemm_pair_kernel<N><<<(unsigned)(2 * clusters), PairCfg<N>::kThreads, PairCfg<N>::kSmemBytes, stream>>>(p);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

template <int BLOCK_N>
static int launch_halo(const ConvGemmParams& p, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        B2R_CUDA(cudaFuncSetAttribute(conv_gemm_halo_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      HaloCfg<BLOCK_N>::kMaxSmem));
        if (dev < 64) attr_set[dev] = true;
    }
    note_conv_kernel(BLOCK_N == 128 ? "conv_gemm_halo_kernel<128>" : "conv_gemm_halo_kernel<256>");
    conv_gemm_halo_kernel<BLOCK_N><<<grid, kNumThreads, smem, stream>>>(p);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

template <int N>
static int launch_pairhalo(const ConvGemmParams& p, int clusters, size_t smem, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        B2R_CUDA(cudaFuncSetAttribute(conv_gemm_pairhalo_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PairHaloCfg<N>::kMaxSmem));
        if (dev < 64) attr_set[dev] = true;
    }
    note_conv_kernel(N == 128 ? "conv_gemm_pairhalo_kernel<128>" : "conv_gemm_pairhalo_kernel<256>");
    conv_gemm_pairhalo_kernel<N><<<(unsigned)(2 * clusters), kNumThreads, smem, stream>>>(p);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

// Fills the halo-mode fields of P when the layer qualifies; returns the dynamic shared memory size or 0.
// Cfg = HaloCfg<N> (one CTA per tile) or PairHaloCfg<N> (cta_group::2: half a weight k-block per CTA and slot).
template <class Cfg>
static size_t plan_halo(const b2r_conv_gemm_desc* d, ConvGemmParams& P, int tw, int th, int tn, int* rc_out) {
    *rc_out = B2R_OK;
    if ((d->flags & B2R_CONV_NO_HALO) || d->kblocks_host == nullptr || d->out_mode != B2R_OUT_NHWC || tn != 1 || tw % 8 != 0)
        return 0;
    const int nk = d->num_kblocks;
    int ns = 0;
    for (int i = 0; i < nk;) {
        const uint32_t e = d->kblocks_host[i];
        const int src = e & 3, dh = int((e >> 2) & 3) - 1, dw = int((e >> 4) & 3) - 1, c64 = int(e >> 8);
        bool group = (i + 9 <= nk);
        for (int t = 0; group && t < 9; ++t)  // dw-major, dh-minor
            group = d->kblocks_host[i + t] == B2R_KBLOCK(src, t % 3 - 1, t / 3 - 1, c64);
        if (group) {
            if (ns + 3 > kN64MaxSlots || c64 > 0xFFF || i + 8 > 0xFFF) return 0;
            for (int j = 0; j < 3; ++j)
                P.slot[ns++] = uint32_t(src) | (uint32_t(j) << 4) | (uint32_t(c64) << 8) | (uint32_t(i + 3 * j) << 20);
            i += 9;
        } else if (dh == 0 && dw == 0) {
            if (ns + 1 > kN64MaxSlots || c64 > 0xFFF || i > 0xFFF) return 0;
            P.slot[ns++] = uint32_t(src) | (1u << 2) | (1u << 4) | (uint32_t(c64) << 8) | (uint32_t(i) << 20);
            i += 1;
        } else {
            return 0;
        }
    }
    const int a_bytes = (th + 2) * tw * 128;
    static const int a_slots_env = [] {   // tuning knob for tools/layer_bench.py, read once
        const char* e = getenv("B2R_HALO_A_SLOTS");
        return e ? atoi(e) : 2;
    }();
    int a_slots = a_slots_env;
    long room = (long)Cfg::kMaxSmem - Cfg::kFixedBytes - (long)a_slots * a_bytes;
    int b_slots = (int)(room / Cfg::kBBytes);
    if (b_slots > kHaloMaxRing) b_slots = kHaloMaxRing;
    if (b_slots < 2) return 0;
    // spend what is left on a third / fourth A slot
    room -= (long)b_slots * Cfg::kBBytes;
    while (a_slots < 4 && room >= a_bytes) {
        ++a_slots;
        room -= a_bytes;
    }
    const uint64_t N = d->N, H = d->H, W = d->W;
    for (int i = 0; i < B2R_MAX_SRC; ++i) {
        const int s = i < d->num_src ? i : 0;
        const uint64_t C = d->src_C[s];
        const uint64_t dims[4] = {C, W, H, N};
        const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
        const uint32_t box3[4] = {64, (uint32_t)tw, (uint32_t)(th + 2), 1};
        int rc = encode_tmap_bf16(&P.a3_map[i], d->src[s], 4, dims, strides, box3);
        if (rc) {
            *rc_out = rc;
            return 0;
        }
    }
    P.num_slots = ns;
    P.a_slot_bytes = a_bytes;
    P.a_slots = a_slots;
    P.b_slots = b_slots;
    return (size_t)Cfg::kFixedBytes + (size_t)a_slots * a_bytes + (size_t)b_slots * Cfg::kBBytes;
}

}  // namespace b2r

extern "C" int b2r_conv_gemm(const b2r_conv_gemm_desc* d, void* stream_v) {
    using namespace b2r;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    B2R_REQUIRE(d != nullptr, "desc is null");
    B2R_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, "bad shape N=%d H=%d W=%d", d->N, d->H, d->W);
    B2R_REQUIRE(d->num_src >= 1 && d->num_src <= B2R_MAX_SRC, "num_src=%d", d->num_src);
    B2R_REQUIRE(d->weights && d->bias, "weights/bias null");
    B2R_REQUIRE(d->cout_total > 0 && d->cout_total % 64 == 0, "cout_total=%d must be a multiple of 64", d->cout_total);
    B2R_REQUIRE(d->num_kblocks >= 1, "num_kblocks=%d", d->num_kblocks);
    B2R_REQUIRE(d->kblocks_host == nullptr || d->num_kblocks <= B2R_MAX_KBLOCKS, "num_kblocks=%d > %d", d->num_kblocks,
                B2R_MAX_KBLOCKS);
    const bool head = d->head_w != nullptr;
    B2R_REQUIRE(d->out || d->out_pool || head, "no output");
    B2R_REQUIRE(!head || (d->head_b && (d->head_out_f32 || d->head_out_u8)), "head_w given without head_b / a head output");
    B2R_REQUIRE(!head || (d->weights_w3 && d->cout_total == 64 && d->out_mode == B2R_OUT_NHWC && !d->out_pool &&
                          !(d->flags & (B2R_CONV_GENERIC_ONLY | B2R_CONV_NO_W3))),
                "the fused 64->3 head needs a C_out = 64 NHWC layer with weights_w3 and no pooled output");
    B2R_REQUIRE(!head || !d->out, "with a fused head the 64-channel tensor is not stored: out must be NULL");
    B2R_REQUIRE(head || (d->out_C > 0 && d->out_C % 64 == 0), "out_C=%d must be a multiple of 64", d->out_C);
    B2R_REQUIRE(d->act >= B2R_ACT_NONE && d->act <= B2R_ACT_PRELU, "act=%d", d->act);
    const bool convt = d->out_mode == B2R_OUT_CONVT2X2;
    B2R_REQUIRE(convt || d->out_mode == B2R_OUT_NHWC, "out_mode=%d", d->out_mode);
    B2R_REQUIRE(!(convt && d->out_pool), "pooling is not available with CONVT2X2");
    B2R_REQUIRE(!convt || d->cout_total % 4 == 0, "CONVT2X2 needs cout_total = 4*C_out");
    const int cout = convt ? d->cout_total / 4 : d->cout_total;
    B2R_REQUIRE(cout % 64 == 0 && (cout <= d->out_C || (head && !d->out)), "C_out=%d vs out_C=%d", cout, d->out_C);
    if (d->out_pool) B2R_REQUIRE(d->H % 2 == 0 && d->W % 2 == 0, "fused pool needs even H, W (got %d x %d)", d->H, d->W);

    bool spatial = false;
    for (int i = 0; i < d->num_src; ++i) {
        B2R_REQUIRE(d->src[i] != nullptr, "src[%d] is null", i);
        B2R_REQUIRE(d->src_C[i] > 0 && d->src_C[i] % 64 == 0, "src_C[%d]=%d must be a multiple of 64", i, d->src_C[i]);
    }
    if (d->kblocks_host) {
        for (int kb = 0; kb < d->num_kblocks; ++kb) {
            const uint32_t e = d->kblocks_host[kb];
            const int src = e & 3, dh = int((e >> 2) & 3) - 1, dw = int((e >> 4) & 3) - 1, c64 = int(e >> 8);
            B2R_REQUIRE(src < d->num_src, "kblock %d: source %d >= num_src", kb, src);
            B2R_REQUIRE(dh >= -1 && dh <= 1 && dw >= -1 && dw <= 1, "kblock %d: bad offset", kb);
            B2R_REQUIRE((c64 + 1) * 64 <= d->src_C[src], "kblock %d: channel chunk %d outside source %d (C=%d)", kb, c64,
                        src, d->src_C[src]);
            if (dh != 0 || dw != 0) spatial = true;
        }
    } else {
        B2R_REQUIRE(d->num_kblocks * 64 <= d->src_C[0], "linear K: %d kblocks vs src_C=%d", d->num_kblocks, d->src_C[0]);
    }

    {
        bool handled = false;
        int rc = try_conv_w3(d, stream, &handled);
        if (rc || handled) return rc;
        B2R_REQUIRE(!head, "the fused head is only implemented on the tap-folded kernel, which rejected this layer "
                           "(weights too large to stay resident in shared memory?)");
        rc = try_conv_halo(d, stream, &handled, 64);
        if (rc || handled) return rc;
    }

    // ---- tile geometry
    int block_n = d->block_n;
    // CONVT2X2: the four taps are four column blocks of ONE GEMM (cout_total = 4 C_out), so an n-tile may span taps:
    // C_out = 64 runs as one 128 x 256 tile per 128 pixels instead of four 128 x 64 tiles that each re-load the input
    const int n_span = convt ? d->cout_total : cout;
    if (block_n == 0) block_n = (n_span % 256 == 0) ? 256 : (n_span % 128 == 0 ? 128 : 64);
    B2R_REQUIRE(block_n == 64 || block_n == 128 || block_n == 256, "block_n=%d", block_n);
    B2R_REQUIRE(n_span % block_n == 0 && (cout % block_n == 0 || block_n % cout == 0),
                "C_out=%d (cout_total=%d) incompatible with block_n=%d", cout, d->cout_total, block_n);
    int tw = d->tile_w, th = d->tile_h, tn = d->tile_n;
    if (tw == 0 && th == 0 && tn == 0) {
        choose_tile(d->N, d->H, d->W, d->out_pool != nullptr, spatial, &tw, &th, &tn);
        if (block_n == 128 && spatial && !(d->flags & B2R_CONV_NO_HALO) && d->out_mode == B2R_OUT_NHWC && (tn != 1 || tw % 8 != 0)) {
            // N = 128 layers are bound by shared-memory fill: a halo-capable tile (one image, TW % 8 == 0) is worth up to
            // 15 % more (partly clipped) tiles (dec3.c1 @56: 8x8x2 -> 8x16x1, 337 -> 315 us; profiles/r01_smem_fill.md)
            const long base = (long)ceil_div(d->W, tw) * ceil_div(d->H, th) * ceil_div(d->N, tn);
            long best = -1;
            int bw = 0, bh = 0;
            for (int cw = 8; cw <= 128; cw <<= 1) {
                const int ch = 128 / cw;
                if (d->out_pool && ch < 2) continue;
                const long t = (long)ceil_div(d->W, cw) * ceil_div(d->H, ch) * d->N;
                if (best < 0 || t < best) {
                    best = t;
                    bw = cw;
                    bh = ch;
                }
            }
            if (best > 0 && best * 100 <= base * 115) {
                tw = bw;
                th = bh;
                tn = 1;
            }
        }
    }
    B2R_REQUIRE(tw > 0 && th > 0 && tn > 0 && tw * th * tn == kBlockM, "tile %dx%dx%d must cover 128 pixels", tw, th, tn);
    B2R_REQUIRE(tw <= 256 && th <= 256 && tn <= 256, "tile dims must be <= 256");
    if (d->out_pool) B2R_REQUIRE(tw % 2 == 0 && th % 2 == 0, "fused pool needs even tile_w, tile_h");

    static thread_local ConvGemmParams tp;  // ~1.6 KB, passed by value as the __grid_constant__ kernel parameter
    ConvGemmParams& P = tp;
    memset(&P, 0, sizeof(P));

    const uint64_t N = d->N, H = d->H, W = d->W;
    for (int i = 0; i < B2R_MAX_SRC; ++i) {
        const int s = i < d->num_src ? i : 0;  // unused slots alias source 0 so that prefetch sees a valid map
        const uint64_t C = d->src_C[s];
        const uint64_t dims[4] = {C, W, H, N};
        const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
        const uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
        int rc = encode_tmap_bf16(&P.a_map[i], d->src[s], 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t K = (uint64_t)d->num_kblocks * 64;
        const uint64_t dims[2] = {K, (uint64_t)d->cout_total};
        const uint64_t strides[1] = {K * 2};
        const uint32_t box[2] = {64, (uint32_t)block_n};
        int rc = encode_tmap_bf16(&P.b_map, d->weights, 2, dims, strides, box);
        if (rc) return rc;
    }
    const uint64_t OC = d->out_C;
    const uint32_t obox[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
    if (convt) {
        B2R_REQUIRE(d->out != nullptr, "CONVT2X2 needs out");
        for (int q = 0; q < 4; ++q) {
            const int i = q >> 1, j = q & 1;
            const uint8_t* base = static_cast<const uint8_t*>(d->out) + ((uint64_t)i * 2 * W + j) * OC * 2;
            const uint64_t dims[4] = {OC, W, H, N};
            const uint64_t strides[3] = {2 * OC * 2, 2 * (2 * W) * OC * 2, (2 * H) * (2 * W) * OC * 2};
            int rc = encode_tmap_bf16(&P.out_map[q], base, 4, dims, strides, obox);
            if (rc) return rc;
        }
    } else {
        const void* obase = d->out ? d->out : d->out_pool;  // placeholder map when only the pooled output is stored
        const uint64_t dims[4] = {OC, W, H, N};
        const uint64_t strides[3] = {OC * 2, W * OC * 2, H * W * OC * 2};
        if (d->out) {
            int rc = encode_tmap_bf16(&P.out_map[0], obase, 4, dims, strides, obox);
            if (rc) return rc;
        } else {
            const uint64_t pd[4] = {OC, W / 2, H / 2, N};
            const uint64_t ps[3] = {OC * 2, (W / 2) * OC * 2, (H / 2) * (W / 2) * OC * 2};
            const uint32_t pb[4] = {64, (uint32_t)(tw / 2), (uint32_t)(th / 2), (uint32_t)tn};
            int rc = encode_tmap_bf16(&P.out_map[0], obase, 4, pd, ps, pb);
            if (rc) return rc;
        }
        for (int q = 1; q < 4; ++q) P.out_map[q] = P.out_map[0];
    }
    if (d->out_pool) {
        const uint64_t pd[4] = {OC, W / 2, H / 2, N};
        const uint64_t ps[3] = {OC * 2, (W / 2) * OC * 2, (H / 2) * (W / 2) * OC * 2};
        const uint32_t pb[4] = {64, (uint32_t)(tw / 2), (uint32_t)(th / 2), (uint32_t)tn};
        int rc = encode_tmap_bf16(&P.pool_map, d->out_pool, 4, pd, ps, pb);
        if (rc) return rc;
    } else {
        P.pool_map = P.out_map[0];
    }

    P.bias = d->bias;
    P.slope = d->slope;
    P.act = d->act;
    P.num_kblocks = d->num_kblocks;
    P.linear_k = d->kblocks_host == nullptr;
    if (d->kblocks_host) memcpy(P.kblk, d->kblocks_host, sizeof(uint32_t) * d->num_kblocks);
    P.tile_w = tw;
    P.tile_h = th;
    P.tile_n = tn;
    P.tiles_w = ceil_div(d->W, tw);
    P.tiles_h = ceil_div(d->H, th);
    P.tiles_n = ceil_div(d->N, tn);
    P.n_tiles = d->cout_total / block_n;
    P.cout_per_out = convt ? cout : d->cout_total;
    P.store_full = d->out != nullptr;
    P.store_pool = d->out_pool != nullptr;

    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const long total_tiles = (long)P.tiles_w * P.tiles_h * P.tiles_n * P.n_tiles;
    B2R_REQUIRE(total_tiles < (1L << 31), "too many tiles");
    int grid = d->max_ctas > 0 ? d->max_ctas : sms;
    if (grid > total_tiles) grid = (int)total_tiles;

    static const bool pair128_env = getenv("B2R_PAIR128") != nullptr;   // experiment switch (tools/layer_bench.py), read once
    const bool pair128 = block_n == 128 && pair128_env;
    if ((block_n == 256 || pair128) && d->out_mode == B2R_OUT_NHWC && !(d->flags & B2R_CONV_NO_PAIR) && sms >= 2 && total_tiles >= 2) {
        {   // each CTA of the pair loads half of a weight k-block: box 64 x block_n / 2
            const uint64_t K = (uint64_t)d->num_kblocks * 64;
            const uint64_t dims[2] = {K, (uint64_t)d->cout_total};
            const uint64_t strides[1] = {K * 2};
            const uint32_t box[2] = {64, (uint32_t)(block_n / 2)};
            int brc = encode_tmap_bf16(&P.b_map, d->weights, 2, dims, strides, box);
            if (brc) return brc;
        }
        const long pairs = ((long)P.tiles_w * P.tiles_h * P.tiles_n + 1) / 2 * P.n_tiles;
        long clusters = d->max_ctas > 0 ? d->max_ctas / 2 : (long)(sms / 2) * (block_n == 256 ? 1 : PairCfg<128>::kCtasPerSm);
        if (clusters < 1) clusters = 1;
        if (clusters > pairs) clusters = pairs;
        return block_n == 256 ? launch_pair<256>(P, (int)clusters, stream) : launch_pair<128>(P, (int)clusters, stream);
    }
    static const bool no_pairhalo_env = getenv("B2R_NO_PAIRHALO") != nullptr;   // A/B switch for tools/layer_bench.py, read once
    if (block_n == 128 && spatial && !pair128 && !no_pairhalo_env && !(d->flags & B2R_CONV_NO_PAIR) && sms >= 2 &&
        (long)P.tiles_w * P.tiles_h * P.tiles_n >= 2) {
        int hrc = B2R_OK;
        const size_t hs = plan_halo<PairHaloCfg<128>>(d, P, tw, th, tn, &hrc);
        if (hrc) return hrc;
        if (hs) {
            const uint64_t K = (uint64_t)d->num_kblocks * 64;      // each CTA of the pair loads half of a weight k-block
            const uint64_t dims[2] = {K, (uint64_t)d->cout_total};
            const uint64_t strides[1] = {K * 2};
            const uint32_t box[2] = {64, 64};
            int brc = encode_tmap_bf16(&P.b_map, d->weights, 2, dims, strides, box);
            if (brc) return brc;
            const long pairs = ((long)P.tiles_w * P.tiles_h * P.tiles_n + 1) / 2 * P.n_tiles;
            long clusters = d->max_ctas > 0 ? d->max_ctas / 2 : (long)(sms / 2) * PairHaloCfg<128>::kCtasPerSm;
            if (clusters < 1) clusters = 1;
            if (clusters > pairs) clusters = pairs;
            return launch_pairhalo<128>(P, (int)clusters, hs, stream);
        }
    }
    if (block_n >= 128 && spatial) {
        int hrc = B2R_OK;
        const size_t hs = block_n == 128 ? plan_halo<HaloCfg<128>>(d, P, tw, th, tn, &hrc) : plan_halo<HaloCfg<256>>(d, P, tw, th, tn, &hrc);
        if (hrc) return hrc;
        if (hs) {
            // co-resident CTAs (N = 128): twice the grid
            const int ctas = block_n == 128 ? HaloCfg<128>::kCtasPerSm : 1;
            int hgrid = d->max_ctas > 0 ? d->max_ctas : sms * ctas;
            if (hgrid > total_tiles) hgrid = (int)total_tiles;
            return block_n == 128 ? launch_halo<128>(P, hgrid, hs, stream) : launch_halo<256>(P, hgrid, hs, stream);
        }
    }
    switch (block_n) {
        case 64: return launch<64>(P, grid, stream);
        case 128: return launch<128>(P, grid, stream);
        default: {
            // short K without a fused pool (the ConvTranspose layers): the epilogue is all there is to hide
            static const bool no_deep_env = getenv("B2R_NO_DEEP_EPI") != nullptr;   // A/B switch, read once
            if (!d->out_pool && d->num_kblocks <= 8 && !no_deep_env) return launch<256, true>(P, grid, stream);
            return launch<256>(P, grid, stream);
        }
    }
}
