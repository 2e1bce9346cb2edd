// tcgen05 conv3x3 for C_out = 64 with the three horizontal taps folded into the MMA N dimension (N = 192).
//
// Measured on B200 (tools/exp/mma_rate.cu, profiles/r01_mma_rate.md): one tcgen05.mma 128 x N x 16 with both
// operands in shared memory takes max(75.6, N/2) cycles.  With N = C_out = 64 the tensor pipe can therefore never
// exceed 32/75.6 = 42 % of its peak, however the operands are staged (csrc/conv_n64.cu reaches 96 % of that floor).
// This kernel makes N = 3 x 64: the B operand of one k-step holds the weights of the three taps (kh, kw = 0..2)
// of one kernel row for 64 input channels, so ONE MMA produces the three partial sums
//     D_kw[(h, c)] = sum_ci in[h + kh - 1][c][ci] * W[co][ci][kh][kw],        c = input column inside the tile,
// at the ideal 96 cycles, and the shift along W moves from the operand side to the epilogue:
//     out[h][w] = D_0[(h, w)] + D_1[(h, w + 1)] + D_2[(h, w + 2)]
// which is two warp shuffles per value because a tile row (16 columns) lives in 16 adjacent lanes.  A tile is
// 8 rows x 16 input columns -> 8 x 14 outputs (224 = 16 * 14, 112 = 8 * 14: no waste at the sizes that matter;
// 12.5 % of the MMA rows are halo).  Per (source, 64-channel chunk) the producer loads ONE box
// 64 ch x 16 x (8 + 2) (20 KB) that feeds the three kernel rows (12 MMAs, 1152 cycles): 11x less L2->SM traffic
// than the generic kernel.  The weights [192][K/3] stay resident in shared memory.  1x1 centre groups (ResidualBlock
// shortcut / identity) are k-steps whose kw = 0 and kw = 2 weight rows are zero.
#include <cstring>

#include "b2r_internal.h"
#include "conv_common.cuh"
#include "ptx_sm100.cuh"

namespace b2r {

constexpr int kW3EpiWarps = 16;         // four warps per TMEM lane quarter, 16 output channels each
constexpr int kW3Threads = (4 + kW3EpiWarps) * 32;   // warp 0 TMA loads, warps 1-2 MMA (alternate tiles), warps 3..18 epilogue, warp 19 TMA stores
constexpr int kW3BStep = 192 * 128;      // weights of one k-step: 192 rows x 128 B
constexpr int kW3Slot = 10 * 16 * 128;   // one halo box: (8 + 2) rows x 16 columns x 128 B
constexpr int kW3Staging = 14336;        // 8 x 14 pixels x 128 B
constexpr int kW3StagingPool = 4096;     // 4 x 7 pixels x 128 B, padded
#ifndef B2R_W3_PREFETCH_TILES
#define B2R_W3_PREFETCH_TILES 4
#endif
constexpr int kW3PrefetchTiles = B2R_W3_PREFETCH_TILES;   // L2 prefetch distance beyond the tile being loaded

// Role timeline (debug builds only: nvcc -DB2R_TIMELINE, see tools/role_timeline.py): CTA 0 stamps clock64() for its
// first B2R_DBG_TILES tiles.  Compiled out of the product: even predicated off, each stamp cost the MMA warp ~10
// instructions on its critical path.
#ifdef B2R_TIMELINE
#define B2R_STAMP(iter, slot) \
    do { if (p.dbg != nullptr && blockIdx.x == 0 && (iter) < B2R_DBG_TILES) p.dbg[(iter) * 8 + (slot)] = clock64(); } while (0)
#else
#define B2R_STAMP(iter, slot) do { } while (0)
#endif

// kHead = true adds the fused 64 -> 3 output head; it is a separate instantiation because even unused, the extra
// epilogue code cost the plain layers ~5 % (A/B in one gpurun call, profiles/r01_w3_timeline.md).
//
// kPair = true (round 2) runs the same kernel as a cluster of two CTAs with tcgen05.mma.cta_group::2: one MMA of M = 256 x
// N = 192 serves two neighbouring tiles, each CTA holding its own tile's halo box (its 128 rows of A, its 128 lanes of D,
// its epilogue and stores) and HALF of the weight rows (96 of the 192 rows of a k-step, 32 of the 64 rows of a compact
// 1x1 k-step).  Why: the resident weights halve (36 KB for 64 -> 64; the 192 -> 64 layer fits without the streamed-weights
// mode: 375 -> 283 us) and the ring gets the space (128 -> 64: 915 -> 805 us); each SM also reads 4 + 3 KB instead of
// 4 + 6 KB of operands per MMA.  Single-group layers (64 -> 64) run at the same speed in both modes.
// Protocol as in conv_gemm_pair_kernel: the leader (cluster rank 0) issues (both issuer warps); its "data landed" and
// "weights landed" barriers count the TMA bytes of both CTAs; tcgen05.commit multicasts to both CTAs' "slot free" /
// "accumulator ready" barriers; the epilogue warps of both CTAs release the accumulator on the leader's barrier.
template <bool kHead, bool kPair = false>
__global__ void __launch_bounds__(kW3Threads, 1) conv_w3_kernel(const __grid_constant__ ConvW3Params p) {
    constexpr uint32_t kIdesc = make_idesc_bf16_f32(kPair ? 256 : 128, 192);
    constexpr uint32_t kIdesc64 = make_idesc_bf16_f32(kPair ? 256 : 128, 64);
    constexpr int kBStepCta = kPair ? kW3BStep / 2 : kW3BStep;   // bytes of one k-step's weights in THIS CTA
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_res = smem;
    uint8_t* ring = b_res + p.b_bytes;                             // resident weights, or a ring of weight k-steps
    uint8_t* sfull = ring + p.ring_slots * kW3Slot;                 // p.stage_bufs (2..4) staging (+ pool) tile buffers, used round-robin
    float* bias_s = reinterpret_cast<float*>(sfull + p.stage_bufs * p.stage_stride);
    float* head_s = bias_s + 64;                                   // [3][64] head weights + [3] head bias (+ pad)
    uint32_t* group_s = reinterpret_cast<uint32_t*>(head_s + 200);  // [kW3MaxGroups] group words (shared-memory copy)
    uint32_t* gboff_s = group_s + kW3MaxGroups;                     // [kW3MaxGroups] weights offset of each group (16-byte units)
    uint64_t* bars = reinterpret_cast<uint64_t*>(gboff_s + kW3MaxGroups);
    // "Data landed" barriers exist once per issuing warp: warp m only ever waits on full_bar[m][.], which the producer
    // arms for tiles of parity m, so it observes EVERY phase of the barriers it waits on.  (With one shared set, a
    // warp that skips the other warp's tile may skip a phase, and a parity wait cannot tell phase n from phase n - 2:
    // it passes immediately on a slot whose previous fill has not even arrived.)
    uint64_t* full_bar = bars;                          // [2][kN64MaxRing]
    uint64_t* empty_bar = bars + 2 * kN64MaxRing;       // [kN64MaxRing]
    uint64_t* tmem_full_bar = empty_bar + kN64MaxRing;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* b_full_bar = tmem_empty_bar + 2;
    uint64_t* bs_full_bar = b_full_bar + 1;                  // [2][kN64MaxRing] streamed-weights ring
    uint64_t* bs_empty_bar = bs_full_bar + 2 * kN64MaxRing;  // [kN64MaxRing]
    uint64_t* staged_bar = bs_empty_bar + kN64MaxRing;   // [kW3MaxStageBufs] staging buffer written by all 16 epilogue warps
    uint64_t* free_bar = staged_bar + kW3MaxStageBufs;   // [kW3MaxStageBufs] its TMA store has been read: reusable
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(free_bar + kW3MaxStageBufs);
    const int NB = p.stage_bufs;

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long total_tiles = long(p.tiles_w) * p.tiles_h * p.n_img;
    const int R = p.ring_slots;
    const int G = p.num_groups;
    // work units: single mode = tiles, one per CTA and step; pair mode = tile pairs, rank r of the cluster takes tile 2 u + r
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const long unit0 = kPair ? long(blockIdx.x >> 1) : long(blockIdx.x);
    const long ustride = kPair ? long(gridDim.x >> 1) : long(gridDim.x);
    const long total_units = kPair ? (total_tiles + 1) / 2 : total_tiles;
    const long tile0 = kPair ? 2 * unit0 + rank : unit0;          // this CTA's first tile and tile stride
    const long tstride = kPair ? 2 * ustride : ustride;

    if (warp_idx == 0 && lane == 0) {
        for (int i = 0; i < B2R_MAX_SRC; ++i) tma_prefetch_desc(&p.a_map[i]);
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.out_map);
        tma_prefetch_desc(&p.pool_map);
    }
    if (warp_idx == 1) {
        if (lane == 0) {
            for (int s = 0; s < R; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&full_bar[kN64MaxRing + s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], kPair ? 2 * kW3EpiWarps : kW3EpiWarps);
            }
            mbar_init(b_full_bar, 1);
            for (int s = 0; s < kW3MaxStageBufs; ++s) {
                mbar_init(&staged_bar[s], kW3EpiWarps);
                mbar_init(&free_bar[s], 1);
            }
            for (int s = 0; s < kN64MaxRing; ++s) {
                mbar_init(&bs_full_bar[s], 1);
                mbar_init(&bs_full_bar[kN64MaxRing + s], 1);
                mbar_init(&bs_empty_bar[s], 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        if constexpr (kPair) tmem_alloc_pair<512>(tmem_ptr_s);
        else tmem_alloc<512>(tmem_ptr_s);   // 2 accumulator stages x 192 columns (power-of-two allocation)
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) bias_s[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
    if (kHead && threadIdx.x >= 128 && threadIdx.x < 128 + 195)
        head_s[threadIdx.x - 128] = threadIdx.x - 128 < 192 ? p.head_w[threadIdx.x - 128] : p.head_b[threadIdx.x - 128 - 192];
    if (threadIdx.x >= 352 && threadIdx.x < 352 + kW3MaxGroups) {
        group_s[threadIdx.x - 352] = p.group[threadIdx.x - 352];
        gboff_s[threadIdx.x - 352] = p.group_boff[threadIdx.x - 352];
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    // Pair mode with an odd tile count: the last pair's second CTA walks one tile past the end (image index n_img).  Its TMA
    // loads are fully out of bounds (zero-filled, bytes still counted), its TMA stores are clipped away, and the head
    // variant's plain stores check the image index.

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            const bool stream_b = p.b_slots > 0;
            if (!stream_b) {
                // resident weights: 3x3 groups hold three 24 KB k-steps; a 1x1 group holds only the 64 kw = 1 rows of its
                // k-step (8 KB) unless it is the first group (whose MMAs must initialise all 192 accumulator columns).
                // Pair mode: this CTA's half of every block (rows 96 r .. of a k-step, rows 64 + 32 r .. of a compact block),
                // counted on the leader's barrier.
                if (leader) mbar_arrive_expect_tx(b_full_bar, uint32_t(p.b_bytes) * (kPair ? 2u : 1u));
                for (int g = 0; g < G; ++g) {
                    const uint32_t e = p.group[g];
                    const int ks0 = int(e >> 20);
                    uint8_t* dst = b_res + size_t(p.group_boff[g]) * 16;
                    if (((e >> 2) & 1) && g > 0) {
                        if constexpr (kPair) tma_load_2d_pair(dst, &p.b_map_c, b_full_bar, ks0 * 64, 64 + 32 * int(rank));
                        else tma_load_2d(dst, &p.b_map_c, b_full_bar, ks0 * 64, 64);
                    } else {
                        const int nk = ((e >> 2) & 1) ? 1 : 3;
                        for (int k = 0; k < nk; ++k) {
                            if constexpr (kPair) tma_load_2d_pair(dst + k * kBStepCta, &p.b_map, b_full_bar, (ks0 + k) * 64, 96 * int(rank));
                            else tma_load_2d(dst + k * kW3BStep, &p.b_map, b_full_bar, (ks0 + k) * 64, 0);
                        }
                    }
                }
            }
            int stage = 0;
            uint32_t phase = 0;
            int bs = 0;
            uint32_t bphase = 0;
            TileWalk tw;
            tw.init(tile0, tstride, p.tiles_w, p.tiles_h);
            // L2 prefetch cursor, kW3PrefetchTiles tiles ahead of the tile being loaded: the ring (2-4 slots, as few as
            // one tile for 128 -> 64) cannot cover HBM latency under load; without this the MMA warps waited ~270-800
            // cycles per tile for their first A box (profiles/r01_w3_timeline.md)
            TileWalk pf;
            long pf_tile = p.prefetch ? tile0 + (long)kW3PrefetchTiles * tstride : total_tiles;   // 0: no L2 prefetch
            pf.init(pf_tile < total_tiles ? pf_tile : 0, tstride, p.tiles_w, p.tiles_h);
            int par = 0;   // which issuing warp consumes this tile
            for (long unit = unit0; unit < total_units; unit += ustride, tw.next(p.tiles_w, p.tiles_h), par ^= 1) {
#ifndef B2R_EXP_NO_PREFETCH
                if (pf_tile < total_tiles) {
                    for (int g = 0; g < G; ++g) {
                        const uint32_t e = group_s[g];
                        tma_prefetch_l2_4d(&p.a_map[e & 3], int((e >> 8) & 0xFFF) * 64, pf.tw * 14 - 1, pf.th * 8 - 1, pf.n);
                    }
                }
                pf_tile += tstride;
                pf.next(p.tiles_w, p.tiles_h);
#endif
                const int n0 = tw.n, w0 = tw.tw * 14, h0 = tw.th * 8;
                uint64_t* full_m = full_bar + par * kN64MaxRing;
                uint64_t* bs_full_m = bs_full_bar + par * kN64MaxRing;
                for (int g = 0; g < G; ++g) {
                    const uint32_t e = group_s[g];
                    const int src = e & 3;
                    const int c0 = int((e >> 8) & 0xFFF) * 64;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
#ifdef B2R_EXP_NO_LOAD   // experiment builds only (tools/exp/README.md): which role paces the tile?
                    mbar_arrive(&full_m[stage]);
#else
                    if constexpr (kPair) {
                        if (leader) mbar_arrive_expect_tx(&full_m[stage], 2u * kW3Slot);   // the boxes of both CTAs
                        tma_load_4d_pair(ring + stage * kW3Slot, &p.a_map[src], &full_m[stage], c0, w0 - 1, h0 - 1, n0);
                    } else {
                        mbar_arrive_expect_tx(&full_m[stage], kW3Slot);
                        tma_load_4d(ring + stage * kW3Slot, &p.a_map[src], &full_m[stage], c0, w0 - 1, h0 - 1, n0);
                    }
#endif
                    if (++stage == R) {
                        stage = 0;
                        phase ^= 1;
                    }
                    if (stream_b) {
                        // weights too large to stay resident (e.g. 192 -> 64): this group's k-steps go through a ring
                        const int nk = ((e >> 2) & 1) ? 1 : 3;
                        const int ks0 = int(e >> 20);
                        for (int k = 0; k < nk; ++k) {
                            mbar_wait(&bs_empty_bar[bs], bphase ^ 1);
                            mbar_arrive_expect_tx(&bs_full_m[bs], kW3BStep);
                            tma_load_2d(b_res + bs * kW3BStep, &p.b_map, &bs_full_m[bs], (ks0 + k) * 64, 0);
                            if (++bs == p.b_slots) {
                                bs = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp_idx <= 2) {
        // ===================================== MMA issuers =====================================
        // TWO issuing warps, one per TMEM stage: warp 1 takes tiles 0, 2, 4, ... of this CTA, warp 2 tiles 1, 3, ...
        // The tensor pipe's queue only covers a few hundred cycles, and a single issuer needed ~650-880 cycles of plain
        // instructions (barrier probes, reconvergence, descriptor set-up; it competes with four busy epilogue warps for
        // its sub-partition's issue slots) between the last MMA of a tile and the first MMA of the next, during which
        // the pipe ran dry (profiles/r01_w3_timeline.md).  With two issuers that gap overlaps the other warp's MMAs.
        // Each warp: warp-uniform control flow with ONE elected lane issuing (descriptor arithmetic stays on the uniform
        // datapath), single-probe waits, group words from shared memory fetched one group ahead.
        const int m = warp_idx - 1;
        const bool stream_b = p.b_slots > 0;
        if (leader) {     // pair mode: the peer CTA's issuer warps have nothing to do
        if (!stream_b) mbar_wait_uniform(b_full_bar, 0);
        tc_fence_after();
        auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_) {
            if constexpr (kPair) umma_bf16_ss_pair(d, a, b, idesc, acc_);
            else umma_bf16_ss(d, a, b, idesc, acc_);
        };
        auto commit = [](uint64_t* bar) {
            if constexpr (kPair) umma_commit_pair(bar);
            else umma_commit(bar);
        };
        const uint64_t desc_hi = make_sdesc_sw128(0, 1024) & 0xFFFFFFFF00000000ull;   // SBO / version / swizzle bits
        const uint32_t a_lo0 = ((smem_u32(ring) >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t b_lo0 = ((smem_u32(b_res) >> 4) & 0x3FFF) | (1u << 16);
        // ring positions: a tile consumes G A-slots and num_ksteps B-slots; this warp starts m tiles in and then
        // skips the other warp's tile after each of its own
        int stage = 0, bs = 0;
        uint32_t phase_bits = 0, bphase_bits = 0;   // bit s: parity of this warp's next wait on slot s of its own barrier set
        uint64_t* full_m = full_bar + m * kN64MaxRing;
        uint64_t* bs_full_m = bs_full_bar + m * kN64MaxRing;
        auto skip_tile = [&]() {
            stage += G;
            while (stage >= R) stage -= R;
            if (stream_b) {
                bs += p.num_ksteps;
                while (bs >= p.b_slots) bs -= p.b_slots;
            }
        };
        if (m == 1) skip_tile();
        const uint32_t acc = uint32_t(m);
        const uint32_t tmem_d = tmem_base + acc * 256u;
        uint32_t acc_phase = 0;
        uint32_t e_next = group_s[0];
        [[maybe_unused]] int iter = m;
        for (long unit = unit0 + (long)m * ustride; unit < total_units; unit += 2L * ustride, iter += 2) {
            mbar_wait_uniform(&tmem_empty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            if (lane == 0) B2R_STAMP(iter, 1);
            uint32_t accum = 0;
            for (int g = 0; g < G; ++g) {
                const uint32_t e = e_next;
                e_next = group_s[g + 1 == G ? 0 : g + 1];
                const bool center = ((e >> 2) & 1) != 0;
                mbar_wait_uniform(&full_m[stage], (phase_bits >> stage) & 1u);
                phase_bits ^= 1u << stage;
                tc_fence_after();
                if (g == 0 && lane == 0) B2R_STAMP(iter, 7);
                const uint32_t a_lo = a_lo0 + uint32_t(stage) * uint32_t(kW3Slot >> 4);
                if (!stream_b) {
                    const uint32_t b_lo = b_lo0 + gboff_s[g];
                    if (elect_one()) {
                        if (center && accum) {
                            // 1x1 k-step (ResidualBlock shortcut) on kernel row 1: A = buffer rows 16 .. 143.  Its kw = 0
                            // and kw = 2 weight rows are zero, so only the kw = 1 third of the accumulator (columns
                            // 64..127, weight rows 64..127 of the k-step) is touched: an N = 64 MMA, a third of the
                            // work and 75.6 instead of 96 cycles.  (Needs an initialised accumulator: accum != 0.)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma(tmem_d + 64u, desc_hi | uint64_t(a_lo + 128u + 2u * k), desc_hi | uint64_t(b_lo + 2u * k),
                                    kIdesc64, 1u);   // compact block = the kw = 1 rows
                        } else if (center) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                mma(tmem_d, desc_hi | uint64_t(a_lo + 128u + 2u * k), desc_hi | uint64_t(b_lo + 2u * k),
                                    kIdesc, accum);
                                accum = 1;
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < 3; ++t) {   // kernel row t: A = buffer rows 16 t .. 16 t + 127 (2048 B apart)
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    mma(tmem_d, desc_hi | uint64_t(a_lo + 128u * t + 2u * k),
                                        desc_hi | uint64_t(b_lo + uint32_t(kBStepCta >> 4) * t + 2u * k), kIdesc, accum);
                                    accum = 1;
                                }
                            }
                        }
                        commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                } else {
                    const int nk = center ? 1 : 3;
                    for (int t = 0; t < nk; ++t) {
                        mbar_wait_uniform(&bs_full_m[bs], (bphase_bits >> bs) & 1u);
                        bphase_bits ^= 1u << bs;
                        tc_fence_after();
                        const uint32_t kh = center ? 1u : uint32_t(t);
                        const uint32_t b_lo = b_lo0 + uint32_t(bs) * uint32_t(kW3BStep >> 4);
                        if (elect_one()) {
                            if (center && accum) {   // see the resident-weights path: N = 64 on the kw = 1 third
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_ss(tmem_d + 64u, desc_hi | uint64_t(a_lo + 128u * kh + 2u * k),
                                                 desc_hi | uint64_t(b_lo + uint32_t((64 * 128) >> 4) + 2u * k), kIdesc64, 1u);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma_bf16_ss(tmem_d, desc_hi | uint64_t(a_lo + 128u * kh + 2u * k), desc_hi | uint64_t(b_lo + 2u * k),
                                                 kIdesc, accum);
                                    accum = 1;
                                }
                            }
                            umma_commit(&bs_empty_bar[bs]);
                            if (t == nk - 1) umma_commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        accum = 1;
                        if (++bs == p.b_slots) bs = 0;
                    }
                }
                accum = 1;
                if (++stage == R) stage = 0;
            }
            if (elect_one()) commit(&tmem_full_bar[acc]);
            __syncwarp();
            if (lane == 0) B2R_STAMP(iter, 2);
            acc_phase ^= 1u;
            skip_tile();   // the other issuer's tile
        }
        }   // leader
    } else if (warp_idx == 3 + kW3EpiWarps) {
        // ===================================== TMA store issuer =====================================
        // One thread: waits until all 16 epilogue warps have staged a tile, stores it, and hands the staging buffer
        // back once the store has read it.  Keeping this off the epilogue warps removed the last CTA-wide barrier and
        // ~340 cycles per tile from their critical path (profiles/r01_w3_timeline.md).
        if (!kHead && lane == 0) {
            TileWalk tw;
            tw.init(tile0, tstride, p.tiles_w, p.tiles_h);
            int iter = 0, buf = 0;
            uint32_t bph = 0;     // staging buffers are used round-robin: buffer `buf`, phase parity `bph` (flips when buf wraps)
            for (long unit = unit0; unit < total_units; unit += ustride, ++iter, tw.next(p.tiles_w, p.tiles_h)) {
                uint8_t* sfull_b = sfull + buf * p.stage_stride;
                mbar_wait(&staged_bar[buf], bph);
#if !defined(B2R_EXP_NO_STAGE) && !defined(B2R_EXP_NO_TMASTORE)
                const int w0 = tw.tw * 14, h0 = tw.th * 8;
                if (p.store_full) tma_store_4d(&p.out_map, sfull_b, 0, w0, h0, tw.n);
                if (p.store_pool) tma_store_4d(&p.pool_map, sfull_b + kW3Staging, 0, w0 >> 1, h0 >> 1, tw.n);
                tma_store_commit();
                tma_store_wait_read<0>();
#endif
                B2R_STAMP(iter, 6);
                mbar_arrive(&free_bar[buf]);
                if (++buf == NB) {
                    buf = 0;
                    bph ^= 1u;
                }
            }
            tma_store_wait_all<0>();
        }
    } else {
        // ===================================== epilogue =====================================
        // All 16 warps work on every tile: warp (quarter, cq) owns TMEM lanes 32 quarter.. (a warp may touch only the
        // lanes 32 (warp_idx % 4)..) and output channels 16 cq.. .  Its slice of the accumulator is 3 x 16 columns, so
        // it is drained with three loads issued back to back and the TMEM stage is handed back to the MMA warp BEFORE
        // any of the shift / activation / staging work (~150 cycles after the accumulator became ready; with two
        // passes per warp the release came ~700 cycles later and the MMA warp waited for it on every tile).
        // The warps never synchronise with each other: each stages its 32 pixels x 16 channels (and, for pooled
        // layers, the 2x2 maxima it can form with two lane exchanges: a warp holds tile rows 2q, 2q+1, i.e. complete
        // pooling windows) and arrives on the staging buffer's mbarrier; the store warp does the rest.
        const int e = warp_idx - 3;
        const int quarter = warp_idx & 3;
        const int cq = e >> 2;                         // channels cq*16 .. cq*16+15
        const int etid = e * 32 + lane;                // 0..511
        const int hh = quarter * 2 + (lane >> 4);      // tile row of this lane's pixel
        const int cc = lane & 15;                      // buffer column; output column w = cc is valid for cc < 14
        const int srow = hh * 14 + cc;                 // row of the 8 x 14 staging tile
        const bool valid = cc < 14;
        const int prow = quarter * 7 + (cc >> 1);      // row of the 4 x 7 pooled staging tile (lanes with even cc, lane < 16)
        const bool pool_lane = lane < 14 && (lane & 1) == 0;
        const uint32_t lane_base = uint32_t(quarter * 32) << 16;
        const int form = act_form(p.act, p.slope);
        const float ns = act_neg_slope(p.act, p.slope);
        float b16[16];
        {
            const uint32_t ba = smem_u32(bias_s + cq * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(b16[4 * i]), "=f"(b16[4 * i + 1]), "=f"(b16[4 * i + 2]), "=f"(b16[4 * i + 3])
                             : "r"(ba + 16 * i));
        }
        TileWalk tw;
        tw.init(tile0, tstride, p.tiles_w, p.tiles_h);
        int iter = 0, sb = 0;
        uint32_t sph = 0;     // staging buffer index / phase parity (round-robin over p.stage_bufs buffers)
        for (long unit = unit0; unit < total_units; unit += ustride, ++iter, tw.next(p.tiles_w, p.tiles_h)) {
            const uint32_t acc = uint32_t(iter) & 1u;
            mbar_wait_uniform(&tmem_full_bar[acc], (uint32_t(iter) >> 1) & 1u);
            tc_fence_after();
            if (etid == 0) B2R_STAMP(iter, 3);
            uint32_t d0[16], d1[16], d2[16];
            const uint32_t tacc = tmem_base + lane_base + acc * 256u + uint32_t(cq * 16);
            tmem_ld_32x16(tacc, d0);            // kw = 0 partial sums
            tmem_ld_32x16(tacc + 64u, d1);      // kw = 1
            tmem_ld_32x16(tacc + 128u, d2);     // kw = 2
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kPair) mbar_arrive_leader(&tmem_empty_bar[acc]);   // the leader's barrier, from either CTA
                else mbar_arrive(&tmem_empty_bar[acc]);
            }
            if (etid == 0) B2R_STAMP(iter, 4);
            uint8_t* sfull_b = sfull + sb * p.stage_stride;
            uint8_t* spool_b = sfull_b + kW3Staging;
#ifndef B2R_EXP_NO_STAGE
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(d1[j]), 1);
                const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[j]), 2);
                x[j] = ((__uint_as_float(d0[j]) + a1) + a2) + b16[j];
            }
            if (form == kActRelu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
            } else if (form == kActPrelu01) {   // max(x, slope x): see act_form()
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], ns * x[j]);
            } else if (form == kActGeneral) {
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = apply_act_ns(x[j], ns);
            }
            if (kHead) {
                // fused 64 -> 3 head on the fp32 activations: partial sums over this warp's 16 channels, parked in the
                // (otherwise unused: a fused head never stores the 64-channel tile) staging buffer as [cq][o][pixel row]
                float hs0 = 0.f, hs1 = 0.f, hs2 = 0.f;
                const uint32_t ha = smem_u32(head_s + cq * 16);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float w0[4], w1[4], w2[4];
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w0[0]), "=f"(w0[1]), "=f"(w0[2]), "=f"(w0[3]) : "r"(ha + 16 * j4));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w1[0]), "=f"(w1[1]), "=f"(w1[2]), "=f"(w1[3]) : "r"(ha + 256 + 16 * j4));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w2[0]), "=f"(w2[1]), "=f"(w2[2]), "=f"(w2[3]) : "r"(ha + 512 + 16 * j4));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        hs0 = fmaf(x[4 * j4 + j], w0[j], hs0);
                        hs1 = fmaf(x[4 * j4 + j], w1[j], hs1);
                        hs2 = fmaf(x[4 * j4 + j], w2[j], hs2);
                    }
                }
                float* part = reinterpret_cast<float*>(sfull_b) + (cq * 3) * 128 + quarter * 32 + lane;
                part[0] = hs0;
                part[128] = hs1;
                part[256] = hs2;
                // the only cross-warp exchange: four partial sums per pixel, all from warps of the SAME TMEM lane quarter, so each
                // quarter synchronises on its own named barrier (4 warps) instead of all 16 epilogue warps on one
                named_barrier_sync(1 + quarter, 4 * 32);
                if (cq < 3) {
                    // warp cq of the quarter finishes output channel o = cq for the quarter's 32 pixels: add the four partial
                    // sums (same order as before: bit-identical) + bias, then the reference's outputs
                    const int o = cq;
                    const int w = tw.tw * 14 + cc;
                    const int h = tw.th * 8 + hh;
                    if (valid && w < p.W && h < p.H && tw.n < p.n_img) {
                        const float* pr = reinterpret_cast<const float*>(sfull_b) + quarter * 32 + lane;
                        const float v = ((pr[o * 128] + pr[(3 + o) * 128]) + (pr[(6 + o) * 128] + pr[(9 + o) * 128])) + head_s[192 + o];
                        const size_t hw = size_t(p.H) * p.W, pix = size_t(h) * p.W + w;
                        if (p.head_f32) p.head_f32[(size_t(tw.n) * 3 + o) * hw + pix] = v;
                        if (p.head_u8)   // torch.clamp(x, 0, 1); (x * 255).astype(np.uint8): truncation (17:86-92)
                            p.head_u8[(size_t(tw.n) * hw + pix) * 3 + o] = static_cast<uint8_t>(fminf(fmaxf(v, 0.f), 1.f) * 255.0f);
                    }
                }
            } else {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
                // the store of this buffer's previous tile (stage_bufs tiles ago) has been read
                mbar_wait_uniform(&free_bar[sb], sph ^ 1u);
                if (p.store_full && valid) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int jj = cq * 2 + q;   // 16-byte chunk of the 128-byte staging row
                        const uint32_t addr = smem_u32(sfull_b) + uint32_t(srow * 128 + ((jj ^ (srow & 7)) << 4));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                                     "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                                     : "memory");
                    }
                }
                if (p.store_pool) {
                    // 2x2 max-pool in registers: the window of pooled pixel (q, cc/2) is lanes {l, l+1, l+16, l+17}
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        pk[j] = bf16x2_max(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 1));
                        pk[j] = bf16x2_max(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 16));
                    }
                    if (pool_lane) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int jj = cq * 2 + q;
                            const uint32_t addr = smem_u32(spool_b) + uint32_t(prow * 128 + ((jj ^ (prow & 7)) << 4));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                                         "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                                         : "memory");
                        }
                    }
                }
                fence_proxy_async_smem();
            }
#else
            (void)valid; (void)srow; (void)form; (void)ns; (void)b16; (void)d0; (void)d1; (void)d2; (void)spool_b; (void)sfull_b;
            (void)prow; (void)pool_lane; (void)hh;
            if (!kHead) mbar_wait_uniform(&free_bar[sb], sph ^ 1u);
#endif
            if (etid == 0) B2R_STAMP(iter, 5);
            if (!kHead) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&staged_bar[sb]);
            }
            if (++sb == NB) {
                sb = 0;
                sph ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still use its shared memory / barriers
    if (warp_idx == 1) {
        tc_fence_after();
        __syncwarp();
        if constexpr (kPair) tmem_dealloc_pair<512>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

size_t conv_w3_smem_bytes(size_t b_bytes, int ring_slots, size_t stage_stride, int stage_bufs) {
    return 1024 + b_bytes + size_t(ring_slots) * kW3Slot + size_t(stage_bufs) * stage_stride + 256 + 800 + 768;
}

template <bool kHead, bool kPair>
static int launch_w3_variant(const ConvW3Params& p, int grid, size_t smem, cudaStream_t stream) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kW3Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kPair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2R_CUDA(cudaLaunchKernelEx(&cfg, conv_w3_kernel<kHead, kPair>, p));
    return B2R_OK;
}

// grid = CTAs (pair mode: an even number, two per cluster)
int launch_conv_w3(const ConvW3Params& p, int grid, cudaStream_t stream, bool pair) {
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        B2R_CUDA(cudaFuncSetAttribute((conv_w3_kernel<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, kN64MaxSmem));
        B2R_CUDA(cudaFuncSetAttribute((conv_w3_kernel<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, kN64MaxSmem));
        B2R_CUDA(cudaFuncSetAttribute((conv_w3_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, kN64MaxSmem));
        B2R_CUDA(cudaFuncSetAttribute((conv_w3_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, kN64MaxSmem));
        if (dev < 64) attr_set[dev] = true;
    }
    const size_t smem = conv_w3_smem_bytes(size_t(p.b_bytes), p.ring_slots, size_t(p.stage_stride), p.stage_bufs);
    const bool head = p.head_w != nullptr;
    note_conv_kernel(head ? (pair ? "conv_w3_kernel<head,pair>" : "conv_w3_kernel<head>") : (pair ? "conv_w3_kernel<pair>" : "conv_w3_kernel"));
    int rc;
    if (head) rc = pair ? launch_w3_variant<true, true>(p, grid, smem, stream) : launch_w3_variant<true, false>(p, grid, smem, stream);
    else rc = pair ? launch_w3_variant<false, true>(p, grid, smem, stream) : launch_w3_variant<false, false>(p, grid, smem, stream);
    if (rc) return rc;
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

}  // namespace b2r
