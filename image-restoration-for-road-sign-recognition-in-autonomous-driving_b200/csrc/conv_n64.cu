// tcgen05 conv3x3 specialised for C_out = 64 (the 224x224 / 112x112 layers of all three networks).
//
// Why a second kernel: with N = 64 an MMA over one 64-channel k-block lasts only 128 cycles, so the generic
// kernel (one 16 KB A box + one 8 KB B box per k-block) needs ~190 B/cycle/SM from L2 to keep the tensor pipe
// busy and is latency/L2-bound at ~33 % of peak (profiles/r01_ncu_conv64_v1.md: 11.1 GB of L2->SM traffic for
// 822 MB of input).  This kernel removes 3.7x of that traffic:
//   * the whole weight matrix [64][K] (K <= 768) is loaded ONCE per CTA and stays resident in shared memory;
//   * for each (source, 64-channel chunk) the producer loads three column-shifted halo boxes
//     (dw = -1, 0, +1; box = 64 ch x TW x (TH+2)) instead of nine tap boxes; the tap (kh, dw) operand is the 128
//     contiguous rows starting at row kh*TW of box dw — a 1024-byte aligned UMMA descriptor start, so no swizzle
//     phase games — i.e. 3 loads feed 9 k-blocks (36 MMAs);
//   * 1x1 centre groups (ResidualBlock shortcut / identity) use a plain TH x TW box.
// Everything else (TMEM double buffering, epilogue, fused 2x2 max-pool, TMA stores) matches conv_gemm.cu.
//
// Requirements checked by the dispatcher in conv_gemm.cu: C_out = cout_total = 64, NHWC output, one image per tile
// (tile_n = 1, tile 16x8 or 8x16), k-blocks ordered as groups of nine (dw-major, dh-minor) followed by centre blocks.
#include <cstring>

#include "b2r_internal.h"
#include "conv_common.cuh"
#include "ptx_sm100.cuh"

namespace b2r {

constexpr int kN64Threads = 192;
constexpr int kN64BBlock = 64 * 128;  // one k-block of weights: 64 rows x 128 B

__global__ void __launch_bounds__(kN64Threads, 1) conv_n64_kernel(const __grid_constant__ ConvN64Params p) {
    constexpr uint32_t kIdesc = make_idesc_bf16_f32(128, 64);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_res = smem;                                    // num_kblocks x 8 KB
    uint8_t* ring = b_res + p.num_kblocks * kN64BBlock;       // ring_slots x slot_bytes
    uint8_t* sfull = ring + p.ring_slots * p.slot_bytes;      // 16 KB
    uint8_t* spool = sfull + 16384;                           // 4 KB
    float* bias_s = reinterpret_cast<float*>(spool + 4096);   // 64 floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 64);
    uint64_t* full_bar = bars;                       // [kN64MaxRing]
    uint64_t* empty_bar = bars + kN64MaxRing;        // [kN64MaxRing]
    uint64_t* tmem_full_bar = bars + 2 * kN64MaxRing;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* b_full_bar = tmem_empty_bar + 2;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(b_full_bar + 1);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int total_tiles = tiles_per_img * p.n_img;
    const int R = p.ring_slots;

    if (warp_idx == 0 && lane == 0) {
        for (int i = 0; i < B2R_MAX_SRC; ++i) {
            tma_prefetch_desc(&p.a3_map[i]);
            tma_prefetch_desc(&p.a1_map[i]);
        }
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.out_map);
        tma_prefetch_desc(&p.pool_map);
    }
    if (warp_idx == 1) {
        if (lane == 0) {
            for (int s = 0; s < R; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);
            }
            mbar_init(b_full_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<128>(tmem_ptr_s);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) bias_s[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            // resident weights: one barrier, num_kblocks boxes of 64 x 64
            mbar_arrive_expect_tx(b_full_bar, uint32_t(p.num_kblocks) * kN64BBlock);
            for (int kb = 0; kb < p.num_kblocks; ++kb)
                tma_load_2d(b_res + kb * kN64BBlock, &p.b_map, b_full_bar, kb * 64, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n0 = tile / tiles_per_img;
                const int t = tile - n0 * tiles_per_img;
                const int w0 = (t % p.tiles_w) * p.tile_w;
                const int h0 = (t / p.tiles_w) * p.tile_h;
                for (int s = 0; s < p.num_slots; ++s) {
                    const uint32_t e = p.slot[s];
                    const int src = e & 3, center = (e >> 2) & 1, dw = int((e >> 4) & 3) - 1;
                    const int c0 = int((e >> 8) & 0xFFF) * 64;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* dst = ring + stage * p.slot_bytes;
                    if (center) {
                        mbar_arrive_expect_tx(&full_bar[stage], 128 * 128);
                        tma_load_4d(dst, &p.a1_map[src], &full_bar[stage], c0, w0, h0, n0);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], uint32_t(p.slot_bytes));
                        tma_load_4d(dst, &p.a3_map[src], &full_bar[stage], c0, w0 + dw, h0 - 1, n0);
                    }
                    if (++stage == R) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            mbar_wait(b_full_bar, 0);
            tc_fence_after();
            const uint32_t b_base = smem_u32(b_res);
            const uint32_t tap_stride = uint32_t(p.tile_w) * 128u;  // rows kh*TW .. kh*TW+127 of the column box
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * 64);
                uint32_t first = 1;
                for (int s = 0; s < p.num_slots; ++s) {
                    const uint32_t e = p.slot[s];
                    const int center = (e >> 2) & 1;
                    const int kb0 = int(e >> 20);
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + stage * p.slot_bytes);
                    const int ntaps = center ? 1 : 3;
                    for (int t = 0; t < ntaps; ++t) {
                        const uint64_t adesc = make_sdesc_sw128(sa + uint32_t(t) * tap_stride, 1024);
                        const uint64_t bdesc = make_sdesc_sw128(b_base + uint32_t(kb0 + t) * kN64BBlock, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, first ? 0u : 1u);
                            first = 0;
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == R) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================================== epilogue =====================================
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;
        const int epi_tid = row;
        const uint32_t lane_base = uint32_t(quarter * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int tw = p.tile_w, th = p.tile_h;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n0 = tile / tiles_per_img;
            const int t = tile - n0 * tiles_per_img;
            const int w0 = (t % p.tiles_w) * p.tile_w;
            const int h0 = (t / p.tiles_w) * p.tile_h;

            mbar_wait_warp(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (epi_tid == 0) tma_store_wait_read<0>();  // single staging buffer: previous tile's store has read it
            named_barrier_sync(1, kEpiThreadsC);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_base + uint32_t(acc * 64 + half * 32), v);
                tmem_ld_wait();
                float b32[32];
                lds_bias32(bias_s + half * 32, b32);
                epilogue_store_half(v, b32, p.act, p.slope, sfull, row, half);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            fence_proxy_async_smem();
            named_barrier_sync(1, kEpiThreadsC);
            if (p.store_pool) {
                epilogue_pool_chunk(sfull, spool, epi_tid, tw, th);
                fence_proxy_async_smem();
                named_barrier_sync(1, kEpiThreadsC);
            }
            if (epi_tid == 0) {
                if (p.store_full) tma_store_4d(&p.out_map, sfull, 0, w0, h0, n0);
                if (p.store_pool) tma_store_4d(&p.pool_map, spool, 0, w0 >> 1, h0 >> 1, n0);
                tma_store_commit();
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (epi_tid == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc<128>(tmem_base);
    }
}

int launch_conv_n64(const ConvN64Params& p, int grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        B2R_CUDA(cudaFuncSetAttribute(conv_n64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kN64MaxSmem));
        if (dev < 64) attr_set[dev] = true;
    }
    note_conv_kernel("conv_n64_kernel");
    conv_n64_kernel<<<grid, kN64Threads, smem_bytes, stream>>>(p);
    B2R_CHECK_LAUNCH();
    return B2R_OK;
}

}  // namespace b2r
