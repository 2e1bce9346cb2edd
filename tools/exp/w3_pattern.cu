// Replicates conv_w3's MMA operand pattern in isolation: N = 192, 12 MMAs per tile (3 kernel rows x 4 k-steps of 16),
// A = halo box of 20 KB (kernel row offset 2 KB), B = 3 resident k-steps of 24 KB, TMEM stage alternates per tile.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__global__ void __launch_bounds__(128, 1) k(long long* out, int tiles, int ring, int same_b, int random_data, int commits) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, dummy[2];
    __shared__ uint32_t tmem_ptr;
    const int total = 3 * 24576 + ring * 20480;
    for (int i = threadIdx.x; i < total / 4; i += 128) {
        // zeros, or pseudo-random bf16 pairs in (-2, 2): operand toggling is what the tensor pipe's power depends on
        uint32_t h = (uint32_t(i) * 2654435761u) ^ (blockIdx.x * 0x9E3779B9u);
        h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13;
        const uint32_t v = (h & 0x807F807Fu) | 0x3F003F00u | ((h >> 8) & 0x00800080u);
        reinterpret_cast<uint32_t*>(smem)[i] = random_data ? v : 0u;
    }
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy[0], 1); mbar_init(&dummy[1], 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&tmem_ptr);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x < 32) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, 192);
        const uint64_t hi = make_sdesc_sw128(0, 1024) & 0xFFFFFFFF00000000ull;
        const uint32_t b0 = ((smem_u32(smem) >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t a0 = (((smem_u32(smem) + 3 * 24576) >> 4) & 0x3FFF) | (1u << 16);
        long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            const uint32_t a_lo = a0 + uint32_t(t % ring) * 1280u;
            const uint32_t tmd = tm + (t & 1) * 256;
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(tmd, hi | uint64_t(a_lo + 128u * r + 2u * kk),
                                     hi | uint64_t(b0 + (same_b ? 0u : 1536u * r) + 2u * kk), idesc, (r | kk) ? 1u : 0u);
                // conv_w3 commits twice per tile (ring slot free, accumulator ready); nobody waits on these here
                if (commits >= 1) umma_commit(&dummy[0]);
                if (commits >= 2) umma_commit(&dummy[1]);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait_warp(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    const int ring = 4;
    size_t smem = 1024 + 3 * 24576 + ring * 20480;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // grid 1 vs 148 CTAs, zero vs random operands, short vs long run: separates the instruction's own rate from
    // chip-level (power) throttling, which shows up as MORE SM CYCLES per MMA, not only as a lower clock
    for (int commits = 0; commits < 3; ++commits) {
        const int grid = 148, tiles = 20000;
        k<<<grid, 128, smem>>>(d, tiles, ring, 0, 1, commits);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("w3 pattern, %3d CTAs, random operands, %d tcgen05.commit per 12 MMAs: %.1f cycles per MMA (ideal 96)  [%s]\n", grid,
               commits, mx / (double(tiles) * 12), cudaGetErrorString(e));
    }
    return 0;
}
