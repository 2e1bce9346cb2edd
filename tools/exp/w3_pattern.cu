// Replicates conv_w3's MMA operand pattern in isolation: N = 192, 12 MMAs per tile (3 kernel rows x 4 k-steps of 16),
// A = halo box of 20 KB (kernel row offset 2 KB), B = 3 resident k-steps of 24 KB, TMEM stage alternates per tile.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__global__ void __launch_bounds__(128, 1) k(long long* out, int tiles, int ring, int same_b) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int total = 3 * 24576 + ring * 20480;
    for (int i = threadIdx.x; i < total / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
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
    for (int same_b = 0; same_b < 2; ++same_b) {
        k<<<148, 128, smem>>>(d, 2000, ring, same_b);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("w3 pattern, %s B: %.1f cycles per MMA (ideal 96)  [%s]\n", same_b ? "same" : "3 different", h[0] / (2000.0 * 12), cudaGetErrorString(e));
    }
    return 0;
}
