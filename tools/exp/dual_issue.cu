// Microbenchmark: two warps issuing tcgen05.mma (N = 192) to alternate TMEM stages.  Each warp: 12 MMAs, commit to its
// own mbarrier, wait for that commit (stands in for "epilogue drained my stage"), repeat.  One warp alone exposes
// the commit round trip after every tile; do two warps overlap each other's round trips?
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__global__ void __launch_bounds__(128, 1) k(long long* out, int tiles, int issuers, int wait_lag) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_ptr;
    const int total = 3 * 24576 + 4 * 20480;
    for (int i = threadIdx.x; i < total / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F003F00u + i;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&tmem_ptr);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp < issuers) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, 192);
        const uint64_t hi = make_sdesc_sw128(0, 1024) & 0xFFFFFFFF00000000ull;
        const uint32_t b0 = ((smem_u32(smem) >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t a0 = (((smem_u32(smem) + 3 * 24576) >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t tmd = tm + warp * 256;
        long long t0 = clock64();
        uint32_t ph = 0;
        int committed = 0, waited = 0;
        for (int t = warp; t < tiles; t += issuers) {
            // wait until the tile issued `wait_lag` tiles ago (by this warp) has completed
            if (committed - waited > wait_lag) { mbar_wait_uniform(&bar[warp], ph); ph ^= 1; ++waited; }
            const uint32_t a_lo = a0 + uint32_t(t & 3) * 1280u;
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(tmd, hi | uint64_t(a_lo + 128u * r + 2u * kk), hi | uint64_t(b0 + 1536u * r + 2u * kk), idesc,
                                     (r | kk) ? 1u : 0u);
                umma_commit(&bar[warp]);
            }
            __syncwarp();
            ++committed;
        }
        while (waited < committed) { mbar_wait_uniform(&bar[warp], ph); ph ^= 1; ++waited; }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x * 2 + warp] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
    long long* d; cudaMalloc(&d, 2 * 148 * sizeof(long long));
    size_t smem = 1024 + 3 * 24576 + 4 * 20480;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = 4000;
    for (int issuers = 1; issuers <= 2; ++issuers)
        for (int lag = 0; lag < 2; ++lag) {
            cudaMemset(d, 0, 2 * 148 * sizeof(long long));
            k<<<148, 128, smem>>>(d, tiles, issuers, lag);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < 296; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%d issuer(s), each waits for its tile %d back: %.1f cycles per tile of 12 MMAs (ideal 1152)  [%s]\n", issuers,
                   lag + 1, mx / double(tiles), cudaGetErrorString(e));
        }
    return 0;
}
