// Latency of (k MMAs + tcgen05.commit + mbarrier wait) round trips, N = 64, in isolation.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

template <int N>
__global__ void __launch_bounds__(128, 1) lat_kernel(long long* out, int rounds, int mmas_per_round) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, N);
        const uint32_t sa = smem_u32(smem);
        const uint64_t ad = make_sdesc_sw128(sa, 1024), bd = make_sdesc_sw128(sa + 16384, 1024);
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            for (int k = 0; k < mmas_per_round; ++k) umma_bf16_ss(tm, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, k > 0);
            umma_commit(&bar);
            mbar_wait(&bar, phase);
            phase ^= 1;
            tc_fence_after();
        }
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    size_t smem = 1024 + 16384 + 256 * 128;
    cudaFuncSetAttribute(lat_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(lat_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int m : {1, 4, 12, 36}) {
        lat_kernel<64><<<148, 128, smem>>>(d, 2000, m);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=64  %2d MMAs + commit + wait: %.0f cycles per round trip (MMA time alone %.0f)\n", m, h[0] / 2000.0, m * 75.6);
        lat_kernel<192><<<148, 128, smem>>>(d, 2000, m);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=192 %2d MMAs + commit + wait: %.0f cycles per round trip (MMA time alone %.0f)\n", m, h[0] / 2000.0, m * 96.0);
    }
    return 0;
}
