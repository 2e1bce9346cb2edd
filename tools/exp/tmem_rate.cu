// Microbenchmark: tcgen05.ld (32x32b.x32) drain rate of TMEM, alone and while one thread keeps the tensor pipe busy
// with N = 192 MMAs (conv_w3's shape).  W reader warps, warp w reads lane quarter w % 4.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__global__ void __launch_bounds__(1024, 1) tmem_rate_kernel(long long* out, int iters, int readers, int mma_on, int mma_iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (16384 + 192 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (warp == 0) {
        if (lane == 0 && mma_on) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(128, 192);
            const uint32_t sa = smem_u32(smem);
            const uint64_t ad = make_sdesc_sw128(sa, 1024), bd = make_sdesc_sw128(sa + 16384, 1024);
            long long t0 = clock64();
            for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tm + (it & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, 1);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            long long t1 = clock64();
            out[0] = t1 - t0;
        }
    } else if (warp <= readers) {
        const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t v[32];
            tmem_ld_32x32(tm + lane_base + uint32_t((it * 32) & 255), v);
            tmem_ld_wait();
            acc ^= v[0] ^ v[31];
        }
        long long t1 = clock64();
        if (lane == 0) out[warp] = t1 - t0;
        if (acc == 0x12345678u) out[63] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 64 * sizeof(long long));
    const size_t smem = 1024 + 16384 + 192 * 128;
    cudaFuncSetAttribute(tmem_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    for (int mma_on = 0; mma_on < 2; ++mma_on)
        for (int readers : {0, 1, 4, 8, 16}) {
            if (!mma_on && readers == 0) continue;
            // MMA stream sized to outlast the readers (96 cycles per MMA, 4 per iteration)
            const int mma_iters = readers ? iters * 64 * (readers < 4 ? 1 : readers / 4) / 384 * 3 / 2 + 200 : 4000;
            cudaMemset(d, 0, 64 * sizeof(long long));
            tmem_rate_kernel<<<1, (1 + (readers ? readers : 1)) * 32, smem>>>(d, iters, readers, mma_on, mma_iters);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[64];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long worst = 0;
            for (int w = 1; w <= readers; ++w) worst = h[w] > worst ? h[w] : worst;
            printf("mma=%d readers=%2d: ", mma_on, readers);
            if (readers) printf("%.1f cycles per LDTM.x32 per warp, %.1f B/clk per SM total;  ", double(worst) / iters,
                                double(readers) * iters * 4096.0 / double(worst));
            if (mma_on) printf("%.1f cycles per N=192 MMA (%d MMAs)", double(h[0]) / (4.0 * mma_iters), 4 * mma_iters);
            printf("\n");
        }
    return 0;
}
