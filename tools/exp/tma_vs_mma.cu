// Microbenchmark: does TMA traffic into / out of shared memory slow down tcgen05.mma (N = 192, SS operands)?
// Warp 0 lane 0 issues MMAs back to back; warp 1 lane 0 streams bulk copies global -> smem (mode 1), smem -> global
// (mode 2) or both (mode 3) with `slots` copies of `bytes` in flight.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(128, 1) tma_mma_kernel(long long* out, const uint8_t* gsrc, uint8_t* gdst, int mma_iters, int mode,
                                                          uint32_t bytes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* buf = smem + 16384 + 192 * 128;   // 4 x 32 KB copy slots
    __shared__ uint64_t bar, full[4];
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < (16384 + 192 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1); stop = 0; fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (warp == 0) {
        if (lane == 0) {
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
            out[0] = clock64() - t0;
            stop = 1;
        }
    } else if (warp == 1 && lane == 0 && mode != 0) {
        long long copied = 0;
        uint32_t ph[4] = {0, 0, 0, 0};
        long long off = (long long)blockIdx.x * (8 << 20);
        long long t0 = clock64();
        if (mode & 1) for (int s = 0; s < 4; ++s) { mbar_arrive_expect_tx(&full[s], bytes); bulk_g2s(buf + s * 32768, gsrc + off + s * 32768, bytes, &full[s]); }
        int s = 0;
        while (!stop) {
            if (mode & 1) {
                mbar_wait(&full[s], ph[s]); ph[s] ^= 1;
                copied += bytes;
            }
            if (mode & 2) {
                bulk_s2g(gdst + off + ((copied + s * 32768) & ((8 << 20) - 1) & ~32767LL), buf + s * 32768, bytes);
                tma_store_commit();
                tma_store_wait_read<2>();
                if (!(mode & 1)) copied += bytes;
            }
            if (mode & 1) {
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(buf + s * 32768, gsrc + off + ((copied + s * 32768) & ((8 << 20) - 1) & ~32767LL), bytes, &full[s]);
            }
            s = (s + 1) & 3;
        }
        out[1] = clock64() - t0;
        out[2] = copied;
        if (mode & 1) for (int i = 0; i < 4; ++i) mbar_wait(&full[i], ph[i]);   // drain before exit
        tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
    long long* d;
    uint8_t *src, *dst;
    cudaMalloc(&d, 64 * sizeof(long long));
    cudaMalloc(&src, 16 << 20);
    cudaMalloc(&dst, 16 << 20);
    cudaMemset(src, 1, 16 << 20);
    const size_t smem = 1024 + 16384 + 192 * 128 + 4 * 32768;
    cudaFuncSetAttribute(tma_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[] = {"no TMA", "global->smem", "smem->global", "both"};
    for (int mode = 0; mode < 4; ++mode)
        for (uint32_t bytes : {8192u, 32768u}) {
            if (mode == 0 && bytes != 8192u) continue;
            cudaMemset(d, 0, 64 * sizeof(long long));
            tma_mma_kernel<<<1, 128, smem>>>(d, src, dst, 4000, mode, bytes);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[3];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%-13s %5u B copies: %.1f cycles per N=192 MMA; TMA moved %.1f B/clk\n", names[mode], bytes, double(h[0]) / 16000.0,
                   h[1] ? double(h[2]) / double(h[1]) : 0.0);
        }
    return 0;
}
