// Microbenchmark: does shared-pipe (MIO) traffic from other warps of the SM slow down (a) tcgen05.mma issue and
// (b) mbarrier probes?  Warp 0 lane 0 issues N = 192 MMAs back to back; warp 4 (same SM sub-partition as warp 0)
// times mbarrier.try_wait on an already-completed barrier; W worker warps run one of:
//   kind 0: nothing   1: SHFL bursts (32 per iteration)   2: LDS.128 bursts   3: STS.128 bursts   4: FFMA only
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__global__ void __launch_bounds__(1024, 1) mio_kernel(long long* out, int mma_iters, int workers, int kind) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, done_bar;
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < (16384 + 192 * 128 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { mbar_init(&bar, 1); mbar_init(&done_bar, 1); stop = 0; fence_mbar_init(); mbar_arrive(&done_bar); }
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
            long long t1 = clock64();
            out[0] = t1 - t0;
            stop = 1;
        }
    } else if (warp == 4) {
        if (lane == 0) {
            long long total = 0; int n = 0;
            while (!stop) {
                long long t0 = clock64();
                uint32_t dep = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {   // dependent chain (the parity operand comes from the previous probe); phase 0 completed at start-up
                    uint32_t ok;
                    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                                 : "=r"(ok) : "r"(smem_u32(&done_bar)), "r"(dep & 2u) : "memory");
                    dep = ok;
                }
                total += clock64() - t0; n += 8;
                if (dep == 7u) out[62] = 1;
                __nanosleep(200);
            }
            out[1] = total; out[2] = n;
        }
    } else if (warp >= 8 && warp < 8 + workers) {
        float acc = float(lane);
        uint32_t x = lane;
        const uint32_t a = smem_u32(smem) + 16384 + 192 * 128 + uint32_t(threadIdx.x & 1023) * 16;
        while (!stop) {
            if (kind == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x += __shfl_down_sync(0xffffffffu, x, 1);
            } else if (kind == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    uint32_t t0, t1, t2, t3; (void)t1; (void)t2;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3) : "r"(a));
                    x += t0 ^ t3;
                }
            } else if (kind == 3) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(x), "r"(x), "r"(x) : "memory");
            } else if (kind == 4) {
#pragma unroll
                for (int i = 0; i < 64; ++i) acc = fmaf(acc, 1.0001f, 0.5f);
            }
            if (kind == 0) __nanosleep(100);
        }
        if (acc == 1.2345f || x == 0x12345u) out[63] = x;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 64 * sizeof(long long));
    const size_t smem = 1024 + 16384 + 192 * 128 + 32768;
    cudaFuncSetAttribute(mio_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[] = {"idle", "SHFL", "LDS.128", "STS.128", "FFMA"};
    for (int kind = 0; kind < 5; ++kind)
        for (int workers : {8, 16}) {
            cudaMemset(d, 0, 64 * sizeof(long long));
            mio_kernel<<<1, (8 + workers) * 32, smem>>>(d, 4000, workers, kind);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[3];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%-8s x %2d warps: %.1f cycles per N=192 MMA; mbarrier.try_wait (completed phase) %.1f cycles each\n", names[kind],
                   workers, double(h[0]) / 16000.0, h[2] ? double(h[1]) / double(h[2]) : 0.0);
        }
    return 0;
}
