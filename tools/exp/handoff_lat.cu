// Microbenchmark: mbarrier hand-off latency between two warps (ping-pong), for the three ways of waiting:
//   0: mbarrier.test_wait spin          1: mbarrier.try_wait spin (default suspend)
//   2: the library's mbar_wait (one failed try_wait, then try_wait with a 20 us suspend-time hint)
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

template <int MODE>
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity) {
    if (MODE == 0) { while (!mbar_test_wait(bar, parity)) {} }
    else if (MODE == 1) { while (!mbar_try_wait(bar, parity)) {} }
    else mbar_wait(bar, parity);
}

template <int MODE>
__global__ void pingpong(long long* out, int iters, int extra_delay) {
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane != 0) return;
    if (warp == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            mbar_arrive(&bar[0]);
            wait_mode<MODE>(&bar[1], i & 1);
        }
        out[MODE] = clock64() - t0;
    } else if (warp == 1) {
        for (int i = 0; i < iters; ++i) {
            wait_mode<MODE>(&bar[0], i & 1);
            if (extra_delay) __nanosleep(extra_delay);   // make the waiter on the other side really go to sleep
            mbar_arrive(&bar[1]);
        }
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8 * sizeof(long long));
    const int iters = 2000;
    for (int delay : {0, 2000}) {
        cudaMemset(d, 0, 64);
        pingpong<0><<<1, 64>>>(d, iters, delay);
        pingpong<1><<<1, 64>>>(d, iters, delay);
        pingpong<2><<<1, 64>>>(d, iters, delay);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[3];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("peer delay %4d ns: round trip (2 hand-offs + delay) cycles: test_wait spin %.0f, try_wait spin %.0f, try_wait+hint %.0f\n",
               delay, double(h[0]) / iters, double(h[1]) / iters, double(h[2]) / iters);
    }
    return 0;
}
