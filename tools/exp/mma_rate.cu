// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 x bf16 -> f32, M = 128, K = 16, SS operands in smem) as a
// function of N, with nothing else using shared memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I ../../image-restoration-for-road-sign-recognition-in-autonomous-driving_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int distinct_stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (16384 + N * 128) * distinct_stages / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int st = it % distinct_stages;
            const uint32_t sa = smem_u32(smem + st * (16384 + N * 128));
            const uint64_t ad = make_sdesc_sw128(sa, 1024), bd = make_sdesc_sw128(sa + 16384, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tm + (it & 1) * N, ad + 2 * k, bd + 2 * k, idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel_uniform(long long* out, int iters, int distinct_stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (16384 + N * 128) * distinct_stages / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x < 32) {   // whole warp runs the loop; one elected lane issues
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, N);
        long long t0 = clock64();
        const uint32_t base = smem_u32(smem);
        for (int it = 0; it < iters; ++it) {
            const int st = it % distinct_stages;
            const uint32_t sa = base + st * (16384 + N * 128);
            const uint64_t ad = make_sdesc_sw128(sa, 1024), bd = make_sdesc_sw128(sa + 16384, 1024);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tm + (it & 1) * N, ad + 2 * k, bd + 2 * k, idesc, 1);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N>
void run(int grid, int iters, int stages) {
    long long* d; cudaMalloc(&d, grid * sizeof(long long));
    size_t smem = 1024 + (16384 + N * 128) * stages;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rate_kernel<N><<<grid, 128, smem>>>(d, iters, stages);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    printf("N=%3d grid=%3d stages=%d: lane0-branch %.1f cycles per MMA (128xNx16), ideal %.0f  [%s]\n", N, grid, stages,
           avg / (iters * 4.0), N / 2.0, cudaGetErrorString(e));
    cudaFuncSetAttribute(rate_kernel_uniform<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rate_kernel_uniform<N><<<grid, 128, smem>>>(d, iters, stages);
    e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    printf("N=%3d grid=%3d stages=%d: warp-uniform %.1f cycles per MMA  [%s]\n", N, grid, stages, avg / (iters * 4.0), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {148}) {
        run<32>(grid, 4000, 1); run<64>(grid, 4000, 1); run<96>(grid, 4000, 1); run<128>(grid, 4000, 1);
        run<160>(grid, 4000, 1); run<192>(grid, 4000, 1); run<224>(grid, 4000, 1); run<256>(grid, 4000, 1);
    }
    return 0;
}
