// Does tcgen05.ld honour a lane base that is NOT a multiple of 32?  If lane (base + i) is what thread i reads, the tap-folded
// conv kernel's epilogue could fetch D1[w + 1] and D2[w + 2] directly instead of with 32 warp shuffles per thread and tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../image-restoration-for-road-sign-recognition-in-autonomous-driving_b200/csrc tmem_lane_offset.cu -o tmem_lane_offset
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

__global__ void k(int* out, int lane_off) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<32>(&tptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tptr + (uint32_t(warp * 32) << 16);
    uint32_t v = 1000u * warp + lane;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base), "r"(v) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(base + (uint32_t(lane_off) << 16)) : "memory");
    tmem_ld_wait();
    out[threadIdx.x] = int(r);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<32>(tptr); }
}

int main() {
    int* d;
    cudaMalloc(&d, 128 * 4);
    for (int off : {0, 1, 2}) {
        cudaMemset(d, 0xFF, 128 * 4);
        k<<<1, 128>>>(d, off);
        cudaError_t e = cudaDeviceSynchronize();
        int h[128];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("lane offset %d [%s]: warp0:", off, cudaGetErrorString(e));
        for (int i : {0, 1, 14, 15, 16, 29, 30, 31}) printf(" t%d=%d", i, h[i]);
        printf(" | warp1:");
        for (int i : {0, 1, 30, 31}) printf(" t%d=%d", i, h[32 + i]);
        printf(" | warp3: t30=%d t31=%d\n", h[96 + 30], h[96 + 31]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
