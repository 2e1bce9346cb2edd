// How fast can a WRITE-ONLY stream go?  The first layers (C_in = 3) write 43x more than they read and reach 3.45 TB/s; torch's
// fill / memset kernels top out at 3.9 TB/s (hbm_rw.py), but the ConvTranspose kernel was seen writing 4.1-4.7 TB/s next to
// 1.1 TB/s of reads.  Persistent CTAs, one per SM, no loads at all:
//   mode 0: one thread TMA-stores 16 KB boxes (8 x 16 pixels x 64 bf16 channels, as the conv epilogues do) with `depth` bulk
//           groups in flight;
//   mode 1: all threads st.global.v4 (fully coalesced 512 B per warp instruction) over the same tensor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../image-restoration-for-road-sign-recognition-in-autonomous-driving_b200/csrc tma_write.cu -o tma_write -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
#include "conv_common.cuh"
using namespace b2r;

struct P {
    CUtensorMap out_map;
    int tiles_w, tiles_h, n_img, depth;
    uint4* out;
    size_t n16;
};

template <int DEPTH>
__device__ void store_loop(const P& p, uint8_t* smem) {
    TileWalk tw;
    tw.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h);
    const long total = long(p.tiles_w) * p.tiles_h * p.n_img;
    int s = 0;
    for (long t = blockIdx.x; t < total; t += gridDim.x, tw.next(p.tiles_w, p.tiles_h)) {
        tma_store_4d(&p.out_map, smem + s * 16384, 0, tw.tw * 16, tw.th * 8, tw.n);
        tma_store_commit();
        tma_store_wait_read<DEPTH - 1>();
        if (++s == DEPTH) s = 0;
    }
    tma_store_wait_all<0>();
}

__global__ void __launch_bounds__(256, 1) k_tma(const __grid_constant__ P p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    for (int i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        switch (p.depth) {
            case 1: store_loop<1>(p, smem); break;
            case 2: store_loop<2>(p, smem); break;
            case 4: store_loop<4>(p, smem); break;
            default: store_loop<8>(p, smem); break;
        }
    }
}

__global__ void __launch_bounds__(256, 1) k_stg(const __grid_constant__ P p) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n16; i += size_t(gridDim.x) * blockDim.x) p.out[i] = v;
}

int main() {
    const int N = 128, H = 224, W = 224;
    const size_t bytes = size_t(N) * H * W * 128;
    void* b;
    cudaMalloc(&b, bytes);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    P p;
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t str[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, 16, 8, 1}, es[4] = {1, 1, 1, 1};
    if (cuTensorMapEncodeTiled(&p.out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, b, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return 1;
    }
    p.tiles_w = 14; p.tiles_h = 28; p.n_img = N; p.out = static_cast<uint4*>(b); p.n16 = bytes / 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (int depth : {1, 2, 4, 8})
            for (int ctas : {148, 296}) {
                if (mode == 1 && depth != 1) continue;
                if (mode == 0 && ctas != 148) continue;
                p.depth = depth;
                auto run = [&]() {
                    if (mode == 0) k_tma<<<148, 256, 1024 + 8 * 16384>>>(p);
                    else k_stg<<<ctas, 256>>>(p);
                };
                for (int it = 0; it < 2; ++it) run();
                cudaEventRecord(e0);
                for (int it = 0; it < 10; ++it) run();
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                ms /= 10;
                printf("%s depth %d ctas %d: %7.1f us per 822 MB, %.2f TB/s written [%s]\n", mode == 0 ? "TMA store 16 KB boxes" : "st.global.v4 coalesced",
                       depth, ctas, ms * 1e3, bytes / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
