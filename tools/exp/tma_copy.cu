// How fast can the tap-folded conv kernel's MEMORY pattern run without any MMA?  Persistent CTAs, one per SM: TMA-load the
// (8 + 2) x 16-pixel x 64-channel halo box of tile t (20 KB, neighbouring boxes overlap by two columns / rows) into a
// shared-memory ring, TMA-store an 8 x 14-pixel box (14 KB) from it to a second NHWC tensor.  224 x 224 x 64 bf16 images.
// Prints TB/s of algorithmic traffic (6.4 MB read + 6.4 MB written per image), to compare with the 3.6 TB/s the 64 -> 64
// layers reach (tools/layer_bench.py) and the 6.5 TB/s of a plain copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../image-restoration-for-road-sign-recognition-in-autonomous-driving_b200/csrc tma_copy.cu -o tma_copy -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
#include "conv_common.cuh"
using namespace b2r;

struct P {
    CUtensorMap in_map, out_map;
    int tiles_w, tiles_h, n_img, ring, load_rows, prefetch;
};

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8], empty[8];
    const int slot = 10 * 16 * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const long total = long(p.tiles_w) * p.tiles_h * p.n_img;
    if (threadIdx.x == 0) {            // producer
        TileWalk tw, pf;
        tw.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h);
        long pft = blockIdx.x + long(p.prefetch) * gridDim.x;
        pf.init(pft < total ? pft : 0, gridDim.x, p.tiles_w, p.tiles_h);
        int s = 0; uint32_t ph = 0;
        for (long t = blockIdx.x; t < total; t += gridDim.x, tw.next(p.tiles_w, p.tiles_h)) {
            if (p.prefetch && pft < total) tma_prefetch_l2_4d(&p.in_map, 0, pf.tw * 14 - 1, pf.th * 8 - 1, pf.n);
            pft += gridDim.x; pf.next(p.tiles_w, p.tiles_h);
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], p.load_rows * 16 * 128);
            tma_load_4d(smem + s * slot, &p.in_map, &full[s], 0, tw.tw * 14 - 1, tw.th * 8 - 1, tw.n);
            if (++s == p.ring) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {    // consumer: store 8 x 14 from the box (rows 1..8, columns 1..14 of it: just bytes here)
        TileWalk tw;
        tw.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h);
        int s = 0; uint32_t ph = 0;
        for (long t = blockIdx.x; t < total; t += gridDim.x, tw.next(p.tiles_w, p.tiles_h)) {
            mbar_wait(&full[s], ph);
            tma_store_4d(&p.out_map, smem + s * slot, 0, tw.tw * 14, tw.th * 8, tw.n);
            tma_store_commit();
            tma_store_wait_read<0>();
            mbar_arrive(&empty[s]);
            if (++s == p.ring) { s = 0; ph ^= 1; }
        }
        tma_store_wait_all<0>();
    }
}

static void enc(CUtensorMap* m, void* base, int N, int H, int W, uint32_t bw, uint32_t bh) {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t str[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, bw, bh, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
    const int N = 128, H = 224, W = 224;
    const size_t bytes = size_t(N) * H * W * 128;
    void *a, *b;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaMemset(a, 1, bytes);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rows : {10, 8}) for (int ring : {2, 4, 8}) for (int pf : {0, 4}) {
        P p;
        enc(&p.in_map, a, N, H, W, 16, rows);
        enc(&p.out_map, b, N, H, W, 14, 8);
        p.tiles_w = 16; p.tiles_h = 28; p.n_img = N; p.ring = ring; p.load_rows = rows; p.prefetch = pf;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 2; ++it) k<<<148, 128, 1024 + ring * 20480>>>(p);
        cudaEventRecord(e0);
        for (int it = 0; it < 10; ++it) k<<<148, 128, 1024 + ring * 20480>>>(p);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("box rows %2d ring %d L2-prefetch %d: %7.1f us per 128 images, %.2f TB/s algorithmic (read + write) [%s]\n", rows, ring, pf,
               ms * 1e3, 2.0 * bytes / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
