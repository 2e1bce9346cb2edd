// Microbenchmark: sustained clock / power of back-to-back tcgen05.mma on all SMs, single-CTA (M = 128) vs CTA pairs
// (cta_group::2, M = 256, each CTA holding half of B), for N = 192 and N = 256, random operands resident in shared
// memory.  Under the board's power cap the clock the GPU settles at IS the energy per MMA: if pairs settle higher, moving
// half the B operand per SM is worth real throughput.  nvidia-smi is polled while the kernel runs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b2r;

template <int N, bool PAIR>
__global__ void __cluster_dims__(PAIR ? 2 : 1, 1, 1) __launch_bounds__(128, 1) k(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    constexpr int kBRows = PAIR ? N / 2 : N;
    constexpr int total = 4 * 16384 + 4 * kBRows * 128;   // four A tiles, four B k-blocks
    for (int i = threadIdx.x; i < total / 4; i += 128) {
        uint32_t h = (uint32_t(i) * 2654435761u) ^ (blockIdx.x * 0x9E3779B9u);
        h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807F807Fu) | 0x3F003F00u | ((h >> 8) & 0x00800080u);
    }
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
        __syncwarp();
        if (PAIR) tmem_alloc_pair<512>(&tmem_ptr); else tmem_alloc<512>(&tmem_ptr);
    }
    tc_fence_before(); __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    const bool issuer = threadIdx.x == 0 && (!PAIR || cluster_ctarank() == 0);
    if (issuer) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(PAIR ? 256 : 128, N);
        const uint32_t sa = smem_u32(smem), sb = sa + 4 * 16384;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint64_t ad = make_sdesc_sw128(sa + (it & 3) * 16384, 1024);
            const uint64_t bd = make_sdesc_sw128(sb + (it & 3) * kBRows * 128, 1024);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (PAIR) umma_bf16_ss_pair(tm + (it & 1) * 256, ad + 2 * kk, bd + 2 * kk, idesc, 1);
                else umma_bf16_ss(tm + (it & 1) * 256, ad + 2 * kk, bd + 2 * kk, idesc, 1);
            }
        }
        if (PAIR) umma_commit_pair(&bar); else umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    } else if (PAIR && threadIdx.x == 0) {
        mbar_wait(&bar, 0);   // the multicast commit arrives here too
    }
    tc_fence_before(); __syncthreads();
    if (PAIR) cluster_sync_all();
    if (threadIdx.x < 32) { tc_fence_after(); if (PAIR) tmem_dealloc_pair<512>(tm); else tmem_dealloc<512>(tm); }
}

static void sample(double* mhz, double* watts) {
    FILE* f = popen("nvidia-smi --id=0 --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits", "r");
    *mhz = *watts = 0;
    if (f) { if (fscanf(f, "%lf, %lf", mhz, watts) != 2) { *mhz = 0; } pclose(f); }
}

template <int N, bool PAIR>
static void run(long long* d, int iters) {
    constexpr int kBRows = PAIR ? N / 2 : N;
    const size_t smem = 1024 + 4 * 16384 + 4 * kBRows * 128;
    cudaFuncSetAttribute(k<N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<N, PAIR><<<148, 128, smem>>>(d, iters);
    cudaEventRecord(e1);
    double mhz[64], w[64]; int ns = 0;
    while (cudaEventQuery(e1) == cudaErrorNotReady && ns < 64) { sample(&mhz[ns], &w[ns]); ++ns; }
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; i += PAIR ? 2 : 1) mx = h[i] > mx ? h[i] : mx;
    // second half of the samples: the governor has settled
    double cm = 0, cw = 0; int c = 0;
    for (int i = ns / 2; i < ns; ++i) if (mhz[i] > 0) { cm += mhz[i]; cw += w[i]; ++c; }
    const double flop = 2.0 * 128 * N * 16 * 4.0 * iters * 148;   // every SM executes 128 x N x 16 per MMA in both modes
    printf("N=%3d %-6s: %.1f cycles per MMA, %.0f ms, %.0f TFLOP/s, settled at %.0f MHz / %.0f W (%d samples)  [%s]\n", N,
           PAIR ? "pair" : "single", double(mx) / (4.0 * iters), ms, flop / (ms * 1e-3) / 1e12, c ? cm / c : 0.0, c ? cw / c : 0.0, c,
           cudaGetErrorString(e));
}

int main(int argc, char** argv) {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    const int iters = argc > 1 ? atoi(argv[1]) : 6000000;   // ~ 1.5 s at N = 256
    run<192, false>(d, iters);
    run<192, true>(d, iters);
    run<256, false>(d, iters * 3 / 4);
    run<256, true>(d, iters * 3 / 4);
    run<192, false>(d, iters);
    return 0;
}
