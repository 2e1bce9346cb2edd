/* Torch-free use of libb2r.so: the C-ABI of include/b2r.h driven from plain C with CUDA runtime allocations.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/cabi_degrade_metrics.c -o /tmp/cabi_demo \
 *       -L /usr/local/cuda/lib64 -lcudart -ldl -lm
 *   /tmp/cabi_demo image-restoration-for-road-sign-recognition-in-autonomous-driving_b200/libb2r.so
 *
 * What it does (the fog stage of 16_gen_compound_data.py:30-31,37 and the metrics of 08_run_inference.py:118-123):
 *   1. b2r_degrade on a batch of u8 images with fog only (t = 0.5, A = 0.9) -> checked byte for byte against the
 *      reference arithmetic restated here in C: float32 x = v / 255; x = x * t + A(1 - t); trunc(clip(x * 255)).
 *   2. b2r_sse_u8 and b2r_ssim_u8 between the clean and the fogged batch -> SSE checked exactly, SSIM checked to be in
 *      (0, 1) and to be exactly 1 for an image against itself.
 * Exit code 0 = all checks passed.  This is the binding a C / C++ / cgo host would write: dlopen + plain pointers. */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "b2r.h"

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                           \
            return 2;                                                                             \
        }                                                                                         \
    } while (0)

typedef int (*degrade_fn)(const uint8_t*, uint8_t*, int, int, int, const float*, const int32_t*, const float*,
                          const float*, const int32_t*, const float*, const double*, uint64_t, uint64_t, int, int, void*);
typedef int (*sse_fn)(const uint8_t*, const uint8_t*, uint64_t*, int, int64_t, void*);
typedef int (*ssim_fn)(const uint8_t*, const uint8_t*, double*, int, int, int, int, double, void*);
typedef const char* (*err_fn)(void);

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "libb2r.so";
    void* lib = dlopen(path, RTLD_NOW);
    if (!lib) {
        fprintf(stderr, "dlopen %s: %s\n", path, dlerror());
        return 2;
    }
    degrade_fn degrade = (degrade_fn)dlsym(lib, "b2r_degrade");
    sse_fn sse = (sse_fn)dlsym(lib, "b2r_sse_u8");
    ssim_fn ssim = (ssim_fn)dlsym(lib, "b2r_ssim_u8");
    err_fn last_error = (err_fn)dlsym(lib, "b2r_last_error");
    if (!degrade || !sse || !ssim || !last_error) {
        fprintf(stderr, "missing symbol\n");
        return 2;
    }

    const int N = 8, H = 64, W = 80;
    const size_t elems = (size_t)H * W * 3, bytes = (size_t)N * elems;
    uint8_t* h_in = (uint8_t*)malloc(bytes);
    uint8_t* h_out = (uint8_t*)malloc(bytes);
    uint32_t s = 12345u;
    for (size_t i = 0; i < bytes; ++i) {   /* smooth-ish content: a ramp plus LCG noise */
        s = s * 1664525u + 1013904223u;
        h_in[i] = (uint8_t)(((i / 3) % W) * 255 / W / 2 + (s >> 25));
    }
    float h_t[8], h_add[8];
    int32_t h_on[8];
    for (int n = 0; n < N; ++n) {
        const double t = 0.5 + 0.04 * n;          /* per-image transmission */
        h_t[n] = (float)t;
        h_add[n] = (float)(0.9 * (1.0 - t));       /* A * (1 - t) evaluated in double, as Python does (16:31) */
        h_on[n] = n != N - 1;                       /* last image passes through untouched */
    }

    uint8_t *d_in, *d_out;
    float *d_t, *d_add;
    int32_t* d_on;
    uint64_t* d_sse;
    double* d_ssim;
    CK(cudaMalloc((void**)&d_in, bytes));
    CK(cudaMalloc((void**)&d_out, bytes));
    CK(cudaMalloc((void**)&d_t, sizeof h_t));
    CK(cudaMalloc((void**)&d_add, sizeof h_add));
    CK(cudaMalloc((void**)&d_on, sizeof h_on));
    CK(cudaMalloc((void**)&d_sse, N * sizeof(uint64_t)));
    CK(cudaMalloc((void**)&d_ssim, N * sizeof(double)));
    CK(cudaMemcpy(d_in, h_in, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_t, h_t, sizeof h_t, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_add, h_add, sizeof h_add, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_on, h_on, sizeof h_on, cudaMemcpyHostToDevice));
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));

    /* fog only: no taps / ksize (nothing blurs), no sigma (no noise) */
    int rc = degrade(d_in, d_out, N, H, W, NULL, NULL, d_t, d_add, d_on, NULL, NULL, 0, 0, B2R_ORDER_BLUR_FOG_NOISE, 0,
                     (void*)stream);
    if (rc) {
        fprintf(stderr, "b2r_degrade: %d %s\n", rc, last_error());
        return 1;
    }
    rc = sse(d_in, d_out, d_sse, N, (int64_t)elems, (void*)stream);
    if (!rc) rc = ssim(d_in, d_out, d_ssim, N, H, W, 3, 255.0, (void*)stream);
    if (rc) {
        fprintf(stderr, "metrics: %d %s\n", rc, last_error());
        return 1;
    }
    uint64_t h_sse[8];
    double h_ssim[8];
    CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_sse, d_sse, sizeof h_sse, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_ssim, d_ssim, sizeof h_ssim, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));

    int bad = 0;
    for (int n = 0; n < N; ++n) {
        uint64_t ref_sse = 0;
        for (size_t i = 0; i < elems; ++i) {
            const uint8_t v = h_in[n * elems + i];
            volatile float x = (float)v / 255.0f;
            if (h_on[n]) {
                volatile float m = x * h_t[n];      /* two roundings, as NumPy evaluates img * t + A * (1 - t) */
                x = m + h_add[n];
            }
            volatile float y = x * 255.0f;
            if (y < 0.f) y = 0.f;
            if (y > 255.f) y = 255.f;
            const uint8_t ref = (uint8_t)y;         /* truncation */
            const uint8_t got = h_out[n * elems + i];
            if (ref != got && bad++ < 5) fprintf(stderr, "image %d byte %zu: %u != %u\n", n, i, got, ref);
            const int d = (int)v - (int)got;
            ref_sse += (uint64_t)(d * d);
        }
        if (ref_sse != h_sse[n]) {
            fprintf(stderr, "image %d: sse %llu != %llu\n", n, (unsigned long long)h_sse[n], (unsigned long long)ref_sse);
            ++bad;
        }
        const int ident = !h_on[n];
        if (ident ? h_ssim[n] != 1.0 : !(h_ssim[n] > 0.0 && h_ssim[n] < 1.0)) {
            fprintf(stderr, "image %d: ssim %.17g out of range\n", n, h_ssim[n]);
            ++bad;
        }
        const double psnr = h_sse[n] ? 10.0 * log10(255.0 * 255.0 * (double)elems / (double)h_sse[n]) : INFINITY;
        printf("image %d  t=%.2f  PSNR %.2f dB  SSIM %.4f\n", n, h_t[n], psnr, h_ssim[n]);
    }
    /* error convention: a bad argument returns a negative code and a message, nothing is launched */
    rc = degrade(d_in, d_out, 0, H, W, NULL, NULL, d_t, d_add, d_on, NULL, NULL, 0, 0, B2R_ORDER_BLUR_FOG_NOISE, 0, (void*)stream);
    if (rc >= 0) {
        fprintf(stderr, "N = 0 was accepted\n");
        ++bad;
    } else {
        printf("N = 0 rejected: %d (%s)\n", rc, last_error());
    }
    printf(bad ? "FAILED (%d)\n" : "ok\n", bad);
    return bad ? 1 : 0;
}
