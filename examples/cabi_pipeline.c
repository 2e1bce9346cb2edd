/* Torch-free END-TO-END use of libb2r.so: degrade -> restore -> classify -> top-1 -> count from plain C.
 *
 *   python examples/export_bundle.py /tmp/bundle.bin            (state_dicts + degradation parameters + images)
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/cabi_pipeline.c -o /tmp/cabi_pipeline \
 *       -L /usr/local/cuda/lib64 -lcudart -ldl
 *   /tmp/cabi_pipeline <libb2r.so> /tmp/bundle.bin
 *
 * This is what 16_gen_compound_data.py -> 17_run_unified_inference.py -> 18_test_unified_benchmark.py do through three
 * processes and two PNG trees, as five library calls on one CUDA stream:
 *   b2r_net_create x2   model = ResUNet(); model.load_state_dict(torch.load(...)); vgg16 + classifier[6] (17:59-64, 18:58-61)
 *   b2r_degrade         apply_compound_distortion per image (16:14-37), Philox noise keyed by the image index
 *   b2r_resunet_forward model(input) + clamp(0, 1) + (x * 255).astype(uint8) (17:85-92), ToTensor fused (17:66)
 *   b2r_vgg16_forward   Resize/ToTensor/Normalize + model(inputs) (18:28-32, 46)
 *   b2r_argmax_count    torch.max(outputs, 1); correct += (predicted == labels).sum() (18:47-49)
 * Prints the counts and FNV-1a digests of the restored bytes and the predictions; tests/test_cabi_example.py compares them
 * with the Python host's results on the same bundle (they must be identical). */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2r.h"

#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) {                                        \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); \
            return 2;                                                   \
        }                                                               \
    } while (0)
#define B2R(call)                                                       \
    do {                                                                \
        int rc_ = (call);                                               \
        if (rc_ != 0) {                                                 \
            fprintf(stderr, "%s: %d %s\n", #call, rc_, last_error());   \
            return 3;                                                   \
        }                                                               \
    } while (0)

typedef const char* (*err_fn)(void);
typedef int (*wbytes_fn)(const b2r_tensor*, int, size_t*);
typedef int (*create_fn)(int, int, const b2r_tensor*, int, void*, size_t, void*, b2r_net**);
typedef void (*destroy_fn)(b2r_net*);
typedef int (*wsbytes_fn)(const b2r_net*, int, int, int, size_t*);
typedef int (*restore_fn)(const b2r_net*, const void*, int, float*, uint8_t*, int, int, int, void*, size_t, void*);
typedef int (*vgg_fn)(const b2r_net*, const void*, int, int, float*, int, int, int, void*, size_t, void*);
typedef int (*degrade_fn)(const uint8_t*, uint8_t*, int, int, int, const float*, const int32_t*, const float*, const float*,
                          const int32_t*, const float*, const double*, uint64_t, uint64_t, int, int, void*);
typedef int (*argmax_fn)(const float*, const int64_t*, int64_t*, float*, int64_t*, int, int, void*);

static uint64_t fnv1a(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

static int read_tensors(FILE* f, int n, b2r_tensor* t) {
    for (int i = 0; i < n; ++i) {
        int32_t len, hdr[2];
        if (fread(&len, 4, 1, f) != 1 || len <= 0 || len > 255) return 1;
        char* name = (char*)calloc((size_t)len + 1, 1);
        if (fread(name, 1, (size_t)len, f) != (size_t)len || fread(hdr, 4, 2, f) != 2 || fread(t[i].shape, 8, 4, f) != 4) return 1;
        t[i].name = name;
        t[i].dtype = hdr[0];
        t[i].ndim = hdr[1];
        size_t numel = 1;
        for (int d = 0; d < hdr[1]; ++d) numel *= (size_t)t[i].shape[d];
        const size_t bytes = numel * (hdr[0] == B2R_DT_F32 ? 4 : 8);
        void* data = malloc(bytes ? bytes : 1);
        if (fread(data, 1, bytes, f) != bytes) return 1;
        t[i].data = data;
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s libb2r.so bundle.bin\n", argv[0]);
        return 2;
    }
    void* lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    err_fn last_error = (err_fn)dlsym(lib, "b2r_last_error");
    wbytes_fn net_weight_bytes = (wbytes_fn)dlsym(lib, "b2r_net_weight_bytes");
    create_fn net_create = (create_fn)dlsym(lib, "b2r_net_create");
    destroy_fn net_destroy = (destroy_fn)dlsym(lib, "b2r_net_destroy");
    wsbytes_fn net_workspace_bytes = (wsbytes_fn)dlsym(lib, "b2r_net_workspace_bytes");
    restore_fn resunet_forward = (restore_fn)dlsym(lib, "b2r_resunet_forward");
    restore_fn unet_forward = (restore_fn)dlsym(lib, "b2r_unet_forward");
    vgg_fn vgg16_forward = (vgg_fn)dlsym(lib, "b2r_vgg16_forward");
    degrade_fn degrade = (degrade_fn)dlsym(lib, "b2r_degrade");
    argmax_fn argmax_count = (argmax_fn)dlsym(lib, "b2r_argmax_count");
    if (!last_error || !net_weight_bytes || !net_create || !net_destroy || !net_workspace_bytes || !resunet_forward || !unet_forward ||
        !vgg16_forward || !degrade || !argmax_count) {
        fprintf(stderr, "missing symbol\n");
        return 2;
    }

    FILE* f = fopen(argv[2], "rb");
    char magic[8];
    int32_t hd[6];
    if (!f || fread(magic, 1, 8, f) != 8 || memcmp(magic, "B2RBNDL1", 8) != 0 || fread(hd, 4, 6, f) != 6) {
        fprintf(stderr, "bad bundle\n");
        return 2;
    }
    const int arch = hd[0] ? B2R_NET_RESUNET : B2R_NET_SIMPLE_UNET, nr = hd[1], nj = hd[2], N = hd[3], H = hd[4], W = hd[5];
    b2r_tensor* tr = (b2r_tensor*)calloc((size_t)nr, sizeof(b2r_tensor));
    b2r_tensor* tj = (b2r_tensor*)calloc((size_t)nj, sizeof(b2r_tensor));
    if (read_tensors(f, nr, tr) || read_tensors(f, nj, tj)) {
        fprintf(stderr, "truncated bundle (tensors)\n");
        return 2;
    }
    const size_t px = (size_t)N * H * W * 3;
    int32_t* ksize = (int32_t*)malloc(4 * (size_t)N);
    float* taps = (float*)malloc(4 * 225 * (size_t)N);
    int32_t* fog_on = (int32_t*)malloc(4 * (size_t)N);
    float* fog_t = (float*)malloc(4 * (size_t)N);
    float* fog_add = (float*)malloc(4 * (size_t)N);
    float* sigma = (float*)malloc(4 * (size_t)N);
    uint8_t* imgs = (uint8_t*)malloc(px);
    int64_t* labels = (int64_t*)malloc(8 * (size_t)N);
    if (fread(ksize, 4, N, f) != (size_t)N || fread(taps, 4, 225 * (size_t)N, f) != 225 * (size_t)N || fread(fog_on, 4, N, f) != (size_t)N ||
        fread(fog_t, 4, N, f) != (size_t)N || fread(fog_add, 4, N, f) != (size_t)N || fread(sigma, 4, N, f) != (size_t)N ||
        fread(imgs, 1, px, f) != px || fread(labels, 8, N, f) != (size_t)N) {
        fprintf(stderr, "truncated bundle (images)\n");
        return 2;
    }
    fclose(f);

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    /* ---- load_state_dict: pack both checkpoints into caller-owned device memory */
    size_t wb_r = 0, wb_j = 0;
    B2R(net_weight_bytes(tr, nr, &wb_r));
    B2R(net_weight_bytes(tj, nj, &wb_j));
    void *dw_r, *dw_j;
    CK(cudaMalloc(&dw_r, wb_r));
    CK(cudaMalloc(&dw_j, wb_j));
    b2r_net *restorer = NULL, *judge = NULL;
    B2R(net_create(arch, 0, tr, nr, dw_r, wb_r, st, &restorer));
    B2R(net_create(B2R_NET_VGG16, 43, tj, nj, dw_j, wb_j, st, &judge));
    {   /* strictness: a missing key is an error that names the key, like load_state_dict (17:63) */
        b2r_net* bad = NULL;
        int rc = net_create(arch, 0, tr + 1, nr - 1, dw_r, wb_r, st, &bad);
        printf("missing key rejected: %d (%s)\n", rc, last_error());
        if (rc == 0 || bad != NULL) return 4;
        net_destroy(restorer);                                          /* dw_r was partly overwritten: pack again */
        B2R(net_create(arch, 0, tr, nr, dw_r, wb_r, st, &restorer));
    }
    size_t ws_r = 0, ws_j = 0;
    B2R(net_workspace_bytes(restorer, N, H, W, &ws_r));
    B2R(net_workspace_bytes(judge, N, H, W, &ws_j));
    const size_t ws_bytes = ws_r > ws_j ? ws_r : ws_j;   /* the two networks run one after the other: share the workspace */
    void* ws;
    CK(cudaMalloc(&ws, ws_bytes));

    /* ---- device buffers */
    uint8_t *d_in, *d_deg, *d_rest;
    int32_t *d_ksize, *d_fog_on;
    float *d_taps, *d_fog_t, *d_fog_add, *d_sigma, *d_logits;
    int64_t *d_labels, *d_pred, *d_counts;
    CK(cudaMalloc((void**)&d_in, px));
    CK(cudaMalloc((void**)&d_deg, px));
    CK(cudaMalloc((void**)&d_rest, px));
    CK(cudaMalloc((void**)&d_ksize, 4 * (size_t)N));
    CK(cudaMalloc((void**)&d_fog_on, 4 * (size_t)N));
    CK(cudaMalloc((void**)&d_taps, 4 * 225 * (size_t)N));
    CK(cudaMalloc((void**)&d_fog_t, 4 * (size_t)N));
    CK(cudaMalloc((void**)&d_fog_add, 4 * (size_t)N));
    CK(cudaMalloc((void**)&d_sigma, 4 * (size_t)N));
    CK(cudaMalloc((void**)&d_logits, 4 * 43 * (size_t)N));
    CK(cudaMalloc((void**)&d_labels, 8 * (size_t)N));
    CK(cudaMalloc((void**)&d_pred, 8 * (size_t)N));
    CK(cudaMalloc((void**)&d_counts, 16));
    CK(cudaMemcpyAsync(d_in, imgs, px, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_ksize, ksize, 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_fog_on, fog_on, 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_taps, taps, 4 * 225 * (size_t)N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_fog_t, fog_t, 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_fog_add, fog_add, 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_sigma, sigma, 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_labels, labels, 8 * (size_t)N, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_counts, 0, 16, st));

    /* ---- the hot path: five calls on one stream, nothing leaves the device in between */
    B2R(degrade(d_in, d_deg, N, H, W, d_taps, d_ksize, d_fog_t, d_fog_add, d_fog_on, d_sigma, NULL, /*seed*/ 2, /*image_index0*/ 0,
                B2R_ORDER_BLUR_FOG_NOISE, 0, st));
    B2R((arch == B2R_NET_RESUNET ? resunet_forward : unet_forward)(restorer, d_deg, B2R_IN_U8_NHWC, NULL, d_rest, N, H, W, ws, ws_bytes, st));
    B2R(vgg16_forward(judge, d_rest, B2R_IN_U8_NHWC, 1, d_logits, N, H, W, ws, ws_bytes, st));
    B2R(argmax_count(d_logits, d_labels, d_pred, NULL, d_counts, N, 43, st));

    uint8_t* rest = (uint8_t*)malloc(px);
    int64_t* pred = (int64_t*)malloc(8 * (size_t)N);
    int64_t counts[2];
    CK(cudaMemcpyAsync(rest, d_rest, px, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pred, d_pred, 8 * (size_t)N, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(counts, d_counts, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    printf("restored fnv1a %016llx\n", (unsigned long long)fnv1a(rest, px));
    printf("pred fnv1a %016llx\n", (unsigned long long)fnv1a(pred, 8 * (size_t)N));
    printf("counts %lld %lld\n", (long long)counts[0], (long long)counts[1]);
    net_destroy(restorer);
    net_destroy(judge);
    printf("ok\n");
    return counts[1] == N ? 0 : 5;
}
