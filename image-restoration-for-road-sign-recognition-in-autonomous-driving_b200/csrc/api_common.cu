// libb2r.so common host code: version, thread-local error string, TMA tensor-map encoding through the driver
// entry point (no link-time dependency on libcuda, so the library loads on a machine without a GPU).
#include <cstring>
#include <mutex>

#include "b2r_internal.h"

namespace b2r {

static thread_local char g_err[512] = "";
static thread_local const char* g_conv_kernel = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(B2R_ECUDA, "cuTensorMapEncodeTiled driver entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
        return set_error(B2R_EINVAL, "tensor base %p is not 16-byte aligned", base);
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (box[i] == 0 || box[i] > 256) return set_error(B2R_EINVAL, "TMA box dim %d = %u out of range", i, box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        if (strides_bytes[i] % 16 != 0)
            return set_error(B2R_EINVAL, "TMA stride %d = %llu not a multiple of 16 bytes", i,
                             (unsigned long long)strides_bytes[i]);
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B2R_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return B2R_OK;
}

void note_conv_kernel(const char* name) { g_conv_kernel = name; }
const char* last_conv_kernel() { return g_conv_kernel; }

int device_sm_count(int* sms) {
    static int cached[64] = {0};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && cached[dev] > 0) {
        *sms = cached[dev];
        return B2R_OK;
    }
    int major = 0, n = 0;
    B2R_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return set_error(B2R_ENODEV, "device %d is compute capability %d.x; libb2r needs sm_100", dev, major);
    B2R_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 64) cached[dev] = n;
    *sms = n;
    return B2R_OK;
}

}  // namespace b2r

extern "C" {

int b2r_version(void) { return B2R_VERSION; }

int b2r_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(b2r_conv_gemm_desc);
        case 1: return (int)sizeof(b2r_tensor);
        default: return -1;
    }
}

const char* b2r_last_error(void) { return b2r::last_error_buf(); }
const char* b2r_last_conv_kernel(void) { return b2r::last_conv_kernel(); }

}  // extern "C"
