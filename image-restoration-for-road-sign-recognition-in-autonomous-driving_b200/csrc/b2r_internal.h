// Host-side internals shared by the .cu files of libb2r.so: error reporting, launch checks, TMA map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/b2r.h"

namespace b2r {

// thread-local message returned by b2r_last_error()
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define B2R_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) return b2r::set_error(B2R_EINVAL, __VA_ARGS__); \
    } while (0)

#define B2R_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return b2r::set_error(B2R_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                                  __LINE__);                                                                 \
    } while (0)

// Kernel launches are asynchronous; this only catches launch-configuration errors.
#define B2R_CHECK_LAUNCH()                                                                                    \
    do {                                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                                 \
        if (e__ != cudaSuccess)                                                                               \
            return b2r::set_error(B2R_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),      \
                                  __FILE__, __LINE__);                                                        \
    } while (0)

// Encode a bf16 tiled tensor map (rank <= 5) with SWIZZLE_128B, zero OOB fill.  dims/box innermost first;
// strides_bytes has rank-1 entries (stride of dim 1..rank-1).  Returns 0 or a negative B2R code.
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);

int device_sm_count(int* sms);

// name of the kernel the last b2r_conv_gemm call of this thread launched (diagnostic: b2r_last_conv_kernel())
void note_conv_kernel(const char* name);
const char* last_conv_kernel();

}  // namespace b2r
