// Debug-only entry points of libb2r.so (tools/c3_timeline.py, tools/role_timeline.py).  Not part of the product ABI:
// include/b2r.h does not declare them and no product code path calls them.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* device buffer int64[B2R_DBG_TILES][8] in which CTA 0 of the NEXT b2r_conv3x3_c3 launches records clock64() stamps per
 * warp role; NULL switches it off (the default). */
void b2r_debug_timeline(int64_t* device_buf);
#ifdef __cplusplus
}
#endif
