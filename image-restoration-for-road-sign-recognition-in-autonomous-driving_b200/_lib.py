"""ctypes binding of libb2r.so (the C-ABI declared in include/b2r.h).

There is no CPU fallback: if the shared object is missing the import of any op raises, and every op raises on a
non-zero return code with the library's own message.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

EXPECTED_ABI = 200   # B2R_VERSION of include/b2r.h this binding matches (bumped on every struct / signature change)
B2R_ACT_NONE, B2R_ACT_RELU, B2R_ACT_PRELU = 0, 1, 2
B2R_ORDER_BLUR_FOG_NOISE, B2R_ORDER_FOG_NOISE_BLUR = 0, 1
B2R_DEG_CLIP_AFTER_NOISE = 1
B2R_IN_F32_NCHW, B2R_IN_U8_NHWC = 0, 1
B2R_OUT_NHWC, B2R_OUT_CONVT2X2 = 0, 1
B2R_MAX_SRC, B2R_MAX_KBLOCKS, B2R_MAX_BLUR = 3, 96, 15
B2R_CONV_GENERIC_ONLY, B2R_CONV_NO_W3, B2R_CONV_NO_HALO, B2R_CONV_NO_PAIR = 1, 2, 4, 8

# every symbol include/b2r.h declares (tests/test_abi.py checks the list against the header and the .so)
SYMBOLS = (
    "b2r_version", "b2r_abi_sizeof", "b2r_last_error", "b2r_last_conv_kernel", "b2r_degrade", "b2r_conv3x3_c3", "b2r_conv_gemm", "b2r_final_conv1x1",
    "b2r_maxpool2x2", "b2r_resize_nearest_bf16", "b2r_adaptive_avgpool7", "b2r_linear_f32out", "b2r_argmax_count",
    "b2r_net_weight_bytes", "b2r_net_create", "b2r_net_destroy", "b2r_net_workspace_bytes", "b2r_unet_forward", "b2r_resunet_forward",
    "b2r_vgg16_forward",
    "b2r_lut_u8", "b2r_minmax_u8", "b2r_normalize_minmax_u8", "b2r_noise02", "b2r_sse_u8", "b2r_ssim_u8", "b2r_mean_bf16", "b2r_resize_bilinear_u8", "b2r_resize_cv_linear_u8",
)


class ConvGemmDesc(C.Structure):
    """struct b2r_conv_gemm_desc (include/b2r.h)."""
    _fields_ = [
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("num_src", C.c_int32),
        ("src", C.c_void_p * B2R_MAX_SRC),
        ("src_C", C.c_int32 * B2R_MAX_SRC),
        ("weights", C.c_void_p),
        ("weights_w3", C.c_void_p),
        ("bias", C.c_void_p),
        ("cout_total", C.c_int32),
        ("num_kblocks", C.c_int32),
        ("kblocks_host", C.POINTER(C.c_uint32)),
        ("act", C.c_int32),
        ("slope", C.c_float),
        ("out_mode", C.c_int32),
        ("out", C.c_void_p),
        ("out_pool", C.c_void_p),
        ("out_C", C.c_int32),
        ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("tile_n", C.c_int32),
        ("block_n", C.c_int32),
        ("max_ctas", C.c_int32),
        ("flags", C.c_int32),
        ("head_w", C.c_void_p),
        ("head_b", C.c_void_p),
        ("head_out_f32", C.c_void_p),
        ("head_out_u8", C.c_void_p),
        ("debug_timeline", C.c_void_p),
    ]


class Tensor(C.Structure):
    """struct b2r_tensor (include/b2r.h): one state_dict entry handed to b2r_net_create."""
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


B2R_NET_SIMPLE_UNET, B2R_NET_RESUNET, B2R_NET_VGG16 = 0, 1, 2
B2R_DT_F32, B2R_DT_I64 = 0, 1


class B2RError(RuntimeError):
    pass


_lib = None


def lib_path() -> Path:
    """In-tree libb2r.so; $B2R_LIB selects another build of the same ABI (A/B kernel experiments only)."""
    import os
    alt = os.environ.get("B2R_LIB")
    return Path(alt) if alt else _build.LIB_PATH


def load() -> C.CDLL:
    """dlopen libb2r.so (building it first when nvcc is available and the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = lib_path()
    in_tree = "B2R_LIB" not in os.environ
    if not path.exists() or (in_tree and _build.needs_build() and _build.have_nvcc()):
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box
            if not path.exists():
                raise B2RError(f"libb2r.so is missing at {path} and could not be built: {e}") from e
    lib = C.CDLL(str(path))
    vp, i32, f32, u64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
    lib.b2r_version.restype = C.c_int
    lib.b2r_version.argtypes = []
    # a stale or foreign build would read a mis-laid descriptor: refuse it instead of corrupting memory
    if lib.b2r_version() != EXPECTED_ABI:
        raise B2RError(f"{path} reports ABI {lib.b2r_version()}, this binding was written for {EXPECTED_ABI}; rebuild "
                       "(python -m b200restore.build --force)")
    lib.b2r_abi_sizeof.restype = C.c_int
    lib.b2r_abi_sizeof.argtypes = [C.c_int]
    if lib.b2r_abi_sizeof(0) != C.sizeof(ConvGemmDesc):
        raise B2RError(f"struct b2r_conv_gemm_desc is {lib.b2r_abi_sizeof(0)} bytes in {path} but {C.sizeof(ConvGemmDesc)} in "
                       "the ctypes binding")
    lib.b2r_last_error.restype = C.c_char_p
    lib.b2r_last_error.argtypes = []
    lib.b2r_last_conv_kernel.restype = C.c_char_p
    lib.b2r_last_conv_kernel.argtypes = []
    if hasattr(lib, "b2r_debug_timeline"):      # debug builds only (csrc/b2r_debug.h), not part of the product ABI
        lib.b2r_debug_timeline.restype = None
        lib.b2r_debug_timeline.argtypes = [vp]
    lib.b2r_degrade.restype = C.c_int
    lib.b2r_degrade.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, u64, u64, i32, i32, vp]
    lib.b2r_conv3x3_c3.restype = C.c_int
    lib.b2r_conv3x3_c3.argtypes = [vp, i32, C.POINTER(f32), C.POINTER(f32), vp, vp, i32, f32, vp, i32, i32, i32, vp]
    lib.b2r_conv_gemm.restype = C.c_int
    lib.b2r_conv_gemm.argtypes = [C.POINTER(ConvGemmDesc), vp]
    lib.b2r_final_conv1x1.restype = C.c_int
    lib.b2r_final_conv1x1.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.b2r_maxpool2x2.restype = C.c_int
    lib.b2r_maxpool2x2.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.b2r_resize_nearest_bf16.restype = C.c_int
    lib.b2r_resize_nearest_bf16.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.b2r_adaptive_avgpool7.restype = C.c_int
    lib.b2r_adaptive_avgpool7.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.b2r_linear_f32out.restype = C.c_int
    lib.b2r_linear_f32out.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.b2r_argmax_count.restype = C.c_int
    lib.b2r_argmax_count.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    i64 = C.c_int64
    lib.b2r_lut_u8.restype = C.c_int
    lib.b2r_lut_u8.argtypes = [vp, vp, vp, i32, i64, vp]
    lib.b2r_minmax_u8.restype = C.c_int
    lib.b2r_minmax_u8.argtypes = [vp, vp, i32, i64, vp]
    lib.b2r_normalize_minmax_u8.restype = C.c_int
    lib.b2r_normalize_minmax_u8.argtypes = [vp, vp, vp, i32, i64, vp]
    lib.b2r_noise02.restype = C.c_int
    lib.b2r_noise02.argtypes = [vp, vp, i32, i64, vp, vp, u64, u64, vp, i32, vp]
    lib.b2r_resize_bilinear_u8.restype = C.c_int
    lib.b2r_resize_bilinear_u8.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, vp]
    lib.b2r_resize_cv_linear_u8.restype = C.c_int
    lib.b2r_resize_cv_linear_u8.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp]
    lib.b2r_mean_bf16.restype = C.c_int
    lib.b2r_mean_bf16.argtypes = [vp, vp, i64, i32, i32, vp]
    lib.b2r_sse_u8.restype = C.c_int
    lib.b2r_sse_u8.argtypes = [vp, vp, vp, i32, i64, vp]
    lib.b2r_ssim_u8.restype = C.c_int
    lib.b2r_ssim_u8.argtypes = [vp, vp, vp, i32, i32, i32, i32, C.c_double, vp]
    if lib.b2r_abi_sizeof(1) != C.sizeof(Tensor):
        raise B2RError("struct b2r_tensor layout differs between the library and the ctypes binding")
    sz = C.c_size_t
    lib.b2r_net_weight_bytes.restype = C.c_int
    lib.b2r_net_weight_bytes.argtypes = [C.POINTER(Tensor), i32, C.POINTER(sz)]
    lib.b2r_net_create.restype = C.c_int
    lib.b2r_net_create.argtypes = [i32, i32, C.POINTER(Tensor), i32, vp, sz, vp, C.POINTER(vp)]
    lib.b2r_net_destroy.restype = None
    lib.b2r_net_destroy.argtypes = [vp]
    lib.b2r_net_workspace_bytes.restype = C.c_int
    lib.b2r_net_workspace_bytes.argtypes = [vp, i32, i32, i32, C.POINTER(sz)]
    for fn in (lib.b2r_unet_forward, lib.b2r_resunet_forward):
        fn.restype = C.c_int
        fn.argtypes = [vp, vp, i32, vp, vp, i32, i32, i32, vp, sz, vp]
    lib.b2r_vgg16_forward.restype = C.c_int
    lib.b2r_vgg16_forward.argtypes = [vp, vp, i32, i32, vp, i32, i32, i32, vp, sz, vp]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b2r_last_error()
        raise B2RError(f"libb2r error {rc}: {msg.decode() if msg else '?'}")
