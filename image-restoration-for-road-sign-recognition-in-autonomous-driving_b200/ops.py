"""Tensor-level wrappers over the C-ABI: argument checking, raw pointers, current CUDA stream.

PyTorch is used here for device memory and streams only; every computation below is one libb2r.so kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # 18_test_unified_benchmark.py:31
IMAGENET_STD = (0.229, 0.224, 0.225)

# launch accounting: every C-ABI call below is exactly one kernel launch of libb2r.so
STATS = {"launches": 0}


class KernelTimer:
    """Optional per-launch CUDA-event timing on the launching stream (used by bench.py for the roofline numbers).

    `with ops.timing(timer):` brackets every launch of the kinds in `kinds` with an event pair and records the
    algorithmic work (FLOPs or bytes) the caller states for it; `summary()` synchronises and sums."""

    def __init__(self, kinds=("conv_gemm",)):
        self.kinds = set(kinds)
        self.records = []   # (kind, work, start_event, end_event, kernel name | None)

    def start(self, kind):
        if kind not in self.kinds:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        return e

    def stop(self, kind, work, e0, sub=None):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(torch.cuda.current_stream())
        self.records.append((kind, float(work), e0, e1, sub))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, work, e0, e1, sub in self.records:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(kind, {"launches": 0, "work": 0.0, "ms": 0.0, "by_kernel": {}})
            for t in (d,) + ((d["by_kernel"].setdefault(sub, {"launches": 0, "work": 0.0, "ms": 0.0}),) if sub else ()):
                t["launches"] += 1
                t["work"] += work
                t["ms"] += ms
        return out


_TIMER: Optional[KernelTimer] = None


class timing:
    def __init__(self, timer: Optional[KernelTimer]):
        self.timer = timer

    def __enter__(self):
        global _TIMER
        self._prev, _TIMER = _TIMER, self.timer
        return self.timer

    def __exit__(self, *a):
        global _TIMER
        _TIMER = self._prev
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str, ndim: Optional[int] = None) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise L.B2RError(f"{name}: tensor must live on a CUDA device (there is no CPU fallback)")
    # every launch goes to the CURRENT device's current stream, and the C side sizes its grids / sets its function
    # attributes for the current device: a tensor on another GPU would be reached over P2P or fault.  The modules and the
    # pipeline enter `torch.cuda.device(x.device)` themselves; a direct op call must be made with the right device current.
    if t.device.index != torch.cuda.current_device():
        raise L.B2RError(f"{name}: tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         "wrap the call in `with torch.cuda.device(tensor.device):`")
    if t.dtype != dtype:
        raise L.B2RError(f"{name}: dtype {t.dtype}, expected {dtype}")
    if not t.is_contiguous():
        raise L.B2RError(f"{name}: tensor must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise L.B2RError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def kblock(src: int, dh: int, dw: int, c64: int) -> int:
    """B2R_KBLOCK from include/b2r.h."""
    return src | ((dh + 1) << 2) | ((dw + 1) << 4) | (c64 << 8)


def conv_gemm(srcs: Sequence[torch.Tensor], weights: torch.Tensor, bias: torch.Tensor,
              kblocks: Optional[Sequence[int]], act: int = L.B2R_ACT_NONE, slope: float = 0.0,
              out: Optional[torch.Tensor] = None, out_pool: Optional[torch.Tensor] = None,
              out_mode: int = L.B2R_OUT_NHWC, tile=(0, 0, 0), block_n: int = 0, max_ctas: int = 0,
              alg_k: Optional[int] = None, flags: int = 0, weights_w3: Optional[torch.Tensor] = None,
              debug_timeline: Optional[torch.Tensor] = None, head_w: Optional[torch.Tensor] = None,
              head_b: Optional[torch.Tensor] = None, head_out_f32: Optional[torch.Tensor] = None,
              head_out_u8: Optional[torch.Tensor] = None) -> None:
    """One fused tensor-core layer (b2r_conv_gemm).  srcs: NHWC bf16 [N,H,W,C_i]; weights bf16 [cout_total, K].
    `alg_k`: K elements that are algorithmic work (excludes e.g. an identity-shortcut block); accounting only."""
    n, h, w = srcs[0].shape[:3]
    d = L.ConvGemmDesc()
    d.N, d.H, d.W = int(n), int(h), int(w)
    d.num_src = len(srcs)
    for i, s in enumerate(srcs):
        _chk(s, torch.bfloat16, f"src[{i}]", 4)
        if tuple(s.shape[:3]) != (n, h, w):
            raise L.B2RError(f"src[{i}] shape {tuple(s.shape)} does not match the pixel grid {(n, h, w)}")
        d.src[i] = s.data_ptr()
        d.src_C[i] = int(s.shape[3])
    _chk(weights, torch.bfloat16, "weights", 2)
    _chk(bias, torch.float32, "bias", 1)
    d.weights = weights.data_ptr()
    if weights_w3 is not None:
        _chk(weights_w3, torch.bfloat16, "weights_w3", 2)
        if weights_w3.shape[0] != 192 or weights_w3.shape[1] % 64 != 0:
            raise L.B2RError(f"weights_w3 must be [192, 64*k-steps], got {tuple(weights_w3.shape)}")
        d.weights_w3 = weights_w3.data_ptr()
    d.bias = bias.data_ptr()
    d.cout_total = int(weights.shape[0])
    if bias.numel() != d.cout_total:
        raise L.B2RError(f"bias has {bias.numel()} entries, weights have {d.cout_total} rows")
    if weights.shape[1] % 64 != 0:
        raise L.B2RError(f"weights K={weights.shape[1]} is not a multiple of 64")
    d.num_kblocks = int(weights.shape[1]) // 64
    arr = None
    if kblocks is not None:
        if len(kblocks) != d.num_kblocks:
            raise L.B2RError(f"{len(kblocks)} k-blocks for a weight matrix with K={weights.shape[1]}")
        arr = (C.c_uint32 * len(kblocks))(*kblocks)
        d.kblocks_host = C.cast(arr, C.POINTER(C.c_uint32))
    d.act, d.slope, d.out_mode = int(act), float(slope), int(out_mode)
    oc = 0
    if out is not None:
        _chk(out, torch.bfloat16, "out", 4)
        exp = (n, 2 * h, 2 * w) if out_mode == L.B2R_OUT_CONVT2X2 else (n, h, w)
        if tuple(out.shape[:3]) != exp:
            raise L.B2RError(f"out shape {tuple(out.shape)} != expected {exp} + (C,)")
        d.out = out.data_ptr()
        oc = int(out.shape[3])
    if out_pool is not None:
        _chk(out_pool, torch.bfloat16, "out_pool", 4)
        if tuple(out_pool.shape[:3]) != (n, h // 2, w // 2):
            raise L.B2RError(f"out_pool shape {tuple(out_pool.shape)} != {(n, h // 2, w // 2)} + (C,)")
        if oc and int(out_pool.shape[3]) != oc:
            raise L.B2RError("out and out_pool must have the same channel count")
        d.out_pool = out_pool.data_ptr()
        oc = int(out_pool.shape[3])
    d.out_C = oc
    if head_w is not None:
        _chk(head_w, torch.float32, "head_w")
        _chk(head_b, torch.float32, "head_b")
        if head_w.numel() != 192 or head_b.numel() != 3:
            raise L.B2RError("head_w / head_b must hold 3x64 and 3 values")
        d.head_w, d.head_b = head_w.data_ptr(), head_b.data_ptr()
        if head_out_f32 is not None:
            _chk(head_out_f32, torch.float32, "head_out_f32", 4)
            if tuple(head_out_f32.shape) != (n, 3, h, w):
                raise L.B2RError(f"head_out_f32 shape {tuple(head_out_f32.shape)} != {(n, 3, h, w)}")
            d.head_out_f32 = head_out_f32.data_ptr()
        if head_out_u8 is not None:
            _chk(head_out_u8, torch.uint8, "head_out_u8", 4)
            if tuple(head_out_u8.shape) != (n, h, w, 3):
                raise L.B2RError(f"head_out_u8 shape {tuple(head_out_u8.shape)} != {(n, h, w, 3)}")
            d.head_out_u8 = head_out_u8.data_ptr()
    d.tile_w, d.tile_h, d.tile_n = (int(x) for x in tile)
    d.block_n, d.max_ctas, d.flags = int(block_n), int(max_ctas), int(flags)
    if debug_timeline is not None:
        _chk(debug_timeline, torch.int64, "debug_timeline")
        if debug_timeline.numel() < 64 * 8:
            raise L.B2RError("debug_timeline must hold 64 x 8 int64")
        d.debug_timeline = debug_timeline.data_ptr()
    e0 = _TIMER.start("conv_gemm") if _TIMER is not None else None
    L.check(L.load().b2r_conv_gemm(C.byref(d), _stream()))
    STATS["launches"] += 1
    if e0 is not None:
        k_alg = int(weights.shape[1]) if alg_k is None else int(alg_k)
        _TIMER.stop("conv_gemm", 2.0 * n * h * w * d.cout_total * k_alg, e0,
                    sub=(L.load().b2r_last_conv_kernel() or b"?").decode())
    del arr


def conv3x3_c3(x: torch.Tensor, weights: torch.Tensor, bias: torch.Tensor, act: int, slope: float = 0.0,
               normalize: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """First layer.  x: f32 [N,3,H,W] or u8 [N,H,W,3]; weights = packing.pack_conv_c3(...) bf16 [64,64]; returns bf16
    NHWC [N,H,W,64]."""
    if x.dtype == torch.uint8:
        _chk(x, torch.uint8, "x", 4)
        n, h, w, c = x.shape
        fmt = L.B2R_IN_U8_NHWC
    else:
        _chk(x, torch.float32, "x", 4)
        n, c, h, w = x.shape
        fmt = L.B2R_IN_F32_NCHW
    if c != 3:
        raise L.B2RError(f"conv3x3_c3 needs 3 input channels, got {c}")
    _chk(weights, torch.bfloat16, "weights", 2)
    _chk(bias, torch.float32, "bias", 1)
    if tuple(weights.shape) != (64, 64) or bias.numel() != 64:
        raise L.B2RError(f"conv3x3_c3 weights must be the packed bf16 [64,64] matrix, got {tuple(weights.shape)}")
    if out is None:
        out = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=x.device)
    _chk(out, torch.bfloat16, "out", 4)
    mean = std = None
    if normalize:
        if fmt != L.B2R_IN_U8_NHWC:
            raise L.B2RError("normalize=True is the u8 judge hand-off; f32 input is taken as is")
        mean = (C.c_float * 3)(*IMAGENET_MEAN)
        std = (C.c_float * 3)(*IMAGENET_STD)
    L.check(L.load().b2r_conv3x3_c3(x.data_ptr(), fmt, mean, std, weights.data_ptr(), bias.data_ptr(), int(act),
                                    float(slope), out.data_ptr(), int(n), int(h), int(w), _stream()))
    STATS["launches"] += 1
    return out


def final_conv1x1(x: torch.Tensor, weights: torch.Tensor, bias: torch.Tensor, want_f32: bool = True,
                  want_u8: bool = False):
    """64 -> 3 head.  x: bf16 NHWC [N,H,W,64]; returns (f32 [N,3,H,W] | None, u8 [N,H,W,3] | None)."""
    _chk(x, torch.bfloat16, "x", 4)
    n, h, w, c = x.shape
    if c != 64:
        raise L.B2RError(f"final_conv1x1 needs 64 channels, got {c}")
    _chk(weights, torch.float32, "weights")
    _chk(bias, torch.float32, "bias", 1)
    if weights.numel() != 192 or bias.numel() != 3:
        raise L.B2RError("final_conv1x1 weights must hold 3x64 values")
    o32 = torch.empty((n, 3, h, w), dtype=torch.float32, device=x.device) if want_f32 else None
    o8 = torch.empty((n, h, w, 3), dtype=torch.uint8, device=x.device) if want_u8 else None
    L.check(L.load().b2r_final_conv1x1(x.data_ptr(), weights.data_ptr(), bias.data_ptr(), _ptr(o32), _ptr(o8), int(n),
                                       int(h), int(w), _stream()))
    STATS["launches"] += 1
    return o32, o8


def maxpool2x2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(x, torch.bfloat16, "x", 4)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=x.device)
    _chk(out, torch.bfloat16, "out", 4)
    L.check(L.load().b2r_maxpool2x2(x.data_ptr(), out.data_ptr(), int(n), int(h), int(w), int(c), _stream()))
    STATS["launches"] += 1
    return out


def resize_nearest(x: torch.Tensor, out_h: int, out_w: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.interpolate(x, size=(out_h, out_w)) (mode 'nearest') on NHWC bf16 (14_train_unified_advanced.py:169-170)."""
    _chk(x, torch.bfloat16, "x", 4)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, out_h, out_w, c), dtype=torch.bfloat16, device=x.device)
    _chk(out, torch.bfloat16, "out", 4)
    if tuple(out.shape) != (n, out_h, out_w, c):
        raise L.B2RError(f"out shape {tuple(out.shape)} != {(n, out_h, out_w, c)}")
    L.check(L.load().b2r_resize_nearest_bf16(x.data_ptr(), out.data_ptr(), int(n), int(h), int(w), int(out_h), int(out_w), int(c),
                                             _stream()))
    STATS["launches"] += 1
    return out


def adaptive_avgpool7(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.bfloat16, "x", 4)
    n, h, w, c = x.shape
    out = torch.empty((n, 7, 7, c), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().b2r_adaptive_avgpool7(x.data_ptr(), out.data_ptr(), int(n), int(h), int(w), int(c), _stream()))
    STATS["launches"] += 1
    return out


def linear_f32out(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.bfloat16, "x", 2)
    _chk(w, torch.bfloat16, "w", 2)
    _chk(b, torch.float32, "b", 1)
    bsz, k = x.shape
    o = w.shape[0]
    if w.shape[1] != k or b.numel() != o:
        raise L.B2RError(f"linear shapes: x {tuple(x.shape)}, w {tuple(w.shape)}, b {tuple(b.shape)}")
    out = torch.empty((bsz, o), dtype=torch.float32, device=x.device)
    L.check(L.load().b2r_linear_f32out(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), int(bsz), int(k),
                                       int(o), _stream()))
    STATS["launches"] += 1
    return out


def argmax_count(logits: torch.Tensor, labels: Optional[torch.Tensor] = None,
                 counts: Optional[torch.Tensor] = None, want_conf: bool = False):
    """Top-1 (+ softmax confidence) and the running (correct, total) counters.  Returns (pred int64 [N], conf|None)."""
    _chk(logits, torch.float32, "logits", 2)
    n, c = logits.shape
    pred = torch.empty((n,), dtype=torch.int64, device=logits.device)
    conf = torch.empty((n,), dtype=torch.float32, device=logits.device) if want_conf else None
    if labels is not None:
        _chk(labels, torch.int64, "labels", 1)
        if labels.numel() != n:
            raise L.B2RError("labels length does not match logits rows")
    if counts is not None:
        _chk(counts, torch.int64, "counts", 1)
        if counts.numel() != 2:
            raise L.B2RError("counts must be int64[2] = (correct, total)")
    L.check(L.load().b2r_argmax_count(logits.data_ptr(), _ptr(labels), pred.data_ptr(), _ptr(conf), _ptr(counts),
                                      int(n), int(c), _stream()))
    STATS["launches"] += 1
    return pred, conf


def degrade(images: torch.Tensor, ksize: Optional[torch.Tensor], taps: Optional[torch.Tensor],
            fog_on: Optional[torch.Tensor], fog_t: Optional[torch.Tensor], fog_add: Optional[torch.Tensor],
            sigma: Optional[torch.Tensor], noise: Optional[torch.Tensor] = None, seed: int = 0,
            image_index0: int = 0, order: int = L.B2R_ORDER_BLUR_FOG_NOISE, flags: int = 0,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """b2r_degrade on u8 NHWC [N,H,W,3] with per-image parameter tensors on the same device."""
    _chk(images, torch.uint8, "images", 4)
    n, h, w, c = images.shape
    if c != 3:
        raise L.B2RError("images must be NHWC with 3 channels")
    if out is None:
        out = torch.empty_like(images)
    _chk(out, torch.uint8, "out", 4)
    for t, dt, nm in ((ksize, torch.int32, "ksize"), (taps, torch.float32, "taps"), (fog_on, torch.int32, "fog_on"),
                      (fog_t, torch.float32, "fog_t"), (fog_add, torch.float32, "fog_add"),
                      (sigma, torch.float32, "sigma"), (noise, torch.float64, "noise")):
        if t is not None:
            _chk(t, dt, nm)
    if taps is not None and taps.numel() != n * 225:
        raise L.B2RError(f"taps must hold N*225 floats, got {taps.numel()}")
    for t, nm in ((ksize, "ksize"), (fog_on, "fog_on"), (fog_t, "fog_t"), (fog_add, "fog_add"), (sigma, "sigma")):
        if t is not None and t.numel() != n:
            raise L.B2RError(f"{nm} must have one entry per image")
    if noise is not None and tuple(noise.shape) != (n, h, w, 3):
        raise L.B2RError(f"noise must be f64 {(n, h, w, 3)}, got {tuple(noise.shape)}")
    L.check(L.load().b2r_degrade(images.data_ptr(), out.data_ptr(), int(n), int(h), int(w), _ptr(taps), _ptr(ksize),
                                 _ptr(fog_t), _ptr(fog_add), _ptr(fog_on), _ptr(sigma), _ptr(noise),
                                 int(seed) & (2 ** 64 - 1), int(image_index0), int(order), int(flags), _stream()))
    STATS["launches"] += 1
    return out


# ---------------------------------------------------------------------------------------------------------------------
# callers either side of the hot path: exact single-degradation generators and the PSNR reduction (csrc/generators.cu)
# ---------------------------------------------------------------------------------------------------------------------
def _images_u8(x: torch.Tensor, name: str):
    _chk(x, torch.uint8, name)
    if x.dim() < 2:
        raise L.B2RError(f"{name} must be [N, ...] uint8")
    return int(x.shape[0]), int(x[0].numel())


def lut_u8(images: torch.Tensor, lut: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[n] = lut[n][images[n]]; lut uint8 [N, 256] on the same device."""
    n, elems = _images_u8(images, "images")
    _chk(lut, torch.uint8, "lut", 2)
    if tuple(lut.shape) != (n, 256):
        raise L.B2RError(f"lut shape {tuple(lut.shape)}; expected ({n}, 256)")
    out = torch.empty_like(images) if out is None else out
    _chk(out, torch.uint8, "out")
    if out.shape != images.shape:
        raise L.B2RError("out shape differs from images")
    L.check(L.load().b2r_lut_u8(images.data_ptr(), lut.data_ptr(), out.data_ptr(), n, elems, _stream()))
    STATS["launches"] += 1
    return out


def minmax_u8(images: torch.Tensor) -> torch.Tensor:
    """Per-image (min, max) over all channels as int32 [N, 2]."""
    n, elems = _images_u8(images, "images")
    mm = torch.empty((n, 2), dtype=torch.int32, device=images.device)
    L.check(L.load().b2r_minmax_u8(images.data_ptr(), mm.data_ptr(), n, elems, _stream()))
    STATS["launches"] += 2
    return mm


def normalize_minmax_u8(images: torch.Tensor, minmax: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cv2.normalize(x, x, 0, 255, cv2.NORM_MINMAX) given the per-image extrema (03_gen_blur.py:29)."""
    n, elems = _images_u8(images, "images")
    _chk(minmax, torch.int32, "minmax", 2)
    if tuple(minmax.shape) != (n, 2):
        raise L.B2RError(f"minmax shape {tuple(minmax.shape)}; expected ({n}, 2)")
    out = torch.empty_like(images) if out is None else out
    _chk(out, torch.uint8, "out")
    L.check(L.load().b2r_normalize_minmax_u8(images.data_ptr(), minmax.data_ptr(), out.data_ptr(), n, elems, _stream()))
    STATS["launches"] += 1
    return out


def noise02(images: torch.Tensor, sigma: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
            image_index0: int = 0, out: Optional[torch.Tensor] = None, clip_rule: int = 0):
    """Float64 noise generators on u8 [N,H,W,3]: clip_rule 0 = 02_gen_noise.py add_gaussian_noise (conditional -1 clip,
    wrap-around), 1 = 13_pipeline_stress_test.py add_noise (clip to [0, 1]).  Returns (out u8, neg_flags int32 [N])."""
    _chk(images, torch.uint8, "images", 4)
    n, elems = _images_u8(images, "images")
    if images.shape[3] != 3:
        raise L.B2RError("images must be NHWC with 3 channels")
    _chk(sigma, torch.float32, "sigma", 1)
    if sigma.numel() != n:
        raise L.B2RError("sigma must have one entry per image")
    if noise is not None:
        _chk(noise, torch.float64, "noise")
        if noise.numel() != images.numel():
            raise L.B2RError("noise must have one float64 per image element")
    out = torch.empty_like(images) if out is None else out
    _chk(out, torch.uint8, "out", 4)
    flags = torch.empty((n,), dtype=torch.int32, device=images.device)
    L.check(L.load().b2r_noise02(images.data_ptr(), out.data_ptr(), n, elems, sigma.data_ptr(), _ptr(noise),
                                 int(seed) & (2 ** 64 - 1), int(image_index0), flags.data_ptr(), int(clip_rule),
                                 _stream()))
    STATS["launches"] += 2 if clip_rule == 0 else 1
    return out, flags


def sse_u8(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Per-image sum of squared differences of two u8 batches, int64 [N] (exact)."""
    n, elems = _images_u8(a, "a")
    _chk(b, torch.uint8, "b")
    if b.shape != a.shape:
        raise L.B2RError("a and b differ in shape")
    sse = torch.empty((n,), dtype=torch.int64, device=a.device)
    L.check(L.load().b2r_sse_u8(a.data_ptr(), b.data_ptr(), sse.data_ptr(), n, elems, _stream()))
    STATS["launches"] += 1
    return sse


def ssim_u8(a: torch.Tensor, b: torch.Tensor, data_range: float = 255.0) -> torch.Tensor:
    """Per-image mean SSIM of two u8 [N, H, W, C] batches (skimage defaults, see b2r_ssim_u8), f64 [N]."""
    _chk(a, torch.uint8, "a", 4)
    _chk(b, torch.uint8, "b", 4)
    if b.shape != a.shape:
        raise L.B2RError("a and b differ in shape")
    n, h, w, c = (int(v) for v in a.shape)
    out = torch.empty((n,), dtype=torch.float64, device=a.device)
    L.check(L.load().b2r_ssim_u8(a.data_ptr(), b.data_ptr(), out.data_ptr(), n, h, w, c, float(data_range), _stream()))
    STATS["launches"] += 1
    return out


def mean_bf16(x: torch.Tensor, outer: int, reduce: int, inner: int) -> torch.Tensor:
    """Mean over the middle axis of a bf16 tensor viewed as [outer][reduce][inner] -> f32 [outer, inner]."""
    _chk(x, torch.bfloat16, "x")
    if x.numel() != outer * reduce * inner:
        raise L.B2RError(f"{x.numel()} elements cannot be viewed as [{outer}][{reduce}][{inner}]")
    out = torch.empty((outer, inner), dtype=torch.float32, device=x.device)
    L.check(L.load().b2r_mean_bf16(x.data_ptr(), out.data_ptr(), int(outer), int(reduce), int(inner), _stream()))
    STATS["launches"] += 1
    return out


def resize_bilinear_u8(src: torch.Tensor, offsets: torch.Tensor, hw: torch.Tensor, xtab_index: torch.Tensor,
                       ytab_index: torch.Tensor, tabs: torch.Tensor, K: int, S: int, out: torch.Tensor, tile_rows: int,
                       max_rows: int) -> torch.Tensor:
    """b2r_resize_bilinear_u8 (Pillow BILINEAR on a ragged packed batch); see imageio.resize_batch for the tables."""
    _chk(src, torch.uint8, "src", 1)
    _chk(offsets, torch.int64, "offsets", 1)
    _chk(hw, torch.int32, "hw", 2)
    _chk(xtab_index, torch.int32, "xtab_index", 1)
    _chk(ytab_index, torch.int32, "ytab_index", 1)
    _chk(tabs, torch.int32, "tabs", 3)
    _chk(out, torch.uint8, "out", 4)
    n, out_h, out_w, c = out.shape
    if c != 3 or offsets.numel() != n or tuple(hw.shape) != (n, 2) or tabs.shape[1] != S or tabs.shape[2] != 2 + K:
        raise L.B2RError("inconsistent resize arguments")
    L.check(L.load().b2r_resize_bilinear_u8(src.data_ptr(), offsets.data_ptr(), hw.data_ptr(), xtab_index.data_ptr(),
                                            ytab_index.data_ptr(), tabs.data_ptr(), int(K), int(S), out.data_ptr(), int(n),
                                            int(out_h), int(out_w), int(tile_rows), int(max_rows), _stream()))
    STATS["launches"] += 1
    return out


def resize_cv_linear_u8(src: torch.Tensor, offsets: torch.Tensor, hw: torch.Tensor, xtab_index: torch.Tensor,
                        ytab_index: torch.Tensor, tabs: torch.Tensor, S: int, out: torch.Tensor) -> torch.Tensor:
    """b2r_resize_cv_linear_u8 (cv2.resize INTER_LINEAR on a ragged packed batch); see imageio.resize_batch_cv."""
    _chk(src, torch.uint8, "src", 1)
    _chk(offsets, torch.int64, "offsets", 1)
    _chk(hw, torch.int32, "hw", 2)
    _chk(xtab_index, torch.int32, "xtab_index", 1)
    _chk(ytab_index, torch.int32, "ytab_index", 1)
    _chk(tabs, torch.int32, "tabs", 3)
    _chk(out, torch.uint8, "out", 4)
    n, out_h, out_w, c = out.shape
    if c != 3 or offsets.numel() != n or tuple(hw.shape) != (n, 2) or tabs.shape[1] != S or tabs.shape[2] != 3:
        raise L.B2RError("inconsistent resize arguments")
    L.check(L.load().b2r_resize_cv_linear_u8(src.data_ptr(), offsets.data_ptr(), hw.data_ptr(), xtab_index.data_ptr(),
                                             ytab_index.data_ptr(), tabs.data_ptr(), int(S), out.data_ptr(), int(n),
                                             int(out_h), int(out_w), _stream()))
    STATS["launches"] += 1
    return out
