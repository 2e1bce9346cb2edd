"""Host side of the fused degradation kernel: blur-tap construction and per-image parameter batches.

Mirrors the reference's degradation functions by name and argument meaning, but produces *parameters* for one
`b2r_degrade` launch over a whole batch instead of looping over images on the CPU:

  reference (per image, NumPy/OpenCV)                         here (per batch, one kernel launch)
  apply_compound_distortion(img)      16_gen_compound_data.py:14-37      compound_params(n)        + degrade()
  apply_random_distortions(img)       14_train_unified_advanced.py:31-64  random_params(n, rng)     + degrade()
  make_compound_distortion(img)       15_test_unified.py:93-120          demo_params(n)            + degrade()
  add_gaussian_noise / apply_motion_blur / add_fog   02:12-27, 03:11-30, 04:12-31   noise_params / blur_params / fog_params

`motion_blur_kernel` restates `cv2.getRotationMatrix2D` + `cv2.warpAffine(np.diag(np.ones(d)), M, (d, d)) / d`
(16:23-25) including OpenCV's fixed-point bilinear sampling, so the taps are bit-identical to the reference's
(tests/test_degrade_host.py checks all 14 x 361 (degree, angle) pairs against cv2 and the committed fixture).
Host-side scalar math only; all pixel work happens in csrc/degrade.cu.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from functools import lru_cache
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import ops

MAX_BLUR = L.B2R_MAX_BLUR
_AB_SCALE = 1024      # OpenCV warpAffine: AB_BITS = 10
_ROUND_DELTA = 16     # AB_SCALE / INTER_TAB_SIZE / 2
_F32 = np.float32


def _rint(v: float) -> int:
    return int(np.rint(v))  # cvRound: round half to even


@lru_cache(maxsize=None)
def _motion_blur_kernel_cached(degree: int, angle: float) -> np.ndarray:
    d = int(degree)
    cx = cy = d / 2
    a = angle * math.pi / 180.0
    alpha, beta = math.cos(a), math.sin(a)
    # cv2.getRotationMatrix2D((d/2, d/2), angle, 1)
    m = [alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy]
    # warpAffine without WARP_INVERSE_MAP inverts the matrix first
    det = m[0] * m[4] - m[1] * m[3]
    det = 1.0 / det if det != 0 else 0.0
    a11, a22 = m[4] * det, m[0] * det
    m[0] = a11
    m[1] *= -det
    m[3] *= -det
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    out = np.zeros((d, d), np.float64)
    inv32 = _F32(1 / 32)

    def src(yy: int, xx: int) -> float:  # np.diag(np.ones(d)) with BORDER_CONSTANT 0
        return 1.0 if (0 <= yy < d and yy == xx) else 0.0

    for y in range(d):
        x0 = _rint((m[1] * y + m[2]) * _AB_SCALE) + _ROUND_DELTA
        y0 = _rint((m[4] * y + m[5]) * _AB_SCALE) + _ROUND_DELTA
        for x in range(d):
            xf = (x0 + _rint(m[0] * x * _AB_SCALE)) >> 5   # 5 fractional bits (INTER_BITS)
            yf = (y0 + _rint(m[3] * x * _AB_SCALE)) >> 5
            sx, sy = xf >> 5, yf >> 5
            fx, fy = _F32(xf & 31) * inv32, _F32(yf & 31) * inv32
            wx0, wy0 = _F32(1) - fx, _F32(1) - fy
            w = (_F32(wy0 * wx0), _F32(wy0 * fx), _F32(fy * wx0), _F32(fy * fx))   # float bilinear table
            out[y, x] = (src(sy, sx) * float(w[0]) + src(sy, sx + 1) * float(w[1])
                         + src(sy + 1, sx) * float(w[2]) + src(sy + 1, sx + 1) * float(w[3]))
    out /= d
    out.setflags(write=False)
    return out


def motion_blur_kernel(degree: int, angle: float) -> np.ndarray:
    """float64 [d, d] motion-blur kernel exactly as the reference builds it (16_gen_compound_data.py:23-25)."""
    if not 1 <= int(degree) <= MAX_BLUR:
        raise ValueError(f"degree must be in 1..{MAX_BLUR}, got {degree}")
    return _motion_blur_kernel_cached(int(degree), float(angle))


@dataclass
class DegradeParams:
    """Per-image parameters of one b2r_degrade launch (host copies; `to(device)` uploads them once)."""
    n: int
    order: int = L.B2R_ORDER_BLUR_FOG_NOISE
    flags: int = 0
    ksize: np.ndarray = field(default=None)      # int32 [n]   0 = no blur
    taps: np.ndarray = field(default=None)       # f32 [n,225] row-major d x d, pitch d
    fog_on: np.ndarray = field(default=None)     # int32 [n]
    fog_t: np.ndarray = field(default=None)      # f32 [n]
    fog_add: np.ndarray = field(default=None)    # f32 [n]   float32(A * (1 - t)) evaluated in double
    sigma: np.ndarray = field(default=None)      # f32 [n]   0 = no noise

    def __post_init__(self):
        n = self.n
        if self.ksize is None:
            self.ksize = np.zeros(n, np.int32)
        if self.taps is None:
            self.taps = np.zeros((n, MAX_BLUR * MAX_BLUR), np.float32)
        if self.fog_on is None:
            self.fog_on = np.zeros(n, np.int32)
        if self.fog_t is None:
            self.fog_t = np.ones(n, np.float32)
        if self.fog_add is None:
            self.fog_add = np.zeros(n, np.float32)
        if self.sigma is None:
            self.sigma = np.zeros(n, np.float32)

    # -- builders for single images ---------------------------------------------------------------------------
    def set_blur(self, i: int, degree: int, angle: float) -> None:
        if degree > 1:                                        # "if degree > 1" (14:56)
            k = motion_blur_kernel(degree, angle).astype(np.float32)   # filter2D on u8 uses a float32 kernel
            self.ksize[i] = degree
            self.taps[i, :] = 0
            self.taps[i, :degree * degree] = k.reshape(-1)
        else:
            self.ksize[i] = 0

    def set_fog(self, i: int, t: float, A: float = 0.9) -> None:
        self.fog_on[i] = 1
        self.fog_t[i] = np.float32(t)
        self.fog_add[i] = np.float32(A * (1 - t))             # Python evaluates A * (1 - t) in double (16:31)

    def set_noise(self, i: int, var: float) -> None:
        self.sigma[i] = np.float32(var ** 0.5)                # np.random.normal(0, var ** 0.5, ...) (16:34)

    # -- the same parameters for every image of the batch (no per-image Python loop) ----------------------------------
    def set_blur_all(self, degree: int, angle: float) -> None:
        self.set_blur(0, degree, angle)
        self.ksize[:] = self.ksize[0]
        self.taps[:] = self.taps[0]

    def set_fog_all(self, t: float, A: float = 0.9) -> None:
        self.set_fog(0, t, A)
        self.fog_on[:], self.fog_t[:], self.fog_add[:] = 1, self.fog_t[0], self.fog_add[0]

    def set_noise_all(self, var: float) -> None:
        self.set_noise(0, var)
        self.sigma[:] = self.sigma[0]

    def to(self, device) -> "DeviceDegradeParams":
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        return DeviceDegradeParams(self.n, self.order, self.flags, t(self.ksize), t(self.taps), t(self.fog_on),
                                   t(self.fog_t), t(self.fog_add), t(self.sigma), bool((self.ksize > 1).any()),
                                   bool((self.sigma > 0).any()))


@dataclass
class DeviceDegradeParams:
    n: int
    order: int
    flags: int
    ksize: torch.Tensor
    taps: torch.Tensor
    fog_on: torch.Tensor
    fog_t: torch.Tensor
    fog_add: torch.Tensor
    sigma: torch.Tensor
    any_blur: bool
    any_noise: bool = True    # False: no image gets noise -> sigma is not passed (point-wise kernel when nothing blurs either)


def compound_params(n: int) -> DegradeParams:
    """apply_compound_distortion (16_gen_compound_data.py:14-37): Blur(10, 45 deg) -> Fog(0.5, A=0.9) -> Noise(var .02)."""
    p = DegradeParams(n, order=L.B2R_ORDER_BLUR_FOG_NOISE)
    p.set_blur_all(10, 45)
    p.set_fog_all(1.0 - 0.5)
    p.set_noise_all(0.02)
    return p


def demo_params(n: int) -> DegradeParams:
    """make_compound_distortion (15_test_unified.py:93-120): Fog -> Noise -> clip -> u8 -> Blur(10, 45 deg)."""
    p = DegradeParams(n, order=L.B2R_ORDER_FOG_NOISE_BLUR, flags=L.B2R_DEG_CLIP_AFTER_NOISE)
    p.set_fog_all(1.0 - 0.5)
    p.set_noise_all(0.02)
    p.set_blur_all(10, 45)
    return p


def random_params(n: int, rng: np.random.Generator, order: int = L.B2R_ORDER_FOG_NOISE_BLUR,
                  p_fog: float = 0.5, p_noise: float = 0.5, p_blur: float = 0.5) -> DegradeParams:
    """apply_random_distortions (14_train_unified_advanced.py:31-64): each stage with probability 0.5; intensity
    U(.3,.7), t = 1 - intensity*U(.8,1.2); var U(.01,.03); degree randint[5,15], angle randint[0,360].
    `order` defaults to script 14's Fog->Noise->Blur; BASELINE config 3 applies the same draws in script 16's order."""
    p = DegradeParams(n, order=order)
    for i in range(n):
        if rng.random() < p_fog:
            intensity = rng.uniform(0.3, 0.7)
            p.set_fog(i, 1.0 - intensity * rng.uniform(0.8, 1.2))
        if rng.random() < p_noise:
            p.set_noise(i, rng.uniform(0.01, 0.03))
        if rng.random() < p_blur:
            p.set_blur(i, int(rng.integers(5, 16)), int(rng.integers(0, 361)))
    return p


def fog_params(n: int, rng: Optional[np.random.Generator] = None, fog_intensity: float = 0.8) -> DegradeParams:
    """add_fog (04_gen_fog.py:12-31): t = clip(1 - intensity*U(.8,1.2), .1, .9) per image, A = 0.9.
    NOTE: script 04 works in float64 (`np.array(image) / 255.0`); this kernel's fog stage is the float32 form used by
    scripts 14/15/16, so results may differ from script 04 by one u8 LSB on truncation boundaries."""
    rng = rng or np.random.default_rng()
    p = DegradeParams(n)
    for i in range(n):
        t = float(np.clip(1.0 - fog_intensity * rng.uniform(0.8, 1.2), 0.1, 0.9))
        p.set_fog(i, t)
    return p


def blur_params(n: int, degree: int = 12, angle: float = 45) -> DegradeParams:
    """apply_motion_blur (03_gen_blur.py:11-30) WITHOUT its cv2.normalize min-max stretch."""
    p = DegradeParams(n)
    p.set_blur_all(degree, angle)
    return p


def noise_params(n: int, var: float = 0.02) -> DegradeParams:
    """add_gaussian_noise (02_gen_noise.py:12-27) with the [0,1] clip; the script's wrap-around for negative values
    (np.uint8 of a negative float) is a dataset-generator quirk outside this path."""
    p = DegradeParams(n, flags=L.B2R_DEG_CLIP_AFTER_NOISE)
    p.set_noise_all(var)
    return p


def degrade(images_u8_nhwc: torch.Tensor, params, seed: int = 0, image_index0: int = 0,
            noise: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fused launch over the batch.  `noise` (f64 [N,H,W,3], optional) injects the reference's noise tensor."""
    if isinstance(params, DegradeParams):
        params = params.to(images_u8_nhwc.device)
    if params.n != images_u8_nhwc.shape[0]:
        raise L.B2RError(f"params describe {params.n} images, batch has {images_u8_nhwc.shape[0]}")
    return ops.degrade(images_u8_nhwc, params.ksize if params.any_blur else None,
                       params.taps if params.any_blur else None, params.fog_on, params.fog_t, params.fog_add,
                       params.sigma if (params.any_noise or noise is not None) else None, noise=noise, seed=seed, image_index0=image_index0, order=params.order,
                       flags=params.flags, out=out)
