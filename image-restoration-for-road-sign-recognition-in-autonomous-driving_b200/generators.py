"""The reference's single-degradation dataset generators and its cascade stress test, batched on the GPU with the
reference's exact arithmetic.

  reference (one image, NumPy / OpenCV on the CPU)                       here (one batch, device resident)
  add_gaussian_noise(image, var=0.01)    02_gen_noise.py:12-27           add_gaussian_noise(images, var=...)
  apply_motion_blur(image, 10, 45)       03_gen_blur.py:11-30            apply_motion_blur(images, degree, angle)
  add_fog(image, fog_intensity=0.8)      04_gen_fog.py:12-31             add_fog(images, fog_intensity, rng=...)
  add_blur / add_fog / add_noise         13_pipeline_stress_test.py:33-56   stress_add_blur / stress_add_fog / stress_add_noise
  cascade Noise -> Fog -> Blur           13_pipeline_stress_test.py:27,175-189   CascadeRestorer
  psnr_metric(clean, out, data_range=255)  08_run_inference.py:118-129   psnr(clean, out)
  ssim_metric(clean, out, data_range=255, channel_axis=2)  08:123        ssim(clean, out)

Images are uint8 [N, H, W, 3] CUDA tensors (the reference's HWC arrays, batched).  Host code here only prepares
per-image scalars and 256-entry tables; every pixel is touched by libb2r.so kernels (csrc/generators.cu, degrade.cu).
There is no CPU fallback.

Exactness notes (pinned in tests/test_generators_host.py against NumPy / cv2 and in tests/test_generators_gpu.py):
* scripts 02, 04 and 13 work on `image / 255.0` in float64.  A point operation on a u8 image has 256 possible results
  per image, so fog is applied as a per-image table that this module evaluates with the reference's own NumPy
  expression; noise runs in float64 on the device.
* script 02 converts with `np.uint8(out * 255)` after clipping to [-1, 1] when any value is negative: negative values
  wrap modulo 256.  Reproduced (B2R_NOISE_CLIP_SCRIPT02).
* script 03 stretches the blurred image with cv2.normalize(NORM_MINMAX) over all channels jointly; OpenCV evaluates
  saturate(rint(fma(src, (float)scale, (float)shift))).  Reproduced.  The blur itself is OpenCV's direct filter2D for
  degree <= 11 (bit-exact) and its DFT path above (<= 1 LSB), see csrc/degrade.cu.
"""
from __future__ import annotations

import math
import random as _random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import degrade as D
from . import ops

NOISE_CLIP_SCRIPT02, NOISE_CLIP_UNIT = 0, 1


# ---------------------------------------------------------------------------------------------------------------------
# fog: float64 point operation -> per-image table
# ---------------------------------------------------------------------------------------------------------------------
def fog_table(t: float, A: float = 0.9) -> np.ndarray:
    """The 256 results of `np.clip((image / 255.0) * t + A * (1 - t)) * 255, 0, 255).astype(np.uint8)` (04:17-30)."""
    image = np.arange(256, dtype=np.uint8)
    img = np.array(image) / 255.0
    fog_img = img * t + A * (1 - t)
    return np.clip(fog_img * 255, 0, 255).astype(np.uint8)


def _apply_tables(images: torch.Tensor, tables: Sequence[np.ndarray]) -> torch.Tensor:
    lut = torch.from_numpy(np.stack(tables).astype(np.uint8)).to(images.device)
    return ops.lut_u8(images, lut)


def add_fog(images: torch.Tensor, fog_intensity: float = 0.8, rng: Optional[_random.Random] = None,
            t: Optional[Sequence[float]] = None) -> Tuple[torch.Tensor, List[float]]:
    """04_gen_fog.py add_fog on a batch: t = clip(1 - fog_intensity * random.uniform(0.8, 1.2), 0.1, 0.9) per image
    (drawn from `rng`, a random.Random like the reference's module-level generator) unless `t` is given.
    Returns (fogged u8 batch, the t used per image)."""
    n = int(images.shape[0])
    if t is None:
        rng = rng or _random.Random()
        t = [float(np.clip(1.0 - fog_intensity * rng.uniform(0.8, 1.2), 0.1, 0.9)) for _ in range(n)]
    t = [float(v) for v in t]
    if len(t) != n:
        raise L.B2RError(f"{len(t)} transmissions for {n} images")
    return _apply_tables(images, [fog_table(v) for v in t]), t


def stress_add_fog(images: torch.Tensor) -> torch.Tensor:
    """13_pipeline_stress_test.py add_fog: fog_intensity 0.1 -> t = 0.9, A = 0.9, float64."""
    n = int(images.shape[0])
    return _apply_tables(images, [fog_table(1.0 - 0.1)] * n)


# ---------------------------------------------------------------------------------------------------------------------
# noise: float64 on the device
# ---------------------------------------------------------------------------------------------------------------------
def _noise(images, var, noise, seed, image_index0, rule):
    n = int(images.shape[0])
    sigma = torch.full((n,), float(var) ** 0.5, dtype=torch.float32, device=images.device)
    out, flags = ops.noise02(images, sigma, noise=noise, seed=seed, image_index0=image_index0, clip_rule=rule)
    return out, flags


def add_gaussian_noise(images: torch.Tensor, mean: float = 0, var: float = 0.01, noise: Optional[torch.Tensor] = None,
                       seed: int = 0, image_index0: int = 0) -> torch.Tensor:
    """02_gen_noise.py add_gaussian_noise (the script calls it with var=0.02), including its wrap-around for negative
    values.  `noise` (float64, same shape) injects the reference's own draw; otherwise Philox keyed by the global image
    index.  `mean` must be 0 as in every call site of the reference."""
    if mean != 0:
        raise L.B2RError("only mean=0 is supported (the reference never uses another value)")
    return _noise(images, var, noise, seed, image_index0, NOISE_CLIP_SCRIPT02)[0]


def stress_add_noise(images: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
                     image_index0: int = 0) -> torch.Tensor:
    """13_pipeline_stress_test.py add_noise: var 0.01, clip to [0, 1], truncate."""
    return _noise(images, 0.01, noise, seed, image_index0, NOISE_CLIP_UNIT)[0]


# ---------------------------------------------------------------------------------------------------------------------
# blur (+ the min-max stretch of script 03)
# ---------------------------------------------------------------------------------------------------------------------
def _blur(images: torch.Tensor, degree: int, angle: float) -> torch.Tensor:
    return D.degrade(images, D.blur_params(int(images.shape[0]), degree, angle))


def apply_motion_blur(images: torch.Tensor, degree: int = 10, angle: float = 45) -> torch.Tensor:
    """03_gen_blur.py apply_motion_blur (the script calls it with degree=12): filter2D, then
    cv2.normalize(blurred, blurred, 0, 255, cv2.NORM_MINMAX) with the extrema taken over all three channels."""
    blurred = _blur(images, degree, angle)
    return ops.normalize_minmax_u8(blurred, ops.minmax_u8(blurred), out=blurred)


def stress_add_blur(images: torch.Tensor) -> torch.Tensor:
    """13_pipeline_stress_test.py add_blur: degree 5, 45 degrees, no stretch."""
    return _blur(images, 5, 45)


def stress_distort(images: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
                   image_index0: int = 0) -> List[torch.Tensor]:
    """Phase 1 of 13_pipeline_stress_test.py:152-171: Blur -> Fog -> Noise with u8 re-quantisation after every stage.
    Returns the three intermediate u8 batches [+Blur, +Fog, +Noise (input of the cascade)]."""
    b = stress_add_blur(images)
    f = stress_add_fog(b)
    z = stress_add_noise(f, noise=noise, seed=seed, image_index0=image_index0)
    return [b, f, z]


# ---------------------------------------------------------------------------------------------------------------------
# PSNR / SSIM
# ---------------------------------------------------------------------------------------------------------------------
def psnr(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """skimage.metrics.peak_signal_noise_ratio(a, b, data_range=255) per image of two u8 batches (08:118-129):
    10 * log10(255^2 / mean((a - b)^2)) in float64; +inf for identical images.  The sum of squares is exact (uint64)."""
    sse = ops.sse_u8(a, b).to(torch.float64)
    mse = sse / float(a[0].numel())
    return 10.0 * torch.log10((255.0 ** 2) / mse)


def ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 255.0) -> torch.Tensor:
    """skimage.metrics.structural_similarity(a, b, data_range=255, channel_axis=2) per image of two u8 [N, H, W, 3]
    batches (08:123): skimage's defaults (7x7 uniform window, sample covariance, K1 = .01, K2 = .03, float64, border
    of 3 cropped, mean over channels).  f64 [N]; images smaller than the window raise, as skimage does.
    Parity note: skimage is not installed in this image, so this is pinned to the oracle's restatement of skimage's published
    algorithm (same scipy.ndimage.uniform_filter calls), not to an execution of skimage itself."""
    return ops.ssim_u8(a, b, data_range)


# ---------------------------------------------------------------------------------------------------------------------
# cascade restoration (13_pipeline_stress_test.py:175-189)
# ---------------------------------------------------------------------------------------------------------------------
RESTORATION_ORDER = ["Noise", "Fog", "Blur"]   # 13:27


def quantize_for_display(x: torch.Tensor) -> torch.Tensor:
    """13:183-186 (= 17:86-92): clamp(0, 1) -> * 255 -> astype(uint8) (truncation), f32 NCHW -> u8 NHWC."""
    return (torch.clamp(x, 0, 1).permute(0, 2, 3, 1) * 255).to(torch.uint8).contiguous()


class CascadeRestorer:
    """Three task-specific SimpleUNets applied back to back, Noise -> Fog -> Blur (13:27), on the compound input.

    As in the reference the hand-off between the networks is the UNCLAMPED float32 tensor (13:180); the clamp and the
    u8 quantisation are applied only to the per-stage visualisation copies (13:183-186).  Missing models are skipped,
    as the reference skips checkpoints that do not exist (13:100-108, 13:178).  Everything stays on the device."""

    def __init__(self, models: Dict[str, torch.nn.Module], order: Sequence[str] = tuple(RESTORATION_ORDER)):
        self.models = {k: m.eval() for k, m in models.items()}
        self.order = list(order)

    @torch.no_grad()
    def __call__(self, images_u8: torch.Tensor) -> Tuple[torch.Tensor, List[Tuple[str, torch.Tensor]]]:
        """u8 NHWC batch -> (final float32 NCHW tensor, [(stage name, u8 NHWC snapshot after that stage), ...])."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3:
            raise L.B2RError("CascadeRestorer takes the distorted batch as uint8 [N, H, W, 3]")
        history: List[Tuple[str, torch.Tensor]] = []
        x: torch.Tensor = images_u8          # ToTensor of the first stage is fused into the network's first layer
        for name in self.order:
            net = self.models.get(name)
            if net is None:
                continue
            x = net(x)                        # u8 NHWC or f32 NCHW in, unclamped f32 NCHW out
            history.append((name, quantize_for_display(x)))
        if x.dtype == torch.uint8:           # no model at all: the reference would carry ToTensor(img) forward
            x = images_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0
        return x, history


@torch.no_grad()
def judge_confidence(judge, images_u8: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_vgg_prediction (13:87-92, = 15:125-129) on a u8 NHWC batch: (predicted int64 [N], confidence f32 [N]) with
    confidence = max softmax probability."""
    logits = judge.forward_u8(images_u8)
    return ops.argmax_count(logits, want_conf=True)
