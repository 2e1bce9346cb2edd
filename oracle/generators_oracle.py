"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

NumPy / OpenCV / PyTorch-fp32 restatement of the reference's single-degradation generators (02 / 03 / 04), of the
stress-test distortions and cascade of 13_pipeline_stress_test.py, and of the PSNR it reports in 08, with every random
draw made an explicit argument.  Third-party arithmetic is called exactly as the reference calls it (cv2 4.13.0,
numpy 2.3.5, torch 2.11 here; un-pinned in the reference).

Parity pinning: tests/test_oracle_generators.py compares this file with the reference's own functions run here
(02 / 03 / 04 imported; 13 cannot be imported because of matplotlib, so tests/golden/make_golden.py compiles ONLY its
add_noise / add_blur / add_fog definitions from the reference's source file) and with their committed outputs
tests/golden/generators_ref.npz for the GPU box, where /root/reference is absent.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import cv2
import numpy as np
import torch

from . import degrade_oracle as DO
from . import models_oracle as MO


def add_gaussian_noise_02(image: np.ndarray, noise: np.ndarray) -> np.ndarray:
    """02_gen_noise.py:12-27 with `noise` = the array np.random.normal(mean, var ** 0.5, image.shape) returned."""
    image = np.array(image / 255, dtype=float)
    out = image + noise
    low_clip = -1. if out.min() < 0 else 0.
    out = np.clip(out, low_clip, 1.0)
    return np.uint8(out * 255)


def apply_motion_blur_03(image: np.ndarray, degree: int = 10, angle: float = 45) -> np.ndarray:
    """03_gen_blur.py:11-30: filter2D, then the in-place min-max stretch over all channels jointly."""
    blurred = cv2.filter2D(np.array(image), -1, DO.blur_kernel(degree, angle))
    cv2.normalize(blurred, blurred, 0, 255, cv2.NORM_MINMAX)
    return np.array(blurred, dtype=np.uint8)


def normalize_minmax_table(smin: int, smax: int) -> np.ndarray:
    """What cv2.normalize(x, x, 0, 255, NORM_MINMAX) does to each of the 256 byte values when the image extrema are
    (smin, smax): scale / shift in double, cast to float, saturate(rint(fma(src, scale, shift))) (cvtScale8u of the
    OpenCV build in this image; checked against cv2 itself in the tests)."""
    scale = 255.0 * (1.0 / (smax - smin) if (smax - smin) > np.finfo(np.float64).eps else 0.0)
    shift = 0.0 - smin * scale
    a, b = np.float32(scale), np.float32(shift)
    src = np.arange(256, dtype=np.float64)
    fused = (src * np.float64(a) + np.float64(b)).astype(np.float32)   # single rounding = fma in float
    return np.clip(np.rint(fused), 0, 255).astype(np.uint8)


def add_fog_04(image: np.ndarray, u: float, fog_intensity: float = 0.8) -> Tuple[np.ndarray, float]:
    """04_gen_fog.py:12-31 with `u` = the value random.uniform(0.8, 1.2) returned.  Returns (image, t)."""
    image = np.array(image) / 255.0
    A = 0.9
    t = 1.0 - fog_intensity * u
    t = np.clip(t, 0.1, 0.9)
    fog_img = image * t + A * (1 - t)
    return np.clip(fog_img * 255, 0, 255).astype(np.uint8), float(t)


# --- 13_pipeline_stress_test.py:33-56 ---------------------------------------------------------------------------------
def stress_add_noise(image: np.ndarray, noise: np.ndarray) -> np.ndarray:
    img = image / 255.0
    out = np.clip(img + noise, 0.0, 1.0)
    return (out * 255).astype(np.uint8)


def stress_add_blur(image: np.ndarray) -> np.ndarray:
    return cv2.filter2D(image, -1, DO.blur_kernel(5, 45))


def stress_add_fog(image: np.ndarray) -> np.ndarray:
    img = image / 255.0
    A, t = 0.9, 1.0 - 0.1
    return np.clip((img * t + A * (1 - t)) * 255, 0, 255).astype(np.uint8)


def cascade_13(sds: Dict[str, MO.SD], distorted_u8_nhwc: torch.Tensor,
               order: Sequence[str] = ("Noise", "Fog", "Blur")) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """13:175-189: ToTensor, then each available SimpleUNet on the UNCLAMPED output of the previous one; the per-stage
    snapshots are clamp(0, 1) * 255 truncated to u8 (NHWC here)."""
    x = MO.to_tensor_u8(distorted_u8_nhwc)
    snaps = []
    for name in order:
        if name not in sds:
            continue
        x = MO.simple_unet_forward(sds[name], x)
        snaps.append(MO.quantize_restored(x))
    return x, snaps


def vgg_prediction(judge_sd: MO.SD, images_u8_nhwc: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_vgg_prediction, 13:87-92: softmax over the logits, (predicted, confidence) = arg-max and its probability."""
    logits = MO.vgg16_forward(judge_sd, MO.normalize_imagenet(MO.to_tensor_u8(images_u8_nhwc)))
    prob = torch.nn.functional.softmax(logits, dim=1)
    conf, pred = torch.max(prob, 1)
    return pred, conf


def psnr_08(a_u8: np.ndarray, b_u8: np.ndarray) -> float:
    """skimage.metrics.peak_signal_noise_ratio(a, b, data_range=255) as 08_run_inference.py:118-129 calls it (skimage is
    not installed here; its published definition: 10 log10(data_range^2 / mean((a - b)^2)) in float64)."""
    err = np.mean((a_u8.astype(np.float64) - b_u8.astype(np.float64)) ** 2)
    return float(10 * np.log10((255.0 ** 2) / err)) if err > 0 else float("inf")


def ssim_08(a_u8: np.ndarray, b_u8: np.ndarray, data_range: float = 255.0, win_size: int = 7, K1: float = 0.01,
            K2: float = 0.03) -> float:
    """skimage.metrics.structural_similarity(a, b, data_range=255, channel_axis=2) as 08_run_inference.py:123 calls it.

    PARITY UNPINNED BY EXECUTION: scikit-image is not installed in this image and the reference pins no version, so
    this restates its published algorithm (skimage/metrics/_structural_similarity.py, unchanged since 0.19) with the
    same scipy.ndimage.uniform_filter calls it makes: per channel, float64 (uint8 input), uniform 7x7 window (mode
    'reflect'), sample covariance NP / (NP - 1), C1 = (K1 R)^2, C2 = (K2 R)^2, S = (2 ux uy + C1)(2 vxy + C2) /
    ((ux^2 + uy^2 + C1)(vx + vy + C2)), mean of S with a border of (win_size - 1) // 2 cropped, then the mean over
    channels.  tests/test_oracle_generators.py cross-checks it against an exact integer-window evaluation."""
    from scipy.ndimage import uniform_filter
    if min(a_u8.shape[0], a_u8.shape[1]) < win_size:
        raise ValueError("win_size exceeds image extent")
    NP = win_size ** 2
    cov_norm = NP / (NP - 1)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    pad = (win_size - 1) // 2
    per_channel = []
    for ch in range(a_u8.shape[2]):
        im1 = a_u8[..., ch].astype(np.float64)
        im2 = b_u8[..., ch].astype(np.float64)
        ux = uniform_filter(im1, size=win_size)
        uy = uniform_filter(im2, size=win_size)
        uxx = uniform_filter(im1 * im1, size=win_size)
        uyy = uniform_filter(im2 * im2, size=win_size)
        uxy = uniform_filter(im1 * im2, size=win_size)
        vx = cov_norm * (uxx - ux * ux)
        vy = cov_norm * (uyy - uy * uy)
        vxy = cov_norm * (uxy - ux * uy)
        A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
        S = (A1 * A2) / (B1 * B2)
        per_channel.append(S[pad:S.shape[0] - pad, pad:S.shape[1] - pad].mean(dtype=np.float64))
    return float(np.mean(np.asarray(per_channel, dtype=np.float64)))
