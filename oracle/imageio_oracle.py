"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

The reference's file -> tensor edge exactly as its scripts call it (17_run_unified_inference.py:66,79-82;
18_test_unified_benchmark.py:28-30): Pillow decode, `.convert('RGB')`, torchvision `transforms.Resize((224, 224))` on the
PIL image.  Third-party arithmetic: Pillow 12.2.0 (BILINEAR resampling with anti-aliasing), torchvision 0.26.0 here,
un-pinned in the reference.  ToTensor (the /255) is models_oracle.to_tensor_u8.

Parity pinning: there is nothing to restate here (these ARE the reference's calls); the product's restatement of Pillow's
arithmetic (imageio.resample_table + csrc/generators.cu) is compared with this file in tests/test_imageio_*.py.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
from PIL import Image
from torchvision import transforms


def resize_pil(img_u8_hwc: np.ndarray, size=(224, 224)) -> np.ndarray:
    """transforms.Resize(size) applied to Image.fromarray(img) -> u8 [size[0], size[1], 3]."""
    return np.asarray(transforms.Resize(size)(Image.fromarray(img_u8_hwc)))


def load_and_resize(files: Sequence, size=(224, 224)) -> np.ndarray:
    """The batch-preparation loop of 17:76-82 up to (not including) ToTensor: -> u8 [N, size[0], size[1], 3]."""
    t = transforms.Resize(size)
    return np.stack([np.asarray(t(Image.open(p).convert("RGB"))) for p in files])


def resize_cv(img_u8_hwc: np.ndarray, size=(224, 224)) -> np.ndarray:
    """`cv2.resize(clean_img, (224, 224))` as 08_run_inference.py:119 calls it (default INTER_LINEAR; dsize is (width,
    height)).  This IS the reference's call (cv2 4.13.0 here, un-pinned in the reference)."""
    import cv2
    return cv2.resize(img_u8_hwc, (int(size[1]), int(size[0])))
