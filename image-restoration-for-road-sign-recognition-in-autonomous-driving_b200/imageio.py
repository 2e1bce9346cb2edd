"""The on-disk edge either side of the hot path: image files -> Resize((224, 224)) -> device
batch, and restored batches -> image files in the reference's tree layout.

  reference (17_run_unified_inference.py)                                  here
  Image.open(p).convert('RGB')                          17:79           load_rgb(p)   (own P6 parser for GTSRB's .ppm,
                                                                                       Pillow for .png and the rest)
  transforms.Resize((224, 224)) + torch.stack           17:66,79-82     resize_batch(images) -> u8 [N,224,224,3] on the
                                                                        device: ONE kernel over the ragged batch,
                                                                        Pillow's BILINEAR arithmetic bit for bit
  ToTensor()                                            17:66           fused into the first conv (u8 NHWC entry)
  save_path = RESTORED_DIR / rel_path; cv2.imwrite      17:89-99        save_batch(batch_u8, files, src_root, dst_root)

File decoding / encoding stays on the host (PNG is an entropy-coded stream); what moves to the GPU is the per-pixel work.
The host part of the resize is Pillow's coefficient construction (`precompute_coeffs` + `normalize_coeffs_8bpc`,
src/libImaging/Resample.c of Pillow, restated here in the same double arithmetic), cached per (input size, output size).
"""
from __future__ import annotations

import math
from functools import lru_cache
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops

PRECISION_BITS = 32 - 8 - 2      # Pillow: 8-bit samples, 2 guard bits


@lru_cache(maxsize=4096)
def resample_table(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's BILINEAR coefficients for resampling `in_size` samples to `out_size`.
    Returns (bounds int32 [out_size, 2] = (first input index, tap count), kk int32 [out_size, ksize] fixed point)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale                      # bilinear filter support = 1
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0.0 else v
            t = 1.0 - v if v < 1.0 else 0.0
            w.append(t)
            ww += t
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    bounds.setflags(write=False)
    kk.setflags(write=False)
    return bounds, kk


def read_ppm(path) -> np.ndarray:
    """Binary PPM (P6, maxval <= 255), the format of the GTSRB training set (02:36, 13:131): -> u8 [H, W, 3] RGB."""
    data = Path(path).read_bytes()
    pos, tokens = 0, []
    while len(tokens) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] not in (b"\n", b""):
                pos += 1
            continue
        start = pos
        while pos < len(data) and not data[pos:pos + 1].isspace():
            pos += 1
        tokens.append(data[start:pos])
    if tokens[0] != b"P6":
        raise L.B2RError(f"{path}: not a binary PPM (magic {tokens[0]!r})")
    w, h, maxval = int(tokens[1]), int(tokens[2]), int(tokens[3])
    if not 0 < maxval <= 255:
        raise L.B2RError(f"{path}: maxval {maxval} not supported")
    pos += 1                                           # exactly one whitespace byte after maxval
    buf = np.frombuffer(data, dtype=np.uint8, count=h * w * 3, offset=pos)
    return buf.reshape(h, w, 3)


def load_rgb(path) -> np.ndarray:
    """`Image.open(p).convert('RGB')` as u8 [H, W, 3]."""
    path = Path(path)
    if path.suffix.lower() == ".ppm":
        try:
            return read_ppm(path)
        except (L.B2RError, ValueError, IndexError):
            pass                                       # ASCII PPM or 16-bit: let Pillow handle it
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"))


def resize_batch(images: Sequence[np.ndarray], size: Tuple[int, int] = (224, 224), device=None,
                 tile_rows: int = 8, _return_plan: bool = False) -> torch.Tensor:
    """transforms.Resize(size) + torch.stack over a list of u8 [H_i, W_i, 3] host arrays -> u8 [N, size[0], size[1], 3]
    on the device.  One pinned host buffer, one H2D copy, one kernel launch for the whole ragged batch."""
    if not torch.cuda.is_available():
        raise L.B2RError("resize_batch needs a CUDA device (there is no CPU fallback)")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out_h, out_w = int(size[0]), int(size[1])
    n = len(images)
    if n == 0:
        return torch.empty((0, out_h, out_w, 3), dtype=torch.uint8, device=device)
    hw = np.zeros((n, 2), np.int32)
    offsets = np.zeros((n,), np.int64)
    total = 0
    for i, im in enumerate(images):
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise L.B2RError(f"image {i}: expected uint8 [H, W, 3], got {im.dtype} {im.shape}")
        hw[i] = im.shape[:2]
        offsets[i] = total
        total += im.size
    packed = torch.empty((total,), dtype=torch.uint8).pin_memory()
    pk = packed.numpy()
    for i, im in enumerate(images):
        pk[offsets[i]:offsets[i] + im.size] = np.ascontiguousarray(im).reshape(-1)
    # coefficient tables: one per distinct (input extent -> output extent)
    keys, xi, yi = {}, np.zeros((n,), np.int32), np.zeros((n,), np.int32)
    for i in range(n):
        xi[i] = keys.setdefault((int(hw[i, 1]), out_w), len(keys))
        yi[i] = keys.setdefault((int(hw[i, 0]), out_h), len(keys))
    tabs_l = [resample_table(*k) for k in keys]
    K = max(kk.shape[1] for _, kk in tabs_l)
    S = max(out_h, out_w)
    tabs = np.zeros((len(tabs_l), S, 2 + K), np.int32)
    for t, (bounds, kk) in enumerate(tabs_l):
        tabs[t, :bounds.shape[0], :2] = bounds
        tabs[t, :kk.shape[0], 2:2 + kk.shape[1]] = kk
    # shared-memory rows a tile of output rows needs, from the vertical tables actually used
    while True:
        max_rows = 1
        for key, t in keys.items():
            if key[1] != out_h or t not in set(yi.tolist()):
                continue
            b = tabs_l[t][0]
            for y0 in range(0, out_h, tile_rows):
                y1 = min(y0 + tile_rows, out_h)
                max_rows = max(max_rows, int((b[y0:y1, 0] + b[y0:y1, 1]).max() - b[y0, 0]))
        if max_rows * out_w * 3 <= 200 * 1024 or tile_rows == 1:
            break
        tile_rows = max(1, tile_rows // 2)
    dev = lambda a: torch.from_numpy(a).to(device, non_blocking=True)   # noqa: E731
    plan = dict(src=packed.to(device, non_blocking=True), offsets=dev(offsets), hw=dev(hw), xtab_index=dev(xi),
                ytab_index=dev(yi), tabs=dev(tabs), K=K, S=S, tile_rows=tile_rows, max_rows=max_rows)
    out = torch.empty((n, out_h, out_w, 3), dtype=torch.uint8, device=device)
    if _return_plan:
        return out, plan
    return ops.resize_bilinear_u8(out=out, **plan)


@lru_cache(maxsize=None)
def cv_linear_table(in_size: int, out_size: int) -> np.ndarray:
    """OpenCV's INTER_LINEAR coordinate arithmetic for u8 images (imgproc/resize.cpp, cv2 4.13) along one axis:
    int32 [out_size, 3] = {first source index s, cvRound((1 - f) * 2048), cvRound(f * 2048)} with
    f = (float)((d + 0.5) * (in / out) - 0.5), s = floor(f), f -= s (s may be -1 or >= in - 1: resize_batch_cv applies
    OpenCV's border rule, which differs between columns and rows)."""
    scale = float(in_size) / float(out_size)
    tab = np.zeros((out_size, 3), np.int32)
    for d in range(out_size):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        tab[d] = (s, int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048)))),
                  int(np.rint(np.float32(f * np.float32(2048)))))
    tab.setflags(write=False)
    return tab


def _pack_ragged(images: Sequence[np.ndarray]):
    """u8 [H_i, W_i, 3] host arrays -> (pinned packed bytes, offsets int64 [N], hw int32 [N, 2])."""
    n = len(images)
    hw = np.zeros((n, 2), np.int32)
    offsets = np.zeros((n,), np.int64)
    total = 0
    for i, im in enumerate(images):
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise L.B2RError(f"image {i}: expected uint8 [H, W, 3], got {im.dtype} {im.shape}")
        hw[i] = im.shape[:2]
        offsets[i] = total
        total += im.size
    packed = torch.empty((total,), dtype=torch.uint8).pin_memory()
    pk = packed.numpy()
    for i, im in enumerate(images):
        pk[offsets[i]:offsets[i] + im.size] = np.ascontiguousarray(im).reshape(-1)
    return packed, offsets, hw


def resize_batch_cv(images: Sequence[np.ndarray], size: Tuple[int, int] = (224, 224), device=None,
                    _return_plan: bool = False) -> torch.Tensor:
    """`cv2.resize(img, (size[1], size[0]))` (INTER_LINEAR, 08_run_inference.py:119) over a list of u8 [H_i, W_i, 3] host
    arrays -> u8 [N, size[0], size[1], 3] on the device, bit for bit; one H2D copy and one launch for the ragged batch.
    (Channel order is irrelevant: the arithmetic is per channel.)"""
    if not torch.cuda.is_available():
        raise L.B2RError("resize_batch_cv needs a CUDA device (there is no CPU fallback)")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out_h, out_w = int(size[0]), int(size[1])
    n = len(images)
    if n == 0:
        return torch.empty((0, out_h, out_w, 3), dtype=torch.uint8, device=device)
    packed, offsets, hw = _pack_ragged(images)
    keys, xi, yi = {}, np.zeros((n,), np.int32), np.zeros((n,), np.int32)
    for i in range(n):
        xi[i] = keys.setdefault((int(hw[i, 1]), out_w, "x"), len(keys))
        yi[i] = keys.setdefault((int(hw[i, 0]), out_h, "y"), len(keys))
    S = max(out_h, out_w)
    tabs = np.zeros((len(keys), S, 3), np.int32)
    for (src, dst, axis), t in keys.items():
        tab = cv_linear_table(src, dst).copy()
        if axis == "x":
            # columns: outside [0, src - 2] OpenCV takes the border pixel alone (index clamped, fraction zeroed)
            outside = (tab[:, 0] < 0) | (tab[:, 0] >= src - 1)
            tab[outside, 1], tab[outside, 2] = 2048, 0
            tab[:, 0] = np.clip(tab[:, 0], 0, src - 1)
        # rows: OpenCV keeps both coefficients and clamps the two row indices (the kernel does that)
        tabs[t, :dst] = tab
    dev = lambda a: torch.from_numpy(a).to(device, non_blocking=True)   # noqa: E731
    out = torch.empty((n, out_h, out_w, 3), dtype=torch.uint8, device=device)
    plan = dict(src=packed.to(device, non_blocking=True), offsets=dev(offsets), hw=dev(hw), xtab_index=dev(xi),
                ytab_index=dev(yi), tabs=dev(tabs), S=S)
    if _return_plan:
        return out, plan
    return ops.resize_cv_linear_u8(out=out, **plan)


def load_batch(files: Sequence, size: Tuple[int, int] = (224, 224), device=None) -> torch.Tensor:
    """The batch-preparation loop of 17_run_unified_inference.py:76-82 without ToTensor (fused downstream)."""
    return resize_batch([load_rgb(p) for p in files], size=size, device=device)


def save_batch(batch_u8: torch.Tensor, files: Sequence, src_root, dst_root, suffix: Optional[str] = None) -> List[Path]:
    """17:89-99: write image i to dst_root / files[i].relative_to(src_root) (parent folders created), RGB u8.
    The reference converts to BGR for cv2.imwrite, i.e. the file holds the same RGB image."""
    from PIL import Image
    if batch_u8.dtype != torch.uint8 or batch_u8.dim() != 4 or batch_u8.shape[3] != 3:
        raise L.B2RError("save_batch takes uint8 [N, H, W, 3]")
    host = batch_u8.cpu().numpy()
    written = []
    for i, f in enumerate(files):
        rel = Path(f).relative_to(src_root)
        if suffix:
            rel = rel.with_suffix(suffix)
        dst = Path(dst_root) / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        Image.fromarray(host[i]).save(dst)
        written.append(dst)
    return written
