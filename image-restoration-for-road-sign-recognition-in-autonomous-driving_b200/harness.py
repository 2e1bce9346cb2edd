"""The reference's driver loops either side of the hot path, as functions over directory trees.

  reference                                                     here
  evaluate_model(model, data_dir, name)   18_test_unified_benchmark.py:22-53 (= 06:23-59, 09:29-65)   evaluate_model(judge, data_dir, name)
  run_inference()                         17_run_unified_inference.py:57-101                          run_inference(model, distorted_dir, restored_dir)
  datasets.ImageFolder(root)              18:35                                                       image_folder(root)
  process_task(task_name)                 08_run_inference.py:55-137                                  process_task(model, task_name, distorted_dir, restored_dir, clean_dir)

Files are decoded on the host (imageio.load_rgb), every pixel after that is touched by libb2r.so kernels: the Pillow
BILINEAR Resize((224, 224)) of the ragged batch, ToTensor (+ Normalize) fused into the first conv, the networks, the
arg-max / count.  Same prints, same return values and the same "skip when the path does not exist" behaviour as the
reference's functions.  There is no CPU fallback: the modules must live on a CUDA device.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import List, Optional, Tuple

import torch

from . import _lib as L
from . import generators, imageio, ops

# torchvision.datasets.folder.IMG_EXTENSIONS
IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def image_folder(root) -> Tuple[List[Tuple[str, int]], List[str]]:
    """`datasets.ImageFolder(root)` (18:35): (samples, classes) with samples = [(path, class_index), ...].
    Classes are the sub-directories of `root` in sorted order; inside each class the tree is walked (links followed)
    in sorted order and files with an image extension are kept (case-insensitive); class folders without any valid file
    and roots without class folders raise FileNotFoundError, as torchvision does."""
    root = os.path.expanduser(os.fspath(root))
    classes = sorted(e.name for e in os.scandir(root) if e.is_dir())
    if not classes:
        raise FileNotFoundError(f"Couldn't find any class folder in {root}.")
    samples: List[Tuple[str, int]] = []
    empty = []
    for idx, cls in enumerate(classes):
        before = len(samples)
        for d, _, fnames in sorted(os.walk(os.path.join(root, cls), followlinks=True)):
            for fname in sorted(fnames):
                if fname.lower().endswith(IMG_EXTENSIONS):
                    samples.append((os.path.join(d, fname), idx))
        if len(samples) == before:
            empty.append(cls)
    if empty:
        raise FileNotFoundError(f"Found no valid file for the classes {', '.join(sorted(empty))}. "
                                f"Supported extensions are: {', '.join(IMG_EXTENSIONS)}")
    return samples, classes


def _device_of(module: torch.nn.Module) -> torch.device:
    dev = next(module.parameters()).device
    if dev.type != "cuda":
        raise L.B2RError("the module must live on a CUDA device (call .cuda() first): there is no CPU fallback")
    return dev


def _on_module_device(fn):
    """Run a driver loop with the module's GPU as the current device: every op launches on the current device's current
    stream, so a module on cuda:1 must not be driven while cuda:0 is current (ops._chk rejects that)."""
    import functools

    @functools.wraps(fn)
    def wrapper(module, *a, **kw):
        try:
            dev = next(module.parameters()).device
        except (AttributeError, StopIteration):
            return fn(module, *a, **kw)      # e.g. a missing path is reported before the module is ever touched (18:23-25)
        if dev.type != "cuda":
            return fn(module, *a, **kw)      # _device_of raises the no-CPU-fallback error where the reference would run
        with torch.cuda.device(dev):
            return fn(module, *a, **kw)
    return wrapper


@_on_module_device
@torch.no_grad()
def evaluate_model(judge, data_dir, name: str, batch_size: int = 64, verbose: bool = True,
                   return_predictions: bool = False):
    """18:22-53: top-1 accuracy of the VGG16 judge over an ImageFolder tree.

    ImageFolder -> Resize((224, 224)) -> ToTensor -> Normalize(ImageNet) -> model -> torch.max(outputs, 1) ->
    correct / total, in batches of 64 (18:13) without shuffling.  Returns None when `data_dir` does not exist (18:23-25).
    The (correct, total) pair is accumulated on the device; one read at the end instead of `.item()` per batch."""
    if not os.path.exists(data_dir):
        if verbose:
            print(f"Skipping {name}: Path does not exist {data_dir}")
        return None
    samples, _ = image_folder(data_dir)
    dev = _device_of(judge)
    judge.eval()
    if verbose:
        print(f"Testing: {name} (Total {len(samples)} images)...")
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    preds = []
    for i in range(0, len(samples), batch_size):
        chunk = samples[i:i + batch_size]
        batch = imageio.load_batch([p for p, _ in chunk], device=dev)
        labels = torch.tensor([c for _, c in chunk], dtype=torch.int64, device=dev)
        pred = ops.argmax_count(judge.forward_u8(batch), labels=labels, counts=counts)[0]
        if return_predictions:
            preds.append(pred)
    correct, total = (int(v) for v in counts.cpu())
    acc = correct / total
    if verbose:
        print(f"-> {name} Accuracy: {acc * 100:.2f}%")
    return (acc, torch.cat(preds)) if return_predictions else acc


@_on_module_device
@torch.no_grad()
def run_inference(model, distorted_dir, restored_dir, batch_size: int = 32, pattern: str = "*/*.png",
                  verbose: bool = True) -> Optional[List[Path]]:
    """17:57-101: restore every `distorted_dir/<class>/<file>.png` and write it to the same relative path under
    `restored_dir`: Image.open + convert('RGB') + Resize((224, 224)) + ToTensor -> model -> clamp(0, 1) -> x 255 ->
    astype(uint8) -> PNG, in batches of 32 (17:15).  `model` is a restorer already loaded and on the device (the
    reference's "model not found" early return, 17:60-62, belongs to the caller that owns the checkpoint path).
    Returns the written paths."""
    distorted_dir, restored_dir = Path(distorted_dir), Path(restored_dir)
    dev = _device_of(model)
    model.eval()
    files = list(distorted_dir.glob(pattern))
    if verbose:
        print(f"Starting restoration of {len(files)} images...")
    written: List[Path] = []
    for i in range(0, len(files), batch_size):
        batch_files = files[i:i + batch_size]
        restored = model.restore_u8(imageio.load_batch(batch_files, device=dev))
        written += imageio.save_batch(restored, batch_files, distorted_dir, restored_dir)
    if verbose:
        print(f"Restoration complete! Please check: {restored_dir}")
    return written


@_on_module_device
@torch.no_grad()
def process_task(model, task_name: str, distorted_dir, restored_dir, clean_dir, batch_size: int = 32,
                 verbose: bool = True) -> Optional[Tuple[float, float, int]]:
    """08:55-137 for one task: restore every `distorted_dir/<class>/*.ppm|*.png`, write it as .png under the same
    relative path in `restored_dir` (08:103-109), and score it against the clean image `clean_dir / rel_path` (falling
    back to the .ppm name, 08:112-114) resized with cv2.resize(clean, (224, 224)) (08:119): PSNR and SSIM with
    data_range 255 (08:121-123), averaged over the images whose clean counterpart exists.
    The reference runs batch 1; here the same per-image arithmetic runs in batches on the device (restoration, both
    resizes, PSNR, SSIM).  `model` is a SimpleUNet already loaded and on the device.  Returns (mean PSNR, mean SSIM,
    count) or None when no image was scored, after the reference's prints."""
    distorted_dir, restored_dir, clean_dir = Path(distorted_dir), Path(restored_dir), Path(clean_dir)
    if verbose:
        print(f"\n=== Starting task processing: {task_name} ===")
    dev = _device_of(model)
    model.eval()
    files = list(distorted_dir.glob("*/*.ppm")) + list(distorted_dir.glob("*/*.png"))      # 08:84
    total_psnr = total_ssim = 0.0
    count = 0
    for i in range(0, len(files), batch_size):
        batch_files = files[i:i + batch_size]
        restored = model.restore_u8(imageio.load_batch(batch_files, device=dev))
        imageio.save_batch(restored, batch_files, distorted_dir, restored_dir, suffix=".png")
        keep, clean = [], []
        for k, f in enumerate(batch_files):
            cp = clean_dir / f.relative_to(distorted_dir)
            if not cp.exists():
                cp = cp.with_suffix(".ppm")
            if cp.exists():
                keep.append(k)
                clean.append(imageio.load_rgb(cp))
        if keep:
            ref = imageio.resize_batch_cv(clean, (224, 224), device=dev)
            out = restored[torch.tensor(keep, device=dev)]
            total_psnr += float(generators.psnr(ref, out).sum())
            total_ssim += float(generators.ssim(ref, out).sum())
            count += len(keep)
    if count > 0:
        if verbose:
            print(f"Task [{task_name}] completed.")
            print(f"Average PSNR: {total_psnr / count:.2f} dB")
            print(f"Average SSIM: {total_ssim / count:.4f}")
        return total_psnr / count, total_ssim / count, count
    if verbose:
        print("No images processed.")
    return None


def benchmark_table(judge, dirs, verbose: bool = True):
    """18:64-79: evaluate_model over {name: path} (TEST_DIRS, 18:15-19) and the report table; returns {name: accuracy}
    without the entries whose path does not exist, as the reference collects them."""
    if verbose:
        print("\n=== Starting Final Benchmark ===")
    results = {}
    for name, path in dict(dirs).items():
        acc = evaluate_model(judge, path, name, verbose=verbose)
        if acc is not None:
            results[name] = acc
    if verbose:
        print("\n" + "=" * 45)
        print("FINAL UNIFIED MODEL REPORT")
        print("=" * 45)
        print(f"{'Dataset Condition':<25} | {'Accuracy':<10}")
        print("-" * 45)
        for name, acc in results.items():
            print(f"{name:<25} | {acc * 100:.2f}%")
        print("=" * 45)
    return results
