"""The reference's OWN code on the host cores: `bench.py --impl reference` and the `cpu_baseline` leg.

Imports the unmodified scripts that tools/install_ref.py copied into git-ignored `baseline/_ref/` (sha256-checked
against MANIFEST.json) and drives them the way the reference's three scripts do, composed in memory (the PNG files the
scripts exchange are lossless, so skipping them changes no value):

    16_gen_compound_data.py  apply_compound_distortion(img)                      one image per call (16:43-47)
    17_run_unified_inference.py  ResUNet, Resize((224,224)) + ToTensor, batches of 32, clamp(0,1),
                                 (out * 255).astype(np.uint8)                     (17:66-92)
    18_test_unified_benchmark.py Resize + ToTensor + Normalize, vgg16 + classifier[6] = Linear(4096, 43),
                                 torch.max(outputs, 1), correct += (predicted == labels).sum().item()   (18:28-51, 58-59)

None of this repo's models, kernels or oracle code is on this path.  Weights are the same seeded synthetic state_dicts
the GPU arm loads (the shipped .pth files are not available offline), loaded with the reference's own strict
`load_state_dict`.  The degradation calls are spread over the host cores with a thread pool (NumPy / OpenCV release the
GIL): the reference runs them serially, this is the fair multi-core form.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF_DIR = ROOT / "baseline" / "_ref"
_MODS = {}


def available() -> bool:
    sys.path.insert(0, str(ROOT / "tools"))
    try:
        import install_ref
        return install_ref.check(REF_DIR)
    finally:
        sys.path.pop(0)


def load(script: str):
    """Import baseline/_ref/<script>.py (names start with digits -> importlib); import-time prints are swallowed
    (07_train_restoration.py:29-32 prints four lines at import)."""
    if script in _MODS:
        return _MODS[script]
    path = REF_DIR / f"{script}.py"
    spec = importlib.util.spec_from_file_location("ref_" + script, path)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    _MODS[script] = mod
    return mod


def build_models(arch: str, classify: bool, seed_restorer: int = 31, seed_judge: int = 32):
    """Reference restorer class + torchvision VGG16-43 with the GPU arm's synthetic checkpoints (strict load)."""
    import torch.nn as nn
    import torchvision
    from b200restore import synth
    if arch == "resunet":
        restorer = load("17_run_unified_inference").ResUNet()
    elif arch == "simple_unet":
        restorer = load("07_train_restoration").SimpleUNet()
    else:
        raise ValueError(arch)
    restorer.load_state_dict(synth.synthetic_state_dict(arch, seed_restorer))
    restorer.eval()                                                            # 17:64
    judge = None
    if classify:
        judge = torchvision.models.vgg16(weights=None)                         # 18:58 ('DEFAULT' would download: offline)
        judge.classifier[6] = nn.Linear(judge.classifier[6].in_features, 43)   # 18:59
        judge.load_state_dict(synth.synthetic_state_dict("vgg16", seed_judge))  # 18:61
        judge.eval()
    return restorer, judge


def images_per_s(arch: str, recipe: str, classify: bool, hw: int, sample: int, repeats: int, warmup: int = 1):
    """Time `sample` images per pass through the reference path; returns (images/s, threads, seconds per pass, correct)."""
    import numpy as np
    import torch
    from PIL import Image
    from torchvision import transforms
    from b200restore import synth
    from concurrent.futures import ThreadPoolExecutor

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    restorer, judge = build_models(arch, classify)
    if recipe == "compound16":
        degrade = load("16_gen_compound_data").apply_compound_distortion
    elif recipe == "random14":
        degrade = load("14_train_unified_advanced").apply_random_distortions
    else:
        raise ValueError(recipe)
    resize = [transforms.Resize((224, 224))] if hw == 224 else []   # identity at 224; other sizes: the sweep's native size
    t17 = transforms.Compose(resize + [transforms.ToTensor()])                                           # 17:66
    t18 = transforms.Compose(resize + [transforms.ToTensor(),
                                       transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])  # 18:28-32
    imgs, labels = synth.sign_like_images(sample, hw, hw, seed=7)
    imgs_np = imgs.numpy()
    pool = ThreadPoolExecutor(max_workers=cores)
    chunk = 32                                                                                         # 17:17 BATCH_SIZE

    def one_chunk(lo, hi):
        bad = list(pool.map(lambda i: degrade(imgs_np[i]), range(lo, hi)))                             # 16:47
        inputs = [t17(Image.fromarray(b)) for b in bad]                                                # 17:78-80
        input_tensor = torch.stack(inputs)                                                             # 17:82
        with torch.no_grad():
            output_tensor = restorer(input_tensor)                                                     # 17:85
            output_tensor = torch.clamp(output_tensor, 0, 1)                                           # 17:86
        restored = []
        for idx in range(hi - lo):
            out_img = output_tensor[idx].cpu().permute(1, 2, 0).numpy()                                # 17:91
            restored.append((out_img * 255).astype(np.uint8))                                          # 17:92
        if judge is None:
            return int(restored[0].sum() > 0)
        x = torch.stack([t18(Image.fromarray(r)) for r in restored])
        with torch.no_grad():
            outputs = judge(x)                                                                         # 18:46
            _, predicted = torch.max(outputs, 1)                                                       # 18:47
        return (predicted == labels[lo:hi]).sum().item()                                               # 18:49

    def one_pass(limit=sample):
        return sum(one_chunk(lo, min(lo + chunk, limit)) for lo in range(0, limit, chunk))

    for _ in range(max(warmup, 1)):
        one_pass(min(sample, chunk))
    times, correct = [], 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        correct = one_pass()
        times.append(time.perf_counter() - t0)
    pool.shutdown()
    dt = statistics.median(times)
    return sample / dt, cores, dt, correct
