"""Seeded synthetic inputs and checkpoints in the reference's exact schemas.

The four shipped `restoration_*.pth` files and `vgg16_baseline.pth` are not available offline (SURVEY.md "Three
facts" #1), so benchmarks and parity tests run on synthetic state_dicts that have the same keys, shapes and dtypes
as `torch.save(model.state_dict(), path)` produces in the reference (07_train_restoration.py:178-180,
14_train_unified_advanced.py:267) — including BatchNorm running statistics and `num_batches_tracked` int64 scalars
— with values chosen so that BN folding, PReLU slopes and biases all matter (default inits would hide bugs).
Generated on the CPU with a seeded torch.Generator, so the same seed gives the same checkpoint everywhere.
"""
from __future__ import annotations

from collections import OrderedDict

import torch


def _fill(name: str, t: torch.Tensor, g: torch.Generator) -> torch.Tensor:
    shape = tuple(t.shape)
    if name.endswith("num_batches_tracked"):
        return torch.tensor(1000, dtype=torch.int64)
    if name.endswith("running_var"):
        return torch.rand(shape, generator=g) + 0.5
    if name.endswith("running_mean"):
        return torch.randn(shape, generator=g) * 0.1
    if t.dim() == 1 and shape == (1,):                      # nn.PReLU() slope
        return torch.rand(shape, generator=g) * 0.3 + 0.1
    if t.dim() == 1:
        if ".conv_block.1." in name or ".conv_block.4." in name or ".shortcut.1." in name:
            if name.endswith("weight"):                      # BN gamma
                gamma = torch.rand(shape, generator=g) + 0.5
                # the two branches that are summed in a ResidualBlock are damped (x0.6, tuned on the oracle) so the
                # activation RMS stays ~constant through the nine blocks instead of doubling per block
                return gamma if ".conv_block.1." in name else gamma * 0.6
            return torch.randn(shape, generator=g) * 0.1     # BN beta
        if name == "final.bias":                             # restorer outputs centred inside [0, 1]
            return 0.5 + torch.randn(shape, generator=g) * 0.05
        return torch.randn(shape, generator=g) * 0.05        # conv / linear bias
    if t.dim() == 4:
        if name.startswith("up"):                            # ConvTranspose2d [C_in, C_out, 2, 2]: one tap per output
            fan_in = shape[0]
            return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        fan_in = shape[1] * shape[2] * shape[3]
        gain = 0.1 if name.startswith("final") else (1.0 if ".shortcut." in name else 2.0)
        return torch.randn(shape, generator=g) * (gain / fan_in) ** 0.5
    if t.dim() == 2:                                         # nn.Linear
        return torch.randn(shape, generator=g) * (2.0 / shape[1]) ** 0.5
    raise ValueError(f"no synthetic rule for {name} {shape}")


def synthetic_state_dict(arch: str, seed: int = 0, num_classes: int = 43) -> "OrderedDict[str, torch.Tensor]":
    """arch in {'simple_unet', 'resunet', 'vgg16'} -> OrderedDict with the reference's keys / shapes / dtypes."""
    from . import models
    if arch == "simple_unet":
        m = models.SimpleUNet()
    elif arch == "resunet":
        m = models.ResUNet()
    elif arch == "vgg16":
        m = models.VGG16Judge(num_classes)
    else:
        raise ValueError(f"unknown arch {arch!r}")
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = OrderedDict()
    for k, v in m.state_dict().items():
        out[k] = _fill(k, v, g).to(v.dtype if v.dtype == torch.int64 else torch.float32)
    return out


def stress_state_dict(arch: str, seed: int = 0, num_classes: int = 43) -> "OrderedDict[str, torch.Tensor]":
    """A deliberately hostile checkpoint in the same schema: what `synthetic_state_dict` tunes away is put back.
    BatchNorm gammas U(0.5, 1.5) on BOTH branches of every ResidualBlock (no damping: the activation RMS grows block by
    block), running_var log-uniform in [1e-3, 10] with the preceding conv scaled to match (pre-BN activations spanning four
    orders of magnitude per channel; eps = 1e-5 is 1 % of the smallest variance), PReLU slopes U(0, 1), conv biases and BN
    betas N(0, 0.5).  Used by tests/test_stress_checkpoints_gpu.py to state the
    bf16 tolerance that survives a wide dynamic range (the shipped restoration_*.pth are not available offline)."""
    import math
    base = synthetic_state_dict(arch, seed, num_classes)
    g = torch.Generator(device="cpu").manual_seed(seed + 7919)
    out = OrderedDict()
    for k, v in base.items():
        shape = tuple(v.shape)
        if k.endswith("running_var"):
            v = torch.exp(torch.rand(shape, generator=g) * (math.log(10.0) - math.log(1e-3)) + math.log(1e-3))
        elif k.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.5
        elif v.dim() == 1 and shape == (1,):
            v = torch.rand(shape, generator=g)
        elif v.dim() == 1 and (".conv_block.1." in k or ".conv_block.4." in k or ".shortcut.1." in k):
            v = (torch.rand(shape, generator=g) + 0.5) if k.endswith("weight") else torch.randn(shape, generator=g) * 0.5
        elif v.dim() == 1 and k.endswith("bias") and not k.startswith("final") and not k.startswith("classifier"):
            v = torch.randn(shape, generator=g) * 0.5
        out[k] = v
    # a trained BatchNorm's running_var IS the variance of the conv output it follows: give every conv in front of a BN the
    # per-channel scale sqrt(running_var), so the BN output stays O(1) per channel while the folded scale spans 1e-3 .. 10
    for k in list(out):
        if k.endswith("running_var"):
            bn = k[:-len(".running_var")]                       # e.g. res1.conv_block.1
            head, idx = bn.rsplit(".", 1)
            conv = f"{head}.{int(idx) - 1}"
            sd_ = torch.sqrt(out[k])
            out[conv + ".weight"] = out[conv + ".weight"] * sd_.view(-1, 1, 1, 1)
            out[conv + ".bias"] = out[conv + ".bias"] * sd_
            out[bn + ".running_mean"] = out[bn + ".running_mean"] * sd_
    return out


def sign_like_images(n: int, h: int = 224, w: int = 224, seed: int = 0, classes: int = 43,
                     index0: int = 0) -> "tuple[torch.Tensor, torch.Tensor]":
    """Low-frequency, class-dependent 'sign-like' u8 NHWC images + labels (label = global index mod classes).

    Flat-spectrum noise is unrepresentative for classifier margins, so each class is a
    fixed 8x8 colour pattern, bicubically upsampled, plus a per-image low-frequency perturbation.
    """
    import torch.nn.functional as F
    gc = torch.Generator(device="cpu").manual_seed(12345)
    protos = torch.rand((classes, 3, 8, 8), generator=gc)
    idx = torch.arange(index0, index0 + n)
    labels = idx % classes
    g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + index0)
    pert = torch.rand((n, 3, 8, 8), generator=g) - 0.5
    low = protos[labels] + 0.35 * pert
    img = F.interpolate(low, size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)
    u8 = (img * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    return u8, labels.to(torch.int64)


def indexed_images(index0: int, n: int, h: int = 224, w: int = 224, seed: int = 0,
                   classes: int = 43) -> "tuple[torch.Tensor, torch.Tensor]":
    """Images [index0, index0 + n) of an unbounded synthetic dataset in which image i is a PURE function of (seed, i):
    the per-image perturbation comes from a counter-based hash (splitmix64) of the global index, not from a sequential
    generator, so any rank / chunking / world size produces the same bytes for the same global index (SURVEY.md
    section 8e: results independent of the world size).  Same look as `sign_like_images`; label = i mod classes."""
    import numpy as np
    import torch.nn.functional as F
    gc = torch.Generator(device="cpu").manual_seed(12345)
    protos = torch.rand((classes, 3, 8, 8), generator=gc)
    idx = np.arange(index0, index0 + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx[:, None] * np.uint64(192) + np.arange(192, dtype=np.uint64)[None, :]
        z = z + np.uint64(seed + 1) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    pert = torch.from_numpy(((z >> np.uint64(11)).astype(np.float64) / float(1 << 53) - 0.5).astype(np.float32))
    labels = torch.from_numpy((idx % np.uint64(classes)).astype(np.int64))
    low = protos[labels] + 0.35 * pert.view(n, 3, 8, 8)
    out = []
    for s in range(0, n, 64):      # fixed inner chunk: the interpolation never sees a batch-size-dependent code path
        img = F.interpolate(low[s:s + 64], size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)
        out.append((img * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous())
    return torch.cat(out) if len(out) > 1 else out[0], labels
