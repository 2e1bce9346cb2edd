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
