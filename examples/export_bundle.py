#!/usr/bin/env python
"""Write the flat binary bundle examples/cabi_pipeline.c reads: two state_dicts (restorer + VGG16 judge) under the reference's
key names, the per-image degradation parameters of b2r_degrade, and a batch of u8 images + labels.

    python examples/export_bundle.py out.bin [--arch resunet] [--restorer restoration_unified_resnet.pth] [--judge vgg16_baseline.pth]
                                             [--n 16] [--hw 64]

With --restorer / --judge the tensors come from `torch.load(path, map_location='cpu')` (the shipped checkpoints, when
available); without them from the seeded synthetic state_dicts the tests use.  Format (little endian):
  "B2RBNDL1" | i32 arch, n_restorer, n_judge, N, H, W | tensors... | ksize i32[N] | taps f32[N*225] | fog_on i32[N] |
  fog_t f32[N] | fog_add f32[N] | sigma f32[N] | images u8[N*H*W*3] | labels i64[N]
  tensor = i32 name_len | name | i32 dtype (0 f32, 1 i64) | i32 ndim | i64 shape[4] | data
"""
import argparse
import struct
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def write_tensors(f, sd):
    for k, v in sd.items():
        t = v.detach().cpu().contiguous()
        if t.dtype not in (torch.float32, torch.int64):
            t = t.float()
        name = k.encode()
        shape = list(t.shape) + [0] * (4 - t.dim())
        f.write(struct.pack("<i", len(name)) + name + struct.pack("<ii4q", 0 if t.dtype == torch.float32 else 1, t.dim(), *shape))
        f.write(t.numpy().tobytes())


def main():
    from b200restore import degrade as D, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--arch", default="resunet", choices=["resunet", "simple_unet"])
    ap.add_argument("--restorer", default=None)
    ap.add_argument("--judge", default=None)
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--hw", type=int, default=64)
    a = ap.parse_args()
    sdr = torch.load(a.restorer, map_location="cpu") if a.restorer else synth.synthetic_state_dict(a.arch, 31)
    sdj = torch.load(a.judge, map_location="cpu") if a.judge else synth.synthetic_state_dict("vgg16", 32)
    imgs, labels = synth.indexed_images(0, a.n, a.hw, a.hw, seed=7)
    p = D.compound_params(a.n)                       # script 16: Blur(10, 45 deg) -> Fog(0.5, A = 0.9) -> Noise(var 0.02)
    with open(a.out, "wb") as f:
        f.write(b"B2RBNDL1")
        f.write(struct.pack("<6i", 1 if a.arch == "resunet" else 0, len(sdr), len(sdj), a.n, a.hw, a.hw))
        write_tensors(f, sdr)
        write_tensors(f, sdj)
        for arr, dt in ((p.ksize, np.int32), (p.taps, np.float32), (p.fog_on, np.int32), (p.fog_t, np.float32),
                        (p.fog_add, np.float32), (p.sigma, np.float32)):
            f.write(np.ascontiguousarray(arr, dtype=dt).tobytes())
        f.write(imgs.numpy().tobytes())
        f.write(labels.numpy().astype(np.int64).tobytes())
    print(f"wrote {a.out}: {len(sdr)} + {len(sdj)} tensors, {a.n} images of {a.hw}x{a.hw}")


if __name__ == "__main__":
    main()
