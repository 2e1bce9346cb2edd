#!/usr/bin/env python
"""Per-launch list of one micro-batch through ResUNet -> VGG16: kernel, algorithmic GFLOP, ms, TFLOP/s, in launch order.

    python tools/launch_list.py [--n 512] [--hw 224] [--arch resunet]

Eager launches with one CUDA-event pair each (ops.KernelTimer); the second pass is the one printed."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from b200restore import models, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--arch", default="resunet")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    r = (models.ResUNet if args.arch == "resunet" else models.SimpleUNet)()
    r.load_state_dict(synth.synthetic_state_dict(args.arch, 31))
    r = r.to(dev).eval()
    j = models.VGG16Judge()
    j.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
    j = j.to(dev).eval()
    x, _ = synth.indexed_images(0, args.n, args.hw, args.hw, seed=5)
    x = x.to(dev)
    with torch.no_grad():
        for rep in range(3):
            t = ops.KernelTimer(kinds=("conv_gemm", "conv3x3_c3", "linear", "elementwise"))
            with ops.timing(t):
                j.forward_u8(r.restore_u8(x))
            torch.cuda.synchronize()
    tot = 0.0
    for kind, work, e0, e1, sub in t.records:
        ms = e0.elapsed_time(e1)
        tot += ms
        print(json.dumps({"kind": kind, "kernel": sub, "gflop": round(work / 1e9, 2), "ms": round(ms, 4),
                          "tflops": round(work / ms / 1e9, 1) if ms > 0 else None}))
    print(json.dumps({"total_ms": tot, "n": args.n}))


if __name__ == "__main__":
    main()
