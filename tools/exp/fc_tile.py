#!/usr/bin/env python
"""Experiment: VGG16 classifier.0 (25088 -> 4096) and classifier.3 (4096 -> 4096) on M = 256 rows with different n-tiles.
M = 256 is two pixel tiles, so the 128 x 256 pair kernel fills only 16 clusters of 74: smaller n-tiles trade MMA rate for
parallelism.    python tools/exp/fc_tile.py [--m 256]"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from b200restore import _lib as L, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=256)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for K, N in ((25088, 4096), (4096, 4096)):
    x = torch.randn((1, 1, args.m, K), device=dev).to(torch.bfloat16)
    w = (torch.randn((N, K), device=dev) * 0.01).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    ref = None
    for bn, flags in ((0, 0), (256, L.B2R_CONV_NO_PAIR), (128, 0), (64, 0)):
        out = torch.empty((1, 1, args.m, N), dtype=torch.bfloat16, device=dev)
        try:
            ops.conv_gemm([x], w, b, None, act=L.B2R_ACT_RELU, out=out, block_n=bn, flags=flags)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"K": K, "N": N, "block_n": bn, "error": str(e)[:120]}))
            continue
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_gemm([x], w, b, None, act=L.B2R_ACT_RELU, out=out, block_n=bn, flags=flags)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        same = None if ref is None else bool(torch.equal(out, ref))
        if ref is None:
            ref = out.clone()
        print(json.dumps({"K": K, "N": N, "m": args.m, "block_n": bn, "flags": flags, "us_median": round(1e3 * ms[len(ms) // 2], 1),
                          "tflops": round(2 * args.m * K * N / ms[len(ms) // 2] / 1e9, 1), "bit_identical_to_default": same}))
