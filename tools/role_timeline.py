#!/usr/bin/env python
"""Per-role timeline of CTA 0 of the tap-folded conv kernel (debug_timeline field of b2r_conv_gemm_desc).

Prints, for tiles in steady state, the cycles between the stamps: producer issue, MMA start / all MMAs issued,
epilogue start (accumulator ready) / staging free / TMEM drained / stores issued.  Tells which role paces the tile."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    from b200restore import ops, packing, _lib as L
    dev = torch.device("cuda", 0)
    n, hw = 128, 224
    for ci_split, pooled in (((64,), False), ((64, 64), False), ((64,), True)):
        ci = sum(ci_split)
        srcs = [torch.randn((n, hw, hw, c), device=dev).mul_(0.5).to(torch.bfloat16) for c in ci_split]
        w = torch.randn((64, ci, 3, 3)) * (2.0 / (9 * ci)) ** 0.5
        plan = packing.plan_conv3x3(w, ci_split)
        wm, kbl = plan.finish(dev)
        w3 = plan.finish_w3(dev)
        out = torch.empty((n, hw, hw, 64), dtype=torch.bfloat16, device=dev)
        pool = torch.empty((n, hw // 2, hw // 2, 64), dtype=torch.bfloat16, device=dev) if pooled else None
        dbg = torch.zeros((64, 8), dtype=torch.int64, device=dev)
        for _ in range(2):
            ops.conv_gemm(srcs, wm, torch.zeros(64, device=dev), kbl, act=L.B2R_ACT_RELU, out=out, out_pool=pool,
                          weights_w3=w3, debug_timeline=dbg)
        torch.cuda.synchronize()
        t = dbg.cpu()
        t0 = int(t[0, 1])
        print(f"--- C_in {ci} pooled={pooled}: per tile (cycles), tiles 20..27 of CTA 0")
        print(" tile | tile period | mma issue span | acc ready->epi start wait | epi: wait staging | drain TMEM+shift+stage | pool+store issue")
        base = int(t[20, 1])
        print(" raw stamps relative to MMA start of tile 20: [producer issue, mma start(tmem free), mma issued, epi start(acc ready), staging free, staged, stores issued, mma A-data ready]")
        for i in range(20, 26):
            print(f"   tile {i}: " + " ".join(f"{int(t[i, k]) - base:7d}" for k in range(8)))
        for i in range(20, 28):
            period = int(t[i + 1, 3] - t[i, 3])
            print(f" {i:4d} | {period:11d} | {int(t[i, 2] - t[i, 1]):14d} | "
                  f"{int(t[i, 3] - t[i - 1, 6]):25d} | {int(t[i, 4] - t[i, 3]):17d} | {int(t[i, 5] - t[i, 4]):22d} | "
                  f"{int(t[i, 6] - t[i, 5]):16d}")


if __name__ == "__main__":
    main()
