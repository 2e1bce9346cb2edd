#!/usr/bin/env python
"""Per-role timeline of CTA 0 of the tap-folded conv kernel (debug_timeline field of b2r_conv_gemm_desc).

Needs a library built with the stamps compiled in (they are not in the product build):
    python -m b200restore.build --define B2R_TIMELINE --out tools/exp/libb2r_timeline.so
    B2R_LIB=tools/exp/libb2r_timeline.so python tools/role_timeline.py
Prints, for tiles in steady state, the MMA warp's and the epilogue's stamps: tells which role paces the tile."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    from b200restore import ops, packing, _lib as L
    dev = torch.device("cuda", 0)
    n, hw = 128, 224
    for ci_split, pooled in (((64,), False), ((64, 64), False), ((64,), True), ((64,), "head")):
        ci = sum(ci_split)
        srcs = [torch.randn((n, hw, hw, c), device=dev).mul_(0.5).to(torch.bfloat16) for c in ci_split]
        w = torch.randn((64, ci, 3, 3)) * (2.0 / (9 * ci)) ** 0.5
        plan = packing.plan_conv3x3(w, ci_split)
        head = {}
        if pooled == "head":   # ResUNet dec1 second conv: + 1x1 shortcut over two 64-channel sources + fused 64 -> 3 head
            for j in range(2):
                srcs.append(srcs[0].clone())
                plan.add_1x1(len(srcs) - 1, torch.randn((64, 64)) * 0.1)
            head = dict(head_w=(torch.randn((3, 64)) * 0.1).to(dev), head_b=torch.zeros(3, device=dev),
                        head_out_u8=torch.empty((n, hw, hw, 3), dtype=torch.uint8, device=dev))
            pooled = False
        wm, kbl = plan.finish(dev)
        w3 = plan.finish_w3(dev)
        out = torch.empty((n, hw, hw, 64), dtype=torch.bfloat16, device=dev)
        pool = torch.empty((n, hw // 2, hw // 2, 64), dtype=torch.bfloat16, device=dev) if pooled else None
        dbg = torch.zeros((64, 8), dtype=torch.int64, device=dev)
        for _ in range(2):
            ops.conv_gemm(srcs, wm, torch.zeros(64, device=dev), kbl, act=L.B2R_ACT_RELU, out=None if head else out,
                          out_pool=pool, weights_w3=w3, debug_timeline=dbg, **head)
        torch.cuda.synchronize()
        t = dbg.cpu()
        print(f"--- C_in {ci} pooled={pooled} head={bool(head)}: tiles 20..27 of CTA 0, cycles relative to the MMA warp's start of tile 20")
        print(" stamps: [1 mma: tmem stage free, 7 mma: first A box ready, 2 mma: all MMAs issued + committed, "
              "3 epi: accumulator ready, 4 epi: TMEM drained + released, 5 epi: staged, 6 epi: stores issued]")
        base = int(t[20, 1])
        for i in range(20, 28):
            r = [int(t[i, k]) - base for k in (1, 7, 2, 3, 4, 5, 6)]
            print(f"   tile {i}: " + " ".join(f"{v:7d}" for v in r) + f"   | period {int(t[i + 1, 1] - t[i, 1]):5d}"
                  f" | mma: wait A {r[1] - r[0]:4d}, issue {r[2] - r[1]:5d}, to next tile {int(t[i + 1, 1]) - base - r[2]:4d}"
                  f" | epi: acc ready {r[3] - r[2]:4d} after last issue, drain {r[4] - r[3]:4d}, math+stage {r[5] - r[4]:4d}, store {r[6] - r[5]:4d}")

if __name__ == "__main__":
    main()
