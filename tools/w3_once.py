#!/usr/bin/env python
"""Launch the tap-folded conv kernel a few times on one 64 -> 64 @224 layer (batch 128): target for `ncu -k regex:conv_w3`."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from b200restore import ops, packing, _lib as L

dev = torch.device("cuda", 0)
n, hw = 128, 224
ci_split = tuple(int(a) for a in sys.argv[1].split(",")) if len(sys.argv) > 1 else (64,)
ci = sum(ci_split)
srcs = [torch.randn((n, hw, hw, c), device=dev).mul_(0.5).to(torch.bfloat16) for c in ci_split]
w = torch.randn((64, ci, 3, 3)) * (2.0 / (9 * ci)) ** 0.5
plan = packing.plan_conv3x3(w, ci_split)
wm, kbl = plan.finish(dev)
w3 = plan.finish_w3(dev)
out = torch.empty((n, hw, hw, 64), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.conv_gemm(srcs, wm, torch.zeros(64, device=dev), kbl, act=L.B2R_ACT_RELU, out=out, weights_w3=w3)
torch.cuda.synchronize()
print("ok")
