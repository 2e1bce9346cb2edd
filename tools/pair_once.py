#!/usr/bin/env python
"""Launch the generic conv on one VGG conv4_2-shaped layer (512 -> 512 @28, batch 128) a few times: ncu target for the
cta_group::2 pair kernel (-k regex:conv_gemm_pair) or, with flags=8, the single-CTA kernel."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from b200restore import ops, packing, _lib as L

dev = torch.device("cuda", 0)
n, hw, ci, co = 128, 28, 512, 512
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
x = torch.randn((n, hw, hw, ci), device=dev).mul_(0.5).to(torch.bfloat16)
w = torch.randn((co, ci, 3, 3)) * (2.0 / (9 * ci)) ** 0.5
wm, kbl = packing.plan_conv3x3(w).finish(dev)
out = torch.empty((n, hw, hw, co), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.conv_gemm([x], wm, torch.zeros(co, device=dev), kbl, act=L.B2R_ACT_RELU, out=out, flags=flags)
torch.cuda.synchronize()
print("ok")
