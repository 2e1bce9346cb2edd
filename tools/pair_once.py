#!/usr/bin/env python
"""Launch the generic conv family on one layer shape a few times (ncu target):

    python tools/pair_once.py [flags] [ci] [co] [hw]      # default 512 -> 512 @28 = cta_group::2 pair kernel
    python tools/pair_once.py 0 128 128 112               # C_out = 128: halo-mode kernel (-k regex:conv_gemm_halo)
    python tools/pair_once.py 8                           # flags = 8 (B2R_CONV_NO_PAIR): single-CTA kernel
"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from b200restore import ops, packing, _lib as L

dev = torch.device("cuda", 0)
a = [int(v) for v in sys.argv[1:]]
flags = a[0] if len(a) > 0 else 0
ci = a[1] if len(a) > 1 else 512
co = a[2] if len(a) > 2 else 512
hw = a[3] if len(a) > 3 else 28
n = 128
x = torch.randn((n, hw, hw, ci), device=dev).mul_(0.5).to(torch.bfloat16)
w = torch.randn((co, ci, 3, 3)) * (2.0 / (9 * ci)) ** 0.5
wm, kbl = packing.plan_conv3x3(w).finish(dev)
out = torch.empty((n, hw, hw, co), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.conv_gemm([x], wm, torch.zeros(co, device=dev), kbl, act=L.B2R_ACT_RELU, out=out, flags=flags)
torch.cuda.synchronize()
print("ok")
