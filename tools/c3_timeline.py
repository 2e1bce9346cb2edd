#!/usr/bin/env python
"""Role timeline of the first-layer kernel (conv_c3.cu), CTA 0, via b2r_debug_timeline.

Needs the stamp-enabled variant build:
    python -m b200restore.build --define B2R_TIMELINE --out tools/exp/libb2r_timeline.so
    B2R_LIB=tools/exp/libb2r_timeline.so python tools/c3_timeline.py"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from b200restore import ops, packing, _lib as L

n, hw = 128, 224
dev = torch.device("cuda", 0)
w = packing.pack_conv_c3(torch.randn(64, 3, 3, 3) * 0.2).to(dev)
b = torch.zeros(64, device=dev)
u8 = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device=dev)
out = torch.empty((n, hw, hw, 64), dtype=torch.bfloat16, device=dev)
dbg = torch.zeros((64, 8), dtype=torch.int64, device=dev)
lib = L.load()
for _ in range(2):
    ops.conv3x3_c3(u8, w, b, act=L.B2R_ACT_RELU, out=out)
lib.b2r_debug_timeline(dbg.data_ptr())
ops.conv3x3_c3(u8, w, b, act=L.B2R_ACT_RELU, out=out)
torch.cuda.synchronize()
lib.b2r_debug_timeline(None)
t = dbg.cpu()
base = int(t[20, 3])
print("stamps rel. to MMA issue of tile 20: [prod row ready, prod slot free, prod arrived, mma issue, epi acc ready, epi ld done, epi staging free, epi store issued]")
for i in range(20, 30):
    print(f" tile {i}: " + " ".join(f"{int(t[i, k]) - base:7d}" for k in range(8)))
