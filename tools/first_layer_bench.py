#!/usr/bin/env python
"""Time the C_in = 3 first layer (b2r_conv3x3_c3) at 224x224, batch 128: u8 NHWC and f32 NCHW entries."""
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
f32 = torch.rand((n, 3, hw, hw), device=dev)
out = torch.empty((n, hw, hw, 64), dtype=torch.bfloat16, device=dev)
for name, x, norm in (("u8 NHWC + normalize", u8, True), ("u8 NHWC", u8, False), ("f32 NCHW", f32, False)):
    for _ in range(3):
        ops.conv3x3_c3(x, w, b, act=L.B2R_ACT_RELU, normalize=norm, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv3x3_c3(x, w, b, act=L.B2R_ACT_RELU, normalize=norm, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    byts = x.numel() * x.element_size() + out.numel() * 2
    print(f"{name:22s} {ms * 1e3:8.1f} us per {n} images = {ms * 1e3 / n:5.2f} us/img; {byts / ms / 1e6:7.1f} GB/s of HBM traffic")
