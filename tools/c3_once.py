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
for _ in range(3):
    ops.conv3x3_c3(u8, w, b, act=L.B2R_ACT_RELU, normalize=True, out=out)
torch.cuda.synchronize()
print("ok")
