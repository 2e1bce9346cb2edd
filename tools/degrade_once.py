"""ncu target for the degradation kernels: python tools/degrade_once.py [compound | blur | fog]  (fog = the point-wise kernel)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from b200restore import degrade as D
n, hw = 256, 224
dev = torch.device("cuda", 0)
img = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device=dev)
out = torch.empty_like(img)
import numpy as np
mode = sys.argv[1] if len(sys.argv) > 1 else "compound"
p = {"blur": lambda: D.blur_params(n, 10, 45), "fog": lambda: D.fog_params(n, np.random.default_rng(1)),
     "compound": lambda: D.compound_params(n)}[mode]().to(dev)
for _ in range(3):
    D.degrade(img, p, seed=1, out=out)
torch.cuda.synchronize()
print("ok")
