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
p = (D.blur_params(n, 10, 45) if len(sys.argv) > 1 and sys.argv[1] == "blur" else D.compound_params(n)).to(dev)
for _ in range(3):
    D.degrade(img, p, seed=1, out=out)
torch.cuda.synchronize()
print("ok")
