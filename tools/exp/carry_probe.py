"""Probe for the tap-folded kernel's carry mode: python tools/exp/carry_probe.py FLAGS N H W [pool]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from b200restore import ops, packing, _lib as L
flags, n, h, w = (int(v) for v in sys.argv[1:5])
pool = len(sys.argv) > 5
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
x = (torch.randn((n, h, w, 64), generator=g) * 0.5).to(torch.bfloat16).to(dev)
wt = torch.randn((64, 64, 3, 3), generator=g) * (2.0 / 576) ** 0.5
plan = packing.plan_conv3x3(wt)
wm, kbl = plan.finish(dev)
w3 = plan.finish_w3(dev)
b = torch.zeros(64, device=dev)
outs = []
for f in (flags, flags | L.B2R_CONV_NO_CARRY):
    out = torch.full((n, h, w, 64), float("nan"), dtype=torch.bfloat16, device=dev)
    pl = torch.full((n, h // 2, w // 2, 64), float("nan"), dtype=torch.bfloat16, device=dev) if pool else None
    ops.conv_gemm([x], wm, b, kbl, act=L.B2R_ACT_RELU, out=out, out_pool=pl, weights_w3=w3, flags=f)
    print("launched", L.load().b2r_last_conv_kernel().decode(), flush=True)
    torch.cuda.synchronize()
    print("  ok, nan count", int(torch.isnan(out).sum()), flush=True)
    outs.append((out, pl))
print("equal:", torch.equal(outs[0][0], outs[1][0]), "" if not pool else torch.equal(outs[0][1], outs[1][1]))
if not torch.equal(outs[0][0], outs[1][0]):
    d = (outs[0][0].float() - outs[1][0].float()).abs()
    bad = (d > 0) | torch.isnan(d)
    idx = bad.nonzero()
    print("mismatches", int(bad.sum()), "first", idx[:5].tolist(), "cols", sorted(set(idx[:, 2].tolist()))[:40])
