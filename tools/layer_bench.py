#!/usr/bin/env python
"""Per-layer microbenchmark of the tcgen05 conv kernel on the shapes the three networks use at 224x224.

    python tools/layer_bench.py [--batch 128] [--iters 20] [--only NAME] [--json out.json]

Each line: algorithmic TFLOP/s (CUDA events on the launching stream, L2 flushed between iterations by the 1.6 GB of
activations the other layers touch — every layer's input is far larger than L2 at batch 128 anyway) and the fraction
of the measured bf16 peaks in MEASURED_PEAKS.json.  Used to steer kernel work and to pick ncu targets; numbers
quoted in DESIGN.md come from here and from bench.py, never from a run under a profiler.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

# name, H(=W), sources (channels), C_out, taps ('3x3' | 'convT' | '3x3+sc'), pooled
LAYERS = [
    ("vgg.conv1_2      64->64   @224 +pool", 224, (64,), 64, "3x3", True),
    ("resunet.res1.c1  64->64   @224", 224, (64,), 64, "3x3", False),
    ("resunet.res1.c2  64->64+I @224 +pool", 224, (64,), 64, "3x3+sc", True),
    ("resunet.dec1.c1  128->64  @224 (cat)", 224, (64, 64), 64, "3x3", False),
    ("resunet.dec1.c2  64->64+sc128+head @224", 224, (64,), 64, "3x3+sc128+head", False),
    ("resunet.up1      64->64   convT@112", 112, (64,), 64, "convT", False),
    ("resunet.dec2.c1  192->64  @112 (cat)", 112, (64, 128), 64, "3x3", False),
    ("vgg.conv2_1      64->128  @112", 112, (64,), 128, "3x3", False),
    ("vgg.conv2_2      128->128 @112 +pool", 112, (128,), 128, "3x3", True),
    ("resunet.dec3.c1  384->128 @56 (cat)", 56, (128, 256), 128, "3x3", False),
    ("vgg.conv3_1      128->256 @56", 56, (128,), 256, "3x3", False),
    ("vgg.conv3_2      256->256 @56", 56, (256,), 256, "3x3", False),
    ("vgg.conv4_1      256->512 @28", 28, (256,), 512, "3x3", False),
    ("vgg.conv4_2      512->512 @28", 28, (512,), 512, "3x3", False),
    ("vgg.conv5_1      512->512 @14", 14, (512,), 512, "3x3", False),
    ("resunet.up3      256->128 convT@28", 28, (256,), 128, "convT", False),
    ("resunet.up2      128->64  convT@56", 56, (128,), 64, "convT", False),
]


class _NvmlSampler:
    """Median SM clock and board power while the timed loop runs (use --iters >= 200 so there is something to sample)."""

    def __init__(self):
        import threading
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.mhz, self.watts, self.go = [], [], True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        import time
        while self.go:
            self.mhz.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.watts.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            time.sleep(0.003)

    def stop(self):
        self.go = False
        self.t.join()
        k = len(self.mhz) // 2   # second half: the governor has settled
        mhz, w = sorted(self.mhz[k:]), sorted(self.watts[k:])
        return f"  | SM {mhz[len(mhz) // 2]} MHz, {w[len(w) // 2]:.0f} W ({len(self.mhz)} samples)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default=None)
    ap.add_argument("--json", default=None)
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--no-w3", action="store_true")
    ap.add_argument("--tile", default=None, help="force the pixel tile, e.g. 8,16,1 (W,H,N)")
    ap.add_argument("--flags", type=int, default=0, help="b2r_conv_gemm flags (4 = B2R_CONV_NO_HALO)")
    ap.add_argument("--clocks", action="store_true", help="sample SM clock / board power (NVML) during the timed loop")
    args = ap.parse_args()
    from b200restore import ops, packing, _lib as L
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    burst, sust = peaks.get("bf16_tflops", 1592.0), peaks.get("bf16_tflops_sustained", 1362.2)
    dev = torch.device("cuda", 0)
    n = args.batch
    g = torch.Generator(device="cpu").manual_seed(0)
    results = []
    for name, hw, srcs_c, co, kind, pooled in LAYERS:
        if args.only and args.only not in name:
            continue
        srcs = [(torch.randn((n, hw, hw, c), generator=g, dtype=torch.float32) * 0.5).to(torch.bfloat16).to(dev)
                if n * hw * hw * c < 2 ** 28 else
                torch.randn((n, hw, hw, c), device=dev, dtype=torch.float32).mul_(0.5).to(torch.bfloat16)
                for c in srcs_c]
        ci = sum(srcs_c)
        if kind == "convT":
            w = torch.randn((ci, co, 2, 2), generator=g) * (1.0 / ci) ** 0.5
            wm, bias = packing.pack_convT2x2(w, torch.zeros(co))
            kbl, mode, w3, head = None, L.B2R_OUT_CONVT2X2, None, {}
            out = torch.empty((n, 2 * hw, 2 * hw, co), dtype=torch.bfloat16, device=dev)
            alg_k = ci
        else:
            w = torch.randn((co, ci, 3, 3), generator=g) * (2.0 / (9 * ci)) ** 0.5
            plan = packing.KPlan(co)
            off = 0
            for s, c in enumerate(srcs_c):
                plan.add_conv3x3(s, w[:, off:off + c])
                off += c
            alg_k = 9 * ci
            if kind == "3x3+sc":
                srcs = srcs + [srcs[0].clone()]
                plan.add_1x1(len(srcs) - 1, torch.eye(co))
            head = {}
            if kind == "3x3+sc128+head":   # ResUNet dec1 second conv: 1x1 shortcut over cat(up1, skip) + fused 64 -> 3 head
                wsc = torch.randn((co, 128), generator=g) * (1.0 / 128) ** 0.5
                for j in range(2):
                    srcs = srcs + [srcs[0].clone()]
                    plan.add_1x1(len(srcs) - 1, wsc[:, 64 * j:64 * j + 64])
                alg_k += 128
                head = dict(head_w=(torch.randn((3, 64), generator=g) * 0.1).to(dev), head_b=torch.zeros(3, device=dev),
                            head_out_u8=torch.empty((n, hw, hw, 3), dtype=torch.uint8, device=dev))
            wm, kbl = plan.finish()
            w3 = plan.finish_w3()
            bias, mode = torch.zeros(co), L.B2R_OUT_NHWC
            out = torch.empty((n, hw, hw, co), dtype=torch.bfloat16, device=dev)
        pool = torch.empty((n, hw // 2, hw // 2, co), dtype=torch.bfloat16, device=dev) if pooled else None
        wm, bias = wm.to(dev), bias.to(dev)
        w3 = w3.to(dev) if (w3 is not None and not args.no_w3) else None
        flops = 2.0 * n * hw * hw * wm.shape[0] * alg_k

        def run():
            ops.conv_gemm(srcs, wm, bias, kbl, act=L.B2R_ACT_NONE if kind == "convT" else L.B2R_ACT_RELU, out=None if head else out, out_pool=pool, out_mode=mode,
                          weights_w3=w3, **head,
                          block_n=args.block_n if args.block_n and co % args.block_n == 0 else 0,
                          tile=tuple(int(v) for v in args.tile.split(",")) if args.tile else (0, 0, 0), flags=args.flags)

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = _NvmlSampler() if args.clocks else None
        e0.record()
        for _ in range(args.iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        clk = sampler.stop() if sampler else ""
        ms = e0.elapsed_time(e1) / args.iters
        tf = flops / (ms * 1e-3) / 1e12
        print(f"{name:42s} {ms * 1e3:9.1f} us  {tf:7.1f} TFLOP/s  {100 * tf / burst:5.1f}% of burst  "
              f"{100 * tf / sust:5.1f}% of sustained{clk}", flush=True)
        results.append({"layer": name, "us": ms * 1e3, "tflops": tf, "frac_burst": tf / burst, "frac_sustained": tf / sust})
        del srcs, out, pool
        torch.cuda.empty_cache()
    if args.json:
        Path(args.json).write_text(json.dumps({"batch": n, "iters": args.iters, "layers": results}, indent=1))


if __name__ == "__main__":
    main()
