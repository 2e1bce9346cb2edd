#!/usr/bin/env python
"""Experiment: does running the full-resolution stages of ResUNet / VGG16 in SMALL sub-batches keep the producer ->
consumer tensors (e1, y1, u1, yd1, c0: 6.4 MB per image each, written once and read once or twice by the next
launches) inside the 126 MB L2, so that they never travel to HBM?

    python tools/exp/l2_subbatch.py [--n 256] [--subs 2,4,8,16,32] [--iters 10] [--stage enc|dec|vgg|all]

The small buffers are REUSED by every sub-batch: a dirty line that is overwritten while still in L2 is never written
back.  Every variant is captured into a CUDA graph (no host launch cost in the numbers) and timed with CUDA events.
Variants: mono (one launch per layer over all n images), sub=s (one stream), sub=s x2 (two streams alternating
sub-batches, each with its own small buffers, so the tail of one sub-batch's launch overlaps the head of the next).
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import b200restore as B  # noqa: E402
from b200restore import _lib as L, models, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--subs", default="2,4,8,16,32")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--stage", default="all")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, H = args.n, args.hw
    W = H
    r = models.ResUNet()
    r.load_state_dict(synth.synthetic_state_dict("resunet", 31))
    r = r.to(dev).eval()
    j = models.VGG16Judge()
    j.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
    j = j.to(dev).eval()
    P, PJ = r._packed(), j._packed()
    imgs, _ = synth.sign_like_images(min(n, 64), H, W, seed=7)
    x = imgs.repeat((n + imgs.shape[0] - 1) // imgs.shape[0], 1, 1, 1)[:n].contiguous().to(dev)
    bf = lambda *s: torch.empty(s, dtype=torch.bfloat16, device=dev)  # noqa: E731
    R, PR = L.B2R_ACT_RELU, L.B2R_ACT_PRELU

    # persistent (HBM) tensors of the stages
    r1, p1 = bf(n, H, W, 64), bf(n, H // 2, W // 2, 64)
    d2 = torch.randn((n, H // 2, W // 2, 64), device=dev).to(torch.bfloat16)
    out_u8 = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    vp = bf(n, H // 2, W // 2, 64)

    def enc_stage(xs, e1, y1, r1s, p1s):
        w0, b0, s0 = P["enc1"]
        ops.conv3x3_c3(xs, w0, b0, act=PR, slope=s0, out=e1)
        c1, slope, c2 = P["res1"]
        ops.conv_gemm([e1], **c1, act=PR, slope=slope, out=y1)
        ops.conv_gemm([y1, e1], **c2, act=R, out=r1s, out_pool=p1s)

    def dec_stage(d2s, r1s, u1, y, o8):
        ops.conv_gemm([d2s], *P["up1"], None, out=u1, out_mode=L.B2R_OUT_CONVT2X2)
        c1, slope, c2 = P["dec1"]
        ops.conv_gemm([u1, r1s], **c1, act=PR, slope=slope, out=y)
        ops.conv_gemm([y, u1, r1s], **c2, act=R, head_w=P["final"][0], head_b=P["final"][1], head_out_u8=o8)

    def vgg_stage(u8s, c0, ps):
        ops.conv3x3_c3(u8s, *PJ["first"], act=R, normalize=True, out=c0)
        cv, pooled, co = PJ["convs"][0]
        ops.conv_gemm([c0], **cv, act=R, out_pool=ps)

    stages = {
        "enc": dict(small=lambda s: (bf(s, H, W, 64), bf(s, H, W, 64)),
                    run=lambda lo, hi, sm: enc_stage(x[lo:hi], sm[0][:hi - lo], sm[1][:hi - lo], r1[lo:hi], p1[lo:hi]),
                    gflop=2 * 3.70 + 0.17, mb_saved="e1 w+2r, y1 w+r = 32 MB/img"),
        "dec": dict(small=lambda s: (bf(s, H, W, 64), bf(s, H, W, 64)),
                    run=lambda lo, hi, sm: dec_stage(d2[lo:hi], r1[lo:hi], sm[0][:hi - lo], sm[1][:hi - lo], out_u8[lo:hi]),
                    gflop=0.41 + 7.4 + 3.7 + 0.82, mb_saved="u1 w+2r, yd1 w+r = 32 MB/img"),
        "vgg": dict(small=lambda s: (bf(s, H, W, 64),),
                    run=lambda lo, hi, sm: vgg_stage(out_u8[lo:hi], sm[0][:hi - lo], vp[lo:hi]),
                    gflop=3.70 + 0.17, mb_saved="c0 w+r = 12.8 MB/img"),
    }
    names = list(stages) if args.stage == "all" else [args.stage]
    # make r1 / out_u8 valid inputs for the later stages
    big = stages["enc"]["small"](n)
    stages["enc"]["run"](0, n, big)
    bigd = stages["dec"]["small"](n)
    stages["dec"]["run"](0, n, bigd)
    torch.cuda.synchronize()
    del big, bigd

    def time_graph(fn, iters):
        fn()                                  # eager once (func attributes, lazy module loads)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    results = []
    for name in names:
        st = stages[name]
        sm = st["small"](n)
        ms = time_graph(lambda: st["run"](0, n, sm), args.iters)
        del sm
        results.append(dict(stage=name, variant="mono", ms=ms, us_per_img=ms * 1e3 / n, tflops=st["gflop"] * n / ms))
        print(json.dumps(results[-1]), flush=True)
        for s in [int(v) for v in args.subs.split(",")]:
            if s >= n:
                continue
            for nstreams in (1, 2):
                smalls = [st["small"](s) for _ in range(nstreams)]
                side = [torch.cuda.Stream(device=dev) for _ in range(nstreams)] if nstreams > 1 else None

                def run():
                    if side is None:
                        for lo in range(0, n, s):
                            st["run"](lo, min(lo + s, n), smalls[0])
                        return
                    main_s = torch.cuda.current_stream()
                    ev = torch.cuda.Event()
                    ev.record(main_s)
                    for k, q in enumerate(side):
                        q.wait_event(ev)
                        with torch.cuda.stream(q):
                            for lo in range(k * s, n, s * nstreams):
                                st["run"](lo, min(lo + s, n), smalls[k])
                        done = torch.cuda.Event()
                        done.record(q)
                        main_s.wait_event(done)

                ms = time_graph(run, args.iters)
                results.append(dict(stage=name, variant=f"sub={s} x{nstreams}", ms=ms, us_per_img=ms * 1e3 / n,
                                    tflops=st["gflop"] * n / ms))
                print(json.dumps(results[-1]), flush=True)
                del smalls
    if args.json:
        Path(args.json).write_text("\n".join(json.dumps(r_) for r_ in results) + "\n")


if __name__ == "__main__":
    main()
