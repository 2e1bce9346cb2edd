#!/usr/bin/env python
"""(correct, total) and a digest of the predictions for the global image range [0, T), computed by WORLD_SIZE ranks.

    python tools/world_invariance.py --total 1001 --hw 64 --out counts.json                      (1 rank)
    python -m torch.distributed.run --nproc-per-node 2 ... tools/world_invariance.py ...          (N ranks, NCCL)

Rank r owns the contiguous block shard_range(T, r, R); images and Philox noise are keyed by the GLOBAL image index; the
only collective is the all-reduce of the int64 (correct, total) pair (+ an all-gather of the predictions, test only).
tests/test_world_invariance_gpu.py asserts that the output does not depend on the world size (18:48-51 semantics)."""
import argparse
import hashlib
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1001)
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--micro-batch", type=int, default=96)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import b200restore as B
    from b200restore import degrade as D, models, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    r, j = models.ResUNet(), models.VGG16Judge()
    r.load_state_dict(synth.synthetic_state_dict("resunet", 31))
    j.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
    pipe = B.RestoreClassifyPipeline(r.to(dev), j.to(dev), micro_batch=a.micro_batch)
    lo, hi = B.shard_range(a.total, rank, world)
    imgs, labels = synth.indexed_images(lo, hi - lo, a.hw, a.hw, seed=7)
    pred, counts = pipe.run(imgs.to(dev), labels.to(dev), D.compound_params(hi - lo), seed=2, image_index0=lo)
    B.all_reduce_counts(counts)
    if world > 1:
        sizes = [B.shard_range(a.total, q, world) for q in range(world)]
        mx = max(h - l for l, h in sizes)
        pad = torch.full((mx,), -1, dtype=torch.int64, device=dev)
        pad[:hi - lo] = pred
        allp = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(allp, pad)
        pred = torch.cat([p[:h - l] for p, (l, h) in zip(allp, sizes)])
    if rank == 0:
        digest = hashlib.sha256(pred.cpu().numpy().tobytes()).hexdigest()
        Path(a.out).write_text(json.dumps({"world": world, "counts": counts.tolist(), "pred_sha256": digest, "n": int(pred.numel())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
