#!/usr/bin/env python
"""bench.py — images/s of the degrade -> restore -> VGG16 classify -> top-1 count path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--batch B]

One "step" = one pass of the hot path over one batch of B synthetic GTSRB-shaped images per GPU (weak scaling: every
rank owns a contiguous block of the global image index range; the only collective is the all-reduce of the int64
(correct, total) pair at the end of the step).  For N > 1 launch under torchrun (one rank per GPU, NCCL).

Prints ONE JSON line on rank 0:
  value      images/s with the batch already resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e        same metric through the public API with HOST buffers (pinned H2D of every micro-batch + D2H of the counts
             inside the timed region)
  roofline   the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOPs / summed CUDA-event duration of its
             launches inside the timed region, against MEASURED_PEAKS.json; `by_kernel` splits the same measurement by
             the kernel each launch actually ran (b2r_last_conv_kernel), `dominant_kernel` names the largest share
  cpu_baseline  the oracle port of the reference path timed on this box's host cores (bounded sample, rank 0, N = 1)

`--impl reference` times the reference's CPU implementation of the same path (oracle port: the reference is Python and
cannot travel to the GPU box; oracle/ is pinned to it bit-for-bit by tests/test_oracle_*.py) with all host threads.
`--impl reference --ref-device cuda [--ref-precision f32|tf32|bf16]` is the opt-in same-box LIBRARY comparator of
BASELINE.md: the same fp32 PyTorch port on cuda:0 through cuDNN / cuBLAS (what the reference does with DEVICE = 'cuda').
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (restorer arch, degradation recipe, classify?)
    "degrade16_resunet_vgg16_top1": ("resunet", "compound16", True),      # BASELINE configs[3] per-GPU shard
    "degrade16_unet_vgg16_top1": ("simple_unet", "compound16", True),
    "degrade_fog_blur_unet": ("simple_unet", "compound16", False),        # BASELINE configs[1]
    "degrade_random_resunet": ("resunet", "random14", False),             # BASELINE configs[2]
    # 13_pipeline_stress_test.py on the device: Blur -> Fog -> Noise (u8 after every stage),
    # three SimpleUNets Noise -> Fog -> Blur with the unclamped f32 hand-off, VGG16 confidence of the result
    "stress13_cascade_vgg16_conf": ("cascade3", "stress13", True),
}
GFLOP_PER_IMAGE_224 = {"simple_unet": 38.831, "resunet": 55.992, "vgg16": 30.933,
                       "cascade3": 3 * 38.831}


CASCADE_SEEDS = {"Noise": 31, "Fog": 33, "Blur": 34}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="degrade16_resunet_vgg16_top1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=512)   # 256 -> 512: +1 % (fewer tail waves on the 14x14 / 28x28 layers)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--cpu-sample", type=int, default=0, help="images per CPU-baseline sample (0 = choose)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=0, choices=[0, 1],
                    help="e2e leg: replay restore -> classify -> count of a micro-batch as one CUDA graph (the device-resident "
                         "`value` leg stays eager: it times every conv launch with CUDA events)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: 'cuda' times the same fp32 PyTorch port on the GPU through cuDNN / cuBLAS "
                         "(the reference with DEVICE='cuda', 17:16) as the same-box library comparator")
    ap.add_argument("--ref-precision", default="f32", choices=["f32", "tf32", "bf16"],
                    help="--ref-device cuda: f32 (TF32 off), tf32, or bf16 autocast with channels_last tensors")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 200 ms during the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# reference / CPU baseline (oracle port; the ONLY place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------------------------
CPU_SAMPLE = 128   # images per CPU pass: ~8 s on 16 cores (bounded sample of the 4096-image step)


def cpu_reference_images_per_s(workload: str, hw: int, sample: int, repeats: int, warmup: int = 1):
    """Time the oracle port of the reference path on the host cores: script-16 degradation per image (as the
    reference runs it, one image per call), then ToTensor -> restorer -> clamp/u8 -> Normalize -> VGG16 -> arg-max in
    fp32 PyTorch with all threads.  Returns (images/s, cores, seconds per sample)."""
    import numpy as np
    import torch
    from b200restore import synth
    from oracle import degrade_oracle as DO, generators_oracle as GO, models_oracle as MO
    arch, recipe, classify = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sdc = {k: synth.synthetic_state_dict("simple_unet", sd) for k, sd in CASCADE_SEEDS.items()} if arch == "cascade3" else None
    sdr = synth.synthetic_state_dict("simple_unet" if arch == "cascade3" else arch, 31)
    sdj = synth.synthetic_state_dict("vgg16", 32) if classify else None
    fn = MO.simple_unet_forward if arch == "simple_unet" else MO.resunet_forward
    imgs, labels = synth.sign_like_images(sample, hw, hw, seed=7)
    imgs_np = imgs.numpy()
    rng = np.random.default_rng(0)

    chunk = 32   # 17_run_unified_inference.py:73 restores in batches of 32
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=cores)

    def one_chunk(lo, hi):
        deg = []
        if recipe == "stress13":
            for i in range(lo, hi):   # 13:152-171, one image per call as the reference runs it
                z = GO.stress_add_noise(GO.stress_add_fog(GO.stress_add_blur(imgs_np[i])),
                                        rng.normal(0, 0.01 ** 0.5, imgs_np[i].shape))
                deg.append(z)
            with torch.no_grad():
                _, snaps = GO.cascade_13(sdc, torch.from_numpy(np.stack(deg)))
                pred, conf = GO.vgg_prediction(sdj, snaps[-1])
            return int((pred == labels[lo:hi]).sum())
        def degrade_one(i):   # one image per call as the reference runs it (16:43-47), own generator per image
            noise = np.random.default_rng(i).normal(0, 0.02 ** 0.5, imgs_np[i].shape)
            if recipe == "compound16":
                return DO.compound_16(imgs_np[i], noise)
            return DO.random_14(imgs_np[i], 0.5, noise, 10, 45)

        # the reference degrades serially on one core; spreading the per-image calls over the host cores (NumPy and
        # OpenCV release the GIL) is the fair form of the baseline
        deg = torch.from_numpy(np.stack(list(pool.map(degrade_one, range(lo, hi)))))
        with torch.no_grad():
            if classify:
                _, _, pred = MO.restore_then_classify(fn, sdr, sdj, deg)
                return int((pred == labels[lo:hi]).sum())
            out = fn(sdr, MO.to_tensor_u8(deg))
            return int(MO.quantize_restored(out).sum() > 0)

    def one_pass(limit=sample):
        return sum(one_chunk(lo, min(lo + chunk, limit)) for lo in range(0, limit, chunk))

    for _ in range(warmup):
        one_pass(min(sample, chunk))   # one batch warms the thread pool and the allocator
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times)
    pool.shutdown()
    return sample / dt, cores, dt


def cuda_reference_images_per_s(workload: str, hw: int, batch: int, micro: int, steps: int, warmup: int, precision: str):
    """BASELINE.md 'same-box GPU comparator': the fp32 PyTorch port of restore -> clamp/u8 -> Normalize -> VGG16 ->
    arg-max on cuda:0, i.e. what the reference's modules do with DEVICE = 'cuda' (17:16): ATen -> cuDNN / cuBLAS.  The
    degradation is NumPy/OpenCV code in the reference (no GPU form), so the batch is degraded once outside the timed
    region and the timed step is restore + classify over `batch` resident images.  None of this repo's kernels run."""
    import torch
    from b200restore import synth
    from oracle import models_oracle as MO
    arch, recipe, classify = WORKLOADS[workload]
    if arch == "cascade3":
        raise SystemExit("--ref-device cuda covers the restore(+classify) workloads")
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    tf32 = precision == "tf32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    cl = precision == "bf16"

    def put(sd):
        out = {}
        for k, v in sd.items():
            v = v.to(dev)
            out[k] = v.contiguous(memory_format=torch.channels_last) if (cl and v.dim() == 4) else v
        return out

    sdr = put(synth.synthetic_state_dict(arch, 31))
    sdj = put(synth.synthetic_state_dict("vgg16", 32)) if classify else None
    fn = MO.simple_unet_forward if arch == "simple_unet" else MO.resunet_forward
    imgs, labels = synth.sign_like_images(min(batch, 512), hw, hw, seed=7)
    reps = (batch + imgs.shape[0] - 1) // imgs.shape[0]
    imgs = imgs.repeat(reps, 1, 1, 1)[:batch].to(dev)
    labels = labels.repeat(reps)[:batch].to(dev)

    def step():
        correct = torch.zeros((), dtype=torch.int64, device=dev)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
            for s0 in range(0, batch, micro):
                u8 = imgs[s0:s0 + micro]
                if classify:
                    _, _, pred = MO.restore_then_classify(fn, sdr, sdj, u8)
                    correct += (pred == labels[s0:s0 + micro]).sum()
                else:
                    x = MO.to_tensor_u8(u8)
                    correct += MO.quantize_restored(fn(sdr, x.contiguous(memory_format=torch.channels_last) if cl else x)).sum()
        return correct

    for _ in range(max(warmup, 1)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return batch / ms * 1e3, ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        batch = min(args.batch, 1024)
        micro = min(args.micro_batch, 64)
        ips, ms = cuda_reference_images_per_s(args.workload, args.hw, batch, micro, args.steps, args.warmup, args.ref_precision)
        arch, recipe, classify = WORKLOADS[args.workload]
        gflop = GFLOP_PER_IMAGE_224[arch] + (GFLOP_PER_IMAGE_224["vgg16"] if classify else 0.0)
        print(json.dumps({
            "impl": "reference-cudnn", "metric": "images/sec restore->VGG16 classify", "value": ips, "unit": "images/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "dtype": args.ref_precision, "data": "synthetic",
            "config": {"workload": args.workload, "hw": args.hw, "images_per_step": batch, "micro_batch": micro,
                       "note": "fp32 PyTorch port of the reference modules on cuda:0 (ATen -> cuDNN / cuBLAS, "
                               "cudnn.benchmark on); degradation outside the timed region (NumPy/OpenCV in the "
                               "reference); none of this repo's kernels"},
            "model_tflops": ips * gflop / 1e3 if args.hw == 224 else None, "gpu_launches": 0}), flush=True)
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or CPU_SAMPLE
    t0 = time.perf_counter()
    ips, cores, dt = cpu_reference_images_per_s(args.workload, args.hw, sample, repeats=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "images/sec restore->VGG16 classify", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "hw": args.hw, "images_per_step": sample,
                   "note": "reference CPU path (oracle port, pinned to the reference's classes), host cores only"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images of {args.hw}x{args.hw} per step in batches of 32 (17:73), median of {args.steps} steps"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import b200restore as B
    from b200restore import degrade as D, models, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: libb2r has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    arch, recipe, classify = WORKLOADS[args.workload]
    B_, hw, mb = args.batch, args.hw, args.micro_batch
    lo = rank * B_                                   # weak scaling: rank r owns global images [r*B, (r+1)*B)

    cascade = None
    if arch == "cascade3":
        from b200restore import generators as G
        nets = {}
        for k, sd_seed in CASCADE_SEEDS.items():
            net = models.SimpleUNet()
            net.load_state_dict(synth.synthetic_state_dict("simple_unet", sd_seed))
            nets[k] = net.to(dev).eval()
        cascade = G.CascadeRestorer(nets)
        restorer = nets["Noise"]
    else:
        restorer = (models.SimpleUNet if arch == "simple_unet" else models.ResUNet)()
        restorer.load_state_dict(synth.synthetic_state_dict(arch, 31))
        restorer = restorer.to(dev).eval()
    judge = models.VGG16Judge()
    judge.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
    judge = judge.to(dev).eval()
    pipe = B.RestoreClassifyPipeline(restorer, judge, micro_batch=mb)

    # synthetic GTSRB-shaped batch, generated in chunks on the host into pinned memory, then copied once to HBM
    host_imgs = torch.empty((B_, hw, hw, 3), dtype=torch.uint8).pin_memory()
    host_labels = torch.empty((B_,), dtype=torch.int64).pin_memory()
    for s in range(0, B_, 512):
        c = min(512, B_ - s)
        im, lb = synth.sign_like_images(c, hw, hw, seed=7, index0=lo + s)
        host_imgs[s:s + c] = im
        host_labels[s:s + c] = lb
    dev_imgs = host_imgs.to(dev)
    dev_labels = host_labels.to(dev)
    params = (D.compound_params(B_) if recipe in ("compound16", "stress13")
              else D.random_params(B_, np.random.default_rng(2 + rank), order=0)).to(dev)

    def cascade_micro_batch(imgs_u8, labels, counts, index0):
        z = G.stress_distort(imgs_u8, seed=2, image_index0=index0)[-1]
        _, hist = cascade(z)
        logits = judge.forward_u8(hist[-1][1])
        ops.argmax_count(logits, labels, counts, want_conf=True)

    def step_device():
        if cascade is not None:
            counts = torch.zeros(2, dtype=torch.int64, device=dev)
            for s in range(0, B_, mb):
                c = min(mb, B_ - s)
                cascade_micro_batch(dev_imgs[s:s + c], dev_labels[s:s + c], counts, lo + s)
        elif classify:
            _, counts = pipe.run(dev_imgs, dev_labels, params, seed=2, image_index0=lo)
        else:
            counts = torch.zeros(2, dtype=torch.int64, device=dev)
            for s in range(0, B_, mb):
                c = min(mb, B_ - s)
                sub = B.pipeline._slice_params(params, s, c)
                deg = D.degrade(dev_imgs[s:s + c], sub, seed=2, image_index0=lo + s)
                restorer.restore_u8(deg)
            counts[1] = B_
        B.all_reduce_counts(counts)
        return counts

    def step_host():
        if cascade is not None:
            counts = torch.zeros(2, dtype=torch.int64, device=dev)
            h2d = 0
            for s in range(0, B_, mb):
                c = min(mb, B_ - s)
                im = host_imgs[s:s + c].to(dev, non_blocking=True)
                lb = host_labels[s:s + c].to(dev, non_blocking=True)
                h2d += im.numel() + lb.numel() * 8
                cascade_micro_batch(im, lb, counts, lo + s)
            hc = counts.cpu()
            cc = torch.tensor([int(hc[0]), int(hc[1])], dtype=torch.int64, device=dev)
            B.all_reduce_counts(cc)
            return cc, h2d, 16
        (correct, total), h2d, d2h = pipe.run_from_host(host_imgs, host_labels, params, seed=2, image_index0=lo)
        c = torch.tensor([correct, total], dtype=torch.int64, device=dev)
        B.all_reduce_counts(c)
        return c, h2d, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            res = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()

    # ---- timed region: K steps, per-launch events on the dominant kernel, clocks sampled
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timer = ops.KernelTimer(kinds=("conv_gemm",))
    launches0 = ops.STATS["launches"]
    with ops.timing(timer):
        total_ms, counts = timed(step_device, args.steps)
    launches = ops.STATS["launches"] - launches0
    ksum = timer.summary().get("conv_gemm", {"launches": 0, "work": 0.0, "ms": 1e-9})
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end (host buffers) : same K steps
    e2e = None
    if classify:
        pipe.use_graph = bool(args.graph) and cascade is None
        step_host()
        e2e_ms, (c2, h2d, d2h) = timed(step_host, args.steps)
        pipe.use_graph = False
        e2e = {"value": B_ * world * args.steps / (e2e_ms / 1e3), "unit": "images/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
               "cuda_graph": bool(args.graph) and cascade is None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
    conv_tf = ksum["work"] / (ksum["ms"] * 1e-3) / 1e12
    traffic = None
    tj = ROOT / "profiles" / "r01_conv_traffic.json"      # dram bytes per launch of these kernels from one ncu capture
    if tj.exists():
        traffic = json.loads(tj.read_text()).get("dram_bytes_per_launch")
        if traffic is not None:   # captured at micro-batch 256, 224x224: activation traffic scales with the launch size
            traffic = traffic * (mb / 256.0) * (hw * hw / (224.0 * 224.0))
    ips = B_ * world * args.steps / (total_ms / 1e3)
    gflop_img = GFLOP_PER_IMAGE_224[arch] + (GFLOP_PER_IMAGE_224["vgg16"] if classify else 0.0)
    line = {
        "metric": "images/sec restore->VGG16 classify", "value": ips, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_gpu_per_step": B_, "hw": hw, "micro_batch": mb,
                   "restorer": arch, "judge": "vgg16-43" if classify else None, "degradation": recipe,
                   "weights": "seeded synthetic state_dicts in the shipped schemas",
                   "l2": "inputs (%.0f MB/step) larger than L2; no flush needed" % (B_ * hw * hw * 3 / 1e6),
                   "parallelism": f"dp{world} (batch sharded, one int64[2] all-reduce per step)"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convs (conv_gemm_pair_kernel<256> cta_group::2, conv_w3_kernel, conv_gemm_halo_kernel<128>, conv_gemm_kernel<N>)",
                     "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf,
                     "peak_source": peak_src, "traffic": traffic,
                     "traffic_note": "dram__bytes_read+write per launch averaged over all tcgen05 conv launches of one "
                                     "ncu capture at micro-batch 256 (profiles/r01_launches_v11_summary.md), scaled to this run's "
                                     "launch size (activations dominate: traffic is linear in images per launch)",
                     "launches": ksum["launches"], "kernel_ms_per_step": ksum["ms"] / args.steps,
                     "dominant_kernel": max(ksum.get("by_kernel", {"?": {"ms": 0}}).items(), key=lambda kv: kv[1]["ms"])[0],
                     "by_kernel": {k: {"launches": v["launches"], "share_of_step": v["ms"] / total_ms,
                                       "achieved": v["work"] / (v["ms"] * 1e-3) / 1e12,
                                       "frac": v["work"] / (v["ms"] * 1e-3) / 1e12 / peak_tf}
                                   for k, v in sorted(ksum.get("by_kernel", {}).items(), key=lambda kv: -kv[1]["ms"])},
                     "share_of_step": ksum["ms"] / total_ms,
                     "pipeline_tflops": ips / world * gflop_img / 1e3 if hw == 224 else None},
        "counts": [int(counts[0]), int(counts[1])],
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or CPU_SAMPLE
        v, cores, dt = cpu_reference_images_per_s(args.workload, hw, sample, repeats=1, warmup=1)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"{sample} images of {hw}x{hw} in batches of 32 (17:73), one timed pass after a warm-up batch "
                                          f"({dt:.1f} s), degradation calls spread over the cores; oracle port of " +
                                          ("script 13 (distortions, cascade, VGG confidence)" if recipe == "stress13"
                                           else "scripts 16 -> 17 -> 18") + " in fp32 PyTorch"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
