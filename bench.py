#!/usr/bin/env python
"""bench.py — images/s of the degrade -> restore -> VGG16 classify -> top-1 count path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--batch B]
                    [--total-images T] [--hw S]

Default mode (what the driver runs): one "step" = one pass of the hot path over B = 4096 synthetic GTSRB-shaped images per
GPU (weak scaling: rank r owns the contiguous block [r*B, (r+1)*B) of the global image index range).  The (correct, total)
counts stay on each device for all K steps; ONE NCCL all-reduce at the end of the timed region combines them
(18_test_unified_benchmark.py:44-51 accumulates per batch on the host; north star: "a single NCCL all-reduce").
For N > 1 launch under torchrun (one rank per GPU).

`--total-images T` is BASELINE.json configs[3] literally: T images (1 000 000) with global index i (Philox counter, label
i mod 43), rank r takes [r*T/R, (r+1)*T/R) (strong scaling), one all-reduce at the end; `value` = T / seconds.

Prints ONE JSON line on rank 0:
  value        images/s with the inputs already resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e          same metric through the public API with HOST buffers (pinned H2D of every micro-batch + D2H of the counts
               inside the timed region); restore -> classify -> count replayed as a CUDA graph per micro-batch
  roofline     the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOPs / summed CUDA-event duration of
               its launches inside the timed region, against MEASURED_PEAKS.json; `by_kernel` splits it by the kernel each
               launch ran
  roofline_hbm the fused degradation kernel (HBM-bound class): algorithmic bytes / CUDA-event time, per recipe
  per_rank     ms per step, conv kernel ms, SM clock, host launch seconds of every rank (min / median / max + the slowest)
  library_comparator  the reference's own modules on cuda:0 through cuDNN / cuBLAS (f32 and bf16 autocast), bounded sample
  cpu_baseline the UNMODIFIED reference scripts (baseline/_ref, tools/install_ref.py) on this box's host cores, bounded
               sample, rank 0, N = 1

`--impl reference` times that reference CPU path alone (kind "reference"; the oracle port is the fallback when
baseline/_ref is absent, and the only other place bench.py executes oracle/).
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
GFLOP_PER_IMAGE = {   # BASELINE.md section 3 (hooked reference modules): hw -> (SimpleUNet, ResUNet, VGG16-43)
    64: (3.170, 4.571, 2.745), 128: (12.679, 18.283, 10.262), 224: (38.831, 55.992, 30.933), 256: (50.718, 73.132, 40.329)}
CASCADE_SEEDS = {"Noise": 31, "Fog": 33, "Blur": 34}
POOL_IMAGES = 43 * 96   # --total-images: resident pool of clean images, a multiple of 43 so that label(i) = i mod 43


def gflop_per_image(arch: str, classify: bool, hw: int):
    if hw not in GFLOP_PER_IMAGE:
        return None
    u, r, v = GFLOP_PER_IMAGE[hw]
    return {"simple_unet": u, "resunet": r, "cascade3": 3 * u}[arch] + (v if classify else 0.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 5 (1 with --total-images)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="degrade16_resunet_vgg16_top1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=0,
                    help="resident images per pass; 0 = 512 x (224 / hw)^2 rounded to a multiple of 256, at most --batch: the same "
                         "number of PIXELS per launch at every resolution (small maps otherwise run a fraction of a wave per "
                         "launch on the 8x8 / 4x4 levels)")
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--total-images", type=int, default=0,
                    help="BASELINE configs[3]: this many images in total, sharded [r*T/R, (r+1)*T/R) over the ranks (strong "
                         "scaling), one all-reduce at the end; --steps / --warmup then count whole passes (default 1 / 0)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images per CPU-baseline sample (0 = choose)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-comparator", action="store_true", help="skip the cuDNN / cuBLAS same-box comparator leg")
    ap.add_argument("--graph", type=int, default=1, choices=[0, 1],
                    help="e2e leg: replay restore -> classify -> count of a micro-batch as one CUDA graph (bit-identical to "
                         "eager; the device-resident `value` leg stays eager: it times every conv launch with CUDA events)")
    ap.add_argument("--pin-cores", type=int, default=1, choices=[0, 1],
                    help="N > 1: give every rank its own slice of the host cores")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: 'cuda' times the reference's modules on the GPU through cuDNN / cuBLAS "
                         "(the reference with DEVICE='cuda', 17:16) as the same-box library comparator")
    ap.add_argument("--ref-precision", default="f32", choices=["f32", "tf32", "bf16"],
                    help="--ref-device cuda: f32 (TF32 off), tf32, or bf16 autocast with channels_last tensors")
    return ap.parse_args()


def config_dict(args, world: int):
    """The workload description both arms print (the reference arm adds what its bounded sample was)."""
    arch, recipe, classify = WORKLOADS[args.workload]
    cfg = {"workload": args.workload, "images_per_gpu_per_step": args.batch, "hw": args.hw,
           "restorer": arch, "judge": "vgg16-43" if classify else None, "degradation": recipe,
           "weights": "seeded synthetic state_dicts in the shipped schemas"}
    if args.total_images:
        cfg["total_images"] = args.total_images
        cfg["images_per_gpu_per_step"] = None
    return cfg


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 200 ms during the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, nvml: bool = False):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = nvml          # ranks > 0 poll NVML in-process (the same counters nvidia-smi prints) instead of forking a poller each
        self._go = False

    def mark(self) -> int:
        return len(self.lines)

    def _poll_nvml(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while self._go:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.lines.append("%d, %d, %.2f, %s" % (nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                                        ", ".join("Active" if r & b else "Not Active" for b in bits.values())))
                time.sleep(0.2)
        except Exception:      # no NVML: this rank reports no clocks
            pass

    def start(self):
        if self.nvml:
            self._go = True
            self.proc = True
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
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
            return
        if self.nvml:
            self._go = False
            self.thread.join(timeout=2)
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def summary(self, lo: int = 0, hi=None):
        """clocks over the samples [lo, hi) (marks taken with mark())"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, watts = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[lo:hi]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                watts.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": statistics.median(watts) if watts else None}


# --------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# --------------------------------------------------------------------------------------------------------------
CPU_SAMPLE = 128   # images per CPU pass: ~8 s on 16 cores (bounded sample of the 4096-image step)


def cpu_reference_images_per_s(workload: str, hw: int, sample: int, repeats: int, warmup: int = 1):
    """The reference path on the host cores.  Returns (images/s, cores, seconds per pass, kind).

    kind "reference": the unmodified scripts in baseline/_ref driven by baseline/reference_arm.py.
    kind "port": oracle/ (pinned bit-for-bit to the reference by tests/test_oracle_*.py) — only for the script-13 cascade
    workload (13_pipeline_stress_test.py imports matplotlib, which this image lacks) or when baseline/_ref is absent."""
    arch, recipe, classify = WORKLOADS[workload]
    from baseline import reference_arm as RA
    if arch != "cascade3" and RA.available():
        ips, cores, dt, _ = RA.images_per_s(arch, recipe, classify, hw, sample, repeats, warmup)
        return ips, cores, dt, "reference"
    ips, cores, dt = _oracle_port_images_per_s(workload, hw, sample, repeats, warmup)
    return ips, cores, dt, "port"


def _oracle_port_images_per_s(workload: str, hw: int, sample: int, repeats: int, warmup: int = 1):
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
        if recipe == "stress13":
            deg = []
            for i in range(lo, hi):   # 13:152-171, one image per call as the reference runs it
                deg.append(GO.stress_add_noise(GO.stress_add_fog(GO.stress_add_blur(imgs_np[i])),
                                               rng.normal(0, 0.01 ** 0.5, imgs_np[i].shape)))
            with torch.no_grad():
                _, snaps = GO.cascade_13(sdc, torch.from_numpy(np.stack(deg)))
                pred, conf = GO.vgg_prediction(sdj, snaps[-1])
            return int((pred == labels[lo:hi]).sum())

        def degrade_one(i):
            noise = np.random.default_rng(i).normal(0, 0.02 ** 0.5, imgs_np[i].shape)
            if recipe == "compound16":
                return DO.compound_16(imgs_np[i], noise)
            return DO.random_14(imgs_np[i], 0.5, noise, 10, 45)

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
        one_pass(min(sample, chunk))
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
    pool.shutdown()
    dt = statistics.median(times)
    return sample / dt, cores, dt


def cuda_reference_images_per_s(workload: str, hw: int, batch: int, micro: int, steps: int, warmup: int, precision: str):
    """BASELINE.md 'same-box GPU comparator': the reference's OWN modules (baseline/_ref ResUNet / SimpleUNet, torchvision
    vgg16 + head swap) on cuda:0, i.e. what the reference does with DEVICE = 'cuda' (17:16): ATen -> cuDNN / cuBLAS,
    restore -> clamp -> u8 -> Normalize -> VGG16 -> arg-max.  The degradation is NumPy/OpenCV code in the reference (no GPU
    form), so the timed step is restore + classify over `batch` resident images.  None of this repo's kernels run."""
    import torch
    from b200restore import synth
    from baseline import reference_arm as RA
    arch, recipe, classify = WORKLOADS[workload]
    if arch == "cascade3":
        raise SystemExit("--ref-device cuda covers the restore(+classify) workloads")
    if not RA.available():
        raise RuntimeError("baseline/_ref is missing (python tools/install_ref.py)")
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    tf32 = precision == "tf32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    cl = precision == "bf16"
    restorer, judge = RA.build_models(arch, classify)
    fmt = torch.channels_last if cl else torch.contiguous_format
    restorer = restorer.to(dev).to(memory_format=fmt)
    judge = judge.to(dev).to(memory_format=fmt) if judge is not None else None
    imgs, labels = synth.sign_like_images(min(batch, 256), hw, hw, seed=7)
    reps = (batch + imgs.shape[0] - 1) // imgs.shape[0]
    imgs = imgs.repeat(reps, 1, 1, 1)[:batch].to(dev)
    labels = labels.repeat(reps)[:batch].to(dev)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)

    def step():
        correct = torch.zeros((), dtype=torch.int64, device=dev)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
            for s0 in range(0, batch, micro):
                x = (imgs[s0:s0 + micro].permute(0, 3, 1, 2).float() / 255.0).contiguous(memory_format=fmt)  # ToTensor (17:66)
                out = torch.clamp(restorer(x), 0, 1)                                                        # 17:85-86
                if judge is None:
                    correct += (out.float() * 255).to(torch.uint8).sum()
                    continue
                u8 = (out.float() * 255).to(torch.uint8)                                                    # 17:92
                xn = ((u8.float() / 255.0 - mean) / std).contiguous(memory_format=fmt)                      # 18:28-32
                _, predicted = torch.max(judge(xn), 1)                                                      # 18:46-47
                correct += (predicted == labels[s0:s0 + micro]).sum()
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
    del restorer, judge
    torch.cuda.empty_cache()
    return batch / ms * 1e3, ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.steps = args.steps or 5
    arch, recipe, classify = WORKLOADS[args.workload]
    if args.ref_device == "cuda":
        batch = min(args.batch, 1024)
        micro = min(args.micro_batch or 64, 64)
        ips, ms = cuda_reference_images_per_s(args.workload, args.hw, batch, micro, args.steps, args.warmup, args.ref_precision)
        gflop = gflop_per_image(arch, classify, args.hw)
        print(json.dumps({
            "impl": "reference-cudnn", "metric": "images/sec restore->VGG16 classify", "value": ips, "unit": "images/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "dtype": args.ref_precision, "data": "synthetic",
            "config": {"workload": args.workload, "hw": args.hw, "images_per_step": batch, "micro_batch": micro,
                       "note": "the reference's own modules (baseline/_ref) on cuda:0 (ATen -> cuDNN / cuBLAS, "
                               "cudnn.benchmark on); degradation outside the timed region (NumPy/OpenCV in the "
                               "reference); none of this repo's kernels"},
            "model_tflops": ips * gflop / 1e3 if gflop else None, "gpu_launches": 0}), flush=True)
        return
    # a bounded sample per step so that the whole K-step run ends within a few minutes (~11 images/s on 16 cores at 224x224)
    sample = args.cpu_sample or (CPU_SAMPLE if args.steps <= 8 else CPU_SAMPLE // 2)
    t0 = time.perf_counter()
    ips, cores, dt, kind = cpu_reference_images_per_s(args.workload, args.hw, sample, repeats=args.steps, warmup=args.warmup)
    cfg = config_dict(args, 1)
    cfg["sample_images_per_step"] = sample
    cfg["note"] = ("the unmodified reference scripts (baseline/_ref: 16 -> 17 -> 18 composed in memory) on the host cores"
                   if kind == "reference" else "oracle port of the reference path on the host cores")
    line = {
        "impl": "reference", "metric": "images/sec restore->VGG16 classify", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.total_images else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} images of {args.hw}x{args.hw} per step in batches of 32 (17:73), median of "
                                   f"{args.steps} steps; a bounded sample of the {args.batch}-image step"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def pin_cores(rank_local: int, world_local: int):
    """Give each rank of the node its own contiguous slice of the cores this process may use."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world_local
        if per >= 2:
            mine = cores[rank_local * per:(rank_local + 1) * per]
            os.sched_setaffinity(0, mine)
            return mine
    except (AttributeError, OSError):
        pass
    return None


def minmedmax(vals):
    vals = [v for v in vals if v is not None]
    if not vals:
        return None
    return {"min": min(vals), "median": statistics.median(vals), "max": max(vals)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import b200restore as B
    from b200restore import degrade as D, models, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: libb2r has no CPU fallback")
    cores_mine = pin_cores(local, local_world) if (world > 1 and args.pin_cores) else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    arch, recipe, classify = WORKLOADS[args.workload]
    if args.micro_batch <= 0:
        args.micro_batch = max(256, min(args.batch, int(round(512 * (224.0 / args.hw) ** 2 / 256.0)) * 256))
    B_, hw, mb = args.batch, args.hw, args.micro_batch
    total_mode = args.total_images > 0
    if total_mode and (not classify or arch == "cascade3"):
        raise SystemExit("--total-images is the degrade -> restore -> classify -> count workload")

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

    # ---- inputs: weak mode = this rank's B images; total mode = a resident pool every rank indexes by global image index
    if total_mode:
        lo, hi = B.shard_range(args.total_images, rank, world)
        n_host, index0 = POOL_IMAGES, 0
    else:
        lo, hi = rank * B_, (rank + 1) * B_
        n_host, index0 = B_, lo
    host_imgs = torch.empty((n_host, hw, hw, 3), dtype=torch.uint8).pin_memory()
    host_labels = torch.empty((n_host,), dtype=torch.int64).pin_memory()
    for s in range(0, n_host, 512):
        c = min(512, n_host - s)
        im, lb = synth.indexed_images(index0 + s, c, hw, hw, seed=7)     # image i is a pure function of its GLOBAL index
        host_imgs[s:s + c] = im
        host_labels[s:s + c] = lb
    dev_imgs = host_imgs.to(dev)
    dev_labels = host_labels.to(dev)
    n_par = mb if total_mode else B_
    params = (D.compound_params(n_par) if recipe in ("compound16", "stress13")
              else D.random_params(n_par, np.random.default_rng(2 + rank), order=0)).to(dev)

    def cascade_micro_batch(imgs_u8, labels, counts, index0_):
        z = G.stress_distort(imgs_u8, seed=2, image_index0=index0_)[-1]
        _, hist = cascade(z)
        logits = judge.forward_u8(hist[-1][1])
        ops.argmax_count(logits, labels, counts, want_conf=True)

    totals = torch.zeros(2, dtype=torch.int64, device=dev)      # (correct, total) of this rank over the timed steps

    def step_device():
        if cascade is not None:
            for s in range(0, B_, mb):
                c = min(mb, B_ - s)
                cascade_micro_batch(dev_imgs[s:s + c], dev_labels[s:s + c], totals, lo + s)
        elif classify:
            _, counts = pipe.run(dev_imgs, dev_labels, params, seed=2, image_index0=lo)
            totals.add_(counts)
        else:
            for s in range(0, B_, mb):
                c = min(mb, B_ - s)
                sub = B.pipeline._slice_params(params, s, c)
                deg = D.degrade(dev_imgs[s:s + c], sub, seed=2, image_index0=lo + s)
                restorer.restore_u8(deg)
            totals[1] += B_

    def pool_slice(imgs, labels, g0, c):
        """c images starting at global index g0 out of the resident pool (image i = pool[i mod P], label = i mod 43)."""
        o = g0 % POOL_IMAGES
        if o + c <= POOL_IMAGES:
            return imgs[o:o + c], labels[o:o + c]
        k = POOL_IMAGES - o
        return torch.cat((imgs[o:], imgs[:c - k])), torch.cat((labels[o:], labels[:c - k]))

    def pass_total_device():
        """BASELINE configs[3]: this rank's contiguous shard [lo, hi) of the T images, device-resident pool."""
        for g0 in range(lo, hi, mb):
            c = min(mb, hi - g0)
            im, lb = pool_slice(dev_imgs, dev_labels, g0, c)
            pipe.run_micro_batch(im, lb, B.pipeline._slice_params(params, 0, c), 2, g0, totals)

    e2e_bytes = [0, 0]

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
            totals.add_(torch.tensor([int(hc[0]), int(hc[1])], dtype=torch.int64, device=dev))
            e2e_bytes[0], e2e_bytes[1] = h2d, 16
            return
        (correct, total), h2d, d2h = pipe.run_from_host(host_imgs, host_labels, params, seed=2, image_index0=lo)
        totals.add_(torch.tensor([correct, total], dtype=torch.int64, device=dev))
        e2e_bytes[0], e2e_bytes[1] = h2d, d2h

    def pass_total_host():
        """The same shard with HOST buffers: every micro-batch is copied from the pinned pool inside the timed region."""
        h2d = 0
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        for g0 in range(lo, hi, mb):
            c = min(mb, hi - g0)
            o = g0 % POOL_IMAGES
            if o + c > POOL_IMAGES:      # wrap: two pinned slices
                k = POOL_IMAGES - o
                im = torch.cat((host_imgs[o:].to(dev, non_blocking=True), host_imgs[:c - k].to(dev, non_blocking=True)))
                lb = torch.cat((host_labels[o:].to(dev, non_blocking=True), host_labels[:c - k].to(dev, non_blocking=True)))
            else:
                im = host_imgs[o:o + c].to(dev, non_blocking=True)
                lb = host_labels[o:o + c].to(dev, non_blocking=True)
            h2d += im.numel() + lb.numel() * 8
            pipe.run_micro_batch(im, lb, B.pipeline._slice_params(params, 0, c), 2, g0, counts)
        hc = counts.cpu()
        totals.add_(torch.tensor([int(hc[0]), int(hc[1])], dtype=torch.int64, device=dev))
        e2e_bytes[0], e2e_bytes[1] = h2d, 16

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; the ONE all-reduce of the counts is inside the timed region.
        Returns (max-over-ranks ms, this rank's ms, host seconds spent launching, global counts)."""
        totals.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_s = time.perf_counter() - t0
        em = torch.cuda.Event(enable_timing=True)
        em.record()                       # this rank's own work ends here; the all-reduce then waits for the slowest rank
        global_counts = totals.clone()
        B.all_reduce_counts(global_counts)
        e1.record()
        barrier()
        own = e0.elapsed_time(e1)
        timed.compute_ms = e0.elapsed_time(em)
        ms = torch.tensor([own], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), own, host_s, global_counts

    # ---- HBM-bound class: the fused degradation kernel timed ALONE, per recipe (CUDA events on the launching stream, inputs
    # 617 MB >> L2), before the long legs pull the board into its power cap: MEASURED_PEAKS.json's hbm_gbs is a kernel-alone
    # figure too.  (Inside the pipeline the kernel is 1.2 % of the step.)
    hbm = None
    if rank == 0 and not total_mode:
        hbm = degrade_roofline(torch, D, dev_imgs, hw)

    if total_mode:
        steps, warm = max(args.steps or 1, 1), 0
        fn_dev, fn_host = pass_total_device, pass_total_host
        images_per_step = args.total_images
        # warm-up: a few micro-batches (lazy module loads, func attributes, workspaces)
        for g0 in range(lo, min(lo + 3 * mb, hi), mb):
            c = min(mb, hi - g0)
            im, lb = pool_slice(dev_imgs, dev_labels, g0, c)
            pipe.run_micro_batch(im, lb, B.pipeline._slice_params(params, 0, c), 2, g0, totals)
    else:
        steps, warm = args.steps or 5, max(args.warmup, 3)
        fn_dev, fn_host = step_device, step_host
        images_per_step = B_ * world
        for _ in range(warm):
            step_device()
    torch.cuda.synchronize()

    # ---- timed region (`value`): K steps, inputs resident in HBM, clocks sampled on every rank.  restore -> classify -> count of
    # each micro-batch is replayed as one CUDA graph (bit-identical to eager, tests/test_pipeline_gpu.py): ~330 launches per
    # micro-batch cost the Python host ~40 ms against 45 ms of GPU work, which made a rank host-bound as soon as its core was
    # shared (the 1 -> 8 efficiency of round 1).  --graph 0 keeps this leg eager.
    sampler = ClockSampler(local, nvml=rank != 0)
    sampler.start()
    use_graph = bool(args.graph) and cascade is None and classify
    pipe.use_graph = use_graph
    if use_graph:                       # capture outside the timed region
        if total_mode:
            im, lb = pool_slice(dev_imgs, dev_labels, lo, min(mb, hi - lo))
            pipe.run_micro_batch(im, lb, B.pipeline._slice_params(params, 0, im.shape[0]), 2, lo, totals)
        else:
            step_device()
        torch.cuda.synchronize()
    mark_start = sampler.mark()
    launches0, replayed0 = ops.STATS["launches"], pipe.graph_launches_replayed
    total_ms, own_ms, host_s, counts = timed(fn_dev, steps)
    compute_ms = timed.compute_ms         # before the all-reduce: shows which rank the others waited for
    graph_launches = pipe.graph_launches_replayed - replayed0
    launches = ops.STATS["launches"] - launches0 + graph_launches
    mark_timed = sampler.mark()
    pipe.use_graph = False

    # ---- instrumented steps (roofline): the SAME step, eager, every tcgen05 conv launch bracketed by a CUDA-event pair on the
    # launching stream; max(1, K / 5) steps right after the timed region (same clocks, same resident inputs)
    inst_steps = 1 if total_mode else max(2, steps // 4)
    timer = ops.KernelTimer(kinds=("conv_gemm",))
    if total_mode:
        def fn_inst():
            for g0 in range(lo, min(lo + 8 * mb, hi), mb):
                c = min(mb, hi - g0)
                im, lb = pool_slice(dev_imgs, dev_labels, g0, c)
                pipe.run_micro_batch(im, lb, B.pipeline._slice_params(params, 0, c), 2, g0, totals)
    else:
        fn_inst = fn_dev
    with ops.timing(timer):
        _, inst_ms, inst_host_s, _ = timed(fn_inst, inst_steps)
    ksum = timer.summary().get("conv_gemm", {"launches": 0, "work": 0.0, "ms": 1e-9})
    sampler.stop()
    clocks = sampler.summary(mark_start, mark_timed)      # the timed region
    clocks_inst = sampler.summary(mark_timed, None)       # the instrumented steps

    # ---- end-to-end (host buffers): same K steps
    e2e = None
    do_e2e = classify and (not total_mode or (hi - lo) <= 300000)
    if do_e2e:
        pipe.use_graph = bool(args.graph) and cascade is None
        if total_mode:      # capture the micro-batch graph outside the timed region
            im, lb = pool_slice(dev_imgs, dev_labels, lo, min(mb, hi - lo))
            pipe.run_micro_batch(im, lb, B.pipeline._slice_params(params, 0, im.shape[0]), 2, lo, totals)
        else:
            step_host()
        e2e_ms, _, _, c2 = timed(fn_host, steps)
        pipe.use_graph = False
        e2e = {"value": images_per_step * steps / (e2e_ms / 1e3), "unit": "images/s",
               "h2d_bytes_per_step": int(e2e_bytes[0]), "d2h_bytes_per_step": int(e2e_bytes[1]),
               "ms_per_step": e2e_ms / steps, "cuda_graph": bool(args.graph) and cascade is None,
               "counts": [int(c2[0]), int(c2[1])]}

    # ---- per-rank record (the 1 -> N curve explains itself)
    mine = {"rank": rank, "ms_per_step": own_ms / steps, "compute_ms_per_step": compute_ms / steps,
            "conv_kernel_ms_per_step": ksum["ms"] / inst_steps,
            "host_launch_s_per_step": host_s / steps, "sm_mhz": clocks.get("sm_mhz"), "power_w": clocks.get("power_w"),
            "reasons": clocks.get("reasons"), "cores": len(cores_mine) if cores_mine else None, "images": int(hi - lo)}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]

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
    traffic, traffic_note = None, None
    for tj in (ROOT / "profiles" / "r02_conv_traffic.json", ROOT / "profiles" / "r01_conv_traffic.json"):
        if tj.exists():      # dram bytes per launch of these kernels from one ncu capture of this workload
            t = json.loads(tj.read_text())
            traffic = t.get("dram_bytes_per_launch")
            if traffic is not None:   # captured at micro-batch 256, 224x224: activation traffic is linear in images per launch
                traffic = traffic * (mb / float(t.get("micro_batch", 256))) * (hw * hw / (224.0 * 224.0))
            traffic_note = (f"dram__bytes_read+write per launch averaged over all tcgen05 conv launches of one ncu capture "
                            f"({tj.name}), scaled to this run's launch size")
            break
    ips = images_per_step * steps / (total_ms / 1e3)
    gflop_img = gflop_per_image(arch, classify, hw)
    slowest = max(gathered, key=lambda g: g["compute_ms_per_step"])
    cfg = config_dict(args, world)
    cfg.update({"micro_batch": mb,
                "l2": "inputs (%.0f MB/step) larger than L2; no flush needed" % ((hi - lo if total_mode else B_) * hw * hw * 3 / 1e6),
                "parallelism": f"dp{world} (contiguous shard of the global image index range per rank, counts stay on the "
                               f"device, ONE int64[2] all-reduce at the end of the timed region)"})
    line = {
        "metric": "images/sec restore->VGG16 classify", "value": ips, "unit": "images/s", "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": total_ms / steps,
        "higher_is_better": True, "scaling": "strong" if total_mode else "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": cfg,
        "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples", "power_w")},
        "cuda_graph": use_graph,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convs (conv_gemm_pair_kernel<256>, conv_w3_kernel<pair>, conv_gemm_pairhalo_kernel<128>: all cta_group::2; conv_gemm_kernel<N>)",
                     "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf,
                     "peak_source": peak_src, "traffic": traffic, "traffic_note": traffic_note,
                     "launches": ksum["launches"], "kernel_ms_per_step": ksum["ms"] / inst_steps,
                     "instrumented": {"steps": inst_steps, "ms_per_step": inst_ms / inst_steps, "sm_mhz": clocks_inst.get("sm_mhz"),
                                      "host_launch_s_per_step": inst_host_s / inst_steps,
                                      "note": "eager steps with one CUDA-event pair per conv launch, run right after the timed "
                                              "region; `value` itself replays CUDA graphs" if use_graph else
                                              "same eager path as the timed region"},
                     "dominant_kernel": max(ksum.get("by_kernel", {"?": {"ms": 0}}).items(), key=lambda kv: kv[1]["ms"])[0],
                     "by_kernel": {k: {"launches": v["launches"], "share_of_step": v["ms"] / inst_ms,
                                       "achieved": v["work"] / (v["ms"] * 1e-3) / 1e12,
                                       "frac": v["work"] / (v["ms"] * 1e-3) / 1e12 / peak_tf}
                                   for k, v in sorted(ksum.get("by_kernel", {}).items(), key=lambda kv: -kv[1]["ms"])},
                     "share_of_step": ksum["ms"] / inst_ms,
                     "pipeline_tflops": ips / world * gflop_img / 1e3 if gflop_img else None},
        "roofline_hbm": hbm,
        "per_rank": {"ms_per_step": minmedmax([g["ms_per_step"] for g in gathered]),
                     "compute_ms_per_step": minmedmax([g["compute_ms_per_step"] for g in gathered]),
                     "conv_kernel_ms_per_step": minmedmax([g["conv_kernel_ms_per_step"] for g in gathered]),
                     "host_launch_s_per_step": minmedmax([g["host_launch_s_per_step"] for g in gathered]),
                     "sm_mhz": minmedmax([g["sm_mhz"] for g in gathered]),
                     "power_w": minmedmax([g["power_w"] for g in gathered]),
                     "slowest": slowest, "cores_per_rank": mine["cores"],
                     "note": "ms_per_step includes the single all-reduce at the end (every rank then waits for the slowest); "
                             "compute_ms_per_step is each rank's own work before it"},
        "counts": [int(counts[0]), int(counts[1])],
    }
    if total_mode:
        line["seconds_for_total"] = total_ms / 1e3 / steps
        assert int(counts[1]) == args.total_images * steps, (int(counts[1]), args.total_images)
    if world == 1 and not total_mode and classify and cascade is None and not args.no_comparator:
        try:
            comp = {}
            for prec in ("f32", "bf16"):
                v, ms = cuda_reference_images_per_s(args.workload, hw, 256, 64, 2, 1, prec)
                comp[prec] = {"value": v, "unit": "images/s", "ms_per_256_images": ms}
            comp["note"] = ("the reference's own modules (baseline/_ref ResUNet, torchvision vgg16 + head swap) on cuda:0 through "
                            "ATen -> cuDNN / cuBLAS, cudnn.benchmark on, 256 resident images in micro-batches of 64, restore + "
                            "classify only (the reference's degradation is NumPy/OpenCV on the host); f32 = TF32 off, bf16 = "
                            "autocast + channels_last.  Same box, same process, after the timed legs")
            line["library_comparator"] = comp
        except Exception as e:   # baseline/_ref absent
            line["library_comparator"] = {"unavailable": str(e)[:200]}
    if world == 1 and not total_mode and not args.no_cpu_baseline:
        sample = args.cpu_sample or CPU_SAMPLE
        v, cores, dt, kind = cpu_reference_images_per_s(args.workload, hw, sample, repeats=1, warmup=1)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": kind,
                                "sample": f"{sample} images of {hw}x{hw} in batches of 32 (17:73), one timed pass after a warm-up batch "
                                          f"({dt:.1f} s), degradation calls spread over the cores; " +
                                          ("the unmodified reference scripts 16 -> 17 -> 18 from baseline/_ref" if kind == "reference"
                                           else "oracle port of " + ("script 13 (distortions, cascade, VGG confidence)"
                                                                     if recipe == "stress13" else "scripts 16 -> 17 -> 18")) +
                                          " in fp32 PyTorch"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def degrade_roofline(torch, D, dev_imgs, hw):
    """Algorithmic bytes (u8 in + u8 out = 6 B/pixel) / CUDA-event time of b2r_degrade per recipe."""
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6500.0))
    n = dev_imgs.shape[0]
    out = torch.empty_like(dev_imgs)
    recipes = {}

    def mk(blur, fog, noise):
        p = D.DegradeParams(n)
        if blur:
            p.set_blur_all(10, 45)
        if fog:
            p.set_fog_all(0.5)
        if noise:
            p.set_noise_all(0.02)
        return p.to(dev_imgs.device)

    for name, p in (("compound16 (blur 10/45 + fog + noise)", mk(1, 1, 1)), ("fog + noise", mk(0, 1, 1)), ("fog", mk(0, 1, 0))):
        for _ in range(2):
            D.degrade(dev_imgs, p, seed=2, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            D.degrade(dev_imgs, p, seed=2, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        gbs = n * hw * hw * 6 / (ms * 1e-3) / 1e9
        recipes[name] = {"us_per_image": ms * 1e3 / n, "achieved": gbs, "frac": gbs / peak}
    head = recipes["compound16 (blur 10/45 + fog + noise)"]
    return {"bound": "hbm", "kernel": "degrade_kernel (b2r_degrade)", "achieved": head["achieved"], "peak": peak, "unit": "GB/s",
            "frac": head["frac"], "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
            "bytes_per_image": hw * hw * 6, "by_recipe": recipes,
            "note": "blurred and noisy recipes are instruction-issue-bound (bit-exact OpenCV tap order, Philox-10 + Box-Muller); "
                    "the fog-only recipe runs at the HBM rate"}


def main():
    args = parse_args()
    # stdout carries exactly ONE line, the JSON record: anything a library writes to file descriptor 1 (NCCL prints its version
    # banner there on rank 0) is sent to stderr; Python-level print() keeps the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
