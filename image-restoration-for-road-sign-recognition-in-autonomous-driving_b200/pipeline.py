"""degrade -> restore -> classify -> count, on-device, one process per GPU.

The reference runs this path as three scripts joined by PNG files on disk (16_gen_compound_data.py ->
17_run_unified_inference.py -> 18_test_unified_benchmark.py); the only in-memory composition is the single-image
demo 15_test_unified.py:170-200, which fixes the semantics kept here:

    u8 image --degrade--> u8 --ToTensor--> restorer --clamp(0,1)*255, truncate--> u8 --ToTensor+Normalize--> VGG16
       --torch.max(outputs, 1)--> predicted;  correct += (predicted == labels).sum()      (18:47-49)

Every arrow is a libb2r.so kernel; images never leave HBM between stages.  Multi-GPU: images are independent, so
each rank takes a contiguous block of the global index range, noise is keyed by the global image index (results
do not depend on the world size) and the only collective is one all-reduce of the int64 (correct, total) pair.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L
from . import degrade as D
from . import ops


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the global image index range owned by `rank`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_counts(counts: torch.Tensor) -> torch.Tensor:
    """Sum the int64 (correct, total) pair over all ranks: the single collective of the path (NCCL on GPUs)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


class RestoreClassifyPipeline:
    """Holds a restorer (SimpleUNet / ResUNet) and the VGG16 judge on one device and streams batches through them."""

    def __init__(self, restorer, judge, micro_batch: int = 256, use_graph: bool = False):
        self.restorer = restorer.eval()
        self.judge = judge.eval()
        self.micro_batch = int(micro_batch)
        # use_graph: restore -> classify -> count of a micro-batch (everything after the degradation launch, ~200 kernel
        # launches and as many host-side TMA descriptor encodings) is captured once per batch shape into a CUDA graph
        # and replayed; the degradation stays a plain launch because its Philox counter base is a by-value argument.
        self.use_graph = bool(use_graph)
        self._graphs = {}
        self.graph_launches_replayed = 0      # libb2r kernel launches executed through graph replays (launch accounting)
        self.device = next(judge.parameters()).device
        if self.device.type != "cuda":
            raise L.B2RError("RestoreClassifyPipeline needs its modules on a CUDA device (no CPU fallback)")
        self._div = 1 if type(restorer).__name__ == "ResUNet" else 4   # ResUNet re-aligns odd sizes (14:169-183)
        self._bufs = {}

    def _buf(self, name, shape, dtype):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    @torch.no_grad()
    def run_micro_batch(self, clean_u8: torch.Tensor, labels: Optional[torch.Tensor], params, seed: int,
                        image_index0: int, counts: Optional[torch.Tensor], noise: Optional[torch.Tensor] = None,
                        keep: bool = False, packs=None):
        """One resident micro-batch through all stages.  Returns (pred int64 [n], extras dict when keep=True).
        `packs` = (restorer._packed(), judge._packed()) when the caller already looked them up (run / run_from_host do it
        once per call instead of walking both state_dicts for every micro-batch)."""
        n, h, w, _ = clean_u8.shape
        with torch.cuda.device(self.device):
            if self.use_graph and not keep and noise is None:
                return self._run_micro_batch_graph(clean_u8, labels, params, seed, image_index0, counts, packs), None
            degraded = D.degrade(clean_u8, params, seed=seed, image_index0=image_index0, noise=noise,
                                 out=self._buf("deg", (n, h, w, 3), torch.uint8)) if params is not None else clean_u8
            restored = self._buf("rest", (n, h, w, 3), torch.uint8)
            pred, logits = self._restore_classify(degraded, restored, labels, counts, want_logits=True, packs=packs)
            if keep:
                return pred, {"degraded": degraded.clone(), "restored": restored.clone(), "logits": logits.clone()}
        return pred, None

    # -- CUDA-graph mode --------------------------------------------------------------------------------------------
    def _restore_classify(self, degraded, restored, labels, counts, want_logits=False, packs=None):
        """restore -> clamp/u8 -> classify -> top-1 (+ counts).  Without labels only `total` can be counted: the kernel
        takes counts together with labels, so a predictions-only call adds n to counts[1] itself."""
        pr, pj = packs if packs is not None else (None, None)
        self.restorer._check_input(degraded, self._div)
        self.restorer._run(degraded, None, restored, pr)
        self.judge._check_input(restored, 32)
        logits = self.judge._run(restored, True, pj)
        pred = ops.argmax_count(logits, labels, counts if labels is not None else None)[0]
        if labels is None and counts is not None:
            counts[1:2].add_(degraded.shape[0])
        return (pred, logits) if want_logits else pred

    def _graph_entry(self, n: int, h: int, w: int, with_labels: bool, packs=None):
        packs = packs if packs is not None else (self.restorer._packed(), self.judge._packed())
        key = (n, h, w, with_labels)
        ent = self._graphs.get(key)
        if ent is not None and ent["packs"][0] is packs[0] and ent["packs"][1] is packs[1]:
            return ent
        dev = self.device
        ent = {"packs": packs,                                             # re-packed weights invalidate the graph
               "deg": torch.zeros((n, h, w, 3), dtype=torch.uint8, device=dev),
               "rest": torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev),
               "labels": torch.zeros((n,), dtype=torch.int64, device=dev) if with_labels else None,
               "counts": torch.zeros((2,), dtype=torch.int64, device=dev)}
        self._restore_classify(ent["deg"], ent["rest"], ent["labels"], ent["counts"], packs=packs)   # eager once: workspaces, attributes
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        launches_before = ops.STATS["launches"]
        with torch.cuda.graph(graph):
            ent["pred"] = self._restore_classify(ent["deg"], ent["rest"], ent["labels"], ent["counts"], packs=packs)
        # the graph holds raw pointers into the modules' activation workspaces: keep those tensors alive even if the
        # workspaces later switch to another batch shape
        ent["pinned"] = (dict(self.restorer._ws._bufs), dict(self.judge._ws._bufs))
        ent["graph"] = graph
        ent["launches"] = ops.STATS["launches"] - launches_before
        self._graphs[key] = ent
        return ent

    def _run_micro_batch_graph(self, clean_u8, labels, params, seed, image_index0, counts, packs=None):
        n, h, w, _ = clean_u8.shape
        ent = self._graph_entry(n, h, w, labels is not None, packs)
        if params is not None:
            D.degrade(clean_u8, params, seed=seed, image_index0=image_index0, out=ent["deg"])
        else:
            ent["deg"].copy_(clean_u8)
        if labels is not None:
            ent["labels"].copy_(labels)
        ent["counts"].zero_()
        ent["graph"].replay()
        self.graph_launches_replayed += ent["launches"]
        if counts is not None:
            counts.add_(ent["counts"])
        return ent["pred"]          # static output of the graph: consume it before the next micro-batch of this shape

    @torch.no_grad()
    def run(self, clean_u8: torch.Tensor, labels: Optional[torch.Tensor], params, seed: int = 0, image_index0: int = 0):
        """Device-resident batch [N,H,W,3] u8 -> (pred int64 [N], counts int64 [2] = (correct, total)); labels = None
        gives predictions only (correct stays 0)."""
        n = clean_u8.shape[0]
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        preds = torch.empty(n, dtype=torch.int64, device=self.device)
        dparams = params.to(self.device) if isinstance(params, D.DegradeParams) else params
        packs = (self.restorer._packed(), self.judge._packed())
        for s in range(0, n, self.micro_batch):
            c = min(self.micro_batch, n - s)
            sub = _slice_params(dparams, s, c) if dparams is not None else None
            p, _ = self.run_micro_batch(clean_u8[s:s + c], None if labels is None else labels[s:s + c], sub, seed,
                                        image_index0 + s, counts, packs=packs)
            preds[s:s + c] = p
        return preds, counts

    @torch.no_grad()
    def run_from_host(self, clean_u8_pinned: torch.Tensor, labels_pinned: torch.Tensor, params, seed: int = 0,
                      image_index0: int = 0):
        """End-to-end call with HOST buffers: per micro-batch H2D copy of the images/labels from pinned memory on a
        copy stream (double-buffered, overlapped with compute), D2H read of the (correct, total) pair at the end.
        Returns (counts as a Python tuple, h2d_bytes, d2h_bytes)."""
        n, h, w, _ = clean_u8_pinned.shape
        mb = self.micro_batch
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        dparams = params.to(self.device) if isinstance(params, D.DegradeParams) else params
        copy_stream = self._bufs.get("_copy_stream") or torch.cuda.Stream(device=self.device)
        self._bufs["_copy_stream"] = copy_stream
        main = torch.cuda.current_stream(self.device)
        stage = [(self._buf(f"h2d_img{i}", (mb, h, w, 3), torch.uint8), self._buf(f"h2d_lab{i}", (mb,), torch.int64))
                 for i in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        h2d = 0
        packs = (self.restorer._packed(), self.judge._packed())
        chunks = [(s, min(mb, n - s)) for s in range(0, n, mb)]

        def issue(i):
            s, c = chunks[i]
            b = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(freed[b])
                stage[b][0][:c].copy_(clean_u8_pinned[s:s + c], non_blocking=True)
                stage[b][1][:c].copy_(labels_pinned[s:s + c], non_blocking=True)
                ready[b].record(copy_stream)

        if chunks:
            issue(0)
        for i, (s, c) in enumerate(chunks):
            if i + 1 < len(chunks):
                issue(i + 1)
            b = i & 1
            main.wait_event(ready[b])
            sub = _slice_params(dparams, s, c) if dparams is not None else None
            self.run_micro_batch(stage[b][0][:c], stage[b][1][:c], sub, seed, image_index0 + s, counts, packs=packs)
            freed[b].record(main)
            h2d += c * h * w * 3 + c * 8
        host_counts = counts.cpu()           # the device -> host read of the step's result (synchronises)
        return (int(host_counts[0]), int(host_counts[1])), h2d, host_counts.numel() * 8


def _slice_params(p: "D.DeviceDegradeParams", s: int, c: int) -> "D.DeviceDegradeParams":
    if s == 0 and c == p.n:
        return p
    return D.DeviceDegradeParams(c, p.order, p.flags, p.ksize[s:s + c], p.taps[s:s + c], p.fog_on[s:s + c],
                                 p.fog_t[s:s + c], p.fog_add[s:s + c], p.sigma[s:s + c], p.any_blur, p.any_noise)
