"""N ranks produce the same global (correct, total) and the same predictions as 1 rank over the same global image index
range (18_test_unified_benchmark.py:48-51 accumulates correct / total over the whole dataset).

* in-process: the shards of world sizes 1, 2, 3 and 8 are run one after another on this GPU (same code path a rank runs:
  shard_range -> indexed images -> degrade keyed by the global index -> restore -> classify -> count);
* NCCL: when >= 2 GPUs are visible, `torchrun --nproc-per-node 2` (and 4 / 8 when available) runs tools/world_invariance.py
  with one rank per GPU and the result is compared with the single-process run."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_shards_of_any_world_size_give_the_single_rank_result():
    import b200restore as B
    from b200restore import degrade as D, models, synth
    dev = torch.device("cuda")
    r, j = models.ResUNet(), models.VGG16Judge()
    r.load_state_dict(synth.synthetic_state_dict("resunet", 31))
    j.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
    pipe = B.RestoreClassifyPipeline(r.to(dev), j.to(dev), micro_batch=64)
    total, hw = 203, 64                                   # odd total: ragged shards

    def run_world(world):
        preds, counts = [], torch.zeros(2, dtype=torch.int64, device=dev)
        for rank in range(world):
            lo, hi = B.shard_range(total, rank, world)
            imgs, labels = synth.indexed_images(lo, hi - lo, hw, hw, seed=7)
            p, c = pipe.run(imgs.to(dev), labels.to(dev), D.compound_params(hi - lo), seed=2, image_index0=lo)
            preds.append(p)
            counts += c                                   # what the all-reduce computes
        return torch.cat(preds), counts

    p1, c1 = run_world(1)
    assert int(c1[1]) == total
    for world in (2, 3, 8):
        p, c = run_world(world)
        assert torch.equal(p, p1) and torch.equal(c, c1), f"world size {world} changes the result"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
def test_nccl_ranks_match_single_process(tmp_path):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 visible GPUs (run through `gpurun --gpus 2`)")
    script = str(ROOT / "tools" / "world_invariance.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    one = tmp_path / "w1.json"
    subprocess.run([sys.executable, script, "--total", "1001", "--out", str(one)], check=True, env=env, timeout=600)
    ref = json.loads(one.read_text())
    assert ref["counts"][1] == 1001
    for world in [w for w in (2, 4, 8) if w <= ngpu]:
        out = tmp_path / f"w{world}.json"
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script, "--total", "1001",
                        "--out", str(out)], check=True, env=env, timeout=600)
        got = json.loads(out.read_text())
        assert got["world"] == world
        assert got["counts"] == ref["counts"] and got["pred_sha256"] == ref["pred_sha256"] and got["n"] == ref["n"]
