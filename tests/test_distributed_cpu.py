"""N > 1 host logic on CPU: two gloo ranks exercise the shard arithmetic and the single collective of the path
(all-reduce of the int64 (correct, total) pair), as bench.py / RestoreClassifyPipeline use them under NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200restore import all_reduce_counts, shard_range, synth
    lo, hi = shard_range(total, rank, world)
    # labels are a pure function of the GLOBAL image index, so every rank can generate its own shard
    _, labels = synth.sign_like_images(hi - lo, 8, 8, seed=1, index0=lo)
    pred = (torch.arange(lo, hi) * 7) % 43                      # stand-in predictions, also keyed by global index
    counts = torch.tensor([int((pred == labels).sum()), hi - lo], dtype=torch.int64)
    all_reduce_counts(counts)
    out_q.put((rank, lo, hi, counts.tolist()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_count_allreduce_matches_single_process():
    world, total = 2, 1001                                      # odd total: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    from b200restore import synth
    _, labels = synth.sign_like_images(total, 8, 8, seed=1, index0=0)
    pred = (torch.arange(total) * 7) % 43
    expect = [int((pred == labels).sum()), total]
    assert [r[3] for r in res] == [expect, expect]              # both ranks hold the global (correct, total)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total   # contiguous, disjoint, complete


def test_shard_range_properties():
    from b200restore import shard_range
    for total in (0, 1, 7, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_all_reduce_counts_is_identity_without_process_group():
    from b200restore import all_reduce_counts
    c = torch.tensor([3, 10], dtype=torch.int64)
    assert all_reduce_counts(c).tolist() == [3, 10]
