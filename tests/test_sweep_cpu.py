"""Host-side sweep logic (clip_spm_b200/sweep.py) on CPU: episode sharding, sufficient statistics and the path's one
collective, exercised with world_size 2 over gloo (the GPU path uses the identical code over NCCL)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clip_spm_b200 import sweep


def test_shards_partition_the_episode_range():
    for n, w in [(10, 1), (10, 2), (10000, 8), (7, 4), (3, 8)]:
        shards = [sweep.shard_episodes(n, r, w) for r in range(w)]
        flat = sorted(e for s in shards for e in s)
        assert flat == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_summary_matches_numpy_formulas_of_the_reference():
    """run/main_run.py:286-289: accuracy = mean*100, confidence = 196*std/sqrt(n) with np.std (population), loss = mean"""
    g = torch.Generator().manual_seed(0)
    acc = (torch.rand(1000, generator=g) * 5).floor() / 5
    loss = torch.rand(1000, generator=g)
    s = sweep.summarize(sweep.make_stats(acc, loss))
    a = acc.double().numpy()
    assert abs(s["accuracy"] - a.mean() * 100) < 1e-9
    assert abs(s["confidence"] - 196.0 * a.std() / math.sqrt(len(a))) < 1e-9
    assert abs(s["loss"] - loss.double().mean().item()) < 1e-12 and s["n"] == 1000


def test_synthetic_batches_are_sharding_invariant():
    """episode content depends on the GLOBAL episode id only"""
    a = sweep.synthetic_episode_batch([5], 2, 1, 1, 2, 24, "cpu")
    b = sweep.synthetic_episode_batch([3, 5], 2, 1, 1, 2, 24, "cpu")
    n = a["context_images"].shape[0]
    assert torch.equal(a["context_images"], b["context_images"][n:])
    assert torch.equal(a["context_labels"][0], b["context_labels"][1])
    assert a["context_labels"].dtype == torch.float32 and a["target_labels"].dtype == torch.int64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # per-episode results are a deterministic function of the global episode id (stands in for the CUDA forward)
    ids = sweep.shard_episodes(n, rank, world)
    acc = torch.tensor([((7 * e) % 6) / 5.0 for e in ids], dtype=torch.float32)
    loss = torch.tensor([1.0 + 0.01 * e for e in ids], dtype=torch.float32)
    stats = sweep.reduce_stats(sweep.make_stats(acc, loss))
    out[rank] = sweep.summarize(stats)
    # the optional prediction gather: every rank ends up with the full table in global episode order
    pred = torch.tensor([[(e + q) % 5 for q in range(3)] for e in ids], dtype=torch.int32).view(len(ids), 3)
    out[("pred", rank)] = sweep.gather_predictions(pred, ids, n).tolist()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 37])
def test_two_rank_sweep_equals_unsharded(n):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    acc = torch.tensor([((7 * e) % 6) / 5.0 for e in range(n)], dtype=torch.float32)
    loss = torch.tensor([1.0 + 0.01 * e for e in range(n)], dtype=torch.float32)
    ref = sweep.summarize(sweep.make_stats(acc, loss))
    for r in range(world):
        assert out[r]["n"] == n
        for k in ("accuracy", "confidence", "loss"):
            assert abs(out[r][k] - ref[k]) < 1e-9, (r, k)
        assert out[("pred", r)] == [[(e + q) % 5 for q in range(3)] for e in range(n)]
