"""(e) multi-GPU on hardware: a sharded sweep must predict exactly what the unsharded sweep predicts.

* one GPU: ranks 0 and 1 of a world of 2 are evaluated one after the other on the same device (no process group) and
  their shards are merged on the host -- checks the sharding + per-episode determinism on the CUDA path;
* two or more GPUs (skipped otherwise): tools/run_sweep.py --check-sharding under torchrun with 2 ranks over NCCL:
  predictions all-gathered (sweep.gather_predictions) == the unsharded run, reduced statistics equal."""
import json
import os
import subprocess
import sys

import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_episode_sets_equal_unsharded_on_one_gpu():
    from clip_spm_b200 import sweep
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci, max_episodes=2)
    n = 7
    full, table, logits = sweep.run_sweep(net, n, 2, 1, 1, 24, 0, 1, 2, True)
    merged, stats = {}, []
    tables = []
    for r in range(2):
        _, t, lg = sweep.run_sweep(net, n, 2, 1, 1, 24, r, 2, 2, True)
        merged.update(lg)
        tables.append(t)
    assert sorted(merged) == list(range(n))
    both = torch.maximum(tables[0], tables[1])
    assert int((both < 0).sum()) == 0
    bit_equal = 0
    for e in range(n):
        # the shards group episodes into calls differently (0,2 | 4,6 vs 0,1 | 2,3 ...): same logits up to fp32 rounding,
        # same prediction wherever the top-1/top-2 margin is not itself at rounding level
        assert torch.allclose(merged[e], logits[e], atol=1e-4, rtol=1e-4), e
        bit_equal += int(torch.equal(merged[e], logits[e]))
        top2 = logits[e].topk(2, dim=-1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-3
        assert torch.equal(both[e][safe], table[e][safe])
    print("\n%d/%d episodes bit-identical across shardings" % (bit_equal, n))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_two_rank_nccl_sweep_equals_unsharded():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tools", "run_sweep.py"), "--episodes", "24",
           "--shot", "1", "--episodes-per-call", "4", "--check-sharding"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{") and "check_sharding" in l]
    assert lines and lines[0]["check_sharding"] == "ok" and lines[0]["predictions_equal"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_two_rank_nccl_training_equals_single_process():
    """tasks of an optimiser step split over 2 ranks + one gradient all-reduce == the same steps by one process"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tools", "run_train.py"), "--steps", "2",
           "--tasks-per-batch", "4", "--seq-len", "2", "--way", "2", "--check"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{") and "check" in l]
    assert lines and lines[0]["check"] == "ok", lines
