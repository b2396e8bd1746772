"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the oracle port on the host
cores) prints one JSON line with the keys the driver reads, non-zero ranks of a multi-rank launch stay silent, and
the CUDA arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-frames", "2", "--no-full-episode"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "episodes_per_sec" and d["unit"] == "episodes/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # both arms describe the same workload with the same `config` dict
    import bench
    assert d["config"] == bench.bench_config(16, 8)


def test_eager_baseline_tower_is_the_reference_tower():
    """tools/eager_baseline.py (the torch.nn restatement bench.py times as the existing-library bar) reproduces the
    executed reference's frame features on the golden case -- same modules, same state_dict keys."""
    import importlib.util
    from tests import helpers as H
    spec = importlib.util.spec_from_file_location("eb", os.path.join(ROOT, "tools", "eager_baseline.py"))
    eb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(eb)
    ci, g = H.case_inputs("vit_2w1s_t2_p0"), H.golden("vit_2w1s_t2_p0")
    tower = eb.build_from(ci["weights"], "cpu")
    with torch.no_grad():
        su = tower(ci["episode"]["context_images"])
    assert H.rel_err(su.view(g["su"].shape), g["su"]) < 1e-4


def test_reference_arm_only_rank0_works():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-frames", "2", "--no-full-episode"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_cuda_arm_refuses_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
