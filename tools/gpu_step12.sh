#!/bin/bash
O=gpurun_out
echo "== tests rn50 + bf16_resid + gemm"; timeout 900 python -m pytest tests -m gpu -x -q -s -k "rn50 or residual_stream or gemm" 2>&1 | tail -n 14
R="python tools/profile_rn50.py 320 4"
timeout 300 $R > $O/r02_s12_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 85 --csv --log-file $O/r02_launches_rn50_s12.csv $R > $O/r02_s12_ncu.log 2>&1
cat $O/r02_s12_plain.log
python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from clip_spm_b200 import CNN, sweep
from clip_spm_b200.config import make_cfg
for prec in ("bf16", "bf16_resid"):
    net = CNN(make_cfg("ViT-B/16", 8, False, 5), max_episodes=8, precision=prec); net.init_random_(0)
    net.text_features_test = torch.randn(24, 512)
    b = sweep.synthetic_episode_batch(list(range(8)), 5, 5, 1, 8, 24, "cuda")
    f = lambda: net.forward_episodes(b["context_images"], b["context_labels"], b["target_images"], b["real_support_labels"], b["real_target_labels"], 8, b["target_labels"])
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print("ViT 5w5s %s: %.1f episodes/s" % (prec, 80 / (e0.elapsed_time(e1) / 1e3)))
    del net
PY
