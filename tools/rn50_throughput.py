"""RN50 backbone, BASELINE config 4 shape (5-way 3-shot, T=8: 160 frames / episode): episodes/s of E episodes per call.
    python tools/rn50_throughput.py [episodes_per_call=8] [calls=6]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import CNN, sweep
from clip_spm_b200.config import make_cfg
E = int(sys.argv[1]) if len(sys.argv) > 1 else 8
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
net = CNN(make_cfg("RN50", 8, False, 5), max_episodes=E)
net.init_random_(0); net.text_features_test = torch.randn(10, 1024)
b = sweep.synthetic_episode_batch(list(range(E)), 5, 3, 1, 8, 10, "cuda")
f = lambda: net.forward_episodes(b["context_images"], b["context_labels"], b["target_images"], b["real_support_labels"], b["real_target_labels"], E, b["target_labels"])
for _ in range(3): out = f()
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(calls): out = f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (calls * E)
print("RN50 5-way 3-shot (160 frames), %d episodes per call [%s]: %.2f ms/episode, %.1f episodes/s, %.0f frames/s, %.0f TFLOP/s (11.59 GFLOP/frame), finite=%s"
      % (E, " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("SPM_")), ms, 1e3 / ms, 160e3 / ms, 160 * 11.59 / ms, bool(torch.isfinite(out["logits"]).all())))
