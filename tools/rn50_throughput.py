import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import CNN, sweep
from clip_spm_b200.config import make_cfg
net = CNN(make_cfg("RN50", 8, False, 5), max_episodes=2)
net.init_random_(0); net.text_features_test = torch.randn(10, 1024)
b = sweep.synthetic_episode_batch([0, 1], 5, 3, 1, 8, 10, "cuda")   # BASELINE config 4 shape: 5-way 3-shot, 160 frames
f = lambda: net.forward_episodes(b["context_images"], b["context_labels"], b["target_images"], b["real_support_labels"], b["real_target_labels"], 2, b["target_labels"])
for _ in range(3): f()
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("RN50 5-way 3-shot (160 frames): %.2f ms/episode, %.1f episodes/s, %.0f frames/s, %.0f TFLOP/s (11.59 GFLOP/frame)" % (ms, 1e3/ms, 160e3/ms, 160*11.59/ms))
