"""216 frames (one back chunk = three front sub-chunks) through the RN50 tower a few times: the command profiled under ncu
for the convolution kernels (launch order per pass: 3 x [stem conv1-3, layer1 (10 GEMMs), layer2 (14)], then layer3 (20),
layer4 (11), attention pool (3))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import CNN
from clip_spm_b200.config import make_cfg
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 216
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net = CNN(make_cfg("RN50", 8, False, 5), max_episodes=1)
net.init_random_(0)
imgs = torch.rand(frames, 3, 224, 224, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    if i == reps - 1:
        e0.record()
    out = net.encode_frames(imgs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("RN50 tower, %d frames: %.2f ms, %.0f frames/s, %.0f TFLOP/s (11.59 GFLOP/frame)" % (frames, ms, frames / ms * 1e3, frames * 11.59 / ms))
