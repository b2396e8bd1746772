"""One frame chunk (default 512 frames, the bench's chunk) through the ViT-B/16 tower a few times: the command profiled
under `ncu --set full` for the encoder kernels (gemm2_tcgen05_kernel flavours, vit_attention_tc_kernel, layernorm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import CNN
from clip_spm_b200.config import make_cfg
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net = CNN(make_cfg("ViT-B/16", 8, False, 5), max_episodes=1)
net.init_random_(0)
imgs = torch.rand(frames, 3, 224, 224, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    if i == reps - 1:
        e0.record()
    out = net.encode_frames(imgs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("ViT-B/16 tower, %d frames: %.2f ms, %.0f frames/s, %.0f TFLOP/s executed" % (frames, ms, frames / ms * 1e3, frames * 33.046 / ms))
