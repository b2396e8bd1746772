"""Micro-benchmark of the two attention kernels (CUDA events, default stream and a side stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops

F = 240
qkv = torch.randn(F * 197, 2304, device="cuda").bfloat16()
for impl in ("tcgen05", "mma"):
    for name, stream in (("default", torch.cuda.default_stream()), ("side", torch.cuda.Stream())):
        with torch.cuda.stream(stream):
            for _ in range(3):
                ops.vit_attention(qkv, F, impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.vit_attention(qkv, F, impl)
            e1.record()
            torch.cuda.synchronize()
            print(impl, name, "%.1f us per call" % (1e3 * e0.elapsed_time(e1) / 20))
