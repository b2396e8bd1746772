"""RN50-like GEMM shapes (1x1 convolutions: huge M, small N and K) against their HBM floors (sustained loop)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops

def t(M, N, K, reps=10, res=False):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    for _ in range(2): ops.gemm(a, b, bias=bias, act="relu", out=out)
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, b, bias=bias, act="relu", out=out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2] * 1e3

for (M, N, K) in ((215296, 64, 256), (215296, 64, 64), (215296, 256, 64), (215296, 64, 576), (57600, 128, 512), (57600, 512, 128),
                  (57600, 128, 1152), (16384, 256, 1024), (16384, 1024, 256), (16384, 256, 2304), (5184, 512, 2048), (5184, 2048, 512), (5184, 512, 4608)):
    us = t(M, N, K)
    byts = M * K * 2 + M * N * 2 + N * K * 2
    print("M=%6d N=%4d K=%4d: %7.1f us  %6.0f TFLOP/s  %5.2f TB/s of A+out (HBM floor %.1f us)  tiles/CTA %.1f" %
          (M, N, K, us, 2.0 * M * N * K / us / 1e6, byts / us / 1e6, byts / 6550.7e3, ((M + 127) // 128) * ((N + 127) // 128) / 148.0))
