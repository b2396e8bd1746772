"""Per-tile cost model of the tcgen05 GEMM: time vs K at fixed M, N (CUDA events, bf16 out)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops

def t(M, N, K, reps=20, **kw):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, b, out=out, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): ops.gemm(a, b, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

M = 47280
for N in (256, 768, 2304):
    for K in (64, 128, 256, 512, 768, 1536, 3072):
        us = t(M, N, K)
        tiles = ((M + 127) // 128) * ((N + 255) // 256)
        per_cta = (tiles + 147) // 148
        print("M=%d N=%4d K=%4d: %7.1f us  %6.0f TFLOP/s  tiles/CTA %3d  us/tile %.2f  (MMA floor %.2f us/tile @1.9GHz)" %
              (M, N, K, us, 2.0 * M * N * K / us / 1e6, per_cta, us / per_cta, (K / 64) * 512 / 1.9e3))
for (M2, N2, K2) in ((802816, 32, 32), (802816, 32, 288), (802816, 64, 288), (200704, 64, 64), (200704, 64, 576), (200704, 256, 64)):
    us = t(M2, N2, K2, reps=5)
    print("M=%d N=%d K=%d: %.1f us, %.2f TB/s of A+out" % (M2, N2, K2, us, (M2 * K2 * 2 + M2 * N2 * 2) / us / 1e6))
