"""Out-projection shape (K = N = 768, fp32 residual stream read + written in the epilogue) against its HBM floor:
which part of the epilogue costs what.  Sustained loop, operands > L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops

def t(M, N, K, reps=20, res=False, out_dtype=torch.float32, bias=True):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=out_dtype)
    bv = torch.randn(N, device="cuda") if bias else None
    kw = dict(bias=bv, out=out)
    if res: kw["residual"] = out
    for _ in range(3): ops.gemm(a, b, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): ops.gemm(a, b, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

M = 100864
for (N, K) in ((768, 768), (768, 3072)):
    f = 2.0 * M * N * K
    for name, kw, byts in (("fp32 out + fp32 residual (in place)", dict(res=True), M * K * 2 + 2 * M * N * 4),
                           ("fp32 out, no residual", dict(), M * K * 2 + M * N * 4),
                           ("bf16 out, no residual", dict(out_dtype=torch.bfloat16), M * K * 2 + M * N * 2)):
        us = t(M, N, K, **kw)
        print("M=%d N=%d K=%d %-38s %7.1f us  %6.0f TFLOP/s  HBM floor %.1f us (%.0f MB)" %
              (M, N, K, name, us, f / us / 1e6, byts / 6550.7e3, byts / 1e6))
