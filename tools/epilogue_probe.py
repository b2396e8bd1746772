import os, sys
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/clip_spm_b200") else os.getcwd())
import torch
from clip_spm_b200 import ops, _lib
import ctypes
M = 47280
def t(N, K, bias=False, reps=30):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    bi = torch.randn(N, device="cuda") if bias else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, b, bias=bi, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): ops.gemm(a, b, bias=bi, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("dbg=%s  N=2304 K=64: %.1f us | K=64+bias: %.1f us | K=768+bias: %.1f us" % (os.environ.get("SPM_GEMM_DBG", "0"), t(2304, 64), t(2304, 64, True), t(2304, 768, True)))
