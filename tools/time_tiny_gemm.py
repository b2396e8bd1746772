import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops
def t(M, N, K, dtype=torch.bfloat16, reps=200):
    a = torch.randn(M, K, device="cuda").to(dtype); b = torch.randn(N, K, device="cuda").to(dtype)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if dtype == torch.bfloat16 else torch.float32)
    for _ in range(5): ops.gemm(a, b, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): ops.gemm(a, b, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (M, N, K) in ((128, 128, 64), (128, 128, 2304), (4096, 256, 2304), (128, 256, 64), (18944, 128, 64), (128*148, 128, 64), (128*148*4, 128, 64)):
    print("bf16 M=%6d N=%4d K=%5d: %7.2f us per launch (back-to-back launches, includes launch overhead)" % (M, N, K, t(M, N, K)))
print("tf32 M=540 N=6144 K=512: %.2f us" % t(540, 6144, 512, torch.float32))
print("tf32 M=540 N=512 K=2048: %.2f us" % t(540, 512, 2048, torch.float32))
# reference point: an empty-ish torch kernel launch
x = torch.zeros(1024, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(200): x.add_(1.0)
e1.record(); torch.cuda.synchronize()
print("torch tiny elementwise kernel: %.2f us per launch" % (e0.elapsed_time(e1) / 200 * 1e3))
