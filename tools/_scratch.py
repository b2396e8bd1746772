import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops
M, N, K = 215296, 64, 256
a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(N, device="cuda")
for _ in range(4): ops.gemm(a, b, bias=bias, act="relu", out=out)
torch.cuda.synchronize()
print("ok")
