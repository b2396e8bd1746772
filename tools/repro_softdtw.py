import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import _lib
lib = _lib.load()
B, N, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
D = torch.rand(B, N, M, device="cuda")
R = torch.empty(B, N + 2, M + 2, device="cuda"); out = torch.empty(B, device="cuda"); E = torch.empty(B, N, M, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
_lib.check(lib.spm_softdtw_forward(st, B, N, M, p(D), 0.5, 0.0, p(R), p(out)))
torch.cuda.synchronize(); print("forward ok", float(out[0]), float(out[-1]))
_lib.check(lib.spm_softdtw_backward(st, B, N, M, p(D), p(R), 0.5, 0.0, p(E)))
torch.cuda.synchronize(); print("backward ok", float(E[0, 0, 0]), float(E[-1, -1, -1]))
