"""Per-shape GEMM / convolution times of the RN50 tower in situ (BASELINE config 4 shape, 8 episodes = 1280 frames per call):
CUDA-event pairs around every GEMM launch of one call (spm_profile_begin / _end with SPM_PROFILE_SHAPES=1 -> stderr)."""
import ctypes, os, sys
os.environ["SPM_PROFILE_SHAPES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import CNN, sweep, _lib
from clip_spm_b200.config import make_cfg
E = 8
net = CNN(make_cfg("RN50", 8, False, 5), max_episodes=E)
net.init_random_(0); net.text_features_test = torch.randn(10, 1024)
b = sweep.synthetic_episode_batch(list(range(E)), 5, 3, 1, 8, 10, "cuda")
f = lambda: net.forward_episodes(b["context_images"], b["context_labels"], b["target_images"], b["real_support_labels"], b["real_target_labels"], E, b["target_labels"])
for _ in range(3): f()
torch.cuda.synchronize()
lib = _lib.load()
_lib.check(lib.spm_profile_begin(4096))
f()
fl, ms, cnt = (ctypes.c_double * 4)(), (ctypes.c_double * 4)(), (ctypes.c_int * 4)()
_lib.check(lib.spm_profile_end(fl, ms, cnt))
print("GEMM launches %d, %.2f ms, %.0f TFLOP/s over the GEMM time" % (sum(cnt), sum(ms), sum(fl) / 1e9 / max(sum(ms), 1e-9)))
