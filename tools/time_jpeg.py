"""JPEG decode on the GPU (csrc/jpeg_decode.cu) at sweep scale: frames/s and bytes/s for one call over N frame files of the
data set's geometry (340 x 256, 4:2:0, quality 90 -- what ffmpeg frame dumps look like), against PIL on one host core.
Algorithmic bytes per frame: the file in + H*W*3 decoded bytes out."""
import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
from clip_spm_b200 import ops
from tests import jpeg_cases as J

N = int(sys.argv[1]) if len(sys.argv) > 1 else 960
files = [J.encode(256, 340, "smooth" if i % 2 else "noise", 90, 2, seed=i) for i in range(16)]
files = [files[i % 16] for i in range(N)]
out = ops.decode_jpegs(files)            # warm-up (module load, buffers)
ref = np.asarray(Image.open(io.BytesIO(files[0])).convert("RGB"))
assert (out[0].cpu().numpy() == ref).all()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); out = ops.decode_jpegs(files); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t = min(ts)
in_b, out_b = sum(len(f) for f in files), N * 256 * 340 * 3
t0 = time.perf_counter()
for f in files[:64]:
    Image.open(io.BytesIO(f)).convert("RGB").load()
tp = (time.perf_counter() - t0) / 64
print("jpeg decode, %d frames 340x256 4:2:0 q90 (%.1f KB/file): %.2f ms per call incl. host parse = %.0f frames/s, %.2f GB/s "
      "decoded bytes (+ %.2f GB/s file bytes); PIL on one host core: %.0f frames/s -> %.0fx"
      % (N, in_b / N / 1e3, t * 1e3, N / t, out_b / t / 1e9, in_b / t / 1e9, 1 / tp, (N / t) * tp))
