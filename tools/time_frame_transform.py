"""Throughput of the frame transform kernel against the HBM roofline.
Algorithmic bytes per frame: H*W*3 read (only the rows/columns the crop needs: counted in full here, conservative for
the crop-only case where 224/340 of each row is used) + 224*224*3*4 written (fp32 CHW) or *2 (bf16 patch matrix)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import _lib, ops

peak = 6550.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (H, W, F) in ((256, 340, 1920), (240, 320, 1920), (720, 1280, 480), (1080, 1920, 240)):
    frames = torch.randint(0, 256, (F, H, W, 3), dtype=torch.uint8, device="cuda")
    oh, ow, y1, x1 = ops.frame_geometry(H, W)
    # input bytes actually needed: the crop window mapped back to the input
    need = (224 * H / oh) * (224 * W / ow) * 3
    ms = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ops.transform_frames(frames); e1.record(); torch.cuda.synchronize()
        if i >= 3: ms.append(e0.elapsed_time(e1))
    t = sum(ms) / len(ms)
    byts = F * (need + 224 * 224 * 3 * 4)
    print("transform %dx%d -> fp32 CHW, %d frames: %.1f us, %.0f frames/s, %.0f GB/s algorithmic (crop window in + out) = %.2f of HBM peak %.0f"
          % (W, H, F, t * 1e3, F / t * 1e3, byts / t / 1e6, byts / t / 1e6 / peak, peak))
    del out, frames
