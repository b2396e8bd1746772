"""Batch-scale timing of the metric-tail kernel (cos_sim + bidirectional OTAM) against the HBM roofline.
Algorithmic bytes per otam_distance problem: (Q+W)*T*D*4 read + Q*W*4 written."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops

peak = 6550.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
shapes = ((1000, 5, 5, 8, 512), (4000, 5, 5, 8, 512), (1000, 5, 5, 16, 512), (1000, 5, 5, 8, 1024))
if "--one" in sys.argv:   # a single shape: the command profiled under ncu
    shapes = shapes[:1]
if "--p4000" in sys.argv:
    shapes = shapes[1:2]
if "--p444" in sys.argv:      # 3 problems per SM, 73 MB of operands: with --noflush they are served from the 126 MB L2
    shapes = ((444, 5, 5, 8, 512),)
NOFLUSH = "--noflush" in sys.argv
for P, W, Q, T, D in shapes:
    sup = torch.randn(P, W, T, D, device="cuda"); tgt = torch.randn(P, Q, T, D, device="cuda")
    out = torch.zeros(P, Q, W, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    ms = []
    for i in range(8):
        if not NOFLUSH:
            flush.zero_()                  # evict the operands from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.otam_distance(sup, tgt, False, out=out); e1.record(); torch.cuda.synchronize()
        if i >= 3: ms.append(e0.elapsed_time(e1))
    t = sum(ms) / len(ms)
    byts = P * ((Q + W) * T * D * 4 + Q * W * 4)
    print("otam P=%d W=%d Q=%d T=%d D=%d: %.1f us, %.0f GB/s algorithmic = %.2f of measured HBM peak (%.0f GB/s); %.2f Gcell/s"
          % (P, W, Q, T, D, t * 1e3, byts / t / 1e6, byts / t / 1e6 / peak, peak, P * 2 * Q * W * T * (T + 2) / t / 1e6))
