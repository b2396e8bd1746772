"""Metric head alone (HSMR motion path, SPM gates / se_te, PADM, OTAM tail, logits) on random frame features:
ms per call by CUDA events, and the command profiled under ncu for the per-kernel HBM figures north_star asks for.
    python tools/time_head.py [--episodes E] [--shot K] [--iters N] [--backbone RN50]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import CNN, _lib
from clip_spm_b200.config import make_cfg

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=8)
ap.add_argument("--shot", type=int, default=5)
ap.add_argument("--seq-len", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--backbone", default="ViT-B/16")
a = ap.parse_args()
D = 512 if a.backbone == "ViT-B/16" else 1024
E, W, T = a.episodes, 5, a.seq_len
S, Q = W * a.shot, W
net = CNN(make_cfg(a.backbone, T, False, W), max_episodes=E)
sd = {k: v for k, v in net.state_dict().items() if not k.startswith("backbone.")}
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    for k, p in list(net.named_parameters()):
        if k.startswith("backbone.") or k in ("scale", "mo_alpha1"):
            continue
        if p.dim() <= 1:
            p.copy_(torch.randn(p.shape, generator=g) * 0.02 + (1.0 if k.endswith("norm.weight") else 0.0))
        else:
            p.copy_(torch.randn(p.shape, generator=g) * p[0].numel() ** -0.5)
net.text_features_test = torch.randn(24 if D == 512 else 10, D, generator=g)
su = torch.randn(E, S, T, D, generator=g).cuda(); qu = torch.randn(E, Q, T, D, generator=g).cuda()
lab = torch.stack([torch.arange(W).repeat_interleave(a.shot)[torch.randperm(S, generator=g)] for _ in range(E)]).float()
rs = lab.clone(); rt = torch.stack([torch.randperm(W, generator=g) for _ in range(E)]).float()
f = lambda: net.head(su, qu, lab, rs, rt, n_episodes=E)
for _ in range(3):
    out = f()
torch.cuda.synchronize()
lib = _lib.load()
l0 = lib.spm_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    out = f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print("head %s E=%d S=%d Q=%d T=%d D=%d: %.3f ms per call (%.1f us per episode), %d launches per call, finite=%s"
      % (a.backbone, E, S, Q, T, D, ms, 1e3 * ms / E, (lib.spm_launch_count() - l0) // a.iters,
         bool(torch.isfinite(out["logits"]).all())))
