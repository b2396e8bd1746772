"""Context numbers for the head's training step (NOT part of bench.py): one train step -- forward, backward, Adam through
the GradScaler -- at BASELINE config-2 head shape (5-way 5-shot, T = 8, D = 512) through this library (clip_spm_b200.train
+ optim) next to the same step as PyTorch eager (the oracle restatement moved to CUDA under torch autograd, TF32 matmuls
allowed, torch.optim.Adam(fused) + torch.amp.GradScaler: how run/main_run.py:245-254,207-209 runs the reference's head on
this GPU).  Lives under tests/ because it imports the oracle.   python tests/train_step_timing.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import clipspm_oracle as O
from clip_spm_b200 import _lib, optim, train

dev = "cuda"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
way, shot, qpc, T, D, ncls = 5, 5, 1, 8, 512, 24
w0 = {k: v for k, v in O.make_weights("ViT-B/16", seed=0, protocol="P1", head_only=True).items() if v.dtype.is_floating_point}
text = O.make_text_features(ncls, D, seed=1).to(dev)
ep = {k: v.to(dev) for k, v in O.make_episode(5001, way, shot, qpc, T, ncls, "P1", images=False).items()}
su, qu = (t.to(dev) for t in O.make_features(5001, way * shot, way * qpc, T, D, ep["context_labels"].cpu(), ep["target_labels"].float().cpu()))


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    l0 = _lib.load().spm_launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, (_lib.load().spm_launch_count() - l0) / n


# ---- this library
w = {k: v.to(dev).clone().requires_grad_(k != "scale") for k, v in w0.items()}
c1, c2 = train.TransformerV1(D), train.TransformerV1(D)
opt = optim.Adam([p for p in w.values() if p.requires_grad], lr=1e-5, betas=(0.5, 0.999))
scaler = optim.GradScaler(dev)
seed = [0]


def mine(dropout=True, backward=True):
    seed[0] += 4
    out = train.spm_head_forward(w, text, su, qu, ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"],
                                 O.DEFAULT_PARAMS, c1, c2, False, False, seed[0] if dropout else None)
    if not backward:
        for b in (c1, c2):
            b.reset()
        return
    loss = train.spm_loss(out, ep["target_labels"], 16.0)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    opt.zero_grad()


with torch.no_grad():
    ms_f, l_f = timed(lambda: mine(True, False), steps)
ms_s, l_s = timed(mine, steps)
print("clip_spm_b200 head train step (dropout on, tf32 products): %.2f ms/step = %.1f steps/s (%d library launches/step); "
      "forward alone (no graph) %.2f ms (%d launches)" % (ms_s, 1e3 / ms_s, l_s, ms_f, l_f))

# ---- PyTorch eager on the same GPU
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
we = {k: v.to(dev).clone().requires_grad_(k != "scale") for k, v in w0.items()}
opt_e = torch.optim.Adam([p for p in we.values() if p.requires_grad], lr=1e-5, betas=(0.5, 0.999), fused=True)
scaler_e = torch.amp.GradScaler(dev)


def eager():
    st = O.head_forward(we, text, su, qu, ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"],
                        O.DEFAULT_PARAMS, False)
    lg = st["logits"][0]
    loss = torch.nn.functional.cross_entropy(lg, ep["target_labels"].long(), reduction="sum") / 16 + 0.001 * st["dists"]
    scaler_e.scale(loss).backward()
    scaler_e.step(opt_e)
    scaler_e.update()
    opt_e.zero_grad()


ms_e, _ = timed(eager, max(3, steps // 5))
print("PyTorch eager (oracle head on CUDA, autograd, TF32 matmuls, fused Adam, GradScaler; dropout p = 0): %.2f ms/step = "
      "%.1f steps/s  ->  %.1fx" % (ms_e, 1e3 / ms_e, ms_e / ms_s))
