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
if os.environ.get("SPM_BIND_CPUS", "1") != "0":      # a Python-driven step launches ~1500 kernels: keep the process next to its GPU
    from clip_spm_b200 import sweep as _sweep
    print("bound to CPUs:", (_sweep.bind_to_gpu_cpus(0) or "unchanged"))
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


ONLY_WHOLE = os.environ.get("SPM_TIMING_ONLY_WHOLE") == "1"      # ncu launch list of one whole step: skip everything else
if ONLY_WHOLE:
    steps = 1
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


ms_e, _ = timed(eager, max(3, steps // 5)) if not ONLY_WHOLE else (float("nan"), 0)
print("PyTorch eager (oracle head on CUDA, autograd, TF32 matmuls, fused Adam, GradScaler; dropout p = 0): %.2f ms/step = "
      "%.1f steps/s  ->  %.1fx" % (ms_e, 1e3 / ms_e, ms_e / ms_s))


# ------------------------------------------------------------------------------------------------------------------
# whole training step from the frames: ViT-B/16 tower + head forward, backward, Adam over all parameters (config-1 shape)
# ------------------------------------------------------------------------------------------------------------------
def whole_step(way, shot, qpc, T, n_steps):
    from clip_spm_b200 import CNN
    from tests.helpers import make_cfg
    wf = O.make_weights("ViT-B/16", seed=0, protocol="P1", head_only=False)
    epi = O.make_episode(6001, way, shot, qpc, T, ncls, "P1", images=True)
    epi = {k: v.to(dev) for k, v in epi.items()}
    frames = (way * shot + way * qpc) * T
    net = CNN(make_cfg("ViT-B/16", T, False, way), text_features_test=text.cpu(), text_features_train=text.cpu(), precision="bf16")
    net.load_state_dict(wf, strict=False)
    net.train_backbone = True
    net.train()
    opt_w = optim.Adam(net.trainable_parameters(), lr=1e-6, betas=(0.5, 0.999))
    sc = optim.GradScaler(dev)

    def step():
        out = net(epi)
        sc.scale(net.loss(out, epi["target_labels"])).backward()
        sc.step(opt_w)
        sc.update()
        opt_w.zero_grad()
    ms, launches = timed(step, n_steps)
    torch.cuda.reset_peak_memory_stats()
    step()
    torch.cuda.synchronize()
    print("clip_spm_b200 WHOLE train step (%d frames: tower + head forward, backward, Adam over %d parameters; tf32 products, "
          "dropout on): %.1f ms/step = %.2f episodes/s (%d library launches/step, torch allocator peak %.1f GB)"
          % (frames, sum(p.numel() for p in net.trainable_parameters()), ms, 1e3 / ms, launches,
             torch.cuda.max_memory_allocated() / 2 ** 30))
    del net, opt_w
    torch.cuda.empty_cache()
    if ONLY_WHOLE:
        return
    # PyTorch eager: the oracle tower + head on CUDA under autograd, fp32 with TF32 matmuls and under autocast(bf16)
    wfe = {k: v.to(dev).clone().requires_grad_(v.dtype.is_floating_point and k != "scale") for k, v in wf.items()}
    oe = torch.optim.Adam([p for p in wfe.values() if p.requires_grad], lr=1e-6, betas=(0.5, 0.999), fused=True)
    se = torch.amp.GradScaler(dev)
    cfg = dict(backbone="ViT-B/16", seq_len=T, mid_dim=512, params=O.DEFAULT_PARAMS, single_direct=False)
    for ac in (False, True):
        def estep():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                su_ = O.vit_forward(wfe, epi["context_images"], chunk=512).reshape(-1, T, 512)
                qu_ = O.vit_forward(wfe, epi["target_images"], chunk=512).reshape(-1, T, 512)
            st = O.head_forward(wfe, text, su_.float(), qu_.float(), epi["context_labels"], epi["real_support_labels"],
                                epi["real_target_labels"], O.DEFAULT_PARAMS, False)
            loss = torch.nn.functional.cross_entropy(st["logits"][0], epi["target_labels"].long(), reduction="sum") / 16 \
                + 0.001 * st["dists"]
            se.scale(loss).backward()
            se.step(oe)
            se.update()
            oe.zero_grad()
        ms_e, _ = timed(estep, max(2, n_steps // 3))
        print("PyTorch eager whole step (%s): %.1f ms/step = %.2f episodes/s -> library is %.2fx"
              % ("autocast bf16 tower" if ac else "fp32 / TF32 matmuls", ms_e, 1e3 / ms_e, ms_e / ms))


whole_step(5, 1, 1, 8, int(os.environ.get("SPM_TIMING_WHOLE_STEPS", "1" if ONLY_WHOLE else "6")))
