"""Context numbers (NOT part of bench.py): the same episode through PyTorch-eager library kernels on the B200
(cuBLAS / SDPA under autocast(bfloat16), i.e. how run/main_run.py:274 would run the reference on this GPU), and the
sustained cuBLAS rate on the four frame-encoder GEMM shapes.  Uses the oracle restatement moved to CUDA."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clipspm_oracle as O

dev = "cuda"
torch.backends.cudnn.benchmark = True
w = {k: v.to(dev) for k, v in O.make_weights("ViT-B/16", seed=0, protocol="P0").items()}
text = O.make_text_features(24, 512).to(dev)
ep = O.make_episode(1000, 5, 5, 1, 8, 24, images=False)
ep = {k: v.to(dev) for k, v in ep.items()}
ep["context_images"] = torch.rand(200, 3, 224, 224, device=dev)
ep["target_images"] = torch.rand(40, 3, 224, 224, device=dev)
cfg = dict(backbone="ViT-B/16", seq_len=8, mid_dim=512, params=O.DEFAULT_PARAMS)


def run(autocast):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        # the head's OTAM python loops dominate eager time; report tower and head separately
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
        t0.record()
        su = O.vit_forward(w, ep["context_images"], chunk=256).reshape(-1, 8, 512).float()
        qu = O.vit_forward(w, ep["target_images"], chunk=256).reshape(-1, 8, 512).float()
        t1.record()
    with torch.no_grad():
        O.head_forward(w, text, su, qu, ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"],
                       O.DEFAULT_PARAMS)
        t2.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1), t1.elapsed_time(t2)


for ac in (True, False):
    run(ac); run(ac)
    r = [run(ac) for _ in range(3)]
    tower = min(x[0] for x in r); head = min(x[1] for x in r)
    print("torch-eager %s: tower %.1f ms (%.0f TFLOP/s), head %.1f ms -> %.1f episodes/s" %
          ("autocast-bf16" if ac else "fp32", tower, 240 * 35.127 / tower, head, 1e3 / (tower + head)))

M = 47280
for name, n, k in (("qkv", 2304, 768), ("out", 768, 768), ("fc", 3072, 768), ("proj", 768, 3072)):
    a = torch.randn(M, k, device=dev, dtype=torch.bfloat16); b = torch.randn(n, k, device=dev, dtype=torch.bfloat16)
    for _ in range(20): a @ b.t()
    torch.cuda.synchronize(); t0 = time.perf_counter(); n_it = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < 1.5:
        for _ in range(50): a @ b.t()
        n_it += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_it
    print("cuBLAS bf16 %-4s [%d x %d x %d]: %.1f us, %.0f TFLOP/s sustained" % (name, M, n, k, ms * 1e3, 2.0 * M * n * k / ms / 1e9))
