"""__graft_entry__.smoke(): one small episode through the CUDA path on cuda:0 -- the evaluation forward, then the same episode
forward + backward in train mode (tower and head) -- checked against the oracle.
Lives outside the product package: it imports the oracle (the checker), which nothing in clip_spm_b200/ may do."""
import os
import sys

import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tools/ -> repo root
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import clipspm_oracle as O  # checker only
    from tests import helpers as H
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs cuda:0")
    ci = H.case_inputs("vit_2w1s_t2_p0")  # 2-way 1-shot, T=2: 8 frames through the full ViT-B/16 + head
    net = H.build_cuda_model(ci)
    ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()}
    out = net(ep)
    torch.cuda.synchronize()
    cfg = dict(backbone="ViT-B/16", seq_len=ci["T"], mid_dim=512, params=O.DEFAULT_PARAMS, single_direct=False)
    with torch.no_grad():
        ref = O.forward(ci["weights"], ci["text"], ci["episode"], cfg)
    err = H.rel_err(out["logits"].cpu(), ref["logits"])
    print("smoke: logits", out["logits"].flatten().tolist(), "oracle", ref["logits"].flatten().tolist(),
          "rel err %.3e" % err)
    if not err < 2e-2:
        raise RuntimeError("smoke: CUDA path disagrees with the oracle (rel err %.3e)" % err)
    # forward + backward: the same 8 frames in train mode (dropout off), every gradient of the tower and the head against the
    # oracle's autograd on the CPU (relative L2; the training path multiplies in tf32)
    net.text_features_train = ci["text"]
    net.train_backbone, net.train_dropout = True, False
    net.train()
    tout = net(ep)
    net.loss(tout, ep["target_labels"]).backward()
    torch.cuda.synchronize()
    loss, grads = O.train_loss_and_grads(ci["weights"], ci["text"], ci["episode"], cfg)
    worst, n = 0.0, 0
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        g = grads[k].reshape(p.grad.shape).double()
        worst = max(worst, float((p.grad.cpu().double() - g).norm() / g.norm().clamp_min(1e-30)))
        n += 1
    print("smoke: train-mode forward + backward, %d gradients, worst relative L2 error %.3e" % (n, worst))
    # (a gross-error gate: the tf32 path measures 2.8e-2 here; the strict gradient parity -- fp32 mode 2e-3, tf32 5e-2 on the
    # goldens of the reference's own backward -- lives in tests/test_train_gpu.py)
    if n != len(grads) or not worst < 0.15:
        raise RuntimeError("smoke: gradients disagree with the oracle's autograd (%d of %d tensors, worst %.3e)" % (n, len(grads), worst))
