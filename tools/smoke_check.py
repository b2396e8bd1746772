"""__graft_entry__.smoke(): one small episode through the CUDA path on cuda:0, checked against the oracle.
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
