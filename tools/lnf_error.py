"""Feature error of the ViT-B/16 tower against the executed reference's golden features (tests/golden/vit_5w5s_t8_p1.npz,
240 frames) for the build / environment it runs in -- used to compare the LayerNorm variants (SPM_LN_FOLD=0/1/2) and the
bf16 residual stream on the same inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import helpers as H
name = sys.argv[1] if len(sys.argv) > 1 else "vit_5w5s_t8_p1"
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
ci, g = H.case_inputs(name), H.golden(name)
net = H.build_cuda_model(ci, precision=precision)
ep = ci["episode"]
su = net.encode_frames(ep["context_images"].cuda()).cpu()
qu = net.encode_frames(ep["target_images"].cuda()).cpu()
part = net.encode_frames(ep["context_images"][1:3].cuda()).cpu()
f = torch.cat([su.view(-1, su.shape[-1]), qu.view(-1, qu.shape[-1])]).double()
r = torch.cat([g["su"].reshape(-1, su.shape[-1]), g["qu"].reshape(-1, qu.shape[-1])]).double()
rms = float(((f - r) ** 2).mean().sqrt() / (r ** 2).mean().sqrt())
print("SPM_LN_FOLD=%s precision=%s %s: max rel err %.3e, rms rel err %.3e, subset-vs-batch max abs %.3e" % (
    os.environ.get("SPM_LN_FOLD", "default"), precision, name, H.rel_err(f, r), rms, float((part - su.view(-1, su.shape[-1])[1:3]).abs().max())))
