"""The existing-library bar: the reference's ViT-B/16 visual tower under PyTorch eager
on the SAME B200 -- built from the torch.nn modules the reference builds it from (models/clip_fsar.py:610-689:
nn.Conv2d patch embedding, nn.MultiheadAttention blocks, nn.Linear MLP with QuickGELU, an fp32 LayerNorm subclass), so
ATen's own attention path and cuBLASLt GEMMs run; fp32 and autocast(bfloat16) with cudnn.benchmark=True as
run/main_run.py:59,274 set them.  It is loaded with the product model's weights (same state_dict keys) and batched as
favourably as the product path.  Baseline harness for bench.py only: nothing in clip_spm_b200/ imports it, and it
imports neither the oracle nor the library."""
import torch
import torch.nn as nn


class _LN(nn.LayerNorm):   # clip_fsar.py:610-616: statistics in fp32 whatever the activation type
    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


class _Block(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(width, heads)
        self.ln_1 = _LN(width)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, 4 * width))
        self.mlp.add_module("c_proj", nn.Linear(4 * width, width))
        self.ln_2 = _LN(width)

    def forward(self, x):   # x: [L, N, C] (sequence first, as the reference feeds nn.MultiheadAttention)
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False)[0]
        h = self.mlp.c_fc(self.ln_2(x))
        return x + self.mlp.c_proj(h * torch.sigmoid(1.702 * h))


class _Stack(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[_Block(width, heads) for _ in range(layers)])


class EagerVisualTower(nn.Module):
    """Parameter names equal the reference's `backbone.*` keys, so the product model's state_dict loads directly."""

    def __init__(self, width=768, layers=12, heads=12, patch=16, res=224, out_dim=512):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, patch, patch, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(width))
        self.positional_embedding = nn.Parameter(torch.zeros((res // patch) ** 2 + 1, width))
        self.ln_pre = _LN(width)
        self.transformer = _Stack(width, layers, heads)
        self.ln_post = _LN(width)
        self.proj = nn.Parameter(torch.zeros(width, out_dim))

    def forward(self, images):
        x = self.conv1(images).flatten(2).transpose(1, 2)
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], 1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x).transpose(0, 1)
        x = self.transformer.resblocks(x).transpose(0, 1)
        return self.ln_post(x[:, 0]) @ self.proj.to(x.dtype)


def build_from(cnn_state_dict, device):
    net = EagerVisualTower()
    sd = {k[len("backbone."):]: v for k, v in cnn_state_dict.items() if k.startswith("backbone.")}
    net.load_state_dict(sd, strict=True)
    return net.to(device).eval()


def tower_frames_per_s(tower, images, autocast, frames_per_forward, steps, warmup=2):
    """frames/s of the eager tower over `images` [F,3,224,224] (device-resident), CUDA events, `steps` passes."""
    torch.backends.cudnn.benchmark = True          # run/main_run.py:59
    F = images.shape[0]

    def one_pass():
        outs = []
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):   # run/main_run.py:258,274
            for i in range(0, F, frames_per_forward):
                outs.append(tower(images[i:i + frames_per_forward]))
        return outs

    for _ in range(warmup):
        one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = one_pass()
    e1.record()
    torch.cuda.synchronize()
    return F * steps / (e0.elapsed_time(e1) / 1e3), torch.cat(out).float()
