"""ORACLE -- TEST INFRASTRUCTURE ONLY.  A CPU restatement (plain torch fp32/fp64 tensor ops, no nn.Module of the
reference, no CUDA) of the algorithm on CLIP-SPM's episode-evaluation hot path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product path (clip_spm_b200/) never
does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference has no golden vectors or tests of its own, so this
restatement is pinned against outputs of the reference itself: oracle/pin_against_reference.py imports the
reference's own modules from /root/reference (CPU), feeds them the same seeded weights and episodes, asserts that
every stage tensor agrees with this file and writes the stage tensors to tests/golden/*.npz.

Every function cites the reference file:line it restates (paths relative to the reference repo root).
"""
import math
import zlib

import torch
import torch.nn.functional as F

VIT = dict(width=768, layers=12, heads=12, patch=16, res=224, out_dim=512)
RN50 = dict(width=64, layers=(3, 4, 6, 3), heads=32, res=224, out_dim=1024)


# =====================================================================================================
# frame encoder: CLIP ViT-B/16 visual tower
# =====================================================================================================
def layer_norm(x, w, b, eps=1e-5):
    """models/clip_fsar.py:610-616 (nn.LayerNorm computed in fp32, eps 1e-5)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def quick_gelu(x):
    """models/clip_fsar.py:618-620."""
    return x * torch.sigmoid(1.702 * x)


def vit_block(x, w, p, heads):
    """models/clip_fsar.py:622-643 ResidualAttentionBlock on x [F, L, C] (batch-first restatement of
    nn.MultiheadAttention(d_model, n_head) self-attention without mask)."""
    Fn, L, C = x.shape
    hd = C // heads
    h = layer_norm(x, w[p + "ln_1.weight"], w[p + "ln_1.bias"])
    qkv = h @ w[p + "attn.in_proj_weight"].t() + w[p + "attn.in_proj_bias"]
    q, k, v = qkv.split(C, dim=-1)
    q = q.view(Fn, L, heads, hd).transpose(1, 2)
    k = k.view(Fn, L, heads, hd).transpose(1, 2)
    v = v.view(Fn, L, heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(Fn, L, C)
    x = x + o @ w[p + "attn.out_proj.weight"].t() + w[p + "attn.out_proj.bias"]
    h = layer_norm(x, w[p + "ln_2.weight"], w[p + "ln_2.bias"])
    h = quick_gelu(h @ w[p + "mlp.c_fc.weight"].t() + w[p + "mlp.c_fc.bias"])
    x = x + h @ w[p + "mlp.c_proj.weight"].t() + w[p + "mlp.c_proj.bias"]
    return x


def vit_forward(w, images, prefix="backbone.", chunk=32):
    """models/clip_fsar.py:672-689 VisionTransformer.forward: images [F,3,224,224] -> [F,512]."""
    outs = []
    for s in range(0, images.shape[0], chunk):
        img = images[s:s + chunk].to(w[prefix + "conv1.weight"].dtype)
        x = F.conv2d(img, w[prefix + "conv1.weight"], stride=VIT["patch"])          # :673
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)                   # :674-675
        cls = w[prefix + "class_embedding"].expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + w[prefix + "positional_embedding"]          # :676-677
        x = layer_norm(x, w[prefix + "ln_pre.weight"], w[prefix + "ln_pre.bias"])    # :678
        for i in range(VIT["layers"]):                                               # :680-682
            x = vit_block(x, w, prefix + "transformer.resblocks.%d." % i, VIT["heads"])
        x = layer_norm(x[:, 0, :], w[prefix + "ln_post.weight"], w[prefix + "ln_post.bias"])  # :684
        outs.append(x @ w[prefix + "proj"])                                          # :686-687
    return torch.cat(outs, 0)


# =====================================================================================================
# frame encoder: CLIP ModifiedResNet-50 visual tower
# =====================================================================================================
def _bn(x, w, p, eps=1e-5):
    """nn.BatchNorm2d in eval mode (running statistics)."""
    scale = w[p + "weight"] / torch.sqrt(w[p + "running_var"] + eps)
    shift = w[p + "bias"] - w[p + "running_mean"] * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def rn50_bottleneck(x, w, p, stride):
    """models/clip_fsar.py:502-547 Bottleneck."""
    out = torch.relu(_bn(F.conv2d(x, w[p + "conv1.weight"]), w, p + "bn1."))
    out = torch.relu(_bn(F.conv2d(out, w[p + "conv2.weight"], padding=1), w, p + "bn2."))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(F.conv2d(out, w[p + "conv3.weight"]), w, p + "bn3.")
    if (p + "downsample.0.weight") in w:
        idn = F.avg_pool2d(x, stride) if stride > 1 else x
        idn = _bn(F.conv2d(idn, w[p + "downsample.0.weight"]), w, p + "downsample.1.")
    else:
        idn = x
    return torch.relu(out + idn)


def rn50_attnpool(x, w, p, heads):
    """models/clip_fsar.py:407-410,481-500 AttentionPool2d: query = mean token only."""
    Fn, C = x.shape[0], x.shape[1]
    t = x.flatten(2).permute(0, 2, 1)                              # [F, HW, C]
    t = torch.cat([t.mean(dim=1, keepdim=True), t], dim=1)         # [F, HW+1, C]
    t = t + w[p + "positional_embedding"]
    hd = C // heads
    q = (t[:, :1] @ w[p + "q_proj.weight"].t() + w[p + "q_proj.bias"]).view(Fn, 1, heads, hd).transpose(1, 2)
    k = (t @ w[p + "k_proj.weight"].t() + w[p + "k_proj.bias"]).view(Fn, -1, heads, hd).transpose(1, 2)
    v = (t @ w[p + "v_proj.weight"].t() + w[p + "v_proj.bias"]).view(Fn, -1, heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(Fn, C)
    return o @ w[p + "c_proj.weight"].t() + w[p + "c_proj.bias"]


def rn50_forward(w, images, prefix="backbone.", chunk=16):
    """models/clip_fsar.py:593-608 ModifiedResNet.forward: images [F,3,224,224] -> [F,1024]."""
    outs = []
    for s in range(0, images.shape[0], chunk):
        x = images[s:s + chunk]
        x = torch.relu(_bn(F.conv2d(x, w[prefix + "conv1.weight"], stride=2, padding=1), w, prefix + "bn1."))
        x = torch.relu(_bn(F.conv2d(x, w[prefix + "conv2.weight"], padding=1), w, prefix + "bn2."))
        x = torch.relu(_bn(F.conv2d(x, w[prefix + "conv3.weight"], padding=1), w, prefix + "bn3."))
        x = F.avg_pool2d(x, 2)
        for li, nb in enumerate(RN50["layers"]):
            for bi in range(nb):
                stride = 2 if (li > 0 and bi == 0) else 1
                x = rn50_bottleneck(x, w, prefix + "layer%d.%d." % (li + 1, bi), stride)
        outs.append(rn50_attnpool(x, w, prefix + "attnpool.", RN50["heads"]))
    return torch.cat(outs, 0)


# =====================================================================================================
# metric head
# =====================================================================================================
def feed_forward(x, w, p, masks=None):
    """models/myRes.py:984-996 FeedForward: Linear -> exact (erf) GELU -> Dropout -> Linear -> Dropout (identity in eval;
    in train mode `masks` = the two keep-masks already scaled by 1 / (1 - p), see dropout_mask)."""
    h = F.gelu(x @ w[p + "net.0.weight"].t() + w[p + "net.0.bias"])
    if masks is not None:
        h = h * masks[0]
    y = h @ w[p + "net.3.weight"].t() + w[p + "net.3.bias"]
    return y * masks[1] if masks is not None else y


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11) on numpy uint64 arrays holding 32-bit words: ctr = 4 arrays, key = 2 ints.
    Known answers (Random123 kat_vectors) are checked in tests/test_train_cpu.py."""
    import numpy as np
    m32 = np.uint64(0xFFFFFFFF)
    c = [np.asarray(v, dtype=np.uint64) for v in ctr]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & m32, p1 & m32, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & m32, p0 & m32]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    return c


def dropout_mask(shape, p, seed, site):
    """The keep-mask (times 1 / (1 - p)) the library's train-mode dropout applies (include/clipspm_b200.h,
    spm_tv1_set_dropout): element i keeps iff (word i % 4 of philox(counter = (i // 4, site), key = seed) >> 8) * 2^-24 >= p.
    Replays the device masks on the CPU so that the oracle can follow a train-mode forward with p > 0."""
    import numpy as np
    n = int(np.prod(shape))
    q = np.arange((n + 3) // 4, dtype=np.uint64)
    words = philox4x32_10([q & np.uint64(0xFFFFFFFF), q >> np.uint64(32), np.full_like(q, site), np.zeros_like(q)],
                          [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
    r = np.stack(words, axis=1).reshape(-1)[:n]
    u = (r >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)
    keep = u >= np.float32(p)
    return torch.from_numpy((keep.astype(np.float32) * (np.float32(1.0) / (np.float32(1.0) - np.float32(p)))).reshape(shape))


def transformer_v1(x, w, p, heads=8, dim_head=256, depth=1, masks=None):
    """models/myRes.py:1066-1075 Transformer_v1.forward with q=k=v=x:
    PreNormattention_qkv (:1039-1040, the SAME LayerNorm on q,k,v, residual is the un-normalised q),
    Attention_qkv.forward (:964-982), then x = ff(x) + x; layers 1.. (depth > 1, :1070-1073) repeat it on the output."""
    B, n, _ = x.shape
    for i in range(depth):
        l = p + "layers.%d." % i
        h = layer_norm(x, w[l + "0.norm.weight"], w[l + "0.norm.bias"])
        q = (h @ w[l + "0.fn.to_q.weight"].t()).view(B, n, heads, dim_head).transpose(1, 2)
        k = (h @ w[l + "0.fn.to_k.weight"].t()).view(B, n, heads, dim_head).transpose(1, 2)
        v = (h @ w[l + "0.fn.to_v.weight"].t()).view(B, n, heads, dim_head).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) * dim_head ** -0.5, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, n, heads * dim_head)
        a = o @ w[l + "0.fn.to_out.0.weight"].t() + w[l + "0.fn.to_out.0.bias"]
        if masks is not None:    # train mode: (to_out dropout :961-962, FeedForward's two :990,992) of layer 0
            a = a * masks[0]
        y = a + x
        x = feed_forward(y, w, l + "1.", None if masks is None else masks[1:]) + y
    return x


def motion_feats(x, w):
    """models/model_clipspm.py:169-191 get_motion_feats for one tensor x [N,T,D] -> [N,D]."""
    xt = x.permute(0, 2, 1)
    c = F.conv1d(xt, w["motion_conv1.weight"], w["motion_conv1.bias"], padding=1)
    c = F.conv1d(c, w["motion_conv2.weight"], w["motion_conv2.bias"], padding=1)
    f = c[:, :, 1:] - xt[:, :, :-1]
    b = c[:, :, :-1] - xt[:, :, 1:]
    return (0.5 * (f + b)).mean(-1)


def gate(x, w, p, slope):
    """models/model_clipspm.py:88-99 gate_text / gate_vision: Linear -> LeakyReLU -> Linear -> Sigmoid."""
    h = F.leaky_relu(x @ w[p + "0.weight"].t() + w[p + "0.bias"], slope)
    return torch.sigmoid(h @ w[p + "2.weight"].t() + w[p + "2.bias"])


def se_te(x, tok, w, params):
    """models/model_clipspm.py:296-314: x [N,T,D], tok [N,1,D] -> (frames [N,T,D], token [N,1,D])."""
    gt = gate(tok, w, "gate_text.", params["negative_slope"])
    gv = gate(x, w, "gate_vision.", params["negative_slope"])
    q = tok * gt * params["alpha"] + x * gv
    z = transformer_v1(torch.cat([tok, q], dim=1), w, "context2.")
    return z[:, 1:, :], z[:, 0:1, :]


def token_trans(token, x, w):
    """models/model_clipspm.py:371-378: token [1,D], x [N,T,D] -> [N,1,D]."""
    t = token.expand(x.size(0), -1, -1)
    return feed_forward(t * x.mean(dim=[1, 2], keepdim=True), w, "token_tr.mlp.")


def dis(x, y):
    """models/model_clipspm.py:341-346 _dis."""
    d = x - y
    dims = [-2, -1] if d.dim() == 3 else [-1]
    return (d ** 2).sum(dim=dims).mean()


def class_means(x, labels):
    """models/model_clipspm.py:133-137,231-239 + models/myRes.py:730-739: per sorted-unique label mean."""
    return torch.stack([x[labels == c].mean(dim=0) for c in torch.unique(labels)])


def cos_sim(x, y, epsilon=0.01):
    """models/myRes.py:756-765."""
    num = x @ y.transpose(-1, -2)
    den = x.norm(dim=-1).unsqueeze(-1) @ y.norm(dim=-1).unsqueeze(-1).transpose(-1, -2) + epsilon
    return num / den


def otam_cum_dist_v2(d, lbda=0.5):
    """models/myRes.py:821-855 OTAM_cum_dist_v2 on d [..., L, M] (cum_dists is always fp32 there, :829)."""
    d = F.pad(d, (1, 1), "constant", 0.0)
    L, M2 = d.shape[-2], d.shape[-1]
    c = torch.zeros(d.shape, dtype=torch.float32 if d.dtype != torch.float64 else torch.float64, device=d.device)
    for m in range(1, M2):
        c[..., 0, m] = d[..., 0, m] + c[..., 0, m - 1]
    for l in range(1, L):
        c[..., l, 1] = d[..., l, 1] - lbda * torch.log(
            torch.exp(-c[..., l - 1, 0] / lbda) + torch.exp(-c[..., l - 1, 1] / lbda) + torch.exp(-c[..., l, 0] / lbda))
        for m in range(2, M2 - 1):
            c[..., l, m] = d[..., l, m] - lbda * torch.log(
                torch.exp(-c[..., l - 1, m - 1] / lbda) + torch.exp(-c[..., l, m - 1] / lbda))
        c[..., l, -1] = d[..., l, -1] - lbda * torch.log(
            torch.exp(-c[..., l - 1, -2] / lbda) + torch.exp(-c[..., l - 1, -1] / lbda)
            + torch.exp(-c[..., l, -2] / lbda))
    return c[..., -1, -1]


def otam_distance(support, target, single_direct=False):
    """models/model_clipspm.py:348-362: support [W,T,D], target [Q,T,D] -> [Q,W]."""
    W, T, D = support.shape
    Q = target.shape[0]
    sim = cos_sim(target.reshape(Q * T, D), support.reshape(W * T, D))
    d = (1 - sim).view(Q, T, W, T).permute(0, 2, 1, 3)  # tb sb ts ss
    out = otam_cum_dist_v2(d)
    if not single_direct:
        out = out + otam_cum_dist_v2(d.transpose(-1, -2))
    return out


def head_forward(w, text_features, su, qu, support_labels, real_support, real_target, params, single_direct=False):
    """models/model_clipspm.py:116-143 after get_feats (eval branch).  Returns every stage tensor."""
    st = {}
    ctx_s = text_features[real_support.long()].unsqueeze(1)   # :120
    ctx_q = text_features[real_target.long()].unsqueeze(1)    # :121
    # ---- mo (:193-206)
    su_mo, qu_mo = motion_feats(su, w), motion_feats(qu, w)
    qu_m, qu_mo2 = se_te(qu, qu_mo.unsqueeze(1), w, params)
    su_m, su_mo2 = se_te(su, su_mo.unsqueeze(1), w, params)
    new_sm, new_qm = motion_feats(su_m, w), motion_feats(qu_m, w)
    mo_dist = dis(new_qm, qu_mo2.squeeze(1)) + dis(new_sm, su_mo2.squeeze(1))
    st.update(su_mo=su_mo, qu_mo=qu_mo, mo_dist_pre=mo_dist)
    # ---- sem / cpt_sem (:208-273); the two se_te calls whose outputs only feed discarded distances are skipped
    token = torch.cat([ctx_q, ctx_s], dim=0).mean(dim=0)       # :213-214  [1,D]
    target_token = token_trans(token, qu, w)                   # :216
    qu_fake, token_q_fake = se_te(qu, target_token, w, params)     # :226
    su_real, token_s_real = se_te(su, ctx_s, w, params)            # :229
    su_pro = class_means(su_real, support_labels)                  # :231-239
    class_dists_l = otam_distance(su_pro, qu_fake, single_direct)  # :269
    st.update(target_token=target_token, qu_fake=qu_fake, su_real=su_real, token_q_fake=token_q_fake,
              token_s_real=token_s_real, su_pro=su_pro, class_dists_l=class_dists_l)
    dists = w["mo_alpha1"] * mo_dist                               # :129 (consist/text distances are 0, :258-259)
    # ---- taskM (:275-294)
    uniq = torch.unique(support_labels)
    suu = torch.stack([su_real[support_labels == c] for c in uniq])          # [W,K,T,D]
    cn = suu.size(0)
    token_s = torch.cat([suu, qu_fake.unsqueeze(0).repeat(cn, 1, 1, 1)], dim=1).mean(dim=1)   # :283
    token_q = token_s.mean(dim=0, keepdim=True)                                              # :284
    su_t = torch.cat([token_s, su_real], dim=0).permute(1, 0, 2)
    qu_t = torch.cat([token_q, qu_fake], dim=0).permute(1, 0, 2)
    _su = transformer_v1(su_t, w, "context1.").permute(1, 0, 2)
    _qu = transformer_v1(qu_t, w, "context1.").permute(1, 0, 2)
    su_2, qu_2, su_t2, qu_t2 = _su[cn:], _qu[1:], _su[:cn], _qu[0:1]
    su_pro2 = class_means(su_2, support_labels)                                 # :133-137
    task_dist = otam_distance(su_pro2, qu_2, single_direct) + otam_distance(su_t2, qu_t2, single_direct)  # :138
    logits = -(0.5 * class_dists_l + task_dist).unsqueeze(0)                    # :141
    st.update(su_2=su_2, qu_2=qu_2, su_t2=su_t2, qu_t2=qu_t2, su_pro2=su_pro2, task_dist=task_dist,
              logits=logits, dists=dists.reshape(()))
    return st


def forward(w, text_features, inputs, cfg):
    """models/model_clipspm.py:111-144 CNN.forward (eval): inputs is the reference's episode dict."""
    T = cfg["seq_len"]
    enc = vit_forward if cfg["backbone"] == "ViT-B/16" else rn50_forward
    su = enc(w, inputs["context_images"]).reshape(-1, T, cfg["mid_dim"])        # :158-163
    qu = enc(w, inputs["target_images"]).reshape(-1, T, cfg["mid_dim"])
    st = head_forward(w, text_features, su, qu, inputs["context_labels"], inputs["real_support_labels"],
                      inputs["real_target_labels"], cfg["params"], cfg.get("single_direct", False))
    st.update(su=su, qu=qu)
    return st


# =====================================================================================================
# sibling head: CLIP-FSAR (models/model_clipfsar.py) -- SURVEY.md 8(f) rank 4, same kernels, other wiring
# =====================================================================================================
def fsar_weight_shapes(D, depth=1):
    """Parameters of CNN_OTAM_CLIPFSAR besides the backbone (models/model_clipfsar.py:137-145): `scale` and
    context2 = Transformer_v1(dim=D, heads=8, dim_head_k=D//8[, depth]) -> inner width D, mlp 2048."""
    s = {"scale": (1,)}
    for i in range(depth):
        c = "context2.layers.%d." % i
        s.update({c + "0.norm.weight": (D,), c + "0.norm.bias": (D,),
                  c + "0.fn.to_q.weight": (D, D), c + "0.fn.to_k.weight": (D, D), c + "0.fn.to_v.weight": (D, D),
                  c + "0.fn.to_out.0.weight": (D, D), c + "0.fn.to_out.0.bias": (D,),
                  c + "1.net.0.weight": (2048, D), c + "1.net.0.bias": (2048,),
                  c + "1.net.3.weight": (D, 2048), c + "1.net.3.bias": (D,)})
    return s


def make_fsar_weights(D, seed=0, scale=1.7, depth=1):
    """Seeded head weights of the CLIP-FSAR sibling (per-tensor generators, like make_weights; the generator key
    carries an 'fsar.' prefix so these never alias CLIP-SPM's context2 of another shape)."""
    w = {}
    for name, shp in fsar_weight_shapes(D, depth).items():
        leaf = name.split(".")[-1]
        if name == "scale":
            w[name] = torch.full((1,), float(scale))
        elif len(shp) == 1 and leaf == "weight":
            w[name] = 1.0 + _normal(seed, "fsar." + name, shp, 0.1)
        elif leaf == "bias":
            w[name] = _normal(seed, "fsar." + name, shp, 0.02)
        else:
            w[name] = _normal(seed, "fsar." + name, shp, (3.0 * shp[1]) ** -0.5)
    return w


def fsar_head_forward(w, text_test, text_train, su, qu, support_labels, real_support, real_target,
                      single_direct=False, merge_before=False, depth=1):
    """models/model_clipfsar.py:325-381, the evaluation branch that works (EVAL_TEXT / COMBINE end in
    `None.unsqueeze(0)`, :384): su [S,T,D], qu [Q,T,D] -> logits [1,Q,W], class_logits [1,S+Q,n_train].
    merge_before (MODEL.MERGE_BEFORE, :341-346): class means of the frames and of the prompts BEFORE context2 instead of
    class means of its outputs; depth (MODEL/TRAIN.TRANSFORMER_DEPTH, :143-144): layers of context2."""
    S, T, D = su.shape
    dh = D // 8
    feat_cls = torch.cat([su, qu], dim=0).mean(1)                      # :329-330 (classification_layer is empty)
    class_logits = cos_sim(feat_cls, text_train) * w["scale"]          # :331
    ctx = text_test[real_support.long()].unsqueeze(1)                  # :338
    qu2 = transformer_v1(qu, w, "context2.", heads=8, dim_head=dh, depth=depth)     # :340
    if merge_before:
        su, ctx = class_means(su, support_labels), class_means(ctx, support_labels)  # :341-346
    su2 = transformer_v1(torch.cat([su, ctx], dim=1), w, "context2.", heads=8, dim_head=dh, depth=depth)[:, :T]   # :347-348
    su_pro = su2 if merge_before else class_means(su2, support_labels)               # :349-354
    cum = otam_distance(su_pro, qu2, single_direct)                    # :360-375
    # :381-383: per-class mean over the (already one-per-class) prototype columns, transposed back -> identity
    return dict(qu_ctx=qu2, su_ctx=su2, su_pro=su_pro, logits=-cum.unsqueeze(0), class_logits=class_logits.unsqueeze(0))


def fsar_loss_and_acc(logits, class_logits, target_labels, real_support, real_target, tasks_per_batch, cls_value):
    """run/main_run.py:355-359 (MODEL.NAME == 'clipfsar'): (CE(logits) + USE_CLASSIFICATION_VALUE * CE(class_logits,
    cat[real_support, real_target])) / TASKS_PER_BATCH with utils/utils.py:174-186 `loss` (sum over rows)."""
    lg = logits[0].double()
    ce = -(lg.log_softmax(-1).gather(1, target_labels.long().view(-1, 1)).squeeze(1)).sum()
    cl = class_logits[0].double()
    real = torch.cat([real_support, real_target]).long()
    ce_cls = -(cl.log_softmax(-1).gather(1, real.view(-1, 1)).squeeze(1)).sum()
    loss = (ce + cls_value * ce_cls) / tasks_per_batch
    pred = lg.argmax(-1)
    acc = (pred == target_labels.long()).double().mean()
    return loss.float(), acc.float(), pred


# -----------------------------------------------------------------------------------------------------
# sibling head CPM2C (models/model_cpm2c.py CLIP_CPMMC_FSAR), evaluation forward
# -----------------------------------------------------------------------------------------------------
CPM2C_PARAMS = dict(mid_dim_vision=0.5, mid_dim_text=1.5, negative_slope=0.0025, alpha=0.2, motion_residual_ratio=0.5,
                    lambdas0=0.5, lambdas1=1.0, lambdas2=0.3, lambdas3=0.0,
                    # constructor-only sizes of the visual-prompt nets the forward never calls (:116-133)
                    prompt_patch=16, hid_dim=8, prompt_patch_2=3, prompt_patch_22=3, hid_dim_2=4)


def cpm2c_weight_shapes(D, params=CPM2C_PARAMS):
    """Head parameters the CPM2C forward reads (models/model_cpm2c.py:73-141): scale, context2 (inner width D), the two
    class tokens, the gates and the multi-scale motion convolutions."""
    ht, hv = int(D * params["mid_dim_text"]), int(D * params["mid_dim_vision"])
    s = dict(fsar_weight_shapes(D))
    s.update({"class_token": (1, 1, D), "class_token_motion": (1, 1, D),
              "gate_text.0.weight": (ht, D), "gate_text.0.bias": (ht,), "gate_text.2.weight": (D, ht), "gate_text.2.bias": (D,),
              "gate_vision.0.weight": (hv, D), "gate_vision.0.bias": (hv,), "gate_vision.2.weight": (D, hv),
              "gate_vision.2.bias": (D,),
              "motion_conv1_1.weight": (D, D, 1), "motion_conv1_1.bias": (D,),
              "motion_conv1_3.weight": (D, D, 3), "motion_conv1_3.bias": (D,),
              "motion_conv1_5.weight": (D, D, 3), "motion_conv1_5.bias": (D,),
              "scale_conv.weight": (D, 3 * D, 1), "scale_conv.bias": (D,)})
    return s


def make_cpm2c_weights(D, seed=0, scale=1.7, params=CPM2C_PARAMS):
    w = {}
    for name, shp in cpm2c_weight_shapes(D, params).items():
        leaf = name.split(".")[-1]
        if name == "scale":
            w[name] = torch.full((1,), float(scale))
        elif name.startswith("class_token"):
            w[name] = _normal(seed, "cpm2c." + name, shp, 0.5)
        elif len(shp) == 1 and leaf == "weight":
            w[name] = 1.0 + _normal(seed, "cpm2c." + name, shp, 0.1)
        elif leaf == "bias":
            w[name] = _normal(seed, "cpm2c." + name, shp, 0.02)
        else:
            fan_in = shp[1] * (shp[2] if len(shp) == 3 else 1)
            w[name] = _normal(seed, "cpm2c." + name, shp, (3.0 * fan_in) ** -0.5)
    return w


def cpm2c_motion(x, w, ratio):
    """models/model_cpm2c.py:165-199: multi-scale temporal convolutions (k=1, k=3, k=3 dilated by 2) -> 1x1 fusion ->
    * ratio + residual; then the forward / backward frame differences.  x [N,T,D] -> [N,T-1,D]."""
    xt = x.permute(0, 2, 1)
    f1 = F.conv1d(xt, w["motion_conv1_1.weight"], w["motion_conv1_1.bias"])
    f3 = F.conv1d(xt, w["motion_conv1_3.weight"], w["motion_conv1_3.bias"], padding=1)
    f5 = F.conv1d(xt, w["motion_conv1_5.weight"], w["motion_conv1_5.bias"], padding=2, dilation=2)
    c = F.conv1d(torch.cat([f1, f3, f5], dim=1), w["scale_conv.weight"], w["scale_conv.bias"]) * ratio + xt
    fwd = c[:, :, 1:] - xt[:, :, :-1]
    bwd = c[:, :, :-1] - xt[:, :, 1:]
    return (0.5 * (fwd + bwd)).permute(0, 2, 1)


def cpm2c_modulate(x, tok, w, params):
    """the gate / mix / context2 step models/model_cpm2c.py repeats four times per call (:337-346 etc.):
    x [N,T,D], tok [N,1,D] -> context2(cat[tok, tok*gate_text(tok)*alpha + x*gate_vision(x)]) [N,T+1,D]"""
    gt = gate(tok, w, "gate_text.", params["negative_slope"])
    gv = gate(x, w, "gate_vision.", params["negative_slope"])
    seq = torch.cat([tok, tok * gt * params["alpha"] + x * gv], dim=1)
    return transformer_v1(seq, w, "context2.", heads=8, dim_head=x.shape[-1] // 8)


def cpm2c_text_eh(ctx_support, ctx_target, su, qu, labels, token, w, params):
    """text_eh_temporal_transformer (:327-418, no MERGE_BEFORE)"""
    qu_contra = cpm2c_modulate(qu, ctx_target, w, params)                      # real target prompt
    su_contra = cpm2c_modulate(su, token.expand(su.shape[0], -1, -1), w, params)   # class token
    qu_fake = cpm2c_modulate(qu, token.expand(qu.shape[0], -1, -1), w, params)
    su_real = cpm2c_modulate(su, ctx_support, w, params)
    return su_real, qu_fake, class_means(su_real, labels), su_contra, qu_contra


def cpm2c_global_distance(su_g, labels, qu):
    """global_distance (:315-325): su_g [S,D] (token rows), qu [Q,T+1,D] -> [W,Q]"""
    d = 1 - cos_sim(qu, su_g)                                                   # [Q,T+1,S]
    out = []
    for c in torch.unique(labels):
        idx = (labels == c).nonzero().reshape(-1)
        out.append(d.index_select(2, idx).sum(2).sum(1))
    return torch.stack(out)


def cpm2c_head_forward(w, text, su, qu, support_labels, real_support, real_target, params=CPM2C_PARAMS,
                       motion_coeff=1.0, normal_coeff=1.0, use_classification=True, single_direct=False):
    """models/model_cpm2c.py:207-312 (evaluation): su [S,T,D], qu [Q,T,D] -> class_logits [1,S+Q,n_cls],
    logits_local [1,Q,W], logits_global [1,Q,W], target_consist_distance []."""
    ctx_s = text[real_support.long()].unsqueeze(1)
    ctx_t = text[real_target.long()].unsqueeze(1)
    su_m, qu_m = cpm2c_motion(su, w, params["motion_residual_ratio"]), cpm2c_motion(qu, w, params["motion_residual_ratio"])
    class_logits = cos_sim(torch.cat([su, qu], dim=0).mean(1), text) * w["scale"] if use_classification else None
    m = cpm2c_text_eh(ctx_s, ctx_t, su_m, qu_m, support_labels, w["class_token_motion"], w, params)
    n = cpm2c_text_eh(ctx_s, ctx_t, su, qu, support_labels, w["class_token"], w, params)

    def consist(r):   # :246-255 / :260-268
        su_real, qu_fake, _, su_contra, qu_contra = r
        return ((su_real - su_contra) ** 2).sum((-2, -1)).mean() + ((qu_fake - qu_contra) ** 2).sum((-2, -1)).mean()
    consist_distance = normal_coeff * consist(n) + motion_coeff * consist(m)
    g = normal_coeff * cpm2c_global_distance(n[0][:, 0, :], support_labels, n[1]) + \
        motion_coeff * cpm2c_global_distance(m[0][:, 0, :], support_labels, m[1])
    loc = normal_coeff * otam_distance(n[2][:, 1:, :], n[1][:, 1:, :], single_direct) + \
        motion_coeff * otam_distance(m[2][:, 1:, :], m[1][:, 1:, :], single_direct)
    return dict(su_motion=su_m, qu_motion=qu_m, su_real=n[0], qu_fake=n[1], su_pro=n[2], su_real_motion=m[0],
                qu_fake_motion=m[1], class_logits=None if class_logits is None else class_logits.unsqueeze(0),
                logits_local=-loc.unsqueeze(0), logits_global=-g.t().unsqueeze(0),
                target_consist_distance=consist_distance)


def cpm2c_loss_and_acc(out, target_labels, real_support, real_target, params=CPM2C_PARAMS, tasks_per_batch=16):
    """run/main_run.py:370-380 (mode 'test'): lambdas-weighted sum of three cross entropies; accuracy of
    lambdas1 * logits_local + lambdas2 * logits_global"""
    def ce(lg, y):
        return -(lg[0].double().log_softmax(-1).gather(1, y.long().view(-1, 1)).squeeze(1)).sum()
    real = torch.cat([real_support, real_target])
    total = params["lambdas1"] * out["logits_local"] + params["lambdas2"] * out["logits_global"]
    loss = (params["lambdas0"] * ce(out["class_logits"], real) + params["lambdas1"] * ce(out["logits_local"], target_labels)
            + params["lambdas2"] * ce(out["logits_global"], target_labels)) / tasks_per_batch
    pred = total[0].argmax(-1)
    return loss.float(), (pred == target_labels.long()).double().mean().float(), pred, total


def sten_head_forward(text_test, su, qu, support_labels, real_support):
    """models/model_sten.py:62-113 as shipped (all learned head modules are commented out there): su [S,8,D],
    qu [Q,8,D] -> logits [1,Q,W]."""
    su_f, qu_f = su.mean(1), qu.mean(1)                       # :65-66
    t_f = text_test[real_support.long()]                      # :71
    t_p = class_means(t_f, support_labels)                    # :97-99
    su_p = class_means(su_f, support_labels)                  # :101-102
    sim = cos_sim(qu_f, t_p).softmax(-1) * cos_sim(qu_f, su_p).softmax(-1)   # :104-106
    return dict(su_f=su_f, qu_f=qu_f, logits=sim.unsqueeze(0))


# =====================================================================================================
# TA2N's soft-DTW (models/OTAM.py) -- the reference's own numba.cuda kernels, restated in numpy (fp64 like the
# reference's CPU path, :234-289)
# =====================================================================================================
def softdtw_forward_np(D, gamma, bandwidth=0.0):
    """models/OTAM.py:234-250 compute_softdtw: D [B,N,M] -> R [B,N+2,M+2] (the value is R[:, -2, -2])."""
    import numpy as np
    D = np.asarray(D, dtype=np.float64)
    B, N, M = D.shape
    R = np.full((B, N + 2, M + 2), np.inf)
    R[:, 0, 0] = 0
    with np.errstate(invalid="ignore", divide="ignore"):
        for j in range(1, M + 1):
            for i in range(1, N + 1):
                if 0 < bandwidth < abs(i - j):
                    continue
                r = np.stack([-R[:, i - 1, j - 1], -R[:, i - 1, j], -R[:, i, j - 1]]) / gamma     # :207-215
                rmax = r.max(0)
                R[:, i, j] = D[:, i - 1, j - 1] - gamma * (np.log(np.exp(r - rmax).sum(0)) + rmax)
    return R


def softdtw_backward_np(D_, R, gamma, bandwidth=0.0):
    """models/OTAM.py:254-289 compute_softdtw_backward: -> E [B,N,M] = d R[:, N, M] / d D."""
    import numpy as np
    D_ = np.asarray(D_, dtype=np.float64)
    R = np.array(R, dtype=np.float64)
    B, N, M = D_.shape
    D = np.zeros((B, N + 2, M + 2))
    E = np.zeros((B, N + 2, M + 2))
    D[:, 1:N + 1, 1:M + 1] = D_
    E[:, -1, -1] = 1
    R[:, :, -1] = -np.inf
    R[:, -1, :] = -np.inf
    R[:, -1, -1] = R[:, -2, -2]
    with np.errstate(invalid="ignore", over="ignore"):
        for j in range(M, 0, -1):
            for i in range(N, 0, -1):
                inf = np.isinf(R[:, i, j])
                R[inf, i, j] = -np.inf
                if 0 < bandwidth < abs(i - j):
                    continue
                a = np.exp((R[:, i + 1, j] - R[:, i, j] - D[:, i + 1, j]) / gamma)
                b = np.exp((R[:, i, j + 1] - R[:, i, j] - D[:, i, j + 1]) / gamma)
                c = np.exp((R[:, i + 1, j + 1] - R[:, i, j] - D[:, i + 1, j + 1]) / gamma)
                E[:, i, j] = E[:, i + 1, j] * a + E[:, i, j + 1] * b + E[:, i + 1, j + 1] * c
    return E[:, 1:N + 1, 1:M + 1]


def softdtw_module(X, Y, gamma=1.0, normalize=False, bandwidth=0.0):
    """models/OTAM.py:390-424 SoftDTW.forward with the default distance 1 - cosine_similarity (:381-388)."""
    def dist(x, y):
        n, m, d = x.size(1), y.size(1), x.size(2)
        return 1 - torch.cosine_similarity(x.unsqueeze(2).expand(-1, n, m, d), y.unsqueeze(1).expand(-1, n, m, d), dim=3)

    def value(D):
        return torch.from_numpy(softdtw_forward_np(D.numpy(), gamma, bandwidth)[:, -2, -2]).float()
    if normalize:
        out = value(dist(torch.cat([X, X, Y]), torch.cat([Y, X, Y])))
        a, b, c = torch.split(out, X.shape[0])
        return a - 0.5 * (b + c)
    d1 = F.pad(dist(X, Y), (0, 0, 1, 1), "constant", 0)
    d2 = F.pad(dist(Y, X), (0, 0, 1, 1), "constant", 0)
    return (value(d1).unsqueeze(-1) + value(d2).unsqueeze(-1)) / 2


def make_softdtw_inputs(B, N, M, d, seed):
    g = _gen(seed, "softdtw")
    return torch.rand(B, N, d, generator=g), torch.rand(B, M, d, generator=g), torch.rand(B, N, M, generator=g)


def loss_and_acc(logits, dists, target_labels, tasks_per_batch=16):
    """utils/utils.py:174-186 loss (CE summed over queries for the single logit sample), :259-264
    aggregate_accuracy, combined as run/main_run.py:390-392."""
    lg = logits[0].double()
    ce = -(lg.log_softmax(-1).gather(1, target_labels.long().view(-1, 1)).squeeze(1))
    loss = ce.sum() / tasks_per_batch + 0.001 * dists.double()
    pred = lg.argmax(-1)
    acc = (pred == target_labels.long()).double().mean()
    return loss.float(), acc.float(), pred


# =====================================================================================================
# synthetic protocol: seeded weights and episodes, shared by the oracle, the reference
# pinning script and the CUDA path so that all three see bit-identical inputs
# =====================================================================================================
DEFAULT_PARAMS = dict(mid_dim_vision=0.5, mid_dim_text=1.5, negative_slope=0.0025, alpha=0.2, motion_alpha=1)


def _gen(seed, name):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))


def _normal(seed, name, shape, std):
    return torch.randn(*shape, generator=_gen(seed, name)) * std


def _uniform(seed, name, shape, lo, hi):
    return torch.rand(*shape, generator=_gen(seed, name)) * (hi - lo) + lo


def head_weight_shapes(D, params=DEFAULT_PARAMS):
    ht, hv = int(D * params["mid_dim_text"]), int(D * params["mid_dim_vision"])
    s = {"scale": (1,), "mo_alpha1": (),
         "motion_conv1.weight": (D, D, 3), "motion_conv1.bias": (D,),
         "motion_conv2.weight": (D, D, 3), "motion_conv2.bias": (D,),
         "token_tr.mlp.net.0.weight": (2048, D), "token_tr.mlp.net.0.bias": (2048,),
         "token_tr.mlp.net.3.weight": (D, 2048), "token_tr.mlp.net.3.bias": (D,),
         "gate_text.0.weight": (ht, D), "gate_text.0.bias": (ht,), "gate_text.2.weight": (D, ht),
         "gate_text.2.bias": (D,),
         "gate_vision.0.weight": (hv, D), "gate_vision.0.bias": (hv,), "gate_vision.2.weight": (D, hv),
         "gate_vision.2.bias": (D,)}
    for c in ("context1.", "context2."):
        s.update({c + "layers.0.0.norm.weight": (D,), c + "layers.0.0.norm.bias": (D,),
                  c + "layers.0.0.fn.to_q.weight": (2048, D), c + "layers.0.0.fn.to_k.weight": (2048, D),
                  c + "layers.0.0.fn.to_v.weight": (2048, D),
                  c + "layers.0.0.fn.to_out.0.weight": (D, 2048), c + "layers.0.0.fn.to_out.0.bias": (D,),
                  c + "layers.0.1.net.0.weight": (2048, D), c + "layers.0.1.net.0.bias": (2048,),
                  c + "layers.0.1.net.3.weight": (D, 2048), c + "layers.0.1.net.3.bias": (D,)})
    return s


def vit_weight_shapes(prefix="backbone."):
    C = VIT["width"]
    s = {prefix + "conv1.weight": (C, 3, 16, 16), prefix + "class_embedding": (C,),
         prefix + "positional_embedding": (197, C), prefix + "ln_pre.weight": (C,), prefix + "ln_pre.bias": (C,),
         prefix + "ln_post.weight": (C,), prefix + "ln_post.bias": (C,), prefix + "proj": (C, VIT["out_dim"])}
    for i in range(VIT["layers"]):
        p = prefix + "transformer.resblocks.%d." % i
        s.update({p + "attn.in_proj_weight": (3 * C, C), p + "attn.in_proj_bias": (3 * C,),
                  p + "attn.out_proj.weight": (C, C), p + "attn.out_proj.bias": (C,),
                  p + "ln_1.weight": (C,), p + "ln_1.bias": (C,), p + "ln_2.weight": (C,), p + "ln_2.bias": (C,),
                  p + "mlp.c_fc.weight": (4 * C, C), p + "mlp.c_fc.bias": (4 * C,),
                  p + "mlp.c_proj.weight": (C, 4 * C), p + "mlp.c_proj.bias": (C,)})
    return s


def rn50_weight_shapes(prefix="backbone."):
    wd = RN50["width"]
    s = {}

    def bn(p, c):
        s.update({p + "weight": (c,), p + "bias": (c,), p + "running_mean": (c,), p + "running_var": (c,),
                  p + "num_batches_tracked": ()})
    s[prefix + "conv1.weight"] = (wd // 2, 3, 3, 3); bn(prefix + "bn1.", wd // 2)
    s[prefix + "conv2.weight"] = (wd // 2, wd // 2, 3, 3); bn(prefix + "bn2.", wd // 2)
    s[prefix + "conv3.weight"] = (wd, wd // 2, 3, 3); bn(prefix + "bn3.", wd)
    inpl = wd
    for li, nb in enumerate(RN50["layers"]):
        planes = wd * (2 ** li)
        for bi in range(nb):
            p = prefix + "layer%d.%d." % (li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            s[p + "conv1.weight"] = (planes, inpl, 1, 1); bn(p + "bn1.", planes)
            s[p + "conv2.weight"] = (planes, planes, 3, 3); bn(p + "bn2.", planes)
            s[p + "conv3.weight"] = (planes * 4, planes, 1, 1); bn(p + "bn3.", planes * 4)
            if stride > 1 or inpl != planes * 4:
                s[p + "downsample.0.weight"] = (planes * 4, inpl, 1, 1); bn(p + "downsample.1.", planes * 4)
            inpl = planes * 4
    E = wd * 32
    p = prefix + "attnpool."
    s[p + "positional_embedding"] = (50, E)
    for n in ("q_proj", "k_proj", "v_proj"):
        s[p + n + ".weight"] = (E, E); s[p + n + ".bias"] = (E,)
    s[p + "c_proj.weight"] = (RN50["out_dim"], E); s[p + "c_proj.bias"] = (RN50["out_dim"],)
    return s


def make_weights(backbone, seed=0, protocol="P1", head_only=False, params=DEFAULT_PARAMS):
    """Seeded synthetic state_dict keyed by the reference's names.  Not the reference's RNG stream:
    each tensor has its own generator (seed, crc32(name)) so any subset regenerates identically anywhere.
      P0: scales of CLIP/PyTorch default init (features of different frames nearly collinear -> near-tied logits)
      P1: "separable": residual-branch weights N(0,1/fan_in), wide attention in-projections, non-trivial
          LayerNorm/BatchNorm affine terms and biases -> logit rows with usable top-1/top-2 margins."""
    D = 512 if backbone == "ViT-B/16" else 1024
    w = {}
    shapes = dict(head_weight_shapes(D, params))
    if not head_only:
        shapes.update(vit_weight_shapes() if backbone == "ViT-B/16" else rn50_weight_shapes())
    for name, shp in shapes.items():
        leaf = name.split(".")[-1]
        if name == "scale":
            w[name] = torch.ones(1)
        elif name == "mo_alpha1":
            w[name] = torch.tensor(1.0)
        elif leaf == "num_batches_tracked":
            w[name] = torch.tensor(0, dtype=torch.long)
        elif leaf == "running_mean":
            w[name] = _normal(seed, name, shp, 0.1)
        elif leaf == "running_var":
            w[name] = _uniform(seed, name, shp, 0.5, 1.5)
        elif len(shp) == 1 and leaf == "weight":   # LayerNorm / BatchNorm scale
            if ".bn3." in name and "layer" in name:
                # last BN of a bottleneck: a small (but non-zero: CLIP's default init zeroes it and hides every
                # conv) scale keeps the residual stream O(1) through 16 blocks
                w[name] = _uniform(seed, name, shp, 0.2, 0.5)
            elif "bn" in name or "downsample" in name:
                w[name] = _uniform(seed, name, shp, 0.5, 1.5)
            else:
                w[name] = 1.0 + _normal(seed, name, shp, 0.1)
        elif leaf == "bias" or leaf == "in_proj_bias":
            w[name] = _normal(seed, name, shp, 0.02)
        elif leaf in ("class_embedding", "positional_embedding", "proj"):
            w[name] = _normal(seed, name, shp, shp[-1 if leaf != "proj" else 0] ** -0.5)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            if protocol == "P1" and leaf == "in_proj_weight":
                std = 3.0 * fan_in ** -0.5
            elif protocol == "P1" and (name.endswith("out_proj.weight") or name.endswith("c_proj.weight")):
                std = fan_in ** -0.5
            elif leaf == "in_proj_weight":
                std = (2.0 / (fan_in + shp[0])) ** 0.5          # xavier_uniform of nn.MultiheadAttention
            elif "backbone." in name and len(shp) == 4:
                std = (2.0 / fan_in) ** 0.5 if backbone == "RN50" else (3.0 * fan_in) ** -0.5
            else:
                std = (3.0 * fan_in) ** -0.5                    # U(-1/sqrt(fan_in), 1/sqrt(fan_in)) of nn.Linear
            w[name] = _normal(seed, name, shp, std)
    return w


def make_text_features(n_cls, D, seed=0):
    return torch.randn(n_cls, D, generator=_gen(seed, "text_features"))


def make_labels(way, shot, query_per_class, n_text_cls, seed):
    """Episode labels as the sampler builds them (video_reader.py:312-326): float tensors, shuffled order."""
    g = _gen(seed, "labels")
    sl = torch.arange(way).repeat_interleave(shot)
    sl = sl[torch.randperm(sl.numel(), generator=g)]
    tl = torch.arange(way).repeat_interleave(query_per_class)
    tl = tl[torch.randperm(tl.numel(), generator=g)]
    cls_map = torch.randperm(n_text_cls, generator=g)[:way]
    return dict(context_labels=sl.float(), target_labels=tl.long(), real_support_labels=cls_map[sl].float(),
                real_target_labels=cls_map[tl].float(), batch_class_list=cls_map.float())


def make_images(labels_support, labels_target, T, seed, protocol="P1", res=224):
    """float32 NCHW frames in [0,1].  P0: iid U[0,1).  P1: class-structured (per-class 14x14x3 prototype
    up-sampled x16, per-video offset +-0.075, per-frame class-specific drift 0.05*t, pixel noise +-0.025)."""
    g = _gen(seed, "images")

    def build(lbl):
        n = lbl.numel()
        if protocol == "P0":
            return torch.rand(n * T, 3, res, res, generator=g)
        out = torch.empty(n, T, 3, res, res)
        for i in range(n):
            c = int(lbl[i])
            gc = _gen(seed, "proto%d" % c)
            proto = torch.rand(3, res // 16, res // 16, generator=gc) * 0.6 + 0.2
            drift = torch.rand(3, res // 16, res // 16, generator=gc) - 0.5
            off = (torch.rand(3, 1, 1, generator=g) - 0.5) * 0.15
            for t in range(T):
                base = (proto + off + 0.05 * t * drift).repeat_interleave(16, 1).repeat_interleave(16, 2)
                out[i, t] = (base + (torch.rand(3, res, res, generator=g) - 0.5) * 0.05).clamp_(0, 1)
        return out.view(n * T, 3, res, res)

    return build(labels_support), build(labels_target)


def make_episode(seed, way=5, shot=1, query_per_class=1, T=8, n_text_cls=24, protocol="P1", images=True):
    ep = make_labels(way, shot, query_per_class, n_text_cls, seed)
    if images:
        s, t = make_images(ep["context_labels"], ep["target_labels"].float(), T, seed, protocol)
        ep["context_images"], ep["target_images"] = s, t
    return ep


def make_features(seed, n_support, n_query, T, D, labels_support=None, labels_target=None, sep=0.12):
    """Synthetic frame features for head-only cases: class centre + per-video + per-frame noise."""
    g = _gen(seed, "features")
    W = int(max(labels_support.max(), labels_target.max())) + 1 if labels_support is not None else 1
    centres = torch.randn(W, 1, D, generator=g)
    drift = torch.randn(W, 1, D, generator=g)

    def build(n, lbl):
        x = torch.randn(n, T, D, generator=g) * 0.5 + torch.randn(n, 1, D, generator=g) * 0.3
        if lbl is not None:
            tt = torch.arange(T).float().view(1, T, 1) / T
            x = x + sep * (centres[lbl.long()] + tt * drift[lbl.long()])
        return x

    return build(n_support, labels_support), build(n_query, labels_target)


# =====================================================================================================
# text-prompt tower: produces the text_features table the hot path consumes
# =====================================================================================================
TEXT = dict(width=512, layers=12, heads=8, ctx=77, vocab=49408)


def text_weight_shapes(embed_dim):
    C = TEXT["width"]
    s = {"token_embedding.weight": (TEXT["vocab"], C), "positional_embedding": (TEXT["ctx"], C),
         "ln_final.weight": (C,), "ln_final.bias": (C,), "text_projection": (C, embed_dim)}
    for i in range(TEXT["layers"]):
        p = "transformer.resblocks.%d." % i
        s.update({p + "attn.in_proj_weight": (3 * C, C), p + "attn.in_proj_bias": (3 * C,),
                  p + "attn.out_proj.weight": (C, C), p + "attn.out_proj.bias": (C,),
                  p + "ln_1.weight": (C,), p + "ln_1.bias": (C,), p + "ln_2.weight": (C,), p + "ln_2.bias": (C,),
                  p + "mlp.c_fc.weight": (4 * C, C), p + "mlp.c_fc.bias": (4 * C,),
                  p + "mlp.c_proj.weight": (C, 4 * C), p + "mlp.c_proj.bias": (C,)})
    return s


def make_text_weights(embed_dim=512, seed=0):
    """Seeded synthetic weights of CLIP's text tower, keyed by the names of the reference's `CLIP` module
    (models/clip_fsar.py:737-746); scales follow CLIP.initialize_parameters (:749-776) with non-trivial LN/biases."""
    C, L = TEXT["width"], TEXT["layers"]
    w = {}
    for name, shp in text_weight_shapes(embed_dim).items():
        leaf = name.split(".")[-1]
        nm = "text." + name
        if name == "token_embedding.weight":
            w[name] = _normal(seed, nm, shp, 0.02)
        elif name == "positional_embedding":
            w[name] = _normal(seed, nm, shp, 0.01)
        elif name == "text_projection":
            w[name] = _normal(seed, nm, shp, C ** -0.5)
        elif len(shp) == 1 and leaf == "weight":
            w[name] = 1.0 + _normal(seed, nm, shp, 0.1)
        elif leaf in ("bias", "in_proj_bias"):
            w[name] = _normal(seed, nm, shp, 0.02)
        elif leaf == "in_proj_weight":
            w[name] = _normal(seed, nm, shp, 2.0 * C ** -0.5)
        elif name.endswith("c_fc.weight"):
            w[name] = _normal(seed, nm, shp, (2 * C) ** -0.5)
        else:  # out_proj / c_proj
            w[name] = _normal(seed, nm, shp, C ** -0.5 * (2 * L) ** -0.5 * 4.0)
    return w


def encode_text(w, tokens):
    """models/clip_fsar.py:793-805 CLIP.encode_text: tokens [B,77] int -> [B, embed_dim]; causal mask of :778-784."""
    C, H = TEXT["width"], TEXT["heads"]
    B, L = tokens.shape
    x = w["token_embedding.weight"][tokens.long()] + w["positional_embedding"]
    mask = torch.full((L, L), float("-inf")).triu_(1)
    hd = C // H
    for i in range(TEXT["layers"]):
        p = "transformer.resblocks.%d." % i
        h = layer_norm(x, w[p + "ln_1.weight"], w[p + "ln_1.bias"])
        qkv = h @ w[p + "attn.in_proj_weight"].t() + w[p + "attn.in_proj_bias"]
        q, k, v = qkv.split(C, dim=-1)
        q = q.view(B, L, H, hd).transpose(1, 2)
        k = k.view(B, L, H, hd).transpose(1, 2)
        v = v.view(B, L, H, hd).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd) + mask, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, L, C)
        x = x + o @ w[p + "attn.out_proj.weight"].t() + w[p + "attn.out_proj.bias"]
        h = layer_norm(x, w[p + "ln_2.weight"], w[p + "ln_2.bias"])
        h = quick_gelu(h @ w[p + "mlp.c_fc.weight"].t() + w[p + "mlp.c_fc.bias"])
        x = x + h @ w[p + "mlp.c_proj.weight"].t() + w[p + "mlp.c_proj.bias"]
    x = layer_norm(x, w["ln_final.weight"], w["ln_final.bias"])
    eot = tokens.argmax(dim=-1)                                   # <|endoftext|> has the highest id
    return x[torch.arange(B), eot] @ w["text_projection"]


def class_text_features(w, tokens_by_template):
    """models/model_clipspm.py:52-70: tokens_by_template [n_templates, n_classes, 77] -> mean over templates of
    encode_text -> [n_classes, embed_dim] (no normalisation)."""
    return torch.stack([encode_text(w, t) for t in tokens_by_template]).mean(dim=0)


def head_loss_and_grads(w, text_features, su, qu, support_labels, real_support, real_target, target_labels, params,
                         single_direct=False, tasks_per_batch=16, dtype=torch.float32):
    """run/main_run.py:245-254 train_task on a head-only episode with dropout p = 0: loss = CE / TASKS_PER_BATCH + 0.001 * dists
    (:390-392) of head_forward (train mode only swaps the prompt table, model_clipspm.py:116-118, so `text_features` is the
    TRAIN table here), differentiated by torch autograd on the CPU with respect to every head parameter and the features.
    Returns (loss, {name: grad}) with the feature gradients under "su" / "qu"; parameters the loss does not reach are absent."""
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in w.items()
              if not k.startswith("backbone.") and v.dtype.is_floating_point}
    su_, qu_ = su.detach().to(dtype).clone().requires_grad_(True), qu.detach().to(dtype).clone().requires_grad_(True)
    st = head_forward(leaves, text_features.to(dtype), su_, qu_, support_labels, real_support, real_target, params,
                      single_direct)
    lg = st["logits"][0]
    ce = -(lg.log_softmax(-1).gather(1, target_labels.long().view(-1, 1)).squeeze(1))
    loss = ce.sum() / tasks_per_batch + 0.001 * st["dists"]
    loss.backward()
    grads = {k: v.grad.detach() for k, v in leaves.items() if v.grad is not None}
    grads["su"], grads["qu"] = su_.grad.detach(), qu_.grad.detach()
    return loss.detach(), grads


def train_loss_and_grads(w, text_features, inputs, cfg, tasks_per_batch=16):
    """run/main_run.py:245-254 train_task from the frames on (dropout p = 0): loss of :390-392 on forward() -- frame encoder
    AND head -- differentiated by torch autograd on the CPU with respect to every floating-point parameter in `w`.
    `text_features` is the TRAIN prompt table (model_clipspm.py:116-118).  Returns (loss, {name: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in w.items() if v.dtype.is_floating_point}
    st = forward(leaves, text_features, inputs, cfg)
    lg = st["logits"][0]
    ce = -(lg.log_softmax(-1).gather(1, inputs["target_labels"].long().view(-1, 1)).squeeze(1))
    loss = ce.sum() / tasks_per_batch + 0.001 * st["dists"]
    loss.backward()
    return loss.detach(), {k: v.grad.detach() for k, v in leaves.items() if v.grad is not None}


def fsar_head_loss_and_grads(w, text_train, su, qu, support_labels, real_support, real_target, target_labels,
                             tasks_per_batch, cls_value, single_direct=False, merge_before=False, depth=1):
    """CLIP-FSAR's training iteration on a head-only episode (models/model_clipfsar.py:183-262 train branch with
    MODEL.USE_CLASSIFICATION, dropout p = 0: the evaluation branch's arithmetic with the prompt rows taken from
    text_features_train, :197-198; loss run/main_run.py:355-356), differentiated by torch autograd.  Returns (loss, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in w.items()
              if not k.startswith("backbone.") and v.dtype.is_floating_point}
    su_, qu_ = su.detach().clone().requires_grad_(True), qu.detach().clone().requires_grad_(True)
    st = fsar_head_forward(leaves, text_train, text_train, su_, qu_, support_labels, real_support, real_target,
                           single_direct, merge_before, depth)
    lg, cl = st["logits"][0], st["class_logits"][0]
    ce = -(lg.log_softmax(-1).gather(1, target_labels.long().view(-1, 1)).squeeze(1)).sum()
    real = torch.cat([real_support, real_target]).long()
    ce_cls = -(cl.log_softmax(-1).gather(1, real.view(-1, 1)).squeeze(1)).sum()
    loss = (ce + cls_value * ce_cls) / tasks_per_batch
    loss.backward()
    grads = {k: v.grad.detach() for k, v in leaves.items() if v.grad is not None}
    grads["su"], grads["qu"] = su_.grad.detach(), qu_.grad.detach()
    return loss.detach(), grads


def grad_sample_index(numel, n=4096, seed=0):
    """fixed sample positions of a large gradient tensor stored in the goldens (the full head gradients are ~70 MB)"""
    if numel <= n:
        return torch.arange(numel)
    g = torch.Generator().manual_seed(seed * 7919 + numel)
    return torch.randperm(numel, generator=g)[:n].sort().values


def make_otam_grad_inputs(W, Q, T, D, seed):
    """seeded inputs of the metric-tail gradient golden (pin_against_reference.py otam_grad): support [W,T,D],
    target [Q,T,D] with a shared component (cosine similarities well above zero) and an upstream gradient [Q,W]"""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1, 1, D, generator=g)
    sup = torch.randn(W, T, D, generator=g) + 0.7 * base
    tgt = torch.randn(Q, T, D, generator=g) + 0.7 * base
    go = torch.randn(Q, W, generator=g)
    return sup, tgt, go
