"""Training step of the CLIP-SPM metric head (run/main_run.py:245-254 `train_task`:
`model(input)` in train mode, `scaler.scale(loss).backward()`), over FROZEN frame-encoder features.

What runs where: every dense contraction of the head and its gradient -- the two `Transformer_v1` blocks (forward and
backward as one library call each, csrc/tv1_train.cu), the gate / FeedForward / temporal-convolution linears (spm_gemm
forward, spm_linear_backward) and the metric tail (spm_otam_distance / _backward) -- is this library's CUDA code behind
`torch.autograd.Function`s; torch's autograd engine only walks the graph between them and does the small elementwise
glue (gathers, concatenations, means, the gating products, `_dis`, the cross-entropy).  The frame encoder's backward is
not built: `CNN` in train mode encodes the frames without a graph and differentiates the head, i.e. it trains the head
parameters of models/model_clipspm.py:72-99 with the CLIP tower frozen (the reference's optimiser also steps the tower).
Dropout: `Transformer_v1` / `FeedForward` carry nn.Dropout(0.2 / 0.05) in the reference (myRes.py:964-996,1053-1064);
with `dropout_seed=None` p = 0 (what the parity oracle pins against the reference's backward); with a seed the library
applies replayable Philox masks (spm_tv1_set_dropout / spm_dropout) at the reference's dropout sites -- the oracle
regenerates the same masks on the CPU (oracle.dropout_mask), torch's own generator stream cannot be reproduced."""
import ctypes

import torch

from . import _lib
from .ops import ACT, _need_cuda, _ptr, _stream, otam_distance

__all__ = ["linear", "dropout", "layer_norm", "TransformerV1", "VitBlock", "vit_forward", "spm_head_forward", "spm_loss",
           "fsar_head_forward", "fsar_loss", "shard_tasks", "allreduce_gradients", "run_listing_training", "GraphedStep"]


class _Linear(torch.autograd.Function):
    """y = act(x W^T + b): forward spm_gemm (tf32 tensor cores, or exact fp32), backward spm_linear_backward."""

    @staticmethod
    def forward(ctx, x, W, b, act, slope, exact):
        lib = _lib.load()
        M, K = x.shape
        N = W.shape[0]
        y = torch.empty(M, N, device=x.device)
        _lib.check(lib.spm_gemm(_stream(), 2 if exact else 1, _ptr(x), K, _ptr(W), K, M, N, K, _ptr(b), act, float(slope),
                                None, 0, 0, 0, 0, 0, 0, _ptr(y), N, 0))
        ctx.save_for_backward(x, W, b, y)
        ctx.cfg = (act, float(slope), bool(exact))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, W, b, y = ctx.saved_tensors
        act, slope, exact = ctx.cfg
        M, K = x.shape
        N = W.shape[0]
        dy = dy.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(W)
        db = torch.empty_like(b) if b is not None else None
        n_ws = lib.spm_linear_backward_workspace(M, N, K)
        ws = torch.empty(n_ws, device=x.device)
        _lib.check(lib.spm_linear_backward(_stream(), 1 if exact else 0, _ptr(x), _ptr(W), _ptr(b), _ptr(y), _ptr(dy), M, N, K,
                                           act, slope, _ptr(dx), _ptr(dW), _ptr(db), _ptr(ws), n_ws))
        return dx, dW, db, None, None, None


def linear(x, W, b=None, act="none", slope=0.0, exact=False):
    """Differentiable y = act(x W^T + b) over the last dim of x (nn.Linear [+ activation]); fp32 CUDA tensors."""
    _need_cuda(x, W, b)
    x2 = x.reshape(-1, x.shape[-1]).contiguous().float()
    y = _Linear.apply(x2, W.contiguous(), None if b is None else b.contiguous(), ACT[act], slope, exact)
    return y.view(*x.shape[:-1], W.shape[0])


class _Dropout(torch.autograd.Function):
    """nn.Dropout(p) with the library's replayable mask; the backward is the same kernel on the upstream gradient."""

    @staticmethod
    def forward(ctx, x, p, seed, site):
        y = torch.empty_like(x)
        _lib.check(_lib.load().spm_dropout(_stream(), _ptr(x), x.numel(), p, seed, site, _ptr(y)))
        ctx.cfg = (p, seed, site)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        _lib.check(_lib.load().spm_dropout(_stream(), _ptr(dy), dy.numel(), *ctx.cfg, _ptr(dx)))
        return dx, None, None, None


def dropout(x, p, seed, site=0):
    """y = x * keep / (1 - p) with keep a pure function of (seed, site, element index) (include/clipspm_b200.h)."""
    _need_cuda(x)
    if p <= 0.0:
        return x
    return _Dropout.apply(x.contiguous().float(), float(p), int(seed) & (2 ** 64 - 1), int(site))


class _TV1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pool, drop, *w):
        lib = _lib.load()
        h = pool._take()
        st = _stream()
        _lib.check(lib.spm_tv1_set_dropout(h, *drop))
        _lib.check(lib.spm_tv1_load_weights(h, st, *[_ptr(t) for t in w]))
        B, n, D = x.shape
        out = torch.empty_like(x)
        _lib.check(lib.spm_tv1_forward(h, st, _ptr(x), B, n, _ptr(out)))
        ctx.save_for_backward(x, *w)   # x must outlive the backward (the handle keeps its address)
        ctx.pool, ctx.h = pool, h
        return out

    @staticmethod
    def backward(ctx, go):
        lib = _lib.load()
        x, *w = ctx.saved_tensors
        ln_g, ln_b, wq, wk, wv, wout, bout, w0, b0, w3, b3 = w
        go = go.contiguous()
        gx = torch.empty_like(x)
        gqkv = torch.empty(3 * wq.shape[0], wq.shape[1], device=x.device)
        inner = wq.shape[0]
        g = [torch.empty_like(t) for t in (ln_g, ln_b)] + [gqkv[:inner], gqkv[inner:2 * inner], gqkv[2 * inner:]] + \
            [torch.empty_like(t) for t in (wout, bout, w0, b0, w3, b3)]
        _lib.check(lib.spm_tv1_backward(ctx.h, _stream(), _ptr(go), _ptr(gx), *[_ptr(t) for t in g]))
        ctx.pool._give(ctx.h)
        return (gx, None, None) + tuple(g)


class TransformerV1:
    """models/myRes.py:1053-1075 `Transformer_v1(dim, heads=8, dim_head_k=256, mlp_dim=2048, depth=1)` called as
    block(x, x, x), differentiable: forward and backward are one library call each (spm_tv1_forward / _backward).
    A library handle holds the activations of ONE forward, so the object keeps a pool: each call inside a graph takes a
    handle, its backward returns it (`reset()` reclaims the handles of graphs that were dropped without a backward)."""

    NAMES = ("0.norm.weight", "0.norm.bias", "0.fn.to_q.weight", "0.fn.to_k.weight", "0.fn.to_v.weight",
             "0.fn.to_out.0.weight", "0.fn.to_out.0.bias", "1.net.0.weight", "1.net.0.bias", "1.net.3.weight", "1.net.3.bias")

    def __init__(self, dim, heads=8, dim_head=256, mlp_dim=2048, exact=False, dropout_atte=0.2, dropout_ffn=0.05):
        self.cfg = (int(dim), int(heads), int(dim_head), int(mlp_dim), 1 if exact else 0)
        self.p = (float(dropout_atte), float(dropout_ffn))   # model_clipspm.py:80-81: dropout_atte=0.2, dropout_ffn default 0.05
        self._free, self._all = [], []

    def _take(self):
        if self._free:
            return self._free.pop()
        h = ctypes.c_void_p()
        _lib.check(_lib.load().spm_tv1_create(*self.cfg, ctypes.byref(h)))
        self._all.append(h)
        return h

    def _give(self, h):
        if not any(f is h for f in self._free):
            self._free.append(h)

    def reset(self):
        # creation order: the k-th call of a forward always gets the k-th handle, so a handle sees ONE shape and its workspace
        # stops growing after the first iteration (a regrow synchronises the device -- fatal inside a CUDA-graph capture)
        self._free = list(reversed(self._all))

    def close(self):
        lib = _lib.load()
        for h in self._all:
            lib.spm_tv1_destroy(h)
        self._free, self._all = [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, x, weights, prefix="layers.0.", dropout_seed=None):
        """x [n_seq, seq_len, D] (seq_len <= 48); weights: mapping with the reference's parameter names under `prefix`.
        dropout_seed: None = no dropout (eval, or the p = 0 parity mode), else the 64-bit seed of this call's masks."""
        _need_cuda(x)
        w = [weights[prefix + n].contiguous() for n in self.NAMES]
        drop = (0.0, 0.0, 0) if dropout_seed is None else (self.p[0], self.p[1], int(dropout_seed) & (2 ** 64 - 1))
        return _TV1.apply(x.contiguous().float(), self, drop, *w)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g, b):
        y = torch.empty_like(x)
        _lib.check(_lib.load().spm_layernorm_forward(_stream(), _ptr(x), x.shape[0], x.shape[1], _ptr(g), _ptr(b), _ptr(y)))
        ctx.save_for_backward(x, g)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g = ctx.saved_tensors
        dy = dy.contiguous()
        dx, dg, db = torch.empty_like(x), torch.empty_like(g), torch.empty_like(g)
        ws = torch.empty((x.shape[0] + 64) * x.shape[1], device=x.device)
        _lib.check(_lib.load().spm_layernorm_backward(_stream(), _ptr(x), _ptr(dy), _ptr(g), x.shape[0], x.shape[1], _ptr(dx),
                                                      _ptr(dg), _ptr(db), _ptr(ws)))
        return dx, dg, db


def layer_norm(x, weight, bias):
    """Differentiable nn.LayerNorm (eps 1e-5) over the last dim."""
    _need_cuda(x, weight, bias)
    y = _LayerNorm.apply(x.reshape(-1, x.shape[-1]).contiguous().float(), weight.contiguous(), bias.contiguous())
    return y.view(x.shape)


class _VitBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pool, *w):
        lib = _lib.load()
        h = pool._take()
        st = _stream()
        _lib.check(lib.spm_vitblock_load_weights(h, st, *[_ptr(t) for t in w]))
        out = torch.empty_like(x)
        _lib.check(lib.spm_vitblock_forward(h, st, _ptr(x), x.shape[0], _ptr(out)))
        ctx.save_for_backward(x, *w)
        ctx.pool, ctx.h = pool, h
        return out

    @staticmethod
    def backward(ctx, go):
        x, *w = ctx.saved_tensors
        go = go.contiguous()
        gx = torch.empty_like(x)
        g = [torch.empty_like(t) for t in w]
        _lib.check(_lib.load().spm_vitblock_backward(ctx.h, _stream(), _ptr(go), _ptr(gx), *[_ptr(t) for t in g]))
        ctx.pool._give(ctx.h)
        return (gx, None) + tuple(g)


class VitBlock(TransformerV1):
    """models/clip_fsar.py:622-643 `ResidualAttentionBlock` of the ViT-B/16 tower on x [F,197,768], differentiable
    (spm_vitblock_forward / _backward); same handle pool as TransformerV1 -- the 12 blocks of a tower take 12 handles."""

    NAMES = ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
             "attn.out_proj.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight",
             "mlp.c_proj.bias")

    def __init__(self, exact=False):
        self.exact = 1 if exact else 0
        self._free, self._all = [], []

    def _take(self):
        if self._free:
            return self._free.pop()
        h = ctypes.c_void_p()
        _lib.check(_lib.load().spm_vitblock_create(self.exact, ctypes.byref(h)))
        self._all.append(h)
        return h

    def __call__(self, x, weights, prefix):
        _need_cuda(x)
        if x.dim() != 3 or x.shape[1] != 197 or x.shape[2] != 768:
            raise RuntimeError("VitBlock takes [frames, 197, 768]")
        w = [weights[prefix + n].contiguous() for n in self.NAMES]
        return _VitBlockFn.apply(x.contiguous().float(), self, *w)


def vit_forward(w, images, blocks, prefix="backbone.", exact=False):
    """models/clip_fsar.py:672-689 VisionTransformer.forward, differentiable with respect to every tower parameter:
    images [F,3,224,224] -> [F,512].  conv1 (16x16 / stride 16, no bias) is a linear over the patch matrix."""
    Fn = images.shape[0]
    patches = images.float().view(Fn, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(Fn * 196, 768)   # column c*256+ky*16+kx
    x = linear(patches, w[prefix + "conv1.weight"].reshape(768, 768), None, exact=exact).view(Fn, 196, 768)      # :673-675
    cls = w[prefix + "class_embedding"].expand(Fn, 1, -1)
    x = torch.cat([cls, x], dim=1) + w[prefix + "positional_embedding"]                                       # :676-677
    x = layer_norm(x, w[prefix + "ln_pre.weight"], w[prefix + "ln_pre.bias"])                                 # :678
    for i in range(12):                                                                                       # :680-682
        x = blocks(x, w, prefix + "transformer.resblocks.%d." % i)
    x = layer_norm(x[:, 0, :], w[prefix + "ln_post.weight"], w[prefix + "ln_post.bias"])                      # :684
    return linear(x, w[prefix + "proj"].t().contiguous(), None, exact=exact)                                  # :686-687


# ------------------------------------------------------------------------------------------------------------------
# the CLIP-SPM head in train mode (models/model_clipspm.py:111-143 with self.training = True)
# ------------------------------------------------------------------------------------------------------------------
def _motion_feats(x, w, exact):
    """models/model_clipspm.py:169-191 get_motion_feats on x [N,T,D] -> [N,D]; the two Conv1d(k=3, padding=1) as im2col
    linears (weight [Cout, Cin, 3] -> [Cout, 3*Cin] with column tap*Cin + c)."""
    def conv(t, name):
        tp = torch.nn.functional.pad(t, (0, 0, 1, 1))
        col = torch.cat([tp[:, :-2], tp[:, 1:-1], tp[:, 2:]], dim=-1)
        W = w[name + ".weight"]
        return linear(col, W.permute(0, 2, 1).reshape(W.shape[0], -1), w[name + ".bias"], exact=exact)
    c = conv(conv(x, "motion_conv1"), "motion_conv2")
    f = c[:, 1:] - x[:, :-1]
    b = c[:, :-1] - x[:, 1:]
    return (0.5 * (f + b)).mean(1)


def _gate(x, w, p, slope, exact):
    """models/model_clipspm.py:88-99: Linear -> LeakyReLU -> Linear -> Sigmoid (activations fused into the GEMM epilogues)."""
    h = linear(x, w[p + "0.weight"], w[p + "0.bias"], "leaky_relu", slope, exact)
    return linear(h, w[p + "2.weight"], w[p + "2.bias"], "sigmoid", 0.0, exact)


def _feed_forward(x, w, p, exact, drop_p=0.0, seed=None):
    """models/myRes.py:984-996 FeedForward (Linear -> GELU -> Dropout -> Linear -> Dropout; seed None: p = 0)."""
    h = linear(x, w[p + "net.0.weight"], w[p + "net.0.bias"], "gelu", 0.0, exact)
    if seed is not None:
        h = dropout(h, drop_p, seed, 1)
    y = linear(h, w[p + "net.3.weight"], w[p + "net.3.bias"], "none", 0.0, exact)
    return dropout(y, drop_p, seed, 2) if seed is not None else y


def _dis(x, y):
    """models/model_clipspm.py:341-346."""
    d = x - y
    return (d * d).sum(dim=[-2, -1] if d.dim() == 3 else [-1]).mean()


def _class_mean_matrix(labels, way=None):
    """[W, S] averaging matrix of the sorted unique labels (models/myRes.py:730-739 extract_class_indices + mean).  way given:
    the labels are the episode's class indices 0 .. way-1 (video_reader.py:312-318), no torch.unique (a host sync)."""
    uniq = torch.unique(labels) if way is None else torch.arange(int(way), device=labels.device, dtype=labels.dtype)
    m = (labels.view(1, -1) == uniq.view(-1, 1)).float()
    return m / m.sum(1, keepdim=True), uniq


def _class_means(cm, x):
    """[W,S] x [S,T,D] -> [W,T,D] as a broadcast product + sum (elementwise glue, no library GEMM)."""
    return (cm.view(cm.shape[0], cm.shape[1], 1, 1) * x.unsqueeze(0)).sum(1)


def spm_head_forward(w, text_features, su, qu, support_labels, real_support, real_target, params, context1, context2,
                     single_direct=False, exact=False, dropout_seed=None, way=None):
    """models/model_clipspm.py:116-143 after get_feats on su [S,T,D], qu [Q,T,D], differentiable with respect to the head
    parameters `w` (reference names) and the features.  The four live `se_te` calls (:296-314) share one `context2` pass;
    the two whose outputs only reach the discarded consistency distances (:258-265) are skipped, as in the evaluation path.
    dropout_seed: None = p = 0; an int = the reference's train-mode dropout (token_tr 0.05, the blocks 0.2 / 0.05) with one
    independent mask stream per call site (seed + k)."""
    sd = (lambda k: None) if dropout_seed is None else (lambda k: int(dropout_seed) + k)
    S, T, D = su.shape
    Q = qu.shape[0]
    slope, alpha = float(params["negative_slope"]), float(params["alpha"])
    ctx_s = text_features[real_support.long()].unsqueeze(1)    # :117 (train table) / :120
    ctx_q = text_features[real_target.long()].unsqueeze(1)
    x = torch.cat([su, qu], dim=0)                                                     # [V,T,D]
    mo = _motion_feats(x, w, exact)                                                    # :194  [V,D]
    su_mo, qu_mo = mo[:S], mo[S:]
    token = torch.cat([ctx_q, ctx_s], dim=0).mean(dim=0)                               # :213-214 [1,D]
    target_token = _feed_forward(token.expand(Q, -1, -1) * qu.mean(dim=[1, 2], keepdim=True), w, "token_tr.mlp.", exact, 0.05, sd(0))
    # se_te x 4 (:196-197 on the motion tokens, :226,229 on the prompt tokens): [qu|su] frames with tokens [qu_mo|su_mo], then
    # [qu|su] with [target_token|ctx_s]
    tok = torch.cat([qu_mo.unsqueeze(1), su_mo.unsqueeze(1), target_token, ctx_s], dim=0)     # [2V,1,D]
    gv = _gate(x, w, "gate_vision.", slope, exact)                                     # once per frame set
    xg = x * gv
    xg = torch.cat([xg[S:], xg[:S], xg[S:], xg[:S]], dim=0)                            # [2V,T,D]
    q = tok * _gate(tok, w, "gate_text.", slope, exact) * alpha + xg
    z = context2(torch.cat([tok, q], dim=1), w, "context2.layers.0.", sd(1))           # [2V,T+1,D]
    zt, zf = z[:, 0, :], z[:, 1:, :]
    qu_m, su_m, qu_fake, su_real = zf[:Q], zf[Q:Q + S], zf[Q + S:2 * Q + S], zf[2 * Q + S:]
    qu_mo2, su_mo2 = zt[:Q], zt[Q:Q + S]
    new_m = _motion_feats(torch.cat([su_m, qu_m], dim=0), w, exact)                    # :199
    mo_dist = _dis(new_m[S:], qu_mo2) + _dis(new_m[:S], su_mo2)                        # :201-205
    cm, uniq = _class_mean_matrix(support_labels, way)
    W = cm.shape[0]
    su_pro = _class_means(cm, su_real)                                                 # :231-239
    class_dists_l = otam_distance(su_pro.unsqueeze(0), qu_fake.unsqueeze(0), single_direct)[0]        # :269 [Q,W]
    dists = w["mo_alpha1"] * mo_dist                                                   # :129
    # taskM (:275-294)
    K = (cm > 0).sum(1).view(-1, 1, 1).float()
    token_s = (su_pro * K + qu_fake.sum(0, keepdim=True)) / (K + Q)                    # :283 mean over [class members | queries]
    token_q = token_s.mean(dim=0, keepdim=True)
    su_t = torch.cat([token_s, su_real], dim=0).permute(1, 0, 2)                       # [T,W+S,D]
    qu_t = torch.cat([token_q, qu_fake], dim=0).permute(1, 0, 2)                       # [T,1+Q,D]
    _su = context1(su_t, w, "context1.layers.0.", sd(2)).permute(1, 0, 2)
    _qu = context1(qu_t, w, "context1.layers.0.", sd(3)).permute(1, 0, 2)
    su_2, qu_2, su_t2, qu_t2 = _su[W:], _qu[1:], _su[:W], _qu[0:1]
    su_pro2 = _class_means(cm, su_2)                                                   # :133-137
    task_dist = otam_distance(su_pro2.unsqueeze(0), qu_2.unsqueeze(0), single_direct)[0] + \
        otam_distance(su_t2.unsqueeze(0), qu_t2.unsqueeze(0), single_direct)[0]        # :138
    logits = -(0.5 * class_dists_l + task_dist).unsqueeze(0)                           # :141
    return {"logits": logits, "dists": dists}


def spm_loss(out, target_labels, tasks_per_batch=16.0):
    """utils/utils.py:174-186 `loss` on the single logit sample + run/main_run.py:390-392: CE summed over the queries /
    TASKS_PER_BATCH + 0.001 * dists."""
    lg = out["logits"][0]
    ce = -(torch.log_softmax(lg, dim=-1).gather(1, target_labels.long().view(-1, 1))).sum()
    return ce / tasks_per_batch + 0.001 * out["dists"]


# ------------------------------------------------------------------------------------------------------------------
# sibling head CLIP-FSAR in train mode (models/model_clipfsar.py:183-262)
# ------------------------------------------------------------------------------------------------------------------
def _cos_sim(x, y, exact, eps=0.01):
    """models/myRes.py:756-765 with the product on the library's GEMM (the table's rows padded to a multiple of 32)."""
    n = y.shape[0]
    yp = torch.nn.functional.pad(y, (0, 0, 0, (-n) % 32))
    num = linear(x, yp, None, exact=exact)[:, :n]
    den = x.norm(dim=-1).unsqueeze(-1) * y.norm(dim=-1).unsqueeze(0) + eps
    return num / den


def fsar_head_forward(w, text_train, su, qu, support_labels, real_support, real_target, context2, depth=1, single_direct=False,
                      merge_before=False, use_classification=True, exact=False, dropout_seed=None, way=None):
    """models/model_clipfsar.py:183-262 (the training branch: prompt rows from text_features_train :197-198) after get_feats on
    su [S,T,D], qu [Q,T,D]: `context2` over each query's frames and over each support's frames + its prompt, class-mean
    prototypes, OTAM; class_text_logits = cos_sim(mean_t feats, text_features_train) * scale when MODEL.USE_CLASSIFICATION."""
    S, T, D = su.shape
    sd = (lambda k: None) if dropout_seed is None else (lambda k: int(dropout_seed) + k)

    def ctx2(x, k):
        for i in range(depth):                                                                # myRes.py:1066-1075
            x = context2(x, w, "context2.layers.%d." % i, sd(10 * k + i))
        return x
    class_logits = None
    if use_classification:                                                                    # :187-190
        class_logits = (_cos_sim(torch.cat([su, qu], dim=0).mean(1), text_train, exact) * w["scale"]).unsqueeze(0)
    ctx = text_train[real_support.long()].unsqueeze(1)                                        # :197
    qu2 = ctx2(qu, 0)                                                                         # :201
    cm, _ = _class_mean_matrix(support_labels, way)
    if merge_before:                                                                          # :203-207
        su, ctx = _class_means(cm, su), _class_means(cm, ctx)
    su2 = ctx2(torch.cat([su, ctx], dim=1), 1)[:, :T]                                         # :208-209
    su_pro = su2 if merge_before else _class_means(cm, su2)                                   # :210-215
    cum = otam_distance(su_pro.unsqueeze(0), qu2.unsqueeze(0), single_direct)[0]              # :221-237
    out = {"logits": -cum.unsqueeze(0)}
    if class_logits is not None:
        out["class_logits"] = class_logits
    return out


def fsar_loss(out, target_labels, real_support, real_target, tasks_per_batch, cls_value):
    """run/main_run.py:355-356: (CE(logits) + USE_CLASSIFICATION_VALUE * CE(class_logits, cat[real_support, real_target])) /
    TASKS_PER_BATCH with utils/utils.py:174-186 `loss` (summed over the rows)."""
    ce = -(torch.log_softmax(out["logits"][0], dim=-1).gather(1, target_labels.long().view(-1, 1))).sum()
    real = torch.cat([real_support, real_target]).long().view(-1, 1)
    ce_cls = -(torch.log_softmax(out["class_logits"][0], dim=-1).gather(1, real)).sum()
    return (ce + cls_value * ce_cls) / tasks_per_batch


# ------------------------------------------------------------------------------------------------------------------
# several GPUs: the TASKS_PER_BATCH tasks of one optimiser step are split over the ranks
# ------------------------------------------------------------------------------------------------------------------
def shard_tasks(tasks_per_batch, rank, world):
    """Which of the TASKS_PER_BATCH tasks between two optimiser steps (run/main_run.py:203-209: gradients accumulate over
    them, the loss already carries the 1 / TASKS_PER_BATCH) this rank runs: a contiguous, balanced range."""
    lo = (tasks_per_batch * rank) // world
    hi = (tasks_per_batch * (rank + 1)) // world
    return range(lo, hi)


def allreduce_gradients(params, group=None, bucket_numel=1 << 24):
    """The training step's one exchange: SUM of the accumulated gradients over the ranks (NCCL all-reduce through
    torch.distributed, before `scaler.step`), in buckets of `bucket_numel` floats, so that every rank then takes the optimiser
    step one process would have taken on all TASKS_PER_BATCH tasks.  The reference's own multi-GPU form is DataParallel over
    the backbone (models/model_clipspm.py:103-109: frames of ONE task split over the GPUs); tasks are independent, so here
    the task batch is split instead and the exchange is one all-reduce per step.  A parameter must have a gradient on
    every rank or on none.  The scaled gradients are reduced as they are: inf / nan propagate to every rank, so all ranks
    skip the same steps."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    i = 0
    while i < len(grads):
        j, n = i, 0
        while j < len(grads) and (n == 0 or n + grads[j].numel() <= bucket_numel):
            n += grads[j].numel()
            j += 1
        if j == i + 1:
            dist.all_reduce(grads[i], op=dist.ReduceOp.SUM, group=group)
        else:
            flat = torch.cat([g.reshape(-1) for g in grads[i:j]])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            o = 0
            for g in grads[i:j]:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        i = j


# ------------------------------------------------------------------------------------------------------------------
# the training loop from a listing of decoded frames (run/main_run.py:180-243 `Learner.run`, training branch)
# ------------------------------------------------------------------------------------------------------------------
def run_listing_training(net, split, load_frame, iterations, way, shot, n_queries, optimizer, scaler, seed=0, flip=True,
                         lr_milestone=None, rank=0, world_size=1, on_iteration=None, graph=False):
    """`Learner.run`'s training branch over `VideoDataset` episodes on this library, from DECODED frames:
    per iteration (numbered from 1 like the reference's `iteration`) one train-mode episode is sampled
    (`frames.sample_episode_plan(train=True)`: class / video / frame-jitter draws AND the draws of the loader's training
    transform, rng seeded `seed + iteration` so that any split over ranks sees the same episodes), its frames come from
    `load_frame(handle) -> uint8 [H, W, 3]` (one size per run), Resize -> RandomHorizontalFlip -> RandomCrop -> ToTensor run on
    the GPU (`ops.transform_frames_train`, bit-exact with the loader's PIL chain), then `train_task` (:245-254):
    `net(inputs)` in train mode, `net.loss`, `scaler.scale(loss).backward()`; the optimiser steps when
    `(iteration + 1) % TASKS_PER_BATCH == 0` or on the last iteration (:203-209, the reference's own off-by-one), and
    `lr_scheduler.MultiStepLR(milestones=[lr_milestone], gamma=0.1).step()` follows every iteration (:99,210).
    world_size > 1: iteration i is computed by rank i % world_size, the gradients are exchanged once per optimiser step
    (`allreduce_gradients`).  `net` must be in train mode; `on_iteration(iteration, loss, acc)` is the logging / validation hook.
    graph=True (single rank, cfg.TRAIN.WAY set): the first TASKS_PER_BATCH window runs Python-driven (it warms the workspaces),
    then every iteration is ONE CUDA-graph replay (`GraphedStep`, re-captured when the learning rate changes); the trajectory is
    the Python-driven one when dropout is off, with dropout the masks come from the graph's device counter.
    Returns [(loss, accuracy)] of the iterations this rank computed (host floats: the reference logs both every iteration)."""
    import random

    from . import frames as F
    from . import ops
    if not net.training:
        raise RuntimeError("run_listing_training: call net.train() first")
    if graph and world_size == 1 and torch.cuda.current_stream(net._dev) == torch.cuda.default_stream(net._dev):
        # the Python-driven iterations that precede the capture must not run on the legacy default stream: autograd binds every
        # parameter's gradient accumulator to the stream of its first use, and a capture may not depend on the legacy stream
        side = torch.cuda.Stream(device=net._dev)
        side.wait_stream(torch.cuda.current_stream(net._dev))
        with torch.cuda.stream(side):
            log = run_listing_training(net, split, load_frame, iterations, way, shot, n_queries, optimizer, scaler, seed, flip,
                                       lr_milestone, rank, world_size, on_iteration, graph)
        torch.cuda.current_stream(net._dev).wait_stream(side)
        return log
    T, tpb = net.seq_len, int(net.tasks_per_batch)
    params = net.trainable_parameters()
    base_lr = optimizer.param_groups[0]["lr"]
    dev = net._dev
    log, size = [], None
    gstep, gstep_lr, graphed = None, None, False
    for iteration in range(1, int(iterations) + 1):
        graphed = False
        if iteration % world_size == rank:
            rng = random.Random(seed + iteration)
            if size is None:
                size = tuple(int(v) for v in torch.as_tensor(load_frame(split.videos[0][0])).shape[:2])
            plan = F.sample_episode_plan(split, way, shot, n_queries, T, train=True, rng=rng, frame_size=size, flip=flip)

            def clip_images(items):
                fr = torch.stack([torch.as_tensor(load_frame(split.videos[v][f])) for v, idx, _ in items for f in idx])
                aug = [[a[0], a[1], int(a[2])] for _, idx, a in items for _f in idx]
                return ops.transform_frames_train(fr.to(dev, non_blocking=True), aug)
            tl = torch.tensor([int(x) for x in plan["target_labels"]], dtype=torch.int64, device=dev)
            inputs = {"context_images": clip_images(plan["support"]), "target_images": clip_images(plan["target"]),
                      "context_labels": torch.tensor(plan["support_labels"], device=dev),
                      "real_support_labels": torch.tensor(plan["real_support_labels"], device=dev),
                      "real_target_labels": torch.tensor(plan["real_target_labels"], device=dev), "target_labels": tl}
            do_step = (iteration + 1) % tpb == 0 or iteration == int(iterations)
            if graph and world_size == 1 and iteration > tpb:
                lr_now = optimizer.param_groups[0]["lr"]
                if gstep is None or gstep_lr != lr_now:        # first graphed iteration, or the schedule moved the learning rate
                    if gstep is not None:
                        gstep.close()
                    gstep, gstep_lr = GraphedStep(net, optimizer, scaler, inputs, warmup=0, accumulate_graph=tpb > 1,
                                                  seed_base=seed + iteration), lr_now
                loss = gstep(inputs, optimizer_step=do_step)
                out = gstep.last_out
                graphed = True
            else:
                out = net(inputs)
                loss = net.loss(out, tl, inputs["real_support_labels"], inputs["real_target_labels"])
                scaler.scale(loss).backward()
                graphed = False
            acc = (out["logits"][0].argmax(-1) == tl).float().mean()
            log.append((float(loss.detach()), float(acc)))          # run/main_run.py:199-201 keeps both as host numbers
            if on_iteration is not None:
                on_iteration(iteration, log[-1][0], log[-1][1])
        if ((iteration + 1) % tpb == 0 or iteration == int(iterations)) and not (graphed and iteration % world_size == rank):
            if world_size > 1:
                allreduce_gradients(params)
            scaler.step(optimizer)
            scaler.update()
            optimizer.zero_grad(set_to_none=not graph)     # graph mode keeps the gradient tensors (stable addresses)
        if lr_milestone is not None:                                 # MultiStepLR, stepped once per iteration
            optimizer.param_groups[0]["lr"] = base_lr * (0.1 if iteration >= int(lr_milestone) else 1.0)
    if gstep is not None:
        gstep.close()
    return log


# ------------------------------------------------------------------------------------------------------------------
# the training iteration as ONE CUDA graph
# ------------------------------------------------------------------------------------------------------------------
class GraphedStep:
    """forward + loss + backward + optimiser step of one training iteration captured into a CUDA graph and replayed: the step is
    600-1500 launches driven from Python (autograd Functions over ctypes calls) and runs close to launch-bound -- a replay issues
    them from the device queue.  What makes the step capturable: the library never synchronises inside it (the GradScaler's
    skip-on-overflow is decided on the device, gradient tensors are kept so that the optimiser's pointer table does not change,
    workspaces are grown by the warm-up iterations), `cfg.TRAIN.WAY` replaces the torch.unique host syncs, and the dropout seed is
    a DEVICE counter incremented inside the graph (spm_dropout_seed_source), so every replay draws fresh masks.
    inputs: dict of tensors with fixed shapes (copied into static buffers at every call); forward(net, inputs) -> the model's
    output dict (default `net(inputs)`); the learning rate is baked in at capture (re-create the object after a schedule step).
    Call it with the next episode's tensors; returns the (device) loss of that iteration."""

    def __init__(self, net, optimizer, scaler, inputs, forward=None, warmup=3, accumulate_graph=False, seed_base=0x5eed):
        """warmup: eager iterations run before the capture (REAL training steps on `inputs`: they grow the workspaces and fill the
        optimiser's pointer table); 0 when the model has already trained on this shape.  accumulate_graph: also capture the
        iteration WITHOUT the optimiser step (gradient accumulation over TASKS_PER_BATCH tasks): `step(x, optimizer_step=False)`."""
        if not net.training:
            raise RuntimeError("GraphedStep: call net.train() first")
        if net.way is None:
            raise RuntimeError("GraphedStep needs cfg.TRAIN.WAY (torch.unique on the labels would synchronise inside the capture)")
        self.net, self.opt, self.scaler = net, optimizer, scaler
        self.forward = forward or (lambda n, x: n(x))
        dev = net._dev
        self.static = {k: v.detach().to(dev).clone() for k, v in inputs.items() if torch.is_tensor(v)}
        if net.text_features_train is not None:
            net.text_features_train = net.text_features_train.to(dev, torch.float32)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.check(_lib.load().spm_dropout_seed_source(_ptr(self.counter)))
        self._dropout = getattr(net, "train_dropout", True)
        net._graph_seed = int(seed_base) if self._dropout else None      # a constant base seed; the device counter moves it
        self.params = net.trainable_parameters()
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        if int(warmup) > 0:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(int(warmup)):
                    self._step(True)
            cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        import gc
        gc.collect()
        gc_was = gc.isenabled()
        gc.disable()     # a collected model would free its block handles (cudaFree) in the middle of the capture
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.out = self._step(True)
            self.graph_acc = None
            if accumulate_graph:
                self.graph_acc = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_acc):
                    self.loss_acc, self.out_acc = self._step(False)
        finally:
            if gc_was:
                gc.enable()
        self.last_out = self.out

    def _step(self, optimizer_step):
        self.counter += 1
        out = self.forward(self.net, self.static)
        loss = self.net.loss(out, self.static["target_labels"], self.static.get("real_support_labels"),
                             self.static.get("real_target_labels"))
        self.scaler.scale(loss).backward()        # accumulates into the kept .grad tensors
        if optimizer_step:                        # run/main_run.py:207-209
            self.scaler.step(self.opt)
            self.scaler.update()
            torch._foreach_zero_([p.grad for p in self.params])
        return loss.detach(), {k: v.detach() for k, v in out.items() if torch.is_tensor(v)}

    def __call__(self, inputs, optimizer_step=True):
        for k, v in inputs.items():
            if torch.is_tensor(v) and k in self.static:
                self.static[k].copy_(v, non_blocking=True)
        if optimizer_step:
            self.graph.replay()
            self.last_out = self.out
            return self.loss
        if self.graph_acc is None:
            raise RuntimeError("GraphedStep: created without accumulate_graph")
        self.graph_acc.replay()
        self.last_out = self.out_acc
        return self.loss_acc

    def close(self):
        """detach the device seed counter (dropout seeds are host arguments again)"""
        _lib.check(_lib.load().spm_dropout_seed_source(None))
        self.net._graph_seed = None
