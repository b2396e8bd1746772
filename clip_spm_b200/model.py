"""Host-side mirror of the reference's operator interface for the episode-evaluation path:

    reference                                   here
    models/model_clipspm.py:13  class CNN       class CNN(cfg, text_features_test=..., text_features_train=...)
    :111  forward(inputs: dict) -> dict         forward(inputs) -> {"logits": [1,Q,W], "dists": 0-d}   (same keys)
    models/model_clipfsar.py:49 forward(support_images, support_labels, target_images)  -> forward3(...)
    run/main_run.py:390-392 loss / accuracy     evaluate(inputs) -> loss, accuracy   (device scalars, no host sync)

`state_dict()` has exactly the reference's keys so `load_state_dict(reference_state_dict)` is the
weight hand-off.  All arithmetic happens in the C-ABI library (include/clipspm_b200.h); torch only owns the memory.
There is no CPU / eager fallback: without a CUDA device or without the built library this raises."""
import ctypes

import torch
import torch.nn as nn

from . import _lib

_FRAME = 3 * 224 * 224


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _fsar_head_shapes(D, depth=1):
    """Parameters of models/model_clipfsar.py::CNN_OTAM_CLIPFSAR besides the backbone (:137-145): `scale` and
    context2 = Transformer_v1(dim=D, heads=8, dim_head_k=D//8[, depth]): inner width D, mlp 2048."""
    s = {"scale": (1,)}
    for i in range(depth):
        p = "context2.layers.%d." % i
        s.update({p + "0.norm.weight": (D,), p + "0.norm.bias": (D,), p + "0.fn.to_q.weight": (D, D),
                  p + "0.fn.to_k.weight": (D, D), p + "0.fn.to_v.weight": (D, D), p + "0.fn.to_out.0.weight": (D, D),
                  p + "0.fn.to_out.0.bias": (D,), p + "1.net.0.weight": (2048, D), p + "1.net.0.bias": (2048,),
                  p + "1.net.3.weight": (D, 2048), p + "1.net.3.bias": (D,)})
    return s


def _param_shapes(backbone, D, params, head="clipspm", depth=1):
    """Names and shapes of the reference CNN's parameters/buffers (models/model_clipspm.py:72-99 and the CLIP visual
    tower of models/clip_fsar.py:549-689), written out here from the module definitions."""
    if head == "clipfsar":
        s = _fsar_head_shapes(D, depth)
    elif head == "cpm2c":
        s = _cpm2c_head_shapes(D, params)
    elif head == "sten":
        s = {}   # models/model_sten.py: every head module is commented out, only the backbone has parameters
    else:
        s = _spm_head_shapes(D, params)
    s.update(_backbone_shapes(backbone))
    return s


def _cpm2c_head_shapes(D, params):
    """Parameters models/model_cpm2c.py::CLIP_CPMMC_FSAR.forward reads (:75-81, :88-89, :103-114, :135-138): scale, context2
    (inner width D), the two class tokens, the gates, and the multi-scale motion convolutions with their 1x1 fusion."""
    ht, hv = int(D * params["mid_dim_text"]), int(D * params["mid_dim_vision"])
    s = _fsar_head_shapes(D)
    s.update({"class_token": (1, 1, D), "class_token_motion": (1, 1, D),
              "gate_text.0.weight": (ht, D), "gate_text.0.bias": (ht,), "gate_text.2.weight": (D, ht),
              "gate_text.2.bias": (D,), "gate_vision.0.weight": (hv, D), "gate_vision.0.bias": (hv,),
              "gate_vision.2.weight": (D, hv), "gate_vision.2.bias": (D,),
              "motion_conv1_1.weight": (D, D, 1), "motion_conv1_1.bias": (D,),
              "motion_conv1_3.weight": (D, D, 3), "motion_conv1_3.bias": (D,),
              "motion_conv1_5.weight": (D, D, 3), "motion_conv1_5.bias": (D,),
              "scale_conv.weight": (D, 3 * D, 1), "scale_conv.bias": (D,)})
    return s


def _spm_head_shapes(D, params):
    ht, hv = int(D * params["mid_dim_text"]), int(D * params["mid_dim_vision"])
    s = {"scale": (1,), "mo_alpha1": ()}
    for n in ("motion_conv1", "motion_conv2"):
        s[n + ".weight"] = (D, D, 3)
        s[n + ".bias"] = (D,)
    s.update({"token_tr.mlp.net.0.weight": (2048, D), "token_tr.mlp.net.0.bias": (2048,),
              "token_tr.mlp.net.3.weight": (D, 2048), "token_tr.mlp.net.3.bias": (D,)})
    for c in ("context1", "context2"):
        p = c + ".layers.0."
        s.update({p + "0.norm.weight": (D,), p + "0.norm.bias": (D,), p + "0.fn.to_q.weight": (2048, D),
                  p + "0.fn.to_k.weight": (2048, D), p + "0.fn.to_v.weight": (2048, D),
                  p + "0.fn.to_out.0.weight": (D, 2048), p + "0.fn.to_out.0.bias": (D,),
                  p + "1.net.0.weight": (2048, D), p + "1.net.0.bias": (2048,),
                  p + "1.net.3.weight": (D, 2048), p + "1.net.3.bias": (D,)})
    s.update({"gate_text.0.weight": (ht, D), "gate_text.0.bias": (ht,), "gate_text.2.weight": (D, ht),
              "gate_text.2.bias": (D,), "gate_vision.0.weight": (hv, D), "gate_vision.0.bias": (hv,),
              "gate_vision.2.weight": (D, hv), "gate_vision.2.bias": (D,)})
    return s


def _backbone_shapes(backbone):
    s = {}
    b = "backbone."
    if backbone == "ViT-B/16":
        C = 768
        s.update({b + "class_embedding": (C,), b + "positional_embedding": (197, C), b + "proj": (C, 512),
                  b + "conv1.weight": (C, 3, 16, 16), b + "ln_pre.weight": (C,), b + "ln_pre.bias": (C,),
                  b + "ln_post.weight": (C,), b + "ln_post.bias": (C,)})
        for i in range(12):
            p = b + "transformer.resblocks.%d." % i
            s.update({p + "attn.in_proj_weight": (3 * C, C), p + "attn.in_proj_bias": (3 * C,),
                      p + "attn.out_proj.weight": (C, C), p + "attn.out_proj.bias": (C,),
                      p + "ln_1.weight": (C,), p + "ln_1.bias": (C,), p + "mlp.c_fc.weight": (4 * C, C),
                      p + "mlp.c_fc.bias": (4 * C,), p + "mlp.c_proj.weight": (C, 4 * C), p + "mlp.c_proj.bias": (C,),
                      p + "ln_2.weight": (C,), p + "ln_2.bias": (C,)})
    else:
        wd = 64

        def bn(p, c):
            s.update({p + "weight": (c,), p + "bias": (c,), p + "running_mean": (c,), p + "running_var": (c,),
                      p + "num_batches_tracked": None})
        s[b + "conv1.weight"] = (wd // 2, 3, 3, 3); bn(b + "bn1.", wd // 2)
        s[b + "conv2.weight"] = (wd // 2, wd // 2, 3, 3); bn(b + "bn2.", wd // 2)
        s[b + "conv3.weight"] = (wd, wd // 2, 3, 3); bn(b + "bn3.", wd)
        inpl = wd
        for li, nb in enumerate((3, 4, 6, 3)):
            planes = wd * 2 ** li
            for bi in range(nb):
                p = b + "layer%d.%d." % (li + 1, bi)
                stride = 2 if (li > 0 and bi == 0) else 1
                s[p + "conv1.weight"] = (planes, inpl, 1, 1); bn(p + "bn1.", planes)
                s[p + "conv2.weight"] = (planes, planes, 3, 3); bn(p + "bn2.", planes)
                s[p + "conv3.weight"] = (planes * 4, planes, 1, 1); bn(p + "bn3.", planes * 4)
                if stride > 1 or inpl != planes * 4:
                    s[p + "downsample.0.weight"] = (planes * 4, inpl, 1, 1); bn(p + "downsample.1.", planes * 4)
                inpl = planes * 4
        E = wd * 32
        p = b + "attnpool."
        s[p + "positional_embedding"] = (50, E)
        for n in ("k_proj", "q_proj", "v_proj"):
            s[p + n + ".weight"] = (E, E); s[p + n + ".bias"] = (E,)
        s[p + "c_proj.weight"] = (1024, E); s[p + "c_proj.bias"] = (1024,)
    return s


class _Node(nn.Module):
    """Parameter container: gives the flat reference names their nested-module structure (no compute)."""


def _register(root, name, tensor, buffer=False):
    parts = name.split(".")
    mod = root
    for part in parts[:-1]:
        if not hasattr(mod, part):
            mod.add_module(part, _Node())
        mod = getattr(mod, part)
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _cfg_get(obj, path, default=None):
    for part in path.split("."):
        if isinstance(obj, dict):
            if part not in obj:
                return default
            obj = obj[part]
        else:
            if not hasattr(obj, part):
                return default
            obj = getattr(obj, part)
    return obj


class CNN(nn.Module):
    """Drop-in for models/model_clipspm.py::CNN on the evaluation path.

    cfg: the reference's config object (attribute tree or nested dict) -- reads MODEL.BACKBONE, DATA.SEQ_LEN,
    params{mid_dim_text, mid_dim_vision, negative_slope, alpha, motion_alpha}, optional MODEL.SINGLE_DIRECT,
    optional TRAIN.WAY (number of classes per episode; derived with torch.unique, a host sync, when absent).
    Prompt features: pass the [n_cls, D] tables the reference keeps in `text_features_test` /
    `text_features_train`, or call `build_text_features(clip_state_dict)` to compute them from the class names with
    the library's own text tower (the constructor work of model_clipspm.py:45-70).
    precision="bf16": bf16 tcgen05 encoder + tf32 head (what autocast(bfloat16) is to the reference);
    precision="bf16_resid" (ViT-B/16): additionally a bf16 residual stream, i.e. exactly what autocast(bfloat16) makes of
    the reference's tower (conv1, every Linear and every `x + ...` yield bf16); ~8 % faster, feature error ~1e-2;
    precision="fp32": every product in fp32 FFMA (the reference's default fp32 arithmetic; slow, for parity)."""

    HEAD = "clipspm"   # which metric head runs behind the shared entry points (SPM_HEAD_*)

    def __init__(self, cfg, text_features_test=None, text_features_train=None, max_episodes=1, device="cuda",
                 precision="bf16"):
        super().__init__()
        self.args = cfg
        if precision not in ("bf16", "fp32", "bf16_resid"):
            raise RuntimeError("precision must be 'bf16' (tensor-core path, fp32 residual stream), 'bf16_resid' (bf16 "
                               "residual stream: the reference's own autocast arithmetic) or 'fp32' (exact parity mode)")
        self.precision = precision
        self.backbone_name = _cfg_get(cfg, "MODEL.BACKBONE")
        if self.backbone_name not in ("ViT-B/16", "RN50"):
            raise RuntimeError("unsupported MODEL.BACKBONE %r" % (self.backbone_name,))
        self.mid_dim = 512 if self.backbone_name == "ViT-B/16" else 1024
        self.params = dict(_cfg_get(cfg, "params", None) or {})
        if self.HEAD in ("clipspm", "cpm2c") and not self.params:
            raise RuntimeError("cfg.params (mid_dim_text, mid_dim_vision, negative_slope, alpha) is required")
        self.seq_len = int(_cfg_get(cfg, "DATA.SEQ_LEN"))
        self.single_direct = bool(_cfg_get(cfg, "MODEL.SINGLE_DIRECT", False))
        self.way = _cfg_get(cfg, "TRAIN.WAY", None)
        self.tasks_per_batch = float(_cfg_get(cfg, "TRAIN.TASKS_PER_BATCH", 16))
        self.max_episodes = int(max_episodes)
        self._dev = torch.device(device)
        self.cls_value = float(_cfg_get(cfg, "MODEL.USE_CLASSIFICATION_VALUE", 0.0) or 0.0)
        # CLIP-FSAR only (models/model_clipfsar.py:143-144, :341): MODEL.TRANSFORMER_DEPTH is the switch, TRAIN.TRANSFORMER_DEPTH
        # the value the constructor reads; MODEL.MERGE_BEFORE averages each class before context2
        self.transformer_depth = 1
        if self.HEAD == "clipfsar" and _cfg_get(cfg, "MODEL.TRANSFORMER_DEPTH", None):
            self.transformer_depth = int(_cfg_get(cfg, "TRAIN.TRANSFORMER_DEPTH"))
        self.merge_before = self.HEAD == "clipfsar" and bool(_cfg_get(cfg, "MODEL.MERGE_BEFORE", False))
        for name, shape in _param_shapes(self.backbone_name, self.mid_dim, self.params, self.HEAD,
                                         self.transformer_depth).items():
            if shape is None:
                _register(self, name, torch.zeros((), dtype=torch.long), buffer=True)
            elif name.split(".")[-1] in ("running_mean", "running_var"):
                _register(self, name, torch.zeros(shape) if name.endswith("mean") else torch.ones(shape), buffer=True)
            else:
                _register(self, name, torch.zeros(shape))
        if hasattr(self, "scale"):
            self.scale.data.fill_(1.0)
        if hasattr(self, "mo_alpha1"):
            self.mo_alpha1.data.fill_(1.0)
        self.text_features_test = text_features_test
        self.text_features_train = text_features_train
        self._h = None
        self._packed_text = None
        self._train_ready = False    # None until here: nn.Module.__init__ -> train() must not touch the parameters
        self.eval()

    # ------------------------------------------------------------------------------------------------ plumbing
    def distribute_model(self):
        """models/model_clipspm.py:103-109 (DataParallel over the backbone) has no equivalent here: episodes are
        sharded across ranks instead (clip_spm_b200.sweep)."""
        return None

    def build_text_features(self, clip_state_dict, test_class_names=None, train_class_names=None, vocab_path=None,
                            tokenizer=None):
        """What the reference constructor does with the CLIP text tower (models/model_clipspm.py:45-70): sets
        `text_features_test` / `text_features_train` from class names (default: cfg.TEST.CLASS_NAME /
        cfg.TRAIN.CLASS_NAME).  clip_state_dict: the CLIP checkpoint's state_dict (text-tower keys)."""
        from .text import TextTower
        test_class_names = test_class_names or _cfg_get(self.args, "TEST.CLASS_NAME")
        train_class_names = train_class_names or _cfg_get(self.args, "TRAIN.CLASS_NAME")
        tower = TextTower(clip_state_dict, precision="fp32" if self.precision == "fp32" else "bf16", device=self._dev, vocab_path=vocab_path,
                          tokenizer=tokenizer)
        if tower.embed_dim != self.mid_dim:
            raise RuntimeError("text tower embed_dim %d does not match backbone %s" % (tower.embed_dim, self.backbone_name))
        try:
            if test_class_names:
                self.text_features_test = tower.class_features(list(test_class_names))
            if train_class_names:
                self.text_features_train = tower.class_features(list(train_class_names))
            torch.cuda.synchronize(self._dev)
        finally:
            tower.close()
        return self

    def init_random_(self, seed=0):
        """Random-init weights of the architecture for benchmarking (checkpoints are unavailable offline): fan-in
        scaled normals for matrices / conv kernels, LayerNorm-BatchNorm scales near 1, small biases."""
        if self._h is not None:
            raise RuntimeError("init_random_ must be called before the first forward")
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in list(self.named_parameters()) + list(self.named_buffers()):
                leaf = name.split(".")[-1]
                if not p.dtype.is_floating_point or name in ("scale", "mo_alpha1"):
                    continue
                if leaf == "running_var":
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5)
                elif p.dim() <= 1:
                    is_scale = leaf == "weight"
                    p.copy_(torch.randn(p.shape, generator=g) * (0.05 if is_scale else 0.02) + (1.0 if is_scale else 0.0))
                else:
                    fan_in = p[0].numel() if leaf != "proj" else p.shape[0]
                    p.copy_(torch.randn(p.shape, generator=g) * fan_in ** -0.5)
        return self

    def _handle(self):
        if self._h is not None:
            return self._h
        if not torch.cuda.is_available():
            raise RuntimeError("clip_spm_b200.CNN needs a CUDA device (sm_100a); there is no CPU fallback")
        lib = _lib.load()
        c = _lib.SpmConfig(
            backbone=0 if self.backbone_name == "ViT-B/16" else 1, seq_len=self.seq_len,
            n_text_classes=0 if self.text_features_test is None else int(self.text_features_test.shape[0]),
            mid_dim_text=float(self.params.get("mid_dim_text", 1.5)),
            mid_dim_vision=float(self.params.get("mid_dim_vision", 0.5)),
            negative_slope=float(self.params.get("negative_slope", 0.0)), alpha=float(self.params.get("alpha", 0.0)),
            single_direct=int(self.single_direct), precision={"bf16": 0, "fp32": 1, "bf16_resid": 2}[self.precision],
            max_episodes=self.max_episodes, max_support=0, max_query=0, max_way=0,
            head={"clipspm": 0, "clipfsar": 1, "sten": 2, "cpm2c": 3}[self.HEAD], cls_value=self.cls_value,
            **self._extra_config())
        h = ctypes.c_void_p()
        with torch.cuda.device(self._dev):
            _lib.check(lib.spm_create(ctypes.byref(c), ctypes.byref(h)))
            sd = {k: v.detach().to(self._dev, torch.float32).contiguous() for k, v in self.state_dict().items()
                  if v.dtype.is_floating_point}
            names = list(sd.keys())
            arr_n = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
            arr_p = (ctypes.c_void_p * len(names))(*[sd[n].data_ptr() for n in names])
            arr_e = (ctypes.c_int64 * len(names))(*[sd[n].numel() for n in names])
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            try:
                _lib.check(lib.spm_load_weights(h, st, len(names), arr_n, arr_p, arr_e))
            except Exception:
                lib.spm_destroy(h)
                raise
        self._h = h
        return h

    def _extra_config(self):
        """Head-specific trailing fields of spm_config (include/clipspm_b200.h)."""
        return {}

    def _text(self):
        tf = self.text_features_train if self.training else self.text_features_test  # model_clipspm.py:116-121
        if tf is None:
            raise RuntimeError("text_features_%s is not set" % ("train" if self.training else "test"))
        if self._packed_text is not tf:
            h = self._handle()
            t = tf.detach().to(self._dev, torch.float32).contiguous()
            _lib.check(_lib.load().spm_set_text_features(
                h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(t), t.shape[0], t.shape[1]))
            torch.cuda.current_stream().synchronize()  # `t` may be a temporary
            self._packed_text = tf
        return tf

    def load_state_dict(self, state_dict, strict=True):
        if self._h is not None:
            raise RuntimeError("load_state_dict must be called before the first forward (weights are packed once)")
        return super().load_state_dict(state_dict, strict=strict)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.load().spm_destroy(self._h)
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------- training
    def _head_param_names(self):
        return [n for n, _ in self.named_parameters() if not n.startswith("backbone.")]

    def train(self, mode=True):
        """nn.Module.train.  Train mode (CLIP-SPM head only) makes the head parameters CUDA leaves with requires_grad and routes
        `forward` / `head` through clip_spm_b200.train (differentiable head over the frozen frame encoder, run/main_run.py:
        245-254); leaving train mode drops the packed weights so that evaluation re-packs the trained ones."""
        was = self.training
        super().train(mode)
        if mode and not was and getattr(self, "_dev", None) is not None and self.HEAD in ("clipspm", "clipfsar") \
                and torch.cuda.is_available() \
                and getattr(self, "_train_ready", None) is not None:
            tower = bool(getattr(self, "train_backbone", False))
            if tower and self.backbone_name != "ViT-B/16":
                raise RuntimeError("train_backbone: only the ViT-B/16 tower has a backward (RN50: head-only training)")
            for n, p in self.named_parameters():
                if n.startswith("backbone.") and not tower:
                    continue
                if not p.is_cuda:
                    p.data = p.data.to(self._dev, torch.float32).contiguous()
                p.requires_grad_(n != "scale" or self.HEAD == "clipfsar")   # CLIP-SPM's forward never reads `scale`
            self._train_ready = True
        if not mode and was and getattr(self, "_train_ready", False) and self._h is not None:
            torch.cuda.synchronize(self._dev)
            _lib.load().spm_destroy(self._h)     # the handle holds a packed copy of the head weights
            self._h, self._packed_text = None, None
        return self

    def trainable_parameters(self):
        """The parameters train mode differentiates (the head of models/model_clipspm.py:72-99; the CLIP tower is frozen)."""
        return [p for p in self.parameters() if p.requires_grad]

    def _train_tower(self, images):
        """models/clip_fsar.py:672-689 with a graph (train_backbone): the differentiable tf32 / fp32 tower of clip_spm_b200.train"""
        from . import train as _train
        if not self.get_parameter("backbone.conv1.weight").requires_grad:
            raise RuntimeError("train_backbone was set after model.train(): set it first (train() prepares the tower's parameters)")
        if getattr(self, "_vitblk", None) is None:
            self._vitblk = _train.VitBlock(exact=self.precision == "fp32")
        return _train.vit_forward(dict(self.named_parameters()), images, self._vitblk, "backbone.", self.precision == "fp32")

    def _train_blocks(self):
        if getattr(self, "_tv1", None) is None:
            from .train import TransformerV1
            exact = self.precision == "fp32"
            if self.HEAD == "clipfsar":    # model_clipfsar.py:143-146: dim_head_k = mid_dim // 8
                self._tv1 = (TransformerV1(self.mid_dim, dim_head=self.mid_dim // 8, exact=exact),)
            else:
                self._tv1 = (TransformerV1(self.mid_dim, exact=exact), TransformerV1(self.mid_dim, exact=exact))
        return self._tv1

    def _reset_train_pools(self):
        """a new train-mode forward starts: block handles still held by a graph that never ran its backward are reclaimed"""
        for b in (getattr(self, "_tv1", None) or ()) + ((self._vitblk,) if getattr(self, "_vitblk", None) is not None else ()):
            b.reset()

    def _train_head(self, su, qu, lab, rs, rt):
        from . import train as _train
        if self.HEAD not in ("clipspm", "clipfsar"):
            raise RuntimeError("train mode is implemented for the CLIP-SPM and CLIP-FSAR heads only")
        if not getattr(self, "_train_ready", False):
            raise RuntimeError("call model.train() on a CUDA machine before the train-mode forward")
        if self.text_features_train is None:
            raise RuntimeError("text_features_train is not set")          # model_clipspm.py:116-118
        text = self.text_features_train.to(self._dev, torch.float32)
        w = dict(self.named_parameters())
        blocks = self._train_blocks()
        # train-mode dropout (myRes.py:961-996): one fresh 62-bit seed per forward from torch's CPU generator, so that
        # torch.manual_seed reproduces a run; `model.train_dropout = False` gives the p = 0 head the parity goldens pin
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if getattr(self, "train_dropout", True) else None
        if getattr(self, "_graph_seed", None) is not None:     # train.GraphedStep: constant base seed + a device counter
            seed = self._graph_seed
        if self.HEAD == "clipfsar":
            return _train.fsar_head_forward(w, text, su, qu, lab, rs, rt, blocks[0], self.transformer_depth, self.single_direct,
                                            self.merge_before, bool(_cfg_get(self.args, "MODEL.USE_CLASSIFICATION", False)),
                                            self.precision == "fp32", seed, self.way)
        c1, c2 = blocks
        return _train.spm_head_forward(w, text, su, qu, lab, rs, rt, self.params, c1, c2, self.single_direct,
                                       self.precision == "fp32", seed, self.way)

    def loss(self, out, target_labels, real_support_labels=None, real_target_labels=None):
        """The runner's loss on a train-mode output (differentiable): CLIP-SPM run/main_run.py:390-392 CE / TASKS_PER_BATCH +
        0.001 * dists; CLIP-FSAR :355-356 (CE + USE_CLASSIFICATION_VALUE * CE(class_logits, real labels)) / TASKS_PER_BATCH."""
        from . import train as _train
        if self.HEAD == "clipfsar":
            return _train.fsar_loss(out, target_labels.to(self._dev), real_support_labels.to(self._dev).view(-1),
                                    real_target_labels.to(self._dev).view(-1), self.tasks_per_batch, self.cls_value)
        return _train.spm_loss(out, target_labels.to(self._dev), self.tasks_per_batch)

    def _way(self, labels):
        if self.way is not None:
            return int(self.way)
        return int(torch.unique(labels).numel())  # host sync, like the reference's own torch.unique (:133)

    def _f32(self, t):
        return t.to(self._dev, torch.float32).contiguous()

    # ------------------------------------------------------------------------------------------------- forward
    def forward(self, inputs):
        """models/model_clipspm.py:111-144: one episode dict -> {"logits": [1,Q,W], "dists": 0-d}.
        HOST image tensors (the DataLoader's pinned batch, before run/main_run.py:prepare_task's `.to(device)`) take the
        copy-overlapped host call: the episode's frames go host->device in chunks while earlier chunks encode, and the
        outputs come back as host tensors (plus "loss" / "acc" when the dict carries target_labels).  An optional
        "next_images" = (context_images, target_images) of the NEXT episode (a look-ahead of one over the DataLoader) lets this
        call copy them behind its own, so that the next call starts computing at once (spm_eval_host_set_next)."""
        ci = inputs["context_images"]
        if not ci.is_cuda and not self.training and self.HEAD == "clipspm":
            c = lambda t, dt=torch.float32: t.detach().to("cpu", dt).contiguous()   # noqa: E731
            lab = c(inputs["context_labels"]).view(-1)
            tl = inputs.get("target_labels")
            rt = c(inputs["real_target_labels"]).view(-1)
            r = self.evaluate_host(c(ci), lab, c(inputs["target_images"]), c(inputs["real_support_labels"]).view(-1), rt,
                                   torch.zeros(rt.numel(), dtype=torch.int64) if tl is None else c(tl, torch.int64).view(-1),
                                   1, self._way(lab), next_images=inputs.get("next_images"))
            out = {"logits": r["logits"][0].unsqueeze(0), "dists": r["dists"][0]}
            if tl is not None:
                out.update(loss=r["loss"][0], acc=r["acc"][0])
            return out
        out = self.forward_episodes(inputs["context_images"], inputs["context_labels"], inputs["target_images"],
                                    inputs["real_support_labels"], inputs["real_target_labels"], n_episodes=1)
        return {"logits": out["logits"][0].unsqueeze(0), "dists": out["dists"][0]}

    def forward3(self, support_images, support_labels, target_images, real_support_labels, real_target_labels):
        """The 3-tensor form of models/model_clipfsar.py:49 plus the two label tensors CLIP-SPM really consumes."""
        return self.forward(dict(context_images=support_images, context_labels=support_labels,
                                 target_images=target_images, real_support_labels=real_support_labels,
                                 real_target_labels=real_target_labels))

    def forward_episodes(self, context_images, context_labels, target_images, real_support_labels,
                         real_target_labels, n_episodes=1, target_labels=None):
        """`n_episodes` episodes stacked along dim 0 of every tensor (images [E*S*T,3,224,224] etc.).
        Returns logits [E,Q,W], dists [E] and, with target_labels, loss [E], acc [E], pred [E,Q]."""
        if self.training:
            # run/main_run.py:245-254: the frame encoder runs without a graph (frozen), the head is differentiable
            if int(n_episodes) != 1:
                raise RuntimeError("train mode takes one episode per call, like the reference's train_task")
            T, D = self.seq_len, self.mid_dim
            self._reset_train_pools()
            if getattr(self, "train_backbone", False):
                ns = self._f32(context_images).view(-1, 3, 224, 224).shape[0]
                feats = self._train_tower(torch.cat([self._f32(context_images).view(-1, 3, 224, 224),
                                                     self._f32(target_images).view(-1, 3, 224, 224)], dim=0))
                su, qu = feats[:ns].view(-1, T, D), feats[ns:].view(-1, T, D)
            else:
                with torch.no_grad():
                    su = self.encode_frames(self._f32(context_images).view(-1, 3, 224, 224)).view(-1, T, D)
                    qu = self.encode_frames(self._f32(target_images).view(-1, 3, 224, 224)).view(-1, T, D)
            out = self._train_head(su, qu, self._f32(context_labels).view(-1), self._f32(real_support_labels).view(-1),
                                   self._f32(real_target_labels).view(-1))
            out = dict(out)
            if target_labels is not None:
                lg = out["logits"][0]
                tl = target_labels.to(self._dev).long().view(-1)
                out.update(loss=self.loss(out, tl, real_support_labels, real_target_labels).view(1),
                           pred=lg.argmax(-1).view(1, -1).int(), acc=(lg.argmax(-1) == tl).float().mean().view(1))
            if "dists" in out:
                out["dists"] = out["dists"].view(1)
            return out
        h = self._handle()
        self._text()
        lib = _lib.load()
        E, T = int(n_episodes), self.seq_len
        lab = self._f32(context_labels).view(E, -1)
        S = lab.shape[1]
        rs, rt = self._f32(real_support_labels).view(E, -1), self._f32(real_target_labels).view(E, -1)
        Q = rt.shape[1]
        W = self._way(lab[0])
        su, qu = self._f32(context_images), self._f32(target_images)
        if su.numel() != E * S * T * _FRAME or qu.numel() != E * Q * T * _FRAME:
            raise RuntimeError("image tensors do not match [E*S*T,3,224,224] / [E*Q*T,3,224,224]")
        logits = torch.empty(E, Q, W, device=self._dev)
        dists = torch.empty(E, device=self._dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        out = {"logits": logits, "dists": dists}
        if target_labels is None:
            _lib.check(lib.spm_forward(h, st, E, S, Q, W, _p(su), _p(qu), _p(lab), _p(rs), _p(rt), _p(logits),
                                       _p(dists)))
        else:
            tl = target_labels.to(self._dev, torch.int64).contiguous()
            loss, acc = torch.empty(E, device=self._dev), torch.empty(E, device=self._dev)
            pred = torch.empty(E, Q, device=self._dev, dtype=torch.int32)
            _lib.check(lib.spm_eval(h, st, E, S, Q, W, _p(su), _p(qu), _p(lab), _p(rs), _p(rt), _p(tl),
                                    self.tasks_per_batch, _p(logits), _p(dists), _p(loss), _p(acc), _p(pred)))
            out.update(loss=loss, acc=acc, pred=pred)
        return out

    def evaluate(self, inputs):
        """forward + run/main_run.py:390-392: returns (loss, accuracy) as device scalars."""
        out = self.forward_episodes(inputs["context_images"], inputs["context_labels"], inputs["target_images"],
                                    inputs["real_support_labels"], inputs["real_target_labels"], 1,
                                    inputs["target_labels"])
        return out["loss"][0], out["acc"][0]

    def _set_next(self, lib, h, next_images, support_frames_per_episode):
        """next_images = (context_images, target_images) host tensors of the NEXT evaluate_host call (a prefetching
        DataLoader knows them; same S, Q and frame format as this call, any number of episodes): their first chunk
        is copied to the device while this call's tail computes."""
        if next_images is not None:
            a, b = next_images
            assert not a.is_cuda and not b.is_cuda and a.is_contiguous() and b.is_contiguous()
            _lib.check(lib.spm_eval_host_set_next(h, _p(a), _p(b), int(a.shape[0]) // int(support_frames_per_episode)))

    def evaluate_host(self, context_images, context_labels, target_images, real_support_labels, real_target_labels,
                      target_labels, n_episodes, way, next_images=None):
        """Episodes in HOST memory (pinned for overlap): chunked H2D on a copy stream overlapped with compute
        (spm_eval_host); returns host tensors.  This is the call the end-to-end benchmark times."""
        h = self._handle()
        self._text()
        lib = _lib.load()
        E, T = int(n_episodes), self.seq_len
        S, Q, W = context_labels.numel() // E, real_target_labels.numel() // E, int(way)
        self._set_next(lib, h, next_images, S * T)
        for t in (context_images, target_images, context_labels, real_support_labels, real_target_labels):
            assert not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        assert target_labels.dtype == torch.int64 and not target_labels.is_cuda
        logits, dists = torch.empty(E, Q, W), torch.empty(E)
        loss, acc, pred = torch.empty(E), torch.empty(E), torch.empty(E, Q, dtype=torch.int32)
        with torch.cuda.device(self._dev):
            # the host call runs on the handle's own streams: earlier device-tensor calls of this handle (caller's stream) must be done
            torch.cuda.current_stream(self._dev).synchronize()
            _lib.check(lib.spm_eval_host(h, E, S, Q, W, _p(context_images), _p(target_images), _p(context_labels),
                                         _p(real_support_labels), _p(real_target_labels), _p(target_labels),
                                         self.tasks_per_batch, _p(logits), _p(dists), _p(loss), _p(acc), _p(pred)))
        return dict(logits=logits, dists=dists, loss=loss, acc=acc, pred=pred)

    def evaluate_host_u8(self, context_frames, context_labels, target_frames, real_support_labels,
                         real_target_labels, target_labels, n_episodes, way, next_images=None):
        """evaluate_host on DECODED frames: uint8 [E*S*T, H, W, 3] / [E*Q*T, H, W, 3] host tensors (what the data
        loader holds before its PIL Resize/CenterCrop/ToTensor chain, video_reader.py:265-272).  The transform runs
        on the GPU, bit-exact with that chain; host->device traffic drops from 602 KB to H*W*3 bytes per frame."""
        h = self._handle()
        self._text()
        lib = _lib.load()
        E = int(n_episodes)
        S, Q, W = context_labels.numel() // E, real_target_labels.numel() // E, int(way)
        self._set_next(lib, h, next_images, S * self.seq_len)
        for t in (context_frames, target_frames):
            assert not t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous() and t.dim() == 4 and t.shape[3] == 3
        assert context_frames.shape[1:] == target_frames.shape[1:]
        assert context_frames.shape[0] == E * S * self.seq_len and target_frames.shape[0] == E * Q * self.seq_len
        for t in (context_labels, real_support_labels, real_target_labels):
            assert not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        assert target_labels.dtype == torch.int64 and not target_labels.is_cuda
        H, Wd = int(context_frames.shape[1]), int(context_frames.shape[2])
        logits, dists = torch.empty(E, Q, W), torch.empty(E)
        loss, acc, pred = torch.empty(E), torch.empty(E), torch.empty(E, Q, dtype=torch.int32)
        with torch.cuda.device(self._dev):
            torch.cuda.current_stream(self._dev).synchronize()      # see evaluate_host
            _lib.check(lib.spm_eval_host_u8(h, E, S, Q, W, H, Wd, _p(context_frames), _p(target_frames),
                                            _p(context_labels), _p(real_support_labels), _p(real_target_labels),
                                            _p(target_labels), self.tasks_per_batch, _p(logits), _p(dists), _p(loss),
                                            _p(acc), _p(pred)))
        return dict(logits=logits, dists=dists, loss=loss, acc=acc, pred=pred)

    def evaluate_frames_u8(self, context_frames, context_labels, target_frames, real_support_labels,
                           real_target_labels, target_labels, n_episodes=1):
        """evaluate on DECODED frames already on the device (uint8 [E*S*T, H, W, 3] / [E*Q*T, H, W, 3], e.g. from
        ops.decode_jpegs): transform + encoder + head + loss / accuracy in one library call (spm_eval_u8)."""
        h = self._handle()
        self._text()
        lib = _lib.load()
        E = int(n_episodes)
        for t in (context_frames, target_frames):
            if not t.is_cuda or t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
                raise RuntimeError("evaluate_frames_u8 takes CUDA uint8 tensors [F, H, W, 3]")
        su, qu = context_frames.contiguous(), target_frames.contiguous()
        lab = self._f32(context_labels).view(E, -1)
        rs, rt = self._f32(real_support_labels).view(E, -1), self._f32(real_target_labels).view(E, -1)
        S, Q, W = lab.shape[1], rt.shape[1], self._way(lab[0])
        if su.shape[0] != E * S * self.seq_len or qu.shape[0] != E * Q * self.seq_len or su.shape[1:] != qu.shape[1:]:
            raise RuntimeError("frame tensors do not match [E*S*T, H, W, 3] / [E*Q*T, H, W, 3]")
        tl = target_labels.to(self._dev, torch.int64).contiguous()
        logits, dists = torch.empty(E, Q, W, device=self._dev), torch.empty(E, device=self._dev)
        loss, acc = torch.empty(E, device=self._dev), torch.empty(E, device=self._dev)
        pred = torch.empty(E, Q, device=self._dev, dtype=torch.int32)
        _lib.check(lib.spm_eval_u8(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), E, S, Q, W,
                                   int(su.shape[1]), int(su.shape[2]), _p(su), _p(qu), _p(lab), _p(rs), _p(rt), _p(tl),
                                   self.tasks_per_batch, _p(logits), _p(dists), _p(loss), _p(acc), _p(pred)))
        return dict(logits=logits, dists=dists, loss=loss, acc=acc, pred=pred)

    def evaluate_jpeg(self, context_files, context_labels, target_files, real_support_labels, real_target_labels,
                      target_labels, n_episodes=1):
        """The reference's whole per-episode input path from the FILES (video_reader.py:227-273: PIL decode ->
        Resize(256) -> CenterCrop(224) -> ToTensor) + forward + loss / accuracy: `context_files` / `target_files` are
        lists of JPEG file contents (bytes) in the sampler's stacking order (frames.sample_episode_plan)."""
        from . import ops
        su = ops.decode_jpegs(list(context_files), self._dev)
        qu = ops.decode_jpegs(list(target_files), self._dev)
        return self.evaluate_frames_u8(su, context_labels, qu, real_support_labels, real_target_labels, target_labels,
                                       n_episodes)

    # --------------------------------------------------------------------------------------------- stage hooks
    def encode_frames_u8(self, frames):
        """decoded RGB frames uint8 [F, H, W, 3] (device) -> Resize/CenterCrop/ToTensor -> encoder -> [F, D]"""
        h = self._handle()
        if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
            raise RuntimeError("encode_frames_u8 takes a CUDA uint8 tensor [F, H, W, 3]")
        frames = frames.contiguous()
        out = torch.empty(frames.shape[0], self.mid_dim, device=self._dev)
        _lib.check(_lib.load().spm_encode_frames_u8(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                                    _p(frames), frames.shape[0], frames.shape[1], frames.shape[2],
                                                    _p(out)))
        return out

    def encode_frames(self, images):
        """models/clip_fsar.py:672-689 / :593-608: [F,3,224,224] -> [F, D] (stage entry point for the tests)."""
        h = self._handle()
        img = self._f32(images)
        out = torch.empty(img.shape[0], self.mid_dim, device=self._dev)
        _lib.check(_lib.load().spm_encode_frames(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(img),
                                                 img.shape[0], _p(out)))
        return out

    def head(self, su, qu, context_labels, real_support_labels, real_target_labels, n_episodes=1):
        """models/model_clipspm.py:125-143 on precomputed features su [E,S,T,D], qu [E,Q,T,D]."""
        if self.training:
            if int(n_episodes) != 1:
                raise RuntimeError("train mode takes one episode per call, like the reference's train_task")
            T, D = self.seq_len, self.mid_dim
            self._reset_train_pools()
            return self._train_head(self._f32(su).view(-1, T, D), self._f32(qu).view(-1, T, D), self._f32(context_labels).view(-1),
                                    self._f32(real_support_labels).view(-1), self._f32(real_target_labels).view(-1))
        h = self._handle()
        self._text()
        E = int(n_episodes)
        su, qu = self._f32(su), self._f32(qu)
        lab = self._f32(context_labels).view(E, -1)
        rs, rt = self._f32(real_support_labels).view(E, -1), self._f32(real_target_labels).view(E, -1)
        S, Q, W = lab.shape[1], rt.shape[1], self._way(lab[0])
        logits, dists = torch.empty(E, Q, W, device=self._dev), torch.empty(E, device=self._dev)
        _lib.check(_lib.load().spm_head(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), E, S, Q, W,
                                        _p(su), _p(qu), _p(lab), _p(rs), _p(rt), _p(logits), _p(dists)))
        return {"logits": logits, "dists": dists}

    def head_stage(self, name):
        """Test hook: a named intermediate tensor of the most recent head pass (spm_head_stage), flat [numel]."""
        h = self._handle()
        lib = _lib.load()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        n = ctypes.c_longlong(0)
        _lib.check(lib.spm_head_stage(h, st, name.encode(), None, 0, ctypes.byref(n)))
        out = torch.empty(n.value, device=self._dev)
        _lib.check(lib.spm_head_stage(h, st, name.encode(), _p(out), n.value, ctypes.byref(n)))
        return out
