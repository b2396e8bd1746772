"""Text-prompt tower: class names -> the `text_features_{train,test}` the episode head consumes.

Mirror of what the reference's constructor does once per class list (models/model_clipspm.py:45-70): every class
name is put into the 16 prompt templates, tokenised (clip_fsar.py:144-180), run through CLIP.encode_text
(clip_fsar.py:793-805) and averaged over the templates.  Tokenisation is host work (tokenizer.py); everything from
token ids on runs in the CUDA library (csrc/text_tower.cu) -- there is no CPU path."""
import ctypes

import torch

from . import _lib
from .tokenizer import CONTEXT_LENGTH, PROMPT_TEMPLATES, ClipTokenizer

_TEXT_PREFIXES = ("token_embedding.", "positional_embedding", "transformer.", "ln_final.", "text_projection")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class TextTower:
    """clip_state_dict: the CLIP checkpoint's state_dict (the keys of clip_fsar.CLIP; the `visual.*` entries are
    ignored).  embed_dim is read from text_projection (512 for ViT-B/16, 1024 for RN50)."""

    def __init__(self, clip_state_dict, precision="bf16", device="cuda", vocab_path=None, tokenizer=None):
        if not torch.cuda.is_available():
            raise RuntimeError("clip_spm_b200.TextTower needs a CUDA device (sm_100a); there is no CPU fallback")
        if precision not in ("bf16", "fp32"):
            raise RuntimeError("precision must be 'bf16' (tf32 tensor-core GEMMs) or 'fp32' (exact fp32 FFMA)")
        lib = _lib.load()
        self._dev = torch.device(device)
        self._h = None
        self._tok = tokenizer            # anything with .tokenize(list of str) -> [n, 77] ids; default: ClipTokenizer
        self._vocab_path = vocab_path
        if "text_projection" not in clip_state_dict:
            raise RuntimeError("clip_state_dict has no 'text_projection' (expected the CLIP checkpoint's state_dict)")
        self.embed_dim = int(clip_state_dict["text_projection"].shape[1])
        h = ctypes.c_void_p()
        with torch.cuda.device(self._dev):
            _lib.check(lib.spm_text_create(self.embed_dim, 0 if precision == "bf16" else 1, ctypes.byref(h)))
            sd = {k: v.detach().to(self._dev, torch.float32).contiguous() for k, v in clip_state_dict.items()
                  if k.startswith(_TEXT_PREFIXES) and v.dtype.is_floating_point}
            names = list(sd)
            arr_n = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
            arr_p = (ctypes.c_void_p * len(names))(*[sd[n].data_ptr() for n in names])
            arr_e = (ctypes.c_int64 * len(names))(*[sd[n].numel() for n in names])
            try:
                _lib.check(lib.spm_text_load_weights(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                                     len(names), arr_n, arr_p, arr_e))
            except Exception:
                lib.spm_text_destroy(h)
                raise
        self._h = h

    def _tokens(self, tokens):
        t = torch.as_tensor(tokens)
        if t.shape[-1] != CONTEXT_LENGTH:
            raise RuntimeError("tokens must end in a dimension of %d ids" % CONTEXT_LENGTH)
        return t.to(self._dev, torch.int32).contiguous()

    def tokenizer(self):
        if self._tok is None:
            self._tok = ClipTokenizer(self._vocab_path)
        return self._tok

    def encode_text(self, tokens):
        """CLIP.encode_text: tokens [n, 77] -> [n, embed_dim] fp32 (not normalised, as in the reference)"""
        t = self._tokens(tokens).reshape(-1, CONTEXT_LENGTH)
        out = torch.empty(t.shape[0], self.embed_dim, device=self._dev)
        with torch.cuda.device(self._dev):
            _lib.check(_lib.load().spm_text_encode(self._h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                                   _p(t), t.shape[0], _p(out)))
        return out

    def class_features_from_tokens(self, tokens):
        """tokens [n_templates, n_classes, 77] -> [n_classes, embed_dim]: mean over the templates"""
        t = self._tokens(tokens)
        if t.dim() != 3:
            raise RuntimeError("tokens must be [n_templates, n_classes, 77]")
        out = torch.empty(t.shape[1], self.embed_dim, device=self._dev)
        with torch.cuda.device(self._dev):
            _lib.check(_lib.load().spm_text_class_features(
                self._h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(t), t.shape[0], t.shape[1],
                _p(out)))
        return out

    def class_features(self, class_names, templates=PROMPT_TEMPLATES):
        """model_clipspm.py:50-70: [template.format(name)] -> tokenize -> encode_text -> mean over templates"""
        tk = self.tokenizer()
        tokens = torch.stack([tk.tokenize([t.format(c) for c in class_names]) for t in templates])
        return self.class_features_from_tokens(tokens)

    def close(self):
        if self._h is not None:
            _lib.load().spm_text_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
