"""Sibling head: host-side mirror of models/model_clipfsar.py::CNN_OTAM_CLIPFSAR (evaluation path) on the same
library -- SURVEY.md 8(f) rank 4 ("sibling heads reusing the same kernels").

    reference                                              here
    models/model_clipfsar.py:105 class CNN_OTAM_CLIPFSAR    class CNN_OTAM_CLIPFSAR(cfg, text_features_test=, text_features_train=)
    :183 forward(inputs) -> {"logits", "class_logits"}      forward(inputs) -> same keys ([1,Q,W], [1,S+Q,n_train])
    run/main_run.py:355-359 loss / accuracy                  evaluate(inputs) -> loss, accuracy

The branch implemented is the one every shipped config takes (configs/clipfsar/*.yaml), model_clipfsar.py:325-383, with
its two working options: MODEL.TRANSFORMER_DEPTH (:143-144, layers of context2; the count is TRAIN.TRANSFORMER_DEPTH) and
MODEL.MERGE_BEFORE (:341-346, class means before context2).  MODEL.EVAL_TEXT and MODEL.COMBINE (:235-322) end in
`None.unsqueeze(0)` (:384) in the reference -- executed and recorded by oracle/pin_against_reference.py
fsar_dead_branches -- so this class raises for them as well.  `state_dict()` has the reference's keys (`scale`,
`context2.layers.0.*` with inner width D, `backbone.*`).  cfg additionally reads MODEL.USE_CLASSIFICATION_VALUE and
TRAIN.TASKS_PER_BATCH for the loss.  No CPU / eager fallback (see model.py)."""
import ctypes

import torch

from . import _lib
from .model import CNN, _cfg_get, _p


class CNN_OTAM_CLIPFSAR(CNN):
    HEAD = "clipfsar"

    def __init__(self, cfg, text_features_test=None, text_features_train=None, max_episodes=1, device="cuda",
                 precision="bf16"):
        super().__init__(cfg, text_features_test=text_features_test, text_features_train=text_features_train,
                         max_episodes=max_episodes, device=device, precision=precision)
        self._packed_train = None
        for flag in ("EVAL_TEXT", "COMBINE"):
            if _cfg_get(cfg, "MODEL." + flag, False):
                raise AttributeError("MODEL.%s: the reference's branch ends in `class_text_logits = None; "
                                     "class_text_logits.unsqueeze(0)` (models/model_clipfsar.py:322,384) and never "
                                     "returns; not part of the path" % flag)

    def _extra_config(self):
        return dict(fsar_depth=self.transformer_depth, fsar_merge_before=int(self.merge_before))

    def _text(self):
        tf = super()._text()   # evaluation: prompts come from text_features_test (model_clipfsar.py:338)
        tr = self.text_features_train
        if tr is not None and self._packed_train is not tr:
            t = tr.detach().to(self._dev, torch.float32).contiguous()
            _lib.check(_lib.load().spm_set_text_features_train(
                self._handle(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(t), t.shape[0], t.shape[1]))
            torch.cuda.current_stream().synchronize()  # `t` may be a temporary
            self._packed_train = tr
        return tf

    def _class_logits(self, E, n_videos):
        """class_text_logits of the call just made (model_clipfsar.py:329-331): [E, S+Q, n_train]"""
        n_train = int(self.text_features_train.shape[0])
        out = torch.empty(E, n_videos, n_train, device=self._dev)
        _lib.check(_lib.load().spm_class_logits(self._handle(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                                E * n_videos, n_train, _p(out)))
        return out

    def forward_episodes(self, context_images, context_labels, target_images, real_support_labels,
                         real_target_labels, n_episodes=1, target_labels=None):
        out = super().forward_episodes(context_images, context_labels, target_images, real_support_labels,
                                       real_target_labels, n_episodes, target_labels)
        if self.training:     # the differentiable branch (clip_spm_b200.train.fsar_head_forward) carries its own class_logits
            return out
        if self.text_features_train is not None:
            E = int(n_episodes)
            out["class_logits"] = self._class_logits(E, (context_labels.numel() + real_target_labels.numel()) // E)
        return out

    def forward(self, inputs):
        """models/model_clipfsar.py:183-385 (eval): {"logits": [1,Q,W], "class_logits": [1,S+Q,n_train]}"""
        out = self.forward_episodes(inputs["context_images"], inputs["context_labels"], inputs["target_images"],
                                    inputs["real_support_labels"], inputs["real_target_labels"], n_episodes=1)
        if self.training:     # models/model_clipfsar.py:183-262: logits [1,Q,W] (+ class_logits [1,S+Q,n_train]) on the tape
            return {k: out[k] for k in ("logits", "class_logits") if k in out}
        res = {"logits": out["logits"][0].unsqueeze(0)}
        if "class_logits" in out:
            res["class_logits"] = out["class_logits"][0].unsqueeze(0)
        return res

    def head(self, su, qu, context_labels, real_support_labels, real_target_labels, n_episodes=1):
        """models/model_clipfsar.py:325-383 on precomputed features su [E,S,T,D], qu [E,Q,T,D]."""
        out = super().head(su, qu, context_labels, real_support_labels, real_target_labels, n_episodes)
        if self.training:
            return out
        if self.text_features_train is not None:
            E = int(n_episodes)
            out["class_logits"] = self._class_logits(E, (context_labels.numel() + real_target_labels.numel()) // E)
        return out


class CNN_STEN(CNN):
    """models/model_sten.py:11 (class CNN_OTAM_CLIPFSAR of that file, cfg.MODEL.NAME == 'sten', run/main_run.py:127-128) as
    shipped: frame features averaged over the 8 frames, class-mean support features and prompts, logits =
    softmax(cos_sim(query, prompts)) * softmax(cos_sim(query, support prototypes)) (:97-108).  No parameters besides
    `backbone.*`; forward returns {"logits": [1,Q,W]}; loss / accuracy are the runner's generic branch
    (run/main_run.py:394-396)."""
    HEAD = "sten"

    def forward(self, inputs):
        out = self.forward_episodes(inputs["context_images"], inputs["context_labels"], inputs["target_images"],
                                    inputs["real_support_labels"], inputs["real_target_labels"], n_episodes=1)
        return {"logits": out["logits"][0].unsqueeze(0)}
