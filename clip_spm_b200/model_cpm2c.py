"""Sibling head: host-side mirror of models/model_cpm2c.py::CLIP_CPMMC_FSAR (evaluation path) on the same library --
SURVEY.md 8(f) rank 4 ("sibling heads reusing the same kernels").

    reference                                               here
    models/model_cpm2c.py:14 class CLIP_CPMMC_FSAR           class CLIP_CPMMC_FSAR(cfg, text_features_test=, ...)
    :207 forward(inputs) -> {"class_logits", "logits_local",  forward(inputs) -> same keys ([1,S+Q,n_cls], [1,Q,W],
          "logits_global", "target_consist_distance"}                          [1,Q,W], 0-d)
    run/main_run.py:370-380 loss / accuracy                   evaluate(inputs) -> loss, accuracy

cfg reads, besides what CNN reads: MODEL.MOTION_COFF, MODEL.NORMAL_COFF, MODEL.USE_CLASSIFICATION,
params{motion_residual_ratio, lambdas0..3}.  The forward implemented is the shipped one (no TRANSFORMER_DEPTH override
needed: context2 has one layer either way in configs/cpm2c/*.yaml; USE_CLASSIFICATION False returns zeros for
class_logits like :222-224).  `state_dict()` holds the parameters the forward reads; `load_state_dict` also accepts a
full reference state_dict and drops the modules the reference constructs but never calls (`transformer.*`,
`frame_position_embeddings.*`, `meta_net*`: model_cpm2c.py:98-99, :123-133).  No CPU / eager fallback (see model.py)."""
import ctypes

import torch

from . import _lib
from .model import CNN, _cfg_get, _p

_UNUSED_PREFIXES = ("transformer.", "frame_position_embeddings.", "meta_net.", "meta_net_2.")


class CLIP_CPMMC_FSAR(CNN):
    HEAD = "cpm2c"

    def __init__(self, cfg, text_features_test=None, text_features_train=None, max_episodes=1, device="cuda",
                 precision="bf16"):
        super().__init__(cfg, text_features_test=text_features_test, text_features_train=text_features_train,
                         max_episodes=max_episodes, device=device, precision=precision)
        for k in ("motion_residual_ratio", "lambdas0", "lambdas1", "lambdas2"):
            if k not in self.params:
                raise RuntimeError("cfg.params[%r] is required by the CPM2C head" % k)
        self.motion_coeff = float(_cfg_get(cfg, "MODEL.MOTION_COFF"))
        self.normal_coeff = float(_cfg_get(cfg, "MODEL.NORMAL_COFF"))
        self.use_classification = bool(_cfg_get(cfg, "MODEL.USE_CLASSIFICATION", False))

    def _extra_config(self):
        lam = (ctypes.c_float * 4)(*[float(self.params.get("lambdas%d" % i, 0.0)) for i in range(4)])
        return dict(motion_residual_ratio=float(self.params["motion_residual_ratio"]), lambdas=lam,
                    motion_coeff=self.motion_coeff, normal_coeff=self.normal_coeff,
                    use_classification=int(self.use_classification))

    def load_state_dict(self, state_dict, strict=True):
        kept = {k: v for k, v in state_dict.items() if not k.startswith(_UNUSED_PREFIXES)}
        return super().load_state_dict(kept, strict=strict)

    def _branches(self, out, E, Q, W, n_videos):
        lib, h = _lib.load(), self._handle()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        loc = torch.empty(E, Q, W, device=self._dev)
        glob = torch.empty(E, Q, W, device=self._dev)
        _lib.check(lib.spm_cpm2c_outputs(h, st, E, Q, W, _p(loc), _p(glob)))
        n_cls = int(self.text_features_test.shape[0])
        if self.use_classification:
            cls = torch.empty(E, n_videos, n_cls, device=self._dev)
            _lib.check(lib.spm_class_logits(h, st, E * n_videos, n_cls, _p(cls)))
        else:
            cls = torch.zeros(E, n_videos, n_cls, device=self._dev)   # model_cpm2c.py:423-424
        out.update(logits_local=loc, logits_global=glob, class_logits=cls, target_consist_distance=out["dists"])
        return out

    def forward_episodes(self, context_images, context_labels, target_images, real_support_labels,
                         real_target_labels, n_episodes=1, target_labels=None):
        """E episodes in one call.  `logits` is the runner's lambdas1 * logits_local + lambdas2 * logits_global
        (run/main_run.py:373), `dists` the consistency distance."""
        out = super().forward_episodes(context_images, context_labels, target_images, real_support_labels,
                                       real_target_labels, n_episodes, target_labels)
        E = int(n_episodes)
        Q = real_target_labels.numel() // E
        return self._branches(out, E, Q, out["logits"].shape[-1], (context_labels.numel() + real_target_labels.numel()) // E)

    def forward(self, inputs):
        """models/model_cpm2c.py:207-242 (eval)."""
        out = self.forward_episodes(inputs["context_images"], inputs["context_labels"], inputs["target_images"],
                                    inputs["real_support_labels"], inputs["real_target_labels"], n_episodes=1)
        return {"class_logits": out["class_logits"][0].unsqueeze(0), "logits_local": out["logits_local"][0].unsqueeze(0),
                "logits_global": out["logits_global"][0].unsqueeze(0),
                "target_consist_distance": out["target_consist_distance"][0]}

    def head(self, su, qu, context_labels, real_support_labels, real_target_labels, n_episodes=1):
        """models/model_cpm2c.py:219-242 on precomputed frame features su [E,S,T,D], qu [E,Q,T,D]."""
        out = super().head(su, qu, context_labels, real_support_labels, real_target_labels, n_episodes)
        E = int(n_episodes)
        Q = real_target_labels.numel() // E
        return self._branches(out, E, Q, out["logits"].shape[-1], (context_labels.numel() + real_target_labels.numel()) // E)
