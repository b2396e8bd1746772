"""Shared helpers of the parity tests: build the CUDA model and the oracle inputs from the same seeds."""
import os
import types

import numpy as np
import torch

from oracle import clipspm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name: (backbone, way, shot, qpc, T, n_text_cls, protocol, head_only, single_direct, seed) -- must match
# oracle/pin_against_reference.py::CASES (the golden files were written by that script from the real reference)
CASES = {
    "vit_5w1s_t8_p1": ("ViT-B/16", 5, 1, 1, 8, 24, "P1", False, False, 1000),
    "vit_2w1s_t2_p0": ("ViT-B/16", 2, 1, 1, 2, 24, "P0", False, False, 1001),
    "head_5w5s_t8": ("ViT-B/16", 5, 5, 1, 8, 24, "P1", True, False, 1002),
    "head_5w1s_t16": ("ViT-B/16", 5, 1, 1, 16, 24, "P1", True, False, 1003),
    "head_5w3s_t8_d1024": ("RN50", 5, 3, 1, 8, 10, "P1", True, False, 1004),
    "head_5w2s_t8_q3_single": ("ViT-B/16", 5, 2, 3, 8, 24, "P1", True, True, 1005),
    "rn50_2w1s_t2_p1": ("RN50", 2, 1, 1, 2, 10, "P1", False, False, 1006),
    "vit_5w5s_t8_p1": ("ViT-B/16", 5, 5, 1, 8, 24, "P1", False, False, 1007),      # BASELINE config 2 / 5
    "vit_5w1s_t16_p1": ("ViT-B/16", 5, 1, 1, 16, 24, "P1", False, False, 1008),    # BASELINE config 3
    "rn50_5w3s_t8_p1": ("RN50", 5, 3, 1, 8, 10, "P1", False, False, 1009),         # BASELINE config 4
}


def golden(name):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}


def make_cfg(backbone, T, single_direct=False, way=None):
    from clip_spm_b200.config import make_cfg as _mk
    return _mk(backbone, T, single_direct, way, params=O.DEFAULT_PARAMS)


def case_inputs(name):
    backbone, way, shot, qpc, T, ncls, proto, head_only, single, seed = CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    w = O.make_weights(backbone, seed=0, protocol=proto, head_only=head_only)
    text = O.make_text_features(ncls, D, seed=0)
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, proto, images=not head_only)
    feats = None
    if head_only:
        feats = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    return dict(backbone=backbone, way=way, shot=shot, qpc=qpc, T=T, D=D, weights=w, text=text, episode=ep,
                feats=feats, single=single, head_only=head_only)


def build_cuda_model(ci, max_episodes=1, precision="bf16"):
    from clip_spm_b200 import CNN
    net = CNN(make_cfg(ci["backbone"], ci["T"], ci["single"], ci["way"]), text_features_test=ci["text"],
              max_episodes=max_episodes, precision=precision)
    missing, unexpected = net.load_state_dict(ci["weights"], strict=False)
    assert not unexpected, unexpected
    if not ci["head_only"]:
        assert not missing, missing
    return net


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---- sibling head CLIP-FSAR: must match oracle/pin_against_reference.py::FSAR_CASES
# name: (backbone, way, shot, qpc, T, n_test_cls, n_train_cls, head_only, single_direct, seed)
FSAR_CASES = {
    "fsar_head_5w5s_t8": ("ViT-B/16", 5, 5, 1, 8, 24, 30, True, False, 2002),
    "fsar_head_5w3s_t8_d1024_q2": ("RN50", 5, 3, 2, 8, 10, 12, True, False, 2004),
    "fsar_head_5w1s_t16_single": ("ViT-B/16", 5, 1, 1, 16, 24, 30, True, True, 2005),
    "fsar_vit_2w1s_t2_p1": ("ViT-B/16", 2, 1, 1, 2, 24, 30, False, False, 2001),
    "fsar_head_5w3s_t8_merge": ("ViT-B/16", 5, 3, 1, 8, 24, 30, True, False, 2006),      # MODEL.MERGE_BEFORE
    "fsar_head_5w2s_t8_depth2": ("ViT-B/16", 5, 2, 2, 8, 24, 30, True, False, 2007),     # TRANSFORMER_DEPTH = 2
}
FSAR_OPTIONS = {"fsar_head_5w3s_t8_merge": dict(merge_before=True), "fsar_head_5w2s_t8_depth2": dict(depth=2)}
FSAR_TASKS_PER_BATCH, FSAR_CLS_VALUE = 4, 3.0


def fsar_case_inputs(name):
    backbone, way, shot, qpc, T, ncls, ntrain, head_only, single, seed = FSAR_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    opt = FSAR_OPTIONS.get(name, {})
    w = O.make_fsar_weights(D, seed=0, depth=opt.get("depth", 1))
    if not head_only:
        w.update({k: v for k, v in O.make_weights(backbone, seed=0, protocol="P1").items() if k.startswith("backbone.")})
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=not head_only)
    feats = None
    if head_only:
        feats = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    return dict(backbone=backbone, way=way, shot=shot, qpc=qpc, T=T, D=D, weights=w, episode=ep, feats=feats,
                text=O.make_text_features(ncls, D, seed=0), text_train=O.make_text_features(ntrain, D, seed=1),
                single=single, head_only=head_only, options=opt)


def build_cuda_fsar_model(ci, max_episodes=1, precision="bf16"):
    from clip_spm_b200 import CNN_OTAM_CLIPFSAR
    from clip_spm_b200.config import make_cfg as _mk
    cfg = _mk(ci["backbone"], ci["T"], ci["single"], ci["way"], params={}, tasks_per_batch=FSAR_TASKS_PER_BATCH,
              cls_value=FSAR_CLS_VALUE)
    opt = ci.get("options", {})
    if opt.get("merge_before"):
        cfg.MODEL.MERGE_BEFORE = True
    if opt.get("depth", 1) > 1:   # the switch and the value the reference constructor reads (model_clipfsar.py:143-144)
        cfg.MODEL.TRANSFORMER_DEPTH = opt["depth"]
        cfg.TRAIN.TRANSFORMER_DEPTH = opt["depth"]
    net = CNN_OTAM_CLIPFSAR(cfg, text_features_test=ci["text"], text_features_train=ci["text_train"],
                            max_episodes=max_episodes, precision=precision)
    missing, unexpected = net.load_state_dict(ci["weights"], strict=False)
    assert not unexpected, unexpected
    if not ci["head_only"]:
        assert not missing, missing
    return net


# ---- sibling head CPM2C: must match oracle/pin_against_reference.py::CPM2C_CASES
# name: (backbone, way, shot, qpc, T, n_test_cls, head_only, single_direct, seed)
CPM2C_CASES = {
    "cpm2c_head_5w3s_t8": ("ViT-B/16", 5, 3, 1, 8, 24, True, False, 2202),
    "cpm2c_head_5w1s_t8_d1024_q2": ("RN50", 5, 1, 2, 8, 10, True, False, 2204),
    "cpm2c_head_5w2s_t6_single": ("ViT-B/16", 5, 2, 1, 6, 24, True, True, 2205),
    "cpm2c_vit_2w1s_t4_p1": ("ViT-B/16", 2, 1, 1, 4, 24, False, False, 2201),
}
CPM2C_TASKS_PER_BATCH, CPM2C_MOTION_COEFF, CPM2C_NORMAL_COEFF = 16, 1.0, 0.7


def cpm2c_case_inputs(name):
    backbone, way, shot, qpc, T, ncls, head_only, single, seed = CPM2C_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    w = O.make_cpm2c_weights(D, seed=0)
    if not head_only:
        w.update({k: v for k, v in O.make_weights(backbone, seed=0, protocol="P1").items() if k.startswith("backbone.")})
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=not head_only)
    feats = None
    if head_only:
        feats = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    return dict(backbone=backbone, way=way, shot=shot, qpc=qpc, T=T, D=D, weights=w, episode=ep, feats=feats,
                text=O.make_text_features(ncls, D, seed=0), single=single, head_only=head_only)


def cpm2c_cfg(ci, use_classification=True):
    from clip_spm_b200.config import make_cfg as _mk
    cfg = _mk(ci["backbone"], ci["T"], ci["single"], ci["way"], params=dict(O.CPM2C_PARAMS),
              tasks_per_batch=CPM2C_TASKS_PER_BATCH)
    cfg.MODEL.MOTION_COFF, cfg.MODEL.NORMAL_COFF = CPM2C_MOTION_COEFF, CPM2C_NORMAL_COEFF
    cfg.MODEL.USE_CLASSIFICATION = use_classification
    return cfg


def build_cuda_cpm2c_model(ci, max_episodes=1, precision="bf16", use_classification=True):
    from clip_spm_b200 import CLIP_CPMMC_FSAR
    net = CLIP_CPMMC_FSAR(cpm2c_cfg(ci, use_classification), text_features_test=ci["text"], max_episodes=max_episodes,
                          precision=precision)
    missing, unexpected = net.load_state_dict(ci["weights"], strict=False)
    assert not unexpected, unexpected
    if not ci["head_only"]:
        assert not missing, missing
    return net


def cpm2c_oracle(ci, su, qu):
    ep = ci["episode"]
    return O.cpm2c_head_forward(ci["weights"], ci["text"], su, qu, ep["context_labels"], ep["real_support_labels"],
                                ep["real_target_labels"], O.CPM2C_PARAMS, motion_coeff=CPM2C_MOTION_COEFF,
                                normal_coeff=CPM2C_NORMAL_COEFF, single_direct=ci["single"])


# ---- sibling head STEN: must match oracle/pin_against_reference.py::STEN_CASES (T = 8)
# name: (backbone, way, shot, qpc, n_test_cls, head_only, seed)
STEN_CASES = {
    "sten_head_5w5s": ("ViT-B/16", 5, 5, 1, 24, True, 2102),
    "sten_head_5w3s_d1024_q2": ("RN50", 5, 3, 2, 10, True, 2104),
}


def sten_case_inputs(name):
    backbone, way, shot, qpc, ncls, head_only, seed = STEN_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    ep = O.make_episode(seed, way, shot, qpc, 8, ncls, "P1", images=False)
    feats = O.make_features(seed, way * shot, way * qpc, 8, D, ep["context_labels"], ep["target_labels"].float())
    return dict(backbone=backbone, way=way, T=8, D=D, episode=ep, feats=feats, text=O.make_text_features(ncls, D, seed=0))


# ---- TA2N soft-DTW (models/OTAM.py): must match oracle/pin_against_reference.py::SOFTDTW_CASES
# name: (B, N, M, d, gamma, bandwidth, seed)
SOFTDTW_CASES = {
    "softdtw_8x8_g01": (6, 8, 8, 64, 0.1, 0.0, 3101),
    "softdtw_17x15_g1": (4, 17, 15, 8, 1.0, 0.0, 3102),
    "softdtw_40x38_bw5": (3, 40, 38, 8, 0.5, 5.0, 3103),
}
