"""Minimal stand-in for the reference's config object (utils/config.py -> attribute tree): only the fields the
episode-evaluation path reads: MODEL.BACKBONE, DATA.SEQ_LEN, params{...}, optional
MODEL.SINGLE_DIRECT, TRAIN.WAY, TRAIN.TASKS_PER_BATCH.  The real reference config object works as well."""
import types

# run/run.py:10-17
DEFAULT_PARAMS = dict(mid_dim_vision=0.5, mid_dim_text=1.5, negative_slope=0.0025, alpha=0.2, motion_alpha=1)


def make_cfg(backbone, seq_len, single_direct=False, way=None, params=None, tasks_per_batch=16, cls_value=None):
    cfg = types.SimpleNamespace(
        MODEL=types.SimpleNamespace(BACKBONE=backbone), DATA=types.SimpleNamespace(SEQ_LEN=seq_len),
        TRAIN=types.SimpleNamespace(TASKS_PER_BATCH=tasks_per_batch), params=dict(params or DEFAULT_PARAMS))
    if single_direct:
        cfg.MODEL.SINGLE_DIRECT = True
    if way is not None:
        cfg.TRAIN.WAY = way
    if cls_value is not None:   # cfg.MODEL.USE_CLASSIFICATION_VALUE (CLIP-FSAR loss, run/main_run.py:356)
        cfg.MODEL.USE_CLASSIFICATION_VALUE = cls_value
    return cfg
