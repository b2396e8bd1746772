"""CPU suite of the sibling head CLIP-FSAR (models/model_clipfsar.py): the oracle restatement
against the golden tensors written from the executed reference class (oracle/pin_against_reference.py fsar_*), and
the host-side mirror's state_dict contract.  No GPU needed."""
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", list(H.FSAR_CASES))
def test_oracle_fsar_head_matches_reference_golden(name):
    ci, g = H.fsar_case_inputs(name), H.golden(name)
    ep = ci["episode"]
    su, qu = ci["feats"] if ci["head_only"] else (g["su"], g["qu"])
    with torch.no_grad():
        st = O.fsar_head_forward(ci["weights"], ci["text"], ci["text_train"], su, qu, ep["context_labels"],
                                 ep["real_support_labels"], ep["real_target_labels"], ci["single"], **ci["options"])
    for k in ("qu_ctx", "su_ctx", "logits", "class_logits"):
        assert H.rel_err(st[k].reshape(g[k].shape), g[k]) < 1e-4, k
    loss, acc, pred = O.fsar_loss_and_acc(st["logits"], st["class_logits"], ep["target_labels"],
                                          ep["real_support_labels"], ep["real_target_labels"], H.FSAR_TASKS_PER_BATCH,
                                          H.FSAR_CLS_VALUE)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    assert float(acc) == float(g["acc"])
    assert torch.equal(pred, g["pred"].long())


def test_oracle_fsar_tower_features_match_reference_golden():
    name = "fsar_vit_2w1s_t2_p1"
    ci, g = H.fsar_case_inputs(name), H.golden(name)
    with torch.no_grad():
        su = O.vit_forward(ci["weights"], ci["episode"]["context_images"])
    assert H.rel_err(su.reshape(g["su"].shape), g["su"]) < 1e-4


def test_fsar_prompt_position_is_free():
    """Transformer_v1 has no positional term, so cat([frames, prompt])[:, :T] (model_clipfsar.py:347-348) equals
    cat([prompt, frames])[:, 1:] -- the layout the CUDA path uses."""
    D, T = 512, 8
    w = O.make_fsar_weights(D, seed=0)
    g = torch.Generator().manual_seed(5)
    x, tok = torch.randn(3, T, D, generator=g), torch.randn(3, 1, D, generator=g)
    with torch.no_grad():
        a = O.transformer_v1(torch.cat([x, tok], 1), w, "context2.", 8, D // 8)[:, :T]
        b = O.transformer_v1(torch.cat([tok, x], 1), w, "context2.", 8, D // 8)[:, 1:]
    assert H.rel_err(b, a) < 1e-5


def test_fsar_state_dict_keys_match_reference_names():
    from clip_spm_b200 import CNN_OTAM_CLIPFSAR
    from clip_spm_b200.config import make_cfg
    for backbone, D in (("ViT-B/16", 512), ("RN50", 1024)):
        net = CNN_OTAM_CLIPFSAR(make_cfg(backbone, 8, params={}))
        ref = O.make_fsar_weights(D, seed=0)
        ref.update({k: v for k, v in O.make_weights(backbone, seed=0).items() if k.startswith("backbone.")})
        sd = net.state_dict()
        assert set(sd.keys()) == set(ref.keys()), set(sd.keys()) ^ set(ref.keys())
        for k in ref:
            assert tuple(sd[k].shape) == tuple(ref[k].shape), k
        net.load_state_dict(ref, strict=True)


def test_fsar_optional_branches_in_state_dict_and_config():
    """MODEL.TRANSFORMER_DEPTH adds context2.layers.1.* (model_clipfsar.py:143-144); MODEL.EVAL_TEXT / MODEL.COMBINE raise
    like the reference does (:384, recorded by pin_against_reference.py fsar_dead_branches)."""
    ci = H.fsar_case_inputs("fsar_head_5w2s_t8_depth2")
    net = H.build_cuda_fsar_model(ci)
    head_keys = {k for k in net.state_dict() if not k.startswith("backbone.")}
    assert head_keys == set(O.make_fsar_weights(512, seed=0, depth=2).keys())
    assert net._extra_config() == dict(fsar_depth=2, fsar_merge_before=0)
    assert H.build_cuda_fsar_model(H.fsar_case_inputs("fsar_head_5w3s_t8_merge"))._extra_config() == \
        dict(fsar_depth=1, fsar_merge_before=1)
    from clip_spm_b200 import CNN_OTAM_CLIPFSAR
    from clip_spm_b200.config import make_cfg
    for flag in ("EVAL_TEXT", "COMBINE"):
        cfg = make_cfg("ViT-B/16", 8, params={})
        setattr(cfg.MODEL, flag, True)
        with pytest.raises(AttributeError):
            CNN_OTAM_CLIPFSAR(cfg)


def test_fsar_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ci = H.fsar_case_inputs("fsar_head_5w5s_t8")
    net = H.build_cuda_fsar_model(ci)
    su, qu = ci["feats"]
    ep = ci["episode"]
    with pytest.raises(RuntimeError):
        net.head(su, qu, ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])


@pytest.mark.parametrize("name", list(H.STEN_CASES))
def test_oracle_sten_head_matches_reference_golden(name):
    ci, g = H.sten_case_inputs(name), H.golden(name)
    ep = ci["episode"]
    su, qu = ci["feats"]
    st = O.sten_head_forward(ci["text"], su, qu, ep["context_labels"], ep["real_support_labels"])
    assert H.rel_err(st["logits"], g["logits"]) < 1e-5
    loss, acc, pred = O.loss_and_acc(st["logits"], torch.zeros(()), ep["target_labels"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5 and float(acc) == float(g["acc"])
    assert torch.equal(pred, g["pred"].long())


def test_sten_state_dict_is_backbone_only():
    from clip_spm_b200 import CNN_STEN
    from clip_spm_b200.config import make_cfg
    net = CNN_STEN(make_cfg("ViT-B/16", 8, params={}))
    assert all(k.startswith("backbone.") for k in net.state_dict())
