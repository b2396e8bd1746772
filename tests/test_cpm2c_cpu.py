"""CPU suite of the sibling head CPM2C (models/model_cpm2c.py): the oracle restatement against the
golden tensors written from the executed reference class (oracle/pin_against_reference.py cpm2c_*), and the host-side
mirror's state_dict contract.  No GPU needed."""
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H

STAGES = ("su_motion", "qu_motion", "su_real_motion", "qu_fake_motion", "su_real", "qu_fake", "su_pro", "class_logits",
          "logits_local", "logits_global", "target_consist_distance")


@pytest.mark.parametrize("name", list(H.CPM2C_CASES))
def test_oracle_cpm2c_head_matches_reference_golden(name):
    ci, g = H.cpm2c_case_inputs(name), H.golden(name)
    ep = ci["episode"]
    su, qu = ci["feats"] if ci["head_only"] else (g["su"], g["qu"])
    with torch.no_grad():
        st = H.cpm2c_oracle(ci, su, qu)
    for k in STAGES:
        assert H.rel_err(st[k].reshape(g[k].shape), g[k]) < 1e-4, k
    loss, acc, pred, total = O.cpm2c_loss_and_acc(st, ep["target_labels"], ep["real_support_labels"],
                                                  ep["real_target_labels"], tasks_per_batch=H.CPM2C_TASKS_PER_BATCH)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    assert float(acc) == float(g["acc"])
    assert torch.equal(pred, g["pred"].long())
    assert H.rel_err(total, g["logits_total"]) < 1e-4


def test_oracle_cpm2c_tower_features_match_reference_golden():
    name = "cpm2c_vit_2w1s_t4_p1"
    ci, g = H.cpm2c_case_inputs(name), H.golden(name)
    with torch.no_grad():
        su = O.vit_forward(ci["weights"], ci["episode"]["context_images"])
    assert H.rel_err(su.reshape(g["su"].shape), g["su"]) < 1e-4


def test_cpm2c_state_dict_keys_and_reference_extras():
    from clip_spm_b200 import CLIP_CPMMC_FSAR
    for backbone, D in (("ViT-B/16", 512), ("RN50", 1024)):
        ci = dict(backbone=backbone, T=8, single=False, way=5)
        net = CLIP_CPMMC_FSAR(H.cpm2c_cfg(ci))
        ref = O.make_cpm2c_weights(D, seed=0)
        ref.update({k: v for k, v in O.make_weights(backbone, seed=0).items() if k.startswith("backbone.")})
        sd = net.state_dict()
        assert set(sd.keys()) == set(ref.keys()), set(sd.keys()) ^ set(ref.keys())
        for k in ref:
            assert tuple(sd[k].shape) == tuple(ref[k].shape), k
        # a full reference state_dict also carries modules its forward never calls (model_cpm2c.py:98-99, :123-133)
        full = dict(ref)
        full.update({"transformer.resblocks.0.ln_1.weight": torch.ones(D), "frame_position_embeddings.weight": torch.zeros(77, D),
                     "meta_net.0.weight": torch.zeros(8, 3), "meta_net_2.0.weight": torch.zeros(4, 3)})
        net.load_state_dict(full, strict=True)
        with pytest.raises(RuntimeError):
            net.load_state_dict(dict(full, bogus=torch.zeros(1)), strict=True)


def test_cpm2c_requires_its_config_fields():
    from clip_spm_b200 import CLIP_CPMMC_FSAR
    ci = dict(backbone="ViT-B/16", T=8, single=False, way=5)
    cfg = H.cpm2c_cfg(ci)
    del cfg.params["motion_residual_ratio"]
    with pytest.raises(RuntimeError):
        CLIP_CPMMC_FSAR(cfg)


def test_cpm2c_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ci = H.cpm2c_case_inputs("cpm2c_head_5w3s_t8")
    net = H.build_cuda_cpm2c_model(ci)
    su, qu = ci["feats"]
    ep = ci["episode"]
    with pytest.raises(RuntimeError):
        net.head(su, qu, ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])
