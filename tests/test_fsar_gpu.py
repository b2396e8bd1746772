"""GPU parity of the sibling head CLIP-FSAR (models/model_clipfsar.py:325-383) through the C ABI, against the golden
tensors of the executed reference class and against the oracle.  Tolerances as in test_stages_gpu.py: the head runs
fp32 data with tf32 tensor-core products -> 5e-3; bf16 tower + head end to end -> 2e-2; fp32 mode -> 1e-4."""
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_HEAD = 5e-3
TOL_BF16 = 2e-2


def _check_pred(out_logits, g):
    err = float((out_logits.cpu() - g["logits"][0]).abs().max())
    pred = out_logits.argmax(-1).cpu()
    safe = g["margin"] > 4 * err
    assert torch.equal(pred[safe], g["pred"].long()[safe])
    return int(safe.sum())


@pytest.mark.parametrize("name", ["fsar_head_5w5s_t8", "fsar_head_5w3s_t8_d1024_q2", "fsar_head_5w1s_t16_single",
                                  "fsar_head_5w3s_t8_merge", "fsar_head_5w2s_t8_depth2"])
def test_fsar_head_matches_reference_golden(name):
    ci, g = H.fsar_case_inputs(name), H.golden(name)
    net = H.build_cuda_fsar_model(ci)
    ep = ci["episode"]
    su, qu = ci["feats"]
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"].cuda(), ep["real_support_labels"].cuda(),
                   ep["real_target_labels"].cuda())
    assert H.rel_err(out["logits"], g["logits"]) < TOL_HEAD
    assert H.rel_err(out["class_logits"], g["class_logits"]) < 1e-4    # pure fp32 kernel
    assert float(out["dists"].abs().max()) == 0.0
    assert _check_pred(out["logits"][0], g) >= 1


@pytest.mark.parametrize("name", ["fsar_head_5w5s_t8", "fsar_head_5w3s_t8_merge", "fsar_head_5w2s_t8_depth2"])
def test_fsar_head_batched_episodes_equal_single(name):
    """E episodes in one call == the same episodes one at a time (bit for bit: same kernels, same order per row)."""
    ci = H.fsar_case_inputs(name)
    net = H.build_cuda_fsar_model(ci, max_episodes=3)
    eps, sus, qus = [], [], []
    for k in range(3):
        ep = O.make_episode(3000 + k, ci["way"], ci["shot"], ci["qpc"], ci["T"], 24, "P1", images=False)
        su, qu = O.make_features(3000 + k, ci["way"] * ci["shot"], ci["way"] * ci["qpc"], ci["T"], ci["D"],
                                 ep["context_labels"], ep["target_labels"].float())
        eps.append(ep); sus.append(su); qus.append(qu)
    cat = lambda key: torch.cat([e[key] for e in eps]).cuda()
    both = net.head(torch.stack(sus).cuda(), torch.stack(qus).cuda(), cat("context_labels"),
                    cat("real_support_labels"), cat("real_target_labels"), n_episodes=3)
    for k in range(3):
        one = net.head(sus[k].cuda(), qus[k].cuda(), eps[k]["context_labels"].cuda(),
                       eps[k]["real_support_labels"].cuda(), eps[k]["real_target_labels"].cuda())
        assert H.rel_err(both["logits"][k], one["logits"][0]) < 1e-6
        assert H.rel_err(both["class_logits"][k], one["class_logits"][0]) < 1e-6
        ref = O.fsar_head_forward(ci["weights"], ci["text"], ci["text_train"], sus[k], qus[k],
                                  eps[k]["context_labels"], eps[k]["real_support_labels"],
                                  eps[k]["real_target_labels"], ci["single"], **ci["options"])
        assert H.rel_err(both["logits"][k], ref["logits"][0]) < TOL_HEAD


@pytest.mark.parametrize("precision,tol", [("bf16", TOL_BF16), ("fp32", 1e-4)])
def test_fsar_forward_and_loss_match_reference_golden(precision, tol):
    """whole operator (ViT-B/16 tower + CLIP-FSAR head + run/main_run.py:355-359 loss) against the reference's output"""
    name = "fsar_vit_2w1s_t2_p1"
    ci, g = H.fsar_case_inputs(name), H.golden(name)
    net = H.build_cuda_fsar_model(ci, precision=precision)
    ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()}
    out = net(ep)
    assert set(out.keys()) == {"logits", "class_logits"}
    assert tuple(out["logits"].shape) == tuple(g["logits"].shape)
    assert H.rel_err(out["logits"], g["logits"]) < tol
    assert H.rel_err(out["class_logits"], g["class_logits"]) < tol
    loss, acc = net.evaluate(ep)
    assert abs(float(loss) - float(g["loss"])) < tol * max(1.0, abs(float(g["loss"])))
    if precision == "fp32":
        assert float(acc) == float(g["acc"])


def test_fsar_wrong_way_fails_loudly():
    ci = H.fsar_case_inputs("fsar_head_5w5s_t8")
    net = H.build_cuda_fsar_model(ci)
    net.way = 4     # the episode has 5 distinct support labels
    ep = ci["episode"]
    su, qu = ci["feats"]
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"].cuda(), ep["real_support_labels"].cuda(),
                   ep["real_target_labels"].cuda())
    assert bool(torch.isnan(out["logits"]).all())


@pytest.mark.parametrize("name", list(H.STEN_CASES))
def test_sten_head_matches_reference_golden(name):
    """models/model_sten.py as shipped: pure fp32 kernels -> 1e-5"""
    from clip_spm_b200 import CNN_STEN
    from clip_spm_b200.config import make_cfg
    ci, g = H.sten_case_inputs(name), H.golden(name)
    net = CNN_STEN(make_cfg(ci["backbone"], 8, False, ci["way"], params={}), text_features_test=ci["text"])
    ep = ci["episode"]
    su, qu = ci["feats"]
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"].cuda(), ep["real_support_labels"].cuda(),
                   ep["real_target_labels"].cuda())
    assert H.rel_err(out["logits"], g["logits"]) < 1e-5
    assert _check_pred(out["logits"][0], g) >= 1
    # two episodes in one call == one at a time
    two = net.head(torch.stack([su, su.flip(0)]).cuda(), torch.stack([qu, qu]).cuda(),
                   torch.cat([ep["context_labels"], ep["context_labels"].flip(0)]).cuda(),
                   torch.cat([ep["real_support_labels"], ep["real_support_labels"].flip(0)]).cuda(),
                   torch.cat([ep["real_target_labels"]] * 2).cuda(), n_episodes=2)
    assert H.rel_err(two["logits"][0], g["logits"][0]) < 1e-5
    assert H.rel_err(two["logits"][1], g["logits"][0]) < 1e-5     # support order does not matter
