"""GPU parity of the sibling head CPM2C (models/model_cpm2c.py:207-312) through the C ABI, against the golden tensors of
the executed reference class and against the oracle.  Tolerances as in test_stages_gpu.py: the head runs fp32 data with
tf32 tensor-core products -> 5e-3; bf16 tower + head end to end -> 2e-2; fp32 mode -> 1e-4."""
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_HEAD = 5e-3
TOL_BF16 = 2e-2
HEAD_CASES = ["cpm2c_head_5w3s_t8", "cpm2c_head_5w1s_t8_d1024_q2", "cpm2c_head_5w2s_t6_single"]


def _run_head(net, ci, **kw):
    ep = ci["episode"]
    su, qu = ci["feats"]
    return net.head(su.cuda(), qu.cuda(), ep["context_labels"].cuda(), ep["real_support_labels"].cuda(),
                    ep["real_target_labels"].cuda(), **kw)


def _check_outputs(out, g, tol):
    for k in ("logits_local", "logits_global"):
        assert H.rel_err(out[k], g[k]) < tol, k
    assert H.rel_err(out["logits"], g["logits_total"]) < tol
    assert H.rel_err(out["target_consist_distance"].reshape(()), g["target_consist_distance"]) < tol
    err = float((out["logits"][0].cpu() - g["logits_total"][0]).abs().max())
    safe = g["margin"] > 4 * err
    assert torch.equal(out["logits"][0].argmax(-1).cpu()[safe], g["pred"].long()[safe])
    return int(safe.sum())


@pytest.mark.parametrize("name", HEAD_CASES)
@pytest.mark.parametrize("precision,tol", [("bf16", TOL_HEAD), ("fp32", 1e-4)])
def test_cpm2c_head_matches_reference_golden(name, precision, tol):
    ci, g = H.cpm2c_case_inputs(name), H.golden(name)
    if precision == "fp32" and ci["backbone"] != "ViT-B/16":
        pytest.skip("the fp32 parity mode exists for the ViT-B/16 backbone only (spm_create says so)")
    net = H.build_cuda_cpm2c_model(ci, precision=precision)
    out = _run_head(net, ci)
    assert _check_outputs(out, g, tol) >= 1
    assert H.rel_err(out["class_logits"], g["class_logits"]) < 1e-4    # pure fp32 kernel
    # per-stage taps (spm_head_stage): motion features and the normal branch's context2 outputs / prototypes
    S, Q, T, D = ci["way"] * ci["shot"], ci["way"] * ci["qpc"], ci["T"], ci["D"]
    for k, shape in (("su_motion", (S, T - 1, D)), ("qu_motion", (Q, T - 1, D)), ("su_real", (S, T + 1, D)),
                     ("qu_fake", (Q, T + 1, D)), ("su_pro", (ci["way"], T + 1, D))):
        assert H.rel_err(net.head_stage(k).view(shape), g[k]) < tol, k


def test_cpm2c_head_batched_episodes_equal_single():
    """E episodes in one call == the same episodes one at a time, and both equal the oracle."""
    name = "cpm2c_head_5w3s_t8"
    ci = H.cpm2c_case_inputs(name)
    net = H.build_cuda_cpm2c_model(ci, max_episodes=3)
    eps, sus, qus = [], [], []
    for k in range(3):
        ep = O.make_episode(3200 + k, ci["way"], ci["shot"], ci["qpc"], ci["T"], 24, "P1", images=False)
        su, qu = O.make_features(3200 + k, ci["way"] * ci["shot"], ci["way"] * ci["qpc"], ci["T"], ci["D"],
                                 ep["context_labels"], ep["target_labels"].float())
        eps.append(ep); sus.append(su); qus.append(qu)
    cat = lambda key: torch.cat([e[key] for e in eps]).cuda()
    both = net.head(torch.stack(sus).cuda(), torch.stack(qus).cuda(), cat("context_labels"),
                    cat("real_support_labels"), cat("real_target_labels"), n_episodes=3)
    for k in range(3):
        one = net.head(sus[k].cuda(), qus[k].cuda(), eps[k]["context_labels"].cuda(),
                       eps[k]["real_support_labels"].cuda(), eps[k]["real_target_labels"].cuda())
        for key in ("logits_local", "logits_global", "class_logits", "logits"):
            assert H.rel_err(both[key][k], one[key][0]) < 1e-6, key
        assert H.rel_err(both["dists"][k], one["dists"][0]) < 1e-6
        ref = H.cpm2c_oracle(dict(ci, episode=eps[k]), sus[k], qus[k])
        assert H.rel_err(both["logits_local"][k], ref["logits_local"][0]) < TOL_HEAD
        assert H.rel_err(both["logits_global"][k], ref["logits_global"][0]) < TOL_HEAD
        assert H.rel_err(both["dists"][k], ref["target_consist_distance"].reshape(())) < TOL_HEAD


@pytest.mark.parametrize("precision,tol", [("bf16", TOL_BF16), ("fp32", 1e-4)])
def test_cpm2c_forward_and_loss_match_reference_golden(precision, tol):
    """whole operator (ViT-B/16 tower + CPM2C head + run/main_run.py:370-380 loss) against the reference's output"""
    name = "cpm2c_vit_2w1s_t4_p1"
    ci, g = H.cpm2c_case_inputs(name), H.golden(name)
    net = H.build_cuda_cpm2c_model(ci, precision=precision)
    ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()}
    out = net(ep)
    assert set(out.keys()) == {"class_logits", "logits_local", "logits_global", "target_consist_distance"}
    for k in out:
        assert tuple(out[k].shape) == tuple(g[k].shape), k
        assert H.rel_err(out[k], g[k]) < tol, k
    loss, acc = net.evaluate(ep)
    assert abs(float(loss) - float(g["loss"])) < tol * max(1.0, abs(float(g["loss"])))
    if precision == "fp32":
        assert float(acc) == float(g["acc"])


def test_cpm2c_loss_matches_golden_on_head_case():
    name = "cpm2c_head_5w3s_t8"
    ci, g = H.cpm2c_case_inputs(name), H.golden(name)
    net = H.build_cuda_cpm2c_model(ci, precision="fp32")
    ep = ci["episode"]
    su, qu = ci["feats"]
    # the library's own loss is reached through evaluate() (test above); here the runner's formula on the head outputs
    out = _run_head(net, ci)
    loss, acc, pred, _ = O.cpm2c_loss_and_acc({k: out[k].cpu() for k in ("class_logits", "logits_local", "logits_global")},
                                              ep["target_labels"], ep["real_support_labels"], ep["real_target_labels"],
                                              tasks_per_batch=H.CPM2C_TASKS_PER_BATCH)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    assert float(acc) == float(g["acc"])


def test_cpm2c_without_classification_returns_zero_class_logits():
    ci = H.cpm2c_case_inputs("cpm2c_head_5w3s_t8")
    g = H.golden("cpm2c_head_5w3s_t8")
    net = H.build_cuda_cpm2c_model(ci, use_classification=False)
    out = _run_head(net, ci)
    assert float(out["class_logits"].abs().max()) == 0.0
    assert H.rel_err(out["logits_local"], g["logits_local"]) < TOL_HEAD


def test_cpm2c_wrong_way_fails_loudly():
    ci = H.cpm2c_case_inputs("cpm2c_head_5w3s_t8")
    net = H.build_cuda_cpm2c_model(ci)
    net.way = 4     # the episode has 5 distinct support labels
    out = _run_head(net, ci)
    assert bool(torch.isnan(out["logits"]).all())
