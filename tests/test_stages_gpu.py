"""GPU parity tests proper: every stage of the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs AND against the golden tensors written from the executed reference.

Tolerances (BASELINE.json north_star): 2e-2 relative (max-abs-error / max-abs-reference) for the bf16 tensor-core
frame encoder; the head runs in fp32 with tf32 tensor-core products -> 5e-3; the OTAM tail is pure fp32 -> 1e-5.
Predicted classes must be identical on every query whose reference top-1/top-2 margin exceeds 4x the measured
logit error."""
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_BF16 = 2e-2
TOL_HEAD = 5e-3
TOL_OTAM = 1e-5


@pytest.mark.parametrize("P,W,Q,T,D,single", [(1, 5, 5, 8, 512, False), (3, 5, 5, 16, 512, False),
                                              (2, 5, 1, 8, 1024, False), (4, 7, 3, 8, 512, True),
                                              (1, 1, 1, 2, 512, False), (2, 32, 2, 30, 512, False),
                                              # >= 2 x 148 problems: one CTA per problem (all queries) on the tensor cores
                                              (300, 5, 5, 8, 512, False), (296, 5, 5, 16, 512, False),
                                              (300, 5, 6, 8, 1024, True), (297, 3, 2, 8, 512, False),
                                              (296, 10, 12, 8, 512, False)])
def test_otam_distance_matches_oracle(P, W, Q, T, D, single):
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(P * 100 + T)
    sup = torch.randn(P, W, T, D, generator=g)
    tgt = torch.randn(P, Q, T, D, generator=g) + 0.5 * sup[:, :1].expand(P, Q, T, D)
    ref = torch.stack([O.otam_distance(sup[p], tgt[p], single) for p in range(P)])
    out = ops.otam_distance(sup.cuda(), tgt.cuda(), single)
    assert H.rel_err(out, ref) < TOL_OTAM
    # accumulate form: out = beta*out + alpha*otam
    acc = ops.otam_distance(sup.cuda(), tgt.cuda(), single, alpha=0.5, beta=2.0, out=out.clone())
    assert H.rel_err(acc, 2.5 * ref) < 1e-5


@pytest.mark.parametrize("env", [{}, {"SPM_OTAM_TC_MINP": "296"}, {"SPM_OTAM_TC_MINP": "296", "SPM_OTAM_TC_MASK": "1", "SPM_OTAM_TC_PF": "1"},
                                 {"SPM_OTAM_DP": "log"}, {"SPM_OTAM": "stream"}, {"SPM_OTAM_TC": "0", "SPM_OTAM_FUSED": "0"},
                                 {"SPM_OTAM_TC": "0", "SPM_OTAM_KC": "16"}])
def test_otam_wavefront_formulations_agree(env):
    """exponent-domain (default, T <= 16) and log-domain wavefronts, in the tcgen05 (default from 12 x #SM problems), persistent mma.sync
    (default from 2 x #SM, both ring geometries), two-kernel and streaming kernels, against
    the oracle at 1e-5 -- including inputs whose distances are all ~0 or all ~2 (the ends of the exponent-domain range)"""
    import os, subprocess, sys
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(H.ROOT, "tools", "otam_dp_check.py")], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_otam_properties_full_size():
    """size-independent properties at the 1000-episode batch size of BASELINE config 2"""
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1000, 5, 8, 512, generator=g).cuda()
    y = torch.randn(1000, 5, 8, 512, generator=g).cuda()
    one = ops.otam_distance(x, y, True)
    two = ops.otam_distance(x, y, False)
    # bidirectional = dir(x,y) + dir(y,x)^T (the transposed DP is the single-direction DP of the swapped pair)
    swapped = ops.otam_distance(y, x, True)
    assert torch.allclose(two, one + swapped.transpose(1, 2), atol=1e-4)
    # scale invariance of the cosine: positive rescaling of a video's frames changes distances only through eps=0.01
    big = ops.otam_distance(x * 1e3, y * 1e3, False)
    assert (big - two).abs().max() < 0.05
    # a video against itself is closer than against an independent one
    self_d = ops.otam_distance(x[:, :1], x[:, :1], False)
    other_d = ops.otam_distance(x[:, :1], y[:, :1], False)
    assert bool((self_d < other_d).all())


@pytest.mark.parametrize("name", ["head_5w5s_t8", "head_5w1s_t16", "head_5w2s_t8_q3_single", "head_5w3s_t8_d1024"])
def test_head_matches_reference_golden(name):
    ci, g = H.case_inputs(name), H.golden(name)
    net = H.build_cuda_model(ci)
    ep = ci["episode"]
    su, qu = ci["feats"]
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])
    torch.cuda.synchronize()
    err = H.rel_err(out["logits"].cpu().unsqueeze(0), g["logits"])
    assert err < TOL_HEAD, err
    assert H.rel_err(out["dists"][0].cpu(), g["dists"]) < TOL_HEAD
    abs_err = float((out["logits"][0].cpu() - g["logits"][0]).abs().max())
    pred = out["logits"][0].argmax(-1).cpu()
    safe = g["margin"] > 4 * abs_err
    assert bool(safe.any())
    assert torch.equal(pred[safe], g["pred"].long()[safe])


def test_head_batched_episodes_equal_single_episodes():
    """E episodes in one call == E separate calls (episodes are independent units: SURVEY.md 8e)"""
    ci = H.case_inputs("head_5w5s_t8")
    net = H.build_cuda_model(ci)
    eps, sus, qus = [], [], []
    for e in range(3):
        ep = O.make_episode(2000 + e, 5, 5, 1, 8, 24, images=False)
        su, qu = O.make_features(2000 + e, 25, 5, 8, 512, ep["context_labels"], ep["target_labels"].float())
        eps.append(ep); sus.append(su); qus.append(qu)
    cat = lambda k: torch.stack([e[k] for e in eps])
    out = net.head(torch.stack(sus).cuda(), torch.stack(qus).cuda(), cat("context_labels"),
                   cat("real_support_labels"), cat("real_target_labels"), n_episodes=3)
    for e in range(3):
        one = net.head(sus[e].cuda(), qus[e].cuda(), eps[e]["context_labels"], eps[e]["real_support_labels"],
                       eps[e]["real_target_labels"])
        assert torch.allclose(out["logits"][e], one["logits"][0], atol=1e-4, rtol=1e-4)
        assert torch.allclose(out["dists"][e], one["dists"][0], atol=1e-4, rtol=1e-4)
        with torch.no_grad():
            ref = O.head_forward(ci["weights"], ci["text"], sus[e], qus[e], eps[e]["context_labels"],
                                 eps[e]["real_support_labels"], eps[e]["real_target_labels"], O.DEFAULT_PARAMS)
        assert H.rel_err(out["logits"][e].cpu(), ref["logits"][0]) < TOL_HEAD


def test_head_wrong_way_fails_loudly():
    ci = H.case_inputs("head_5w5s_t8")
    ci["way"] = 4  # episode really has 5 classes
    net = H.build_cuda_model(ci)
    ep = ci["episode"]
    su, qu = ci["feats"]
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])
    assert bool(torch.isnan(out["logits"]).all())
    # the failure is per call, not sticky: the same handle with the right W works again
    net.way = 5
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])
    assert bool(torch.isfinite(out["logits"]).all())


@pytest.mark.parametrize("name", ["vit_2w1s_t2_p0", "vit_5w1s_t8_p1"])
def test_vit_encoder_matches_reference_golden(name):
    ci, g = H.case_inputs(name), H.golden(name)
    net = H.build_cuda_model(ci)
    ep = ci["episode"]
    su = net.encode_frames(ep["context_images"].cuda())
    qu = net.encode_frames(ep["target_images"].cuda())
    torch.cuda.synchronize()
    assert H.rel_err(su.cpu().view(g["su"].shape), g["su"]) < TOL_BF16
    assert H.rel_err(qu.cpu().view(g["qu"].shape), g["qu"]) < TOL_BF16
    # frames are independent: encoding a subset gives the same rows (chunking / tile-boundary independence)
    part = net.encode_frames(ep["context_images"][1:3].cuda())
    assert torch.allclose(part, su[1:3], atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("name", ["vit_2w1s_t2_p0", "vit_5w1s_t8_p1"])
def test_forward_matches_reference_golden(name):
    """CNN.forward / loss / accuracy end to end (BASELINE config 1 is vit_5w1s_t8_p1)"""
    ci, g = H.case_inputs(name), H.golden(name)
    net = H.build_cuda_model(ci)
    ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()}
    out = net(ep)
    assert out["logits"].shape == g["logits"].shape and out["dists"].dim() == 0
    err = H.rel_err(out["logits"].cpu(), g["logits"])
    assert err < TOL_BF16, err
    assert H.rel_err(out["dists"].cpu(), g["dists"]) < 5e-2
    abs_err = float((out["logits"].cpu() - g["logits"]).abs().max())
    pred = out["logits"][0].argmax(-1).cpu()
    safe = g["margin"] > 4 * abs_err
    assert torch.equal(pred[safe], g["pred"].long()[safe])
    print("\n%s: logits rel err %.2e abs err %.4f; %d/%d queries above the margin filter, all agree; raw agreement %d/%d"
          % (name, err, abs_err, int(safe.sum()), safe.numel(), int((pred == g["pred"].long()).sum()), safe.numel()))
    loss, acc = net.evaluate(ep)
    if bool(safe.all()):
        assert float(acc) == float(g["acc"])
    assert abs(float(loss) - float(g["loss"])) < 5e-2 * max(1.0, abs(float(g["loss"])))
    # forward3 convenience form == dict form
    o3 = net.forward3(ep["context_images"], ep["context_labels"], ep["target_images"], ep["real_support_labels"],
                      ep["real_target_labels"])
    assert torch.equal(o3["logits"], out["logits"])


def test_forward_on_host_dict_equals_forward_on_device_dict():
    """the drop-in call with the DataLoader's host tensors (no prepare_task `.to(device)`): same logits / dists, host outputs,
    loss and accuracy attached"""
    ci = H.case_inputs("vit_5w1s_t8_p1")
    net = H.build_cuda_model(ci)
    ep_h = {k: (v.pin_memory() if torch.is_tensor(v) and v.dtype == torch.float32 and v.dim() == 4 else v)
            for k, v in ci["episode"].items()}
    out_h = net(ep_h)
    out_d = net({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()})
    assert not out_h["logits"].is_cuda and out_h["logits"].shape == out_d["logits"].shape
    assert torch.allclose(out_h["logits"], out_d["logits"].cpu(), atol=1e-5, rtol=1e-5)
    assert torch.allclose(out_h["dists"], out_d["dists"].cpu(), atol=1e-6, rtol=1e-5)
    loss, acc = net.evaluate({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()})
    assert abs(float(out_h["loss"]) - float(loss)) < 1e-5 and float(out_h["acc"]) == float(acc)


def test_eval_host_matches_device_path():
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci)
    eps = [O.make_episode(3000 + e, 2, 1, 1, 2, 24, "P0") for e in range(3)]
    cat = lambda k: torch.cat([e[k] for e in eps]).contiguous()
    host = net.evaluate_host(cat("context_images").pin_memory(), cat("context_labels"), cat("target_images").pin_memory(),
                             cat("real_support_labels"), cat("real_target_labels"), cat("target_labels"), 3, 2)
    for e in range(3):
        ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in eps[e].items()}
        out = net(ep)
        assert torch.allclose(host["logits"][e], out["logits"][0].cpu(), atol=1e-5)
        loss, acc = net.evaluate(ep)
        assert abs(float(host["loss"][e]) - float(loss)) < 1e-5 and float(host["acc"][e]) == float(acc)


def test_eval_host_next_call_prefetch_changes_nothing():
    """spm_eval_host_set_next: the first chunk of the next call is copied behind the current call's copies; a call
    that consumes the prefetch, one whose buffers do not match it, and one without any hint give identical results."""
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci, max_episodes=2)
    def batch(seed0):
        eps = [O.make_episode(seed0 + e, 2, 1, 1, 2, 24, "P0") for e in range(4)]
        cat = lambda k: torch.cat([e[k] for e in eps]).contiguous()
        return dict(su=cat("context_images").pin_memory(), qu=cat("target_images").pin_memory(), lab=cat("context_labels"),
                    rs=cat("real_support_labels"), rt=cat("real_target_labels"), tl=cat("target_labels"))
    A, B = batch(4000), batch(5000)
    run = lambda b, nxt=None: net.evaluate_host(b["su"], b["lab"], b["qu"], b["rs"], b["rt"], b["tl"], 4, 2,
                                                next_images=None if nxt is None else (nxt["su"], nxt["qu"]))
    plain_a, plain_b = run(A), run(B)
    r1 = run(A, nxt=B)          # prefetches B's first chunk
    r2 = run(B, nxt=B)          # consumes it, prefetches B again
    r3 = run(A)                 # prefetch present but for other buffers: ignored
    r4 = run(B)                 # no prefetch left (a hint lasts one call)
    for got, want in ((r1, plain_a), (r2, plain_b), (r3, plain_a), (r4, plain_b)):
        for k in ("logits", "dists", "loss", "acc", "pred"):
            assert torch.equal(got[k], want[k]), k
    assert not torch.equal(plain_a["logits"], plain_b["logits"])


@pytest.mark.parametrize("impl", ["tcgen05", "mma"])
@pytest.mark.parametrize("F,scale", [(1, 1.0), (3, 1.0), (40, 3.0), (160, 1.0)])
def test_vit_attention_matches_torch(impl, F, scale):
    """both attention kernels against fp32 softmax attention on the same bf16 q, k, v"""
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(F)
    qkv = (torch.randn(F * 197, 2304, generator=g) * scale).cuda().bfloat16()
    out = ops.vit_attention(qkv, F, impl)
    x = qkv.float().view(F, 197, 3, 12, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(x[0] @ x[1].transpose(-1, -2) / 8.0, dim=-1)
    ref = (att @ x[2]).permute(0, 2, 1, 3).reshape(F * 197, 768)
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-2, (impl, F, err)


@pytest.mark.parametrize("ln_fold", ["0", "1"])
def test_last_block_cls_pruning_is_exact(monkeypatch, ln_fold):
    """running the last transformer block on the class-token rows only (default) == running it on all tokens: the same
    arithmetic row by row with separate LayerNorm kernels (1e-5); with the LayerNorms folded into the GEMM epilogues the
    class-token rows take ln_2 -> c_fc through the LayerNorm kernel instead, equal up to bf16 rounding (5e-3)"""
    ci = H.case_inputs("vit_2w1s_t2_p0")
    imgs = ci["episode"]["context_images"].cuda()
    monkeypatch.setenv("SPM_LN_FOLD", ln_fold)
    pruned = H.build_cuda_model(ci).encode_frames(imgs)
    monkeypatch.setenv("SPM_PRUNE_LAST", "0")
    full = H.build_cuda_model(ci).encode_frames(imgs)
    tol = 1e-5 if ln_fold == "0" else 5e-3
    assert torch.allclose(pruned, full, atol=tol, rtol=tol)


@pytest.mark.parametrize("ln_fold", ["0", "2"])
def test_layernorm_variants_match_reference_golden(monkeypatch, ln_fold):
    """SPM_LN_FOLD=0 (separate LayerNorm kernels) and =2 (centred folded weights) against the executed reference, like the
    default (=1, exact folding) in test_vit_encoder_matches_reference_golden; a frame's features do not depend on the batch"""
    monkeypatch.setenv("SPM_LN_FOLD", ln_fold)
    ci, g = H.case_inputs("vit_5w1s_t8_p1"), H.golden("vit_5w1s_t8_p1")
    net = H.build_cuda_model(ci)
    su = net.encode_frames(ci["episode"]["context_images"].cuda())
    assert H.rel_err(su.cpu().view(g["su"].shape), g["su"]) < TOL_BF16
    part = net.encode_frames(ci["episode"]["context_images"][1:3].cuda())
    assert torch.equal(part, su[1:3])


def test_rn50_encoder_and_forward_match_reference_golden():
    """CLIP ModifiedResNet-50 tower (clip_fsar.py:593-608) + head with D=1024 (BASELINE config 4 architecture)"""
    name = "rn50_2w1s_t2_p1"
    ci, g = H.case_inputs(name), H.golden(name)
    net = H.build_cuda_model(ci)
    ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()}
    su = net.encode_frames(ep["context_images"])
    err_f = H.rel_err(su.cpu().view(g["su"].shape), g["su"])
    out = net(ep)
    err_l = H.rel_err(out["logits"].cpu(), g["logits"])
    print("\nrn50: feature rel err %.3e, logits rel err %.3e" % (err_f, err_l))
    assert err_f < TOL_BF16, err_f
    assert err_l < TOL_BF16, err_l
    abs_err = float((out["logits"].cpu() - g["logits"]).abs().max())
    safe = g["margin"] > 4 * abs_err
    assert torch.equal(out["logits"][0].argmax(-1).cpu()[safe], g["pred"].long()[safe])


TOL_FP32 = 1e-4   # north_star: logits and features within 1e-4 relative in the fp32 mode


@pytest.mark.parametrize("name", ["vit_2w1s_t2_p0", "vit_5w1s_t8_p1"])
def test_fp32_mode_matches_reference_golden(name):
    """SPM_PRECISION_FP32 (exact FFMA arithmetic): features, logits, dists, loss and EVERY predicted class"""
    ci, g = H.case_inputs(name), H.golden(name)
    net = H.build_cuda_model(ci, precision="fp32")
    ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()}
    su = net.encode_frames(ep["context_images"])
    err_f = H.rel_err(su.cpu().view(g["su"].shape), g["su"])
    out = net(ep)
    err_l = H.rel_err(out["logits"].cpu(), g["logits"])
    print("\n%s fp32 mode: feature rel err %.2e, logits rel err %.2e" % (name, err_f, err_l))
    assert err_f < TOL_FP32 and err_l < TOL_FP32
    assert H.rel_err(out["dists"].cpu(), g["dists"]) < 1e-3
    abs_err = float((out["logits"].cpu() - g["logits"]).abs().max())
    safe = g["margin"] > 4 * abs_err
    assert bool(safe.all()), "fp32 mode must resolve every query of the golden episodes"
    assert torch.equal(out["logits"][0].argmax(-1).cpu(), g["pred"].long())
    loss, acc = net.evaluate(ep)
    assert float(acc) == float(g["acc"])
    assert abs(float(loss) - float(g["loss"])) < 1e-3 * max(1.0, abs(float(g["loss"])))


@pytest.mark.parametrize("name", ["head_5w5s_t8", "head_5w1s_t16", "head_5w2s_t8_q3_single"])
def test_fp32_mode_head_matches_reference_golden(name):
    ci, g = H.case_inputs(name), H.golden(name)
    net = H.build_cuda_model(ci, precision="fp32")
    ep = ci["episode"]
    su, qu = ci["feats"]
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])
    assert H.rel_err(out["logits"].cpu().unsqueeze(0), g["logits"]) < TOL_FP32
    assert H.rel_err(out["dists"][0].cpu(), g["dists"]) < 1e-3
    assert torch.equal(out["logits"][0].argmax(-1).cpu(), g["pred"].long())


def test_otam_matches_c_restatement():
    """CUDA wavefront kernel against the plain-C oracle (oracle/otam_ref.c)"""
    from clip_spm_b200 import ops
    from tests.test_oracle_cpu import c_otam_distance
    g = torch.Generator().manual_seed(17)
    sup, tgt = torch.randn(5, 8, 512, generator=g), torch.randn(5, 8, 512, generator=g)
    out = ops.otam_distance(sup[None].cuda(), tgt[None].cuda(), False)[0].cpu()
    assert H.rel_err(out, c_otam_distance(sup, tgt)) < TOL_OTAM


def test_encoder_edge_cases():
    """single frame, frame counts that are not tile multiples, more frames than one workspace chunk (256)"""
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci)
    g = torch.Generator().manual_seed(5)
    imgs = torch.rand(3, 3, 224, 224, generator=g).cuda()
    one = net.encode_frames(imgs[:1])
    three = net.encode_frames(imgs)
    assert one.shape == (1, 512) and torch.allclose(one, three[:1], atol=2e-3, rtol=2e-3)
    big = imgs.repeat(87, 1, 1, 1)[:259]                      # 259 frames: chunks of 256 + 3
    out = net.encode_frames(big)
    assert out.shape == (259, 512)
    assert torch.allclose(out[256:259], three[(torch.arange(256, 259) % 3)], atol=2e-3, rtol=2e-3)
    assert torch.isfinite(out).all()
    assert net.encode_frames(imgs[:0]).shape == (0, 512)      # empty input: no launch, empty result


def test_head_ragged_episode_shapes():
    """1 query, unequal shots per class (the reference's torch.stack in taskM cannot even run this), W == S"""
    ci = H.case_inputs("head_5w5s_t8")
    g = torch.Generator().manual_seed(9)
    for labels, Q in ((torch.tensor([2., 0., 1., 0., 2., 2., 1.]), 1), (torch.tensor([0., 1., 2.]), 4)):
        ci2 = dict(ci); ci2["way"] = 3
        net = H.build_cuda_model(ci2)
        S = labels.numel()
        su, qu = torch.randn(S, 8, 512, generator=g), torch.randn(Q, 8, 512, generator=g)
        rs, rt = labels.clone(), torch.zeros(Q)
        out = net.head(su.cuda(), qu.cuda(), labels, rs, rt)
        assert out["logits"].shape == (1, Q, 3) and torch.isfinite(out["logits"]).all()
        if S == 3:   # equal shots: the oracle (like the reference) can run it
            with torch.no_grad():
                ref = O.head_forward(ci["weights"], ci["text"], su, qu, labels, rs, rt, O.DEFAULT_PARAMS)
            assert H.rel_err(out["logits"].cpu(), ref["logits"]) < TOL_HEAD


def test_full_size_episodes_bf16_path_agrees_with_fp32_mode():
    """BASELINE config 2 size (5-way 5-shot, 240 frames / episode), several episodes in one call: the bf16 tensor-core
    path against the exact fp32 mode (which is itself pinned to the executed reference at the golden sizes)."""
    E = 3
    backbone, way, shot, qpc, T, ncls = "ViT-B/16", 5, 5, 1, 8, 24
    w = O.make_weights(backbone, seed=0, protocol="P1")
    text = O.make_text_features(ncls, 512, seed=0)
    eps = [O.make_episode(4000 + e, way, shot, qpc, T, ncls, "P1") for e in range(E)]
    cat = lambda k: torch.cat([e[k] for e in eps]).cuda()
    stack = lambda k: torch.stack([e[k] for e in eps]).cuda()
    outs = {}
    for prec in ("fp32", "bf16"):
        from clip_spm_b200 import CNN
        net = CNN(H.make_cfg(backbone, T, False, way), text_features_test=text, max_episodes=E, precision=prec)
        net.load_state_dict(w, strict=True)
        outs[prec] = net.forward_episodes(cat("context_images"), stack("context_labels"), cat("target_images"),
                                          stack("real_support_labels"), stack("real_target_labels"), E,
                                          stack("target_labels"))
        torch.cuda.synchronize()
        del net
    ref, got = outs["fp32"], outs["bf16"]
    err = H.rel_err(got["logits"], ref["logits"])
    assert err < TOL_BF16, err
    abs_err = float((got["logits"] - ref["logits"]).abs().max())
    top2 = ref["logits"].topk(2, dim=-1).values
    safe = (top2[..., 0] - top2[..., 1]) > 4 * abs_err
    agree = got["logits"].argmax(-1) == ref["logits"].argmax(-1)
    assert bool(agree[safe].all())
    assert H.rel_err(got["dists"], ref["dists"]) < 5e-2
    print("\nfull-size: logits rel err %.2e (abs %.4f); %d/%d queries pass the margin filter, all agree; raw agreement %d/%d"
          % (err, abs_err, int(safe.sum()), safe.numel(), int(agree.sum()), agree.numel()))


@pytest.mark.parametrize("E,chunk", [(5, 4), (8, 6), (3, 8), (2, 3)])
def test_pipelined_forward_equals_serial(monkeypatch, E, chunk):
    """Several frame chunks + several episodes: the forward alternates chunks between two streams and runs the head
    of each episode group on a third stream while later groups are still being encoded.  The results must equal the
    single-stream schedule bit for bit (same kernels, same operands; only the overlap differs)."""
    ci = H.case_inputs("vit_2w1s_t2_p0")
    eps = [O.make_episode(4000 + e, 2, 1, 1, 2, 24, "P0") for e in range(E)]
    cat = lambda k: torch.cat([e[k] for e in eps]).contiguous().cuda()
    outs = []
    for streams in ("1", "2"):
        monkeypatch.setenv("SPM_ENC_STREAMS", streams)
        monkeypatch.setenv("SPM_FRAME_CHUNK", str(chunk))      # 4 frames per episode -> several chunks per call
        net = H.build_cuda_model(ci, E)
        for _ in range(2):   # second call reuses every workspace / plan / event
            out = net.forward_episodes(cat("context_images"), cat("context_labels"), cat("target_images"),
                                       cat("real_support_labels"), cat("real_target_labels"), E, cat("target_labels"))
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in out.items() if torch.is_tensor(v)})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
    assert torch.isfinite(outs[0]["logits"]).all()


# ------------------------------------------------------------------------------------------------------------------
# r02: BASELINE configs 2 / 3 / 4 at their full shapes against goldens written from the executed reference, and
# per-stage taps of the head (every SURVEY 8a row has its own assertion)
# ------------------------------------------------------------------------------------------------------------------
FULL_CASES = ["vit_5w5s_t8_p1", "vit_5w1s_t16_p1", "rn50_5w3s_t8_p1"]


def _check_forward_against_golden(net, ci, g, tol, tol_dists, all_argmax):
    ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ci["episode"].items()}
    su = net.encode_frames(ep["context_images"])
    qu = net.encode_frames(ep["target_images"])
    err_f = max(H.rel_err(su.cpu().view(g["su"].shape), g["su"]), H.rel_err(qu.cpu().view(g["qu"].shape), g["qu"]))
    out = net(ep)
    torch.cuda.synchronize()
    err_l = H.rel_err(out["logits"].cpu(), g["logits"])
    abs_err = float((out["logits"].cpu() - g["logits"]).abs().max())
    pred = out["logits"][0].argmax(-1).cpu()
    safe = g["margin"] > 4 * abs_err
    print("\nfeatures rel err %.2e, logits rel err %.2e (abs %.4f), %d/%d queries above the margin filter, raw agreement "
          "%d/%d" % (err_f, err_l, abs_err, int(safe.sum()), safe.numel(), int((pred == g["pred"].long()).sum()),
                     safe.numel()))
    assert err_f < tol, err_f
    assert err_l < tol, err_l
    assert H.rel_err(out["dists"].cpu(), g["dists"]) < tol_dists
    assert torch.equal(pred[safe], g["pred"].long()[safe])
    if all_argmax:
        assert bool(safe.all()) and torch.equal(pred, g["pred"].long())
    loss, acc = net.evaluate(ep)
    if bool(safe.all()):
        assert float(acc) == float(g["acc"])
    assert abs(float(loss) - float(g["loss"])) < max(tol_dists, tol) * max(1.0, abs(float(g["loss"])))


@pytest.mark.parametrize("name", FULL_CASES)
def test_full_shape_forward_matches_reference_golden(name):
    """tower + head together at the BASELINE shape: 240 frames (config 2), 160 frames T=16 (config 3), RN50 160 frames
    (config 4; its 120 support frames cross the tower's frame-chunk boundary)"""
    ci, g = H.case_inputs(name), H.golden(name)
    _check_forward_against_golden(H.build_cuda_model(ci), ci, g, TOL_BF16, 5e-2, False)


@pytest.mark.parametrize("name", ["vit_5w5s_t8_p1", "vit_5w1s_t16_p1"])
def test_full_shape_fp32_mode_matches_reference_golden(name):
    ci, g = H.case_inputs(name), H.golden(name)
    _check_forward_against_golden(H.build_cuda_model(ci, precision="fp32"), ci, g, TOL_FP32, 1e-3, True)


def test_rn50_encoder_frame_count_edges():
    """frame counts around the RN50 tower's chunk size: rows are independent of how many frames share a launch"""
    ci = H.case_inputs("rn50_2w1s_t2_p1")
    net = H.build_cuda_model(ci)
    g = torch.Generator().manual_seed(11)
    imgs = torch.rand(5, 3, 224, 224, generator=g).cuda()
    base = net.encode_frames(imgs)
    assert net.encode_frames(imgs[:1]).shape == (1, 1024)
    for n in (63, 64, 65, 129):
        big = imgs.repeat((n + 4) // 5, 1, 1, 1)[:n]
        out = net.encode_frames(big)
        assert out.shape == (n, 1024) and torch.isfinite(out).all()
        want = base[torch.arange(n) % 5]
        assert H.rel_err(out, want) < 2e-3, n
    assert net.encode_frames(imgs[:0]).shape == (0, 1024)


# golden key -> (spm_head_stage name)
STAGES = [("su_mo", "su_mo", "b1"), ("qu_mo", "qu_mo", "b1"), ("target_token", "target_token", "c3"),
          ("su_real", "su_real", "c1"), ("qu_fake", "qu_fake", "c1"), ("token_s_real", "token_s_real", "c1"),
          ("token_q_fake", "token_q_fake", "c1"), ("su_pro", "su_pro", "c6"), ("su_2", "su_2", "c5"),
          ("qu_2", "qu_2", "c5"), ("su_t2", "su_t2", "c5"), ("qu_t2", "qu_t2", "c5")]


def _check_head_stages(net, ci, g, tol):
    ep = ci["episode"]
    su, qu = ci["feats"] if ci["feats"] is not None else (g["su"], g["qu"])
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])
    worst = {}
    for key, stage, row in STAGES:
        got = net.head_stage(stage).cpu()
        assert got.numel() == g[key].numel(), (key, got.numel(), tuple(g[key].shape))
        err = H.rel_err(got.view(g[key].shape), g[key])
        worst[row] = max(worst.get(row, 0.0), err)
        assert err < tol, (key, row, err)
    # b2 / b3: mo() returns the distance before mo_alpha1 (model_clipspm.py:205,141)
    alpha1 = float(ci["weights"]["mo_alpha1"])
    err = abs(float(out["dists"][0]) / alpha1 - float(g["mo_dist_pre"])) / abs(float(g["mo_dist_pre"]))
    worst["b2/b3"] = err
    assert err < tol, ("mo_dist_pre", err)
    # a4: the support prompts are a plain gather of the text table
    tok = net.head_stage("support_token").cpu().view(-1, ci["D"])
    assert torch.equal(tok, ci["text"][ep["real_support_labels"].long()])
    print("\nper-row worst rel err: " + ", ".join("%s %.1e" % kv for kv in sorted(worst.items())))


@pytest.mark.parametrize("name", ["head_5w5s_t8", "head_5w1s_t16", "head_5w2s_t8_q3_single", "head_5w3s_t8_d1024"])
def test_head_stage_tensors_match_reference_golden(name):
    """rows b1-b3, c1, c3, c5, c6, a4 one by one (tf32 tensor-core products: 5e-3)"""
    ci, g = H.case_inputs(name), H.golden(name)
    _check_head_stages(H.build_cuda_model(ci), ci, g, TOL_HEAD)


@pytest.mark.parametrize("name", ["head_5w5s_t8", "head_5w1s_t16", "head_5w2s_t8_q3_single"])
def test_head_stage_tensors_fp32_mode(name):
    ci, g = H.case_inputs(name), H.golden(name)
    _check_head_stages(H.build_cuda_model(ci, precision="fp32"), ci, g, TOL_FP32)


@pytest.mark.parametrize("name", FULL_CASES)
def test_head_stage_tensors_on_reference_features_full_shape(name):
    """the head alone on the REFERENCE's own frame features of the full-shape goldens (isolates it from the tower)"""
    ci, g = H.case_inputs(name), H.golden(name)
    _check_head_stages(H.build_cuda_model(ci), ci, g, TOL_HEAD)


def test_head_stage_batched_and_errors():
    ci, g = H.case_inputs("head_5w5s_t8"), H.golden("head_5w5s_t8")
    net = H.build_cuda_model(ci)
    with pytest.raises(RuntimeError):
        net.head_stage("su_real")            # no head pass yet
    ep = ci["episode"]
    su, qu = ci["feats"]
    rep = lambda t: torch.stack([t, t])
    net.head(rep(su).cuda(), rep(qu).cuda(), rep(ep["context_labels"]), rep(ep["real_support_labels"]),
             rep(ep["real_target_labels"]), n_episodes=2)
    got = net.head_stage("su_2").cpu().view(2, *g["su_2"].shape)
    assert H.rel_err(got[0], g["su_2"]) < TOL_HEAD and H.rel_err(got[1], g["su_2"]) < TOL_HEAD
    with pytest.raises(RuntimeError):
        net.head_stage("no_such_stage")


def test_class_id_outside_text_table_fails_loudly():
    """a real_* label beyond the text table is an IndexError in the reference (model_clipspm.py:116-121): NaN logits
    here (device path, no host sync) instead of an out-of-bounds read"""
    ci = H.case_inputs("head_5w5s_t8")
    net = H.build_cuda_model(ci)
    ep = ci["episode"]
    su, qu = ci["feats"]
    bad = ep["real_target_labels"].clone()
    bad[0] = 1000.0
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"], ep["real_support_labels"], bad)
    assert bool(torch.isnan(out["logits"]).all())
    out = net.head(su.cuda(), qu.cuda(), ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"])
    assert bool(torch.isfinite(out["logits"]).all())


def test_eval_host_regrows_every_staging_ring():
    """a later call with the same frame counts but more logits per episode (larger Q*W) must not write past
    the staging rings sized by an earlier call; results must equal the device path for both shapes."""
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci, max_episodes=2)
    net.way = None
    def run(way, shot, qpc, seed0, E=3):
        eps = [O.make_episode(seed0 + e, way, shot, qpc, 2, 24, "P0") for e in range(E)]
        cat = lambda k: torch.cat([e[k] for e in eps]).contiguous()
        host = net.evaluate_host(cat("context_images").pin_memory(), cat("context_labels"),
                                 cat("target_images").pin_memory(), cat("real_support_labels"),
                                 cat("real_target_labels"), cat("target_labels"), E, way)
        for e in range(E):
            ep = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in eps[e].items()}
            out = net(ep)
            assert torch.allclose(host["logits"][e], out["logits"][0].cpu(), atol=1e-5)
    run(2, 2, 1, 6000)      # S=4, Q=2, W=2: Q*W = 4
    run(4, 1, 1, 6100)      # S=4, Q=4, W=4: same support frames, Q*W = 16
    run(2, 1, 2, 6200, 5)   # S=2, Q=4, W=2, more episodes than before
    # device path: scratch logits / dists (loss-only callers) regrow independently
    eps = [O.make_episode(6300 + e, 2, 1, 1, 2, 24, "P0") for e in range(4)]
    cat = lambda k: torch.cat([e[k] for e in eps]).contiguous().cuda()
    out = net.forward_episodes(cat("context_images"), cat("context_labels"), cat("target_images"),
                               cat("real_support_labels"), cat("real_target_labels"), 4, cat("target_labels"))
    assert torch.isfinite(out["loss"]).all()


@pytest.mark.parametrize("name", ["vit_2w1s_t2_p0", "vit_5w1s_t8_p1", "vit_5w5s_t8_p1"])
def test_bf16_residual_stream_mode_matches_reference_golden(name):
    """precision="bf16_resid": bf16 residual stream, i.e. the arithmetic of the reference's own autocast(bfloat16)
    forward (run/main_run.py:274); same 2e-2 tolerance and margin-filtered argmax as the default bf16 mode"""
    ci, g = H.case_inputs(name), H.golden(name)
    _check_forward_against_golden(H.build_cuda_model(ci, precision="bf16_resid"), ci, g, TOL_BF16, 5e-2, False)
