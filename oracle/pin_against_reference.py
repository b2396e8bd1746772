"""Pins oracle/clipspm_oracle.py against the REAL reference and writes the golden fixtures.

Runs only in the build container (needs /root/reference, CPU is enough):
    python oracle/pin_against_reference.py            # all cases, rewrites tests/golden/*.npz
It imports the reference's own modules through the import shim of SURVEY.md 8(c) (stub `ftfy`; `load` replaced by a
random-init CLIP(...) because checkpoints cannot be downloaded; Tensor.cuda a no-op on this CPU box), loads the
seeded synthetic state_dict of oracle.make_weights() key-for-key (strict), runs the reference's CNN.forward on the
seeded episodes, records every stage tensor by wrapping the reference's methods, asserts the oracle restatement
reproduces them, and stores them under tests/golden/ (small .npz files; weights and images are NOT stored -- they
regenerate from the seeds).  Nothing here is imported by the product path."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import clipspm_oracle as O  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    ftfy = types.ModuleType("ftfy")
    ftfy.fix_text = lambda s: s
    sys.modules.setdefault("ftfy", ftfy)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import models.clip_fsar as clip_fsar
    import models.model_clipspm as m

    def fake_load(name, device="cpu", cfg=None, jit=False):
        if name == "ViT-B/16":
            net = clip_fsar.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12)
        else:
            net = clip_fsar.CLIP(1024, 224, (3, 4, 6, 3), 64, None, 77, 49408, 512, 8, 12)
        return net.float().eval(), None

    m.load = fake_load
    return m


class NS(types.SimpleNamespace):
    pass


def build_reference(m, backbone, T, single_direct=False):
    cfg = NS(MODEL=NS(BACKBONE=backbone), TRAIN=NS(CLASS_NAME=["run"]), TEST=NS(CLASS_NAME=["run"]),
             DATA=NS(SEQ_LEN=T), DEVICE=NS(NUM_GPUS=1), params=dict(O.DEFAULT_PARAMS))
    if single_direct:
        cfg.MODEL.SINGLE_DIRECT = True
    torch.manual_seed(0)
    net = m.CNN(cfg).eval()
    return net


def record(net):
    """Wrap the reference's own methods to capture stage tensors (no reference source is modified)."""
    st = {}
    orig = dict(get_feats=net.get_feats, mo=net.mo, sem=net.sem, taskM=net.taskM, token_tr=net.token_tr.forward,
                get_motion_feats=net.get_motion_feats)

    def get_feats(*a, **k):
        r = orig["get_feats"](*a, **k)
        st["su"], st["qu"] = r[0].clone(), r[1].clone()
        return r

    def gmf(*a, **k):
        r = orig["get_motion_feats"](*a, **k)
        if "su_mo" not in st:
            st["su_mo"], st["qu_mo"] = r[0].clone(), r[1].clone()
        return r

    def mo(*a, **k):
        r = orig["mo"](*a, **k)
        st["mo_dist_pre"] = r.clone()
        return r

    def sem(*a, **k):
        r = orig["sem"](*a, **k)
        st["su_real"], st["qu_fake"], st["su_pro"] = r[0].clone(), r[1].clone(), r[2].clone()
        st["target_token"] = r[6].clone()
        st["token_q_fake"], st["token_s_real"] = r[8].clone().unsqueeze(1) if r[8].dim() == 2 else r[8].clone(), r[9].clone()
        return r

    def taskM(*a, **k):
        r = orig["taskM"](*a, **k)
        st["su_2"], st["qu_2"], st["su_t2"], st["qu_t2"] = [x.clone() for x in r]
        return r

    net.get_feats, net.get_motion_feats, net.mo, net.sem, net.taskM = get_feats, gmf, mo, sem, taskM
    return st


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


CASES = {
    # name: (backbone, way, shot, qpc, T, n_text_cls, protocol, head_only, single_direct, seed)
    "vit_5w1s_t8_p1": ("ViT-B/16", 5, 1, 1, 8, 24, "P1", False, False, 1000),      # BASELINE config 1
    "vit_2w1s_t2_p0": ("ViT-B/16", 2, 1, 1, 2, 24, "P0", False, False, 1001),      # tiny tower case, default-scale init
    "head_5w5s_t8": ("ViT-B/16", 5, 5, 1, 8, 24, "P1", True, False, 1002),         # config 2/5 head shape
    "head_5w1s_t16": ("ViT-B/16", 5, 1, 1, 16, 24, "P1", True, False, 1003),       # config 3 (OTAM 16x18 grid)
    "head_5w3s_t8_d1024": ("RN50", 5, 3, 1, 8, 10, "P1", True, False, 1004),       # config 4 head shape
    "head_5w2s_t8_q3_single": ("ViT-B/16", 5, 2, 3, 8, 24, "P1", True, True, 1005),  # SINGLE_DIRECT, 3 queries/class
    "rn50_2w1s_t2_p1": ("RN50", 2, 1, 1, 2, 10, "P1", False, False, 1006),         # RN50 tower
    # full BASELINE shapes, tower + head together (r02): ~20 s / 15 s / 10 s of reference CPU time each
    "vit_5w5s_t8_p1": ("ViT-B/16", 5, 5, 1, 8, 24, "P1", False, False, 1007),      # BASELINE config 2 / 5 (240 frames)
    "vit_5w1s_t16_p1": ("ViT-B/16", 5, 1, 1, 16, 24, "P1", False, False, 1008),    # BASELINE config 3 (160 frames, OTAM 16x18)
    "rn50_5w3s_t8_p1": ("RN50", 5, 3, 1, 8, 10, "P1", False, False, 1009),         # BASELINE config 4 (160 frames; 120 support
                                                                                   # frames cross the tower's frame-chunk boundary)
}


def run_case(m, name):
    backbone, way, shot, qpc, T, ncls, proto, head_only, single, seed = CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    net = build_reference(m, backbone, T, single)
    w = O.make_weights(backbone, seed=0, protocol=proto, head_only=False)
    missing, unexpected = net.load_state_dict(w, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    text = O.make_text_features(ncls, D, seed=0)
    net.text_features_test = text
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, proto, images=not head_only)
    cfg = dict(backbone=backbone, seq_len=T, mid_dim=D, params=O.DEFAULT_PARAMS, single_direct=single)
    st_ref = record(net)
    if head_only:
        su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
        net.get_feats = lambda *a, **k: (su, qu, None)
        ep["context_images"] = torch.zeros(1)
        ep["target_images"] = torch.zeros(1)
    with torch.no_grad():
        out = net(ep)
        if head_only:
            st_ref["su"], st_ref["qu"] = su, qu
            st = O.head_forward(w, text, su, qu, ep["context_labels"], ep["real_support_labels"],
                                ep["real_target_labels"], O.DEFAULT_PARAMS, single)
            st.update(su=su, qu=qu)
        else:
            st = O.forward(w, text, ep, cfg)
    st_ref["logits"], st_ref["dists"] = out["logits"], out["dists"].reshape(())
    # loss / accuracy through the reference's own utils (matplotlib stubbed: it is imported but unused there)
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import utils.utils as U
    loss_ref = U.loss(out["logits"], ep["target_labels"].long(), "cpu") / 16 + 0.001 * out["dists"]
    acc_ref = U.aggregate_accuracy(out["logits"], ep["target_labels"])
    loss, acc, pred = O.loss_and_acc(st["logits"], st["dists"], ep["target_labels"])
    st_ref["loss"], st_ref["acc"] = loss_ref.reshape(()), acc_ref.reshape(())
    st["loss"], st["acc"] = loss.reshape(()), acc.reshape(())
    worst = 0.0
    for k, v in st_ref.items():
        r = rel(st[k].reshape(v.shape), v)
        worst = max(worst, r)
        assert r < 2e-4, "oracle disagrees with the reference on %s/%s: rel err %.3e" % (name, k, r)
    lg = st_ref["logits"][0]
    top2 = lg.topk(2, dim=-1).values
    margin = (top2[:, 0] - top2[:, 1])
    print("%-24s oracle==reference, worst stage rel err %.2e | logits row spread %.3f, min top1-top2 margin %.4f, "
          "ref acc %.2f" % (name, worst, float((lg.max(-1).values - lg.min(-1).values).mean()), float(margin.min()),
                            float(acc_ref)))
    gold = {k: v.detach().float().numpy() for k, v in st_ref.items()}
    gold["pred"] = lg.argmax(-1).numpy()
    gold["margin"] = margin.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **gold)


TEXT_CLASSES = ["run", "jumping jacks", "pour water into a glass", "riding a bike"]


def run_text_case(m):
    """Text-prompt tower: the reference's own constructor (model_clipspm.py:45-70: 16 templates x class names ->
    tokenize -> CLIP.encode_text -> mean) with seeded text-tower weights, against the oracle + this repo's tokenizer."""
    import models.clip_fsar as clip_fsar
    from clip_spm_b200.tokenizer import ClipTokenizer, PROMPT_TEMPLATES
    wt = O.make_text_weights(512, seed=0)

    def fake_load(name, device="cpu", cfg=None, jit=False):
        net = clip_fsar.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).float().eval()
        missing, unexpected = net.load_state_dict(wt, strict=False)
        assert not unexpected and all(k.startswith("visual.") or k == "logit_scale" for k in missing), (missing, unexpected)
        return net, None

    old = m.load
    m.load = fake_load
    cfg = NS(MODEL=NS(BACKBONE="ViT-B/16"), TRAIN=NS(CLASS_NAME=["run"]), TEST=NS(CLASS_NAME=TEXT_CLASSES),
             DATA=NS(SEQ_LEN=8), DEVICE=NS(NUM_GPUS=1), params=dict(O.DEFAULT_PARAMS))
    with torch.no_grad():
        net = m.CNN(cfg)
    m.load = old
    ref = net.text_features_test
    tk = ClipTokenizer()
    tokens = torch.stack([tk.tokenize([t.format(c) for c in TEXT_CLASSES]) for t in PROMPT_TEMPLATES])   # [16, n_cls, 77]
    for ti, t in enumerate(PROMPT_TEMPLATES):
        assert torch.equal(tokens[ti], clip_fsar.tokenize([t.format(c) for c in TEXT_CLASSES]).int())
    with torch.no_grad():
        mine = O.class_text_features(wt, tokens)
    r = rel(mine, ref)
    assert r < 2e-4, r
    print("%-24s oracle==reference (text_features_test of CNN.__init__), rel err %.2e, |feat| max %.3f" %
          ("text_tower_4cls", r, float(ref.abs().max())))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "text_tower_4cls.npz"), tokens=tokens.numpy(),
                        text_features=ref.float().numpy())


def run_otam_grad_case(m):
    """Gradients of the metric tail: autograd through the REFERENCE's own CNN.otam_distance (model_clipspm.py:348-362:
    cos_sim + OTAM_cum_dist_v2 in both directions) against autograd through the oracle restatement; the reference
    gradients become the golden of the CUDA backward kernel (inputs regenerate from the seed)."""
    W, Q, T, D, seed = 3, 2, 8, 512, 77
    sup, tgt, go = O.make_otam_grad_inputs(W, Q, T, D, seed)
    net = types.SimpleNamespace(args=NS(MODEL=NS()))
    s1, t1 = sup.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    out_ref = m.CNN.otam_distance(net, s1, t1)
    (out_ref * go).sum().backward()
    s2, t2 = sup.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    out = O.otam_distance(s2, t2, False)
    (out * go).sum().backward()
    r = max(rel(out, out_ref.detach()), rel(s2.grad, s1.grad), rel(t2.grad, t1.grad))
    assert r < 2e-4, r
    print("%-24s oracle autograd == reference autograd (CNN.otam_distance), worst rel err %.2e" % ("otam_grad_3w2q_t8", r))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "otam_grad_3w2q_t8.npz"), out=out_ref.detach().numpy(),
                        grad_support=s1.grad.numpy(), grad_target=t1.grad.numpy(),
                        shape=np.array([W, Q, T, D, seed], np.int32))


# name: backbone (-> D), way, shot, queries per class, T, n text classes, SINGLE_DIRECT, seed
HEAD_GRAD_CASES = {
    "head_grad_5w2s_t8": ("ViT-B/16", 5, 2, 1, 8, 24, False, 4001),
    "head_grad_3w1s_t4_d1024_single": ("RN50", 3, 1, 2, 4, 10, True, 4002),
}


def run_head_grad_case(m, name):
    """Training step through the head (run/main_run.py:245-254): the REFERENCE's own CNN in train mode (prompt rows from
    text_features_train, model_clipspm.py:116-118; every nn.Dropout set to p = 0 on the instance, no source is modified),
    frame features in place of get_feats, loss of :390-392 through the reference's utils.loss, `.backward()`.  The oracle's
    autograd (head_loss_and_grads) must reproduce the loss and every parameter / feature gradient; the reference gradients
    become the golden (full tensors up to 4096 entries, else a fixed sample of 4096 + the L2 norm)."""
    backbone, way, shot, qpc, T, ncls, single, seed = HEAD_GRAD_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    net = build_reference(m, backbone, T, single)
    w = O.make_weights(backbone, seed=0, protocol="P1", head_only=False)
    missing, unexpected = net.load_state_dict(w, strict=False)
    assert not missing and not unexpected
    net.train()
    n_drop = 0
    for mod in net.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
            n_drop += 1
    assert n_drop >= 8   # context1 / context2 (3 each) + token_tr (2)
    text_train = O.make_text_features(ncls, D, seed=1)
    net.text_features_train = text_train
    net.text_features_test = None          # the train branch must not touch it
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=False)
    su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    su_r, qu_r = su.clone().requires_grad_(True), qu.clone().requires_grad_(True)
    net.get_feats = lambda *a, **k: (su_r, qu_r, None)
    ep["context_images"] = torch.zeros(1)
    ep["target_images"] = torch.zeros(1)
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import utils.utils as U
    out = net(ep)
    loss_ref = U.loss(out["logits"], ep["target_labels"].long(), "cpu") / 16 + 0.001 * out["dists"]
    loss_ref.backward()
    ref = {k: p.grad.detach() for k, p in net.named_parameters() if not k.startswith("backbone.") and p.grad is not None}
    ref["su"], ref["qu"] = su_r.grad.detach(), qu_r.grad.detach()
    loss, grads = O.head_loss_and_grads(w, text_train, su, qu, ep["context_labels"], ep["real_support_labels"],
                                        ep["real_target_labels"], ep["target_labels"], O.DEFAULT_PARAMS, single)
    assert set(ref) == set(grads), (sorted(set(ref) ^ set(grads)))
    worst = rel(loss.reshape(()), loss_ref.detach().reshape(()))
    gold = {"loss": loss_ref.detach().numpy().reshape(()), "logits": out["logits"].detach().numpy()}
    for k, g in ref.items():
        r = rel(grads[k].reshape(g.shape), g)
        worst = max(worst, r)
        assert r < 2e-4, "oracle autograd disagrees with the reference on d loss / d %s: rel err %.3e" % (k, r)
        flat = g.reshape(-1)
        gold["g:" + k] = flat[O.grad_sample_index(flat.numel())].numpy()
        gold["n:" + k] = np.float64(flat.double().norm())
    print("%-24s oracle autograd == reference autograd through CNN.forward (train mode, p = 0): %d gradients, worst rel err "
          "%.2e, loss %.5f" % (name, len(ref), worst, float(loss_ref)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **gold)


# name: backbone (-> D), way, shot, queries per class, T, n train classes, SINGLE_DIRECT, options, seed
FSAR_GRAD_CASES = {
    "fsar_grad_5w2s_t8": ("ViT-B/16", 5, 2, 1, 8, 30, False, {}, 4201),
    "fsar_grad_3w2s_t4_d1024_depth2": ("RN50", 3, 2, 2, 4, 12, True, dict(depth=2), 4202),
}


def run_fsar_grad_case(m, name):
    """CLIP-FSAR's training iteration: the REFERENCE's CNN_OTAM_CLIPFSAR in train mode (MODEL.USE_CLASSIFICATION, dropout
    p = 0 on the instances, prompt rows from text_features_train :197-198), frame features in place of get_feats, loss of
    run/main_run.py:355-356, `.backward()`; the oracle's autograd must reproduce every gradient (context2.*, scale, features)."""
    backbone, way, shot, qpc, T, ntrain, single, opt, seed = FSAR_GRAD_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    f = import_reference_fsar(m)
    cfg = NS(MODEL=NS(BACKBONE=backbone, USE_CLASSIFICATION=True), TRAIN=NS(CLASS_NAME=["run"]), TEST=NS(CLASS_NAME=["run"]),
             DATA=NS(SEQ_LEN=T), DEVICE=NS(NUM_GPUS=1))
    if single:
        cfg.MODEL.SINGLE_DIRECT = True
    if opt.get("depth", 1) > 1:
        cfg.MODEL.TRANSFORMER_DEPTH = opt["depth"]
        cfg.TRAIN.TRANSFORMER_DEPTH = opt["depth"]
    torch.manual_seed(0)
    net = f.CNN_OTAM_CLIPFSAR(cfg)
    w = O.make_fsar_weights(D, seed=0, depth=opt.get("depth", 1))
    missing, unexpected = net.load_state_dict(w, strict=False)
    assert not unexpected and all(k.startswith("backbone.") for k in missing)
    net.train()
    for mod in net.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    text_train = O.make_text_features(ntrain, D, seed=1)
    net.text_features_train, net.text_features_test = text_train, None
    ep = O.make_episode(seed, way, shot, qpc, T, ntrain, "P1", images=False)
    su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    su_r, qu_r = su.clone().requires_grad_(True), qu.clone().requires_grad_(True)
    net.get_feats = lambda *a, **k: (su_r, qu_r, None)
    ep["context_images"], ep["target_images"] = torch.zeros(1), torch.zeros(1)
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import utils.utils as U
    out = net(ep)
    real = torch.cat([ep["real_support_labels"], ep["real_target_labels"]], 0).long()
    loss_ref = (U.loss(out["logits"], ep["target_labels"].long(), "cpu")
                + FSAR_CLS_VALUE * U.loss(out["class_logits"], real, "cpu")) / FSAR_TASKS_PER_BATCH
    loss_ref.backward()
    ref = {k: p.grad.detach() for k, p in net.named_parameters() if not k.startswith("backbone.") and p.grad is not None}
    ref["su"], ref["qu"] = su_r.grad.detach(), qu_r.grad.detach()
    loss, grads = O.fsar_head_loss_and_grads(w, text_train, su, qu, ep["context_labels"], ep["real_support_labels"],
                                             ep["real_target_labels"], ep["target_labels"], FSAR_TASKS_PER_BATCH,
                                             FSAR_CLS_VALUE, single, **opt)
    assert set(ref) == set(grads), sorted(set(ref) ^ set(grads))
    worst = rel(loss.reshape(()), loss_ref.detach().reshape(()))
    gold = {"loss": loss_ref.detach().numpy().reshape(()), "logits": out["logits"].detach().numpy(),
            "class_logits": out["class_logits"].detach().numpy()}
    for k, g in ref.items():
        r = rel(grads[k].reshape(g.shape), g)
        worst = max(worst, r)
        assert r < 2e-4, "oracle autograd disagrees with the reference on d loss / d %s: rel err %.3e" % (k, r)
        flat = g.reshape(-1)
        gold["g:" + k] = flat[O.grad_sample_index(flat.numel())].numpy()
        gold["n:" + k] = np.float64(flat.double().norm())
    print("%-24s oracle autograd == reference autograd through CNN_OTAM_CLIPFSAR.forward (train mode, p = 0): %d gradients, "
          "worst rel err %.2e, loss %.5f" % (name, len(ref), worst, float(loss_ref.detach())))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **gold)


# name: backbone, way, shot, queries per class, T, n text classes, seed  (frames -> loss -> every gradient, tower included)
TRAIN_CASES = {
    "train_vit_2w1s_t2": ("ViT-B/16", 2, 1, 1, 2, 24, 4101),
}


def run_train_case(m, name):
    """The whole training step's backward: the REFERENCE's CNN in train mode (dropout p = 0 on the instances) from the frames,
    loss through utils.loss, `.backward()` -> gradients of every parameter of the CLIP visual tower and of the head.  The
    oracle's autograd (train_loss_and_grads) must reproduce them; the reference gradients become the golden (samples)."""
    backbone, way, shot, qpc, T, ncls, seed = TRAIN_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    net = build_reference(m, backbone, T, False)
    w = O.make_weights(backbone, seed=0, protocol="P1", head_only=False)
    missing, unexpected = net.load_state_dict(w, strict=False)
    assert not missing and not unexpected
    net.train()
    for mod in net.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    text_train = O.make_text_features(ncls, D, seed=1)
    net.text_features_train, net.text_features_test = text_train, None
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=True)
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import utils.utils as U
    out = net(ep)
    loss_ref = U.loss(out["logits"], ep["target_labels"].long(), "cpu") / 16 + 0.001 * out["dists"]
    loss_ref.backward()
    ref = {k: p.grad.detach() for k, p in net.named_parameters() if p.grad is not None}
    cfg = dict(backbone=backbone, seq_len=T, mid_dim=D, params=O.DEFAULT_PARAMS, single_direct=False)
    loss, grads = O.train_loss_and_grads(w, text_train, ep, cfg)
    assert set(ref) == set(grads), sorted(set(ref) ^ set(grads))
    assert any(k.startswith("backbone.") for k in ref)
    worst = rel(loss.reshape(()), loss_ref.detach().reshape(()))
    gold = {"loss": loss_ref.detach().numpy().reshape(()), "logits": out["logits"].detach().numpy()}
    for k, g in ref.items():
        r = rel(grads[k].reshape(g.shape), g)
        worst = max(worst, r)
        assert r < 5e-4, "oracle autograd disagrees with the reference on d loss / d %s: rel err %.3e" % (k, r)
        flat = g.reshape(-1)
        gold["g:" + k] = flat[O.grad_sample_index(flat.numel(), 1024)].numpy()
        gold["n:" + k] = np.float64(flat.double().norm())
    print("%-24s oracle autograd == reference autograd, frames -> loss (train mode, p = 0): %d gradients (%d of the tower), "
          "worst rel err %.2e, loss %.5f" % (name, len(ref), sum(k.startswith("backbone.") for k in ref), worst,
                                             float(loss_ref.detach())))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **gold)


FSAR_CASES = {
    # name: (backbone, way, shot, qpc, T, n_test_cls, n_train_cls, head_only, single_direct, seed)
    "fsar_head_5w5s_t8": ("ViT-B/16", 5, 5, 1, 8, 24, 30, True, False, 2002),
    "fsar_head_5w3s_t8_d1024_q2": ("RN50", 5, 3, 2, 8, 10, 12, True, False, 2004),
    "fsar_head_5w1s_t16_single": ("ViT-B/16", 5, 1, 1, 16, 24, 30, True, True, 2005),
    "fsar_vit_2w1s_t2_p1": ("ViT-B/16", 2, 1, 1, 2, 24, 30, False, False, 2001),
    # the two optional branches that work in the reference (EVAL_TEXT / COMBINE raise at model_clipfsar.py:384)
    "fsar_head_5w3s_t8_merge": ("ViT-B/16", 5, 3, 1, 8, 24, 30, True, False, 2006),      # MODEL.MERGE_BEFORE
    "fsar_head_5w2s_t8_depth2": ("ViT-B/16", 5, 2, 2, 8, 24, 30, True, False, 2007),     # TRANSFORMER_DEPTH = 2
}
FSAR_OPTIONS = {"fsar_head_5w3s_t8_merge": dict(merge_before=True), "fsar_head_5w2s_t8_depth2": dict(depth=2)}
FSAR_TASKS_PER_BATCH, FSAR_CLS_VALUE = 4, 3.0   # configs/clipfsar/ssv2_otam.yaml: TASKS_PER_BATCH, USE_CLASSIFICATION_VALUE


def import_reference_fsar(m):
    """models/model_clipfsar.py as shipped only binds load / tokenize / Transformer_v1 / cos_sim /
    extract_class_indices when run as __main__ (its module-level imports are commented out, :8-9, :402-403); the
    runner therefore cannot construct it unmodified.  The shim binds exactly those names from the reference's own
    modules -- what the commented lines say -- before the class is used."""
    import models.clip_fsar as clip_fsar
    import models.myRes as myres
    import models.model_clipfsar as f
    f.load, f.tokenize = m.load, clip_fsar.tokenize
    f.Transformer_v1, f.cos_sim, f.extract_class_indices = myres.Transformer_v1, myres.cos_sim, myres.extract_class_indices
    return f


def run_fsar_case(m, name):
    backbone, way, shot, qpc, T, ncls, ntrain, head_only, single, seed = FSAR_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    f = import_reference_fsar(m)
    cfg = NS(MODEL=NS(BACKBONE=backbone), TRAIN=NS(CLASS_NAME=["run"]), TEST=NS(CLASS_NAME=["run"]),
             DATA=NS(SEQ_LEN=T), DEVICE=NS(NUM_GPUS=1))
    if single:
        cfg.MODEL.SINGLE_DIRECT = True
    opt = FSAR_OPTIONS.get(name, {})
    if opt.get("merge_before"):
        cfg.MODEL.MERGE_BEFORE = True
    if opt.get("depth", 1) > 1:
        cfg.MODEL.TRANSFORMER_DEPTH = opt["depth"]   # the flag (model_clipfsar.py:143) ...
        cfg.TRAIN.TRANSFORMER_DEPTH = opt["depth"]   # ... and the value the constructor actually reads (:144)
    torch.manual_seed(0)
    with torch.no_grad():
        net = f.CNN_OTAM_CLIPFSAR(cfg).eval()
    w = O.make_fsar_weights(D, seed=0, depth=opt.get("depth", 1))
    if not head_only:
        w.update({k: v for k, v in O.make_weights(backbone, seed=0, protocol="P1").items() if k.startswith("backbone.")})
    missing, unexpected = net.load_state_dict(w, strict=False)
    assert not unexpected and all(k.startswith("backbone.") for k in missing) and (head_only or not missing), (missing, unexpected)
    text_test = O.make_text_features(ncls, D, seed=0)
    text_train = O.make_text_features(ntrain, D, seed=1)
    net.text_features_test, net.text_features_train = text_test, text_train
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=not head_only)
    st_ref = {}
    ctx_calls = []
    orig_ctx = net.context2.forward
    net.context2.forward = lambda q, k, v: (ctx_calls.append(orig_ctx(q, k, v)), ctx_calls[-1])[1]
    if head_only:
        su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
        net.get_feats = lambda *a, **k: (su, qu, None)
        ep["context_images"] = torch.zeros(1)
        ep["target_images"] = torch.zeros(1)
    else:
        orig_gf = net.get_feats

        def get_feats(*a, **k):
            r = orig_gf(*a, **k)
            st_ref["su"], st_ref["qu"] = r[0].clone(), r[1].clone()
            return r
        net.get_feats = get_feats
    with torch.no_grad():
        out = net(ep)
        if not head_only:
            su, qu = st_ref["su"], st_ref["qu"]
            enc = O.vit_forward if backbone == "ViT-B/16" else O.rn50_forward
            mine_su = enc(w, ep["context_images"]).reshape(-1, T, D)
            mine_qu = enc(w, ep["target_images"]).reshape(-1, T, D)
            assert rel(mine_su, su) < 2e-4 and rel(mine_qu, qu) < 2e-4
        st = O.fsar_head_forward(w, text_test, text_train, su, qu, ep["context_labels"], ep["real_support_labels"],
                                 ep["real_target_labels"], single, **opt)
    st_ref["qu_ctx"], st_ref["su_ctx"] = ctx_calls[0], ctx_calls[1][:, :T]
    st_ref["logits"], st_ref["class_logits"] = out["logits"], out["class_logits"]
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import utils.utils as U
    real = torch.cat([ep["real_support_labels"], ep["real_target_labels"]], 0).long()
    loss_ref = (U.loss(out["logits"], ep["target_labels"].long(), "cpu")
                + FSAR_CLS_VALUE * U.loss(out["class_logits"], real, "cpu")) / FSAR_TASKS_PER_BATCH   # main_run.py:355-356
    acc_ref = U.aggregate_accuracy(out["logits"], ep["target_labels"])
    loss, acc, pred = O.fsar_loss_and_acc(st["logits"], st["class_logits"], ep["target_labels"], ep["real_support_labels"],
                                          ep["real_target_labels"], FSAR_TASKS_PER_BATCH, FSAR_CLS_VALUE)
    st_ref["loss"], st_ref["acc"] = loss_ref.reshape(()), acc_ref.reshape(())
    st["loss"], st["acc"] = loss.reshape(()), acc.reshape(())
    st.update(su=su, qu=qu)
    worst = 0.0
    for k, v in st_ref.items():
        r = rel(st[k].reshape(v.shape), v)
        worst = max(worst, r)
        assert r < 2e-4, "oracle disagrees with the reference on %s/%s: rel err %.3e" % (name, k, r)
    lg = st_ref["logits"][0]
    top2 = lg.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    print("%-28s oracle==reference (CNN_OTAM_CLIPFSAR), worst stage rel err %.2e | min top1-top2 margin %.4f, ref acc %.2f"
          % (name, worst, float(margin.min()), float(acc_ref)))
    gold = {k: v.detach().float().numpy() for k, v in st_ref.items()}
    gold["pred"] = lg.argmax(-1).numpy()
    gold["margin"] = margin.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **gold)


def run_fsar_dead_branches(m):
    """MODEL.EVAL_TEXT and MODEL.COMBINE (model_clipfsar.py:235-322) both finish with `class_text_logits = None` and then
    `class_text_logits.unsqueeze(0)` at :384: executed here to record that they raise, i.e. are not part of the path."""
    f = import_reference_fsar(m)
    D, T = 512, 8
    for flag in ("EVAL_TEXT", "COMBINE"):
        cfg = NS(MODEL=NS(BACKBONE="ViT-B/16"), TRAIN=NS(CLASS_NAME=["run"]), TEST=NS(CLASS_NAME=["run"]),
                 DATA=NS(SEQ_LEN=T), DEVICE=NS(NUM_GPUS=1))
        setattr(cfg.MODEL, flag, True)
        torch.manual_seed(0)
        with torch.no_grad():
            net = f.CNN_OTAM_CLIPFSAR(cfg).eval()
        net.load_state_dict(O.make_fsar_weights(D, seed=0), strict=False)
        net.text_features_test, net.text_features_train = O.make_text_features(24, D, seed=0), O.make_text_features(30, D, seed=1)
        ep = O.make_episode(2002, 5, 1, 1, T, 24, "P1", images=False)
        su, qu = O.make_features(2002, 5, 5, T, D, ep["context_labels"], ep["target_labels"].float())
        text = net.text_features_test[ep["real_support_labels"].long()]
        net.get_feats = lambda *a, **k: (su, qu, text)
        ep["context_images"] = ep["target_images"] = torch.zeros(1)
        try:
            with torch.no_grad():
                net(ep)
        except AttributeError as e:
            print("%-28s reference raises %s: %s" % ("fsar MODEL." + flag, type(e).__name__, e))
        else:
            raise AssertionError("MODEL.%s unexpectedly works in the reference" % flag)


STEN_CASES = {
    # name: (backbone, way, shot, qpc, n_test_cls, head_only, seed)   -- T is 8 (models/model_sten.py:65 hard-codes it)
    "sten_head_5w5s": ("ViT-B/16", 5, 5, 1, 24, True, 2102),
    "sten_head_5w3s_d1024_q2": ("RN50", 5, 3, 2, 10, True, 2104),
}


def run_sten_case(m, name):
    """models/model_sten.py (its relative imports work as shipped).  Its constructor needs `load`; the forward only
    uses the backbone, so head-only cases replace the backbone by a lookup of seeded features."""
    backbone, way, shot, qpc, ncls, head_only, seed = STEN_CASES[name]
    T, D = 8, 512 if backbone == "ViT-B/16" else 1024
    import models.model_sten as sten
    sten.load = m.load
    cfg = NS(MODEL=NS(BACKBONE=backbone), TRAIN=NS(CLASS_NAME=["run"]), TEST=NS(CLASS_NAME=["run"]),
             DATA=NS(SEQ_LEN=T), DEVICE=NS(NUM_GPUS=1))
    torch.manual_seed(0)
    with torch.no_grad():
        net = sten.CNN_OTAM_CLIPFSAR(cfg).eval()
    text = O.make_text_features(ncls, D, seed=0)
    net.text_features_test = text
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=False)
    su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    feats = {"s": su.reshape(-1, D), "q": qu.reshape(-1, D)}
    ep["context_images"], ep["target_images"] = "s", "q"
    class Lookup(torch.nn.Module):
        def forward(self, key):
            return feats[key]
    net.backbone = Lookup()
    with torch.no_grad():
        out = net(ep)
        st = O.sten_head_forward(text, su, qu, ep["context_labels"], ep["real_support_labels"])
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import utils.utils as U
    loss_ref = U.loss(out["logits"], ep["target_labels"].long(), "cpu") / 16          # run/main_run.py:394-395
    acc_ref = U.aggregate_accuracy(out["logits"], ep["target_labels"])
    loss, acc, pred = O.loss_and_acc(st["logits"], torch.zeros(()), ep["target_labels"])
    r = max(rel(st["logits"], out["logits"]), rel(loss.reshape(()), loss_ref.reshape(())))
    assert r < 2e-4 and float(acc) == float(acc_ref), (r, acc, acc_ref)
    lg = out["logits"][0]
    top2 = lg.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    print("%-28s oracle==reference (models/model_sten.py), rel err %.2e | min top1-top2 margin %.4f, ref acc %.2f"
          % (name, r, float(margin.min()), float(acc_ref)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), logits=out["logits"].numpy(),
                        loss=loss_ref.reshape(()).numpy(), acc=acc_ref.reshape(()).numpy(), pred=lg.argmax(-1).numpy(),
                        margin=margin.numpy())


CPM2C_CASES = {
    # name: (backbone, way, shot, qpc, T, n_test_cls, head_only, single_direct, seed)
    "cpm2c_head_5w3s_t8": ("ViT-B/16", 5, 3, 1, 8, 24, True, False, 2202),          # configs/cpm2c/hmdb.yaml shape (3-shot)
    "cpm2c_head_5w1s_t8_d1024_q2": ("RN50", 5, 1, 2, 8, 10, True, False, 2204),     # the shipped backbone (RN50, D=1024)
    "cpm2c_head_5w2s_t6_single": ("ViT-B/16", 5, 2, 1, 6, 24, True, True, 2205),    # SINGLE_DIRECT, even T-1... odd motion length 5
    "cpm2c_vit_2w1s_t4_p1": ("ViT-B/16", 2, 1, 1, 4, 24, False, False, 2201),       # tower + head
}
CPM2C_TASKS_PER_BATCH = 16


def run_cpm2c_case(m, name):
    """models/model_cpm2c.py::CLIP_CPMMC_FSAR.  run/run.py's params dict lacks the keys this class reads
    (prompt_patch, hid_dim, ..., motion_residual_ratio, lambdas0-3), so the shipped runner cannot construct it; the
    values used here are O.CPM2C_PARAMS (the constructor-only ones size nets the forward never calls)."""
    backbone, way, shot, qpc, T, ncls, head_only, single, seed = CPM2C_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    import models.model_cpm2c as c2
    c2.load = m.load
    cfg = NS(MODEL=NS(BACKBONE=backbone, MOTION_COFF=1.0, NORMAL_COFF=0.7, USE_CLASSIFICATION=True),
             TRAIN=NS(CLASS_NAME=["run"]), TEST=NS(CLASS_NAME=["run"]), DATA=NS(SEQ_LEN=T), DEVICE=NS(NUM_GPUS=1),
             params=dict(O.CPM2C_PARAMS))
    if single:
        cfg.MODEL.SINGLE_DIRECT = True
    torch.manual_seed(0)
    with torch.no_grad():
        net = c2.CLIP_CPMMC_FSAR(cfg).eval()
    w = O.make_cpm2c_weights(D, seed=0)
    if not head_only:
        w.update({k: v for k, v in O.make_weights(backbone, seed=0, protocol="P1").items() if k.startswith("backbone.")})
    missing, unexpected = net.load_state_dict(w, strict=False)
    head_keys = set(O.cpm2c_weight_shapes(D))
    assert not unexpected and not (set(missing) & head_keys) and (head_only or not [k for k in missing if k.startswith("backbone.")]), \
        (missing, unexpected)
    text = O.make_text_features(ncls, D, seed=0)
    net.text_features_test = text
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=not head_only)
    st_ref = {}
    if head_only:
        su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
        net.get_feats = lambda *a, **k: (su, qu, None)
        ep["context_images"] = torch.zeros(1)
        ep["target_images"] = torch.zeros(1)
    else:
        orig_gf = net.get_feats

        def get_feats(*a, **k):
            r = orig_gf(*a, **k)
            st_ref["su"], st_ref["qu"] = r[0].clone(), r[1].clone()
            return r
        net.get_feats = get_feats
    orig_mo, orig_eh = net.get_motion_feats, net.text_eh_temporal_transformer
    eh_calls = []

    def gmf(*a, **k):
        r = orig_mo(*a, **k)
        st_ref["su_motion"], st_ref["qu_motion"] = r[0].clone(), r[1].clone()
        return r

    def eh(*a, **k):
        r = orig_eh(*a, **k)
        eh_calls.append([x.clone() for x in r])
        return r
    net.get_motion_feats, net.text_eh_temporal_transformer = gmf, eh
    with torch.no_grad():
        out = net(ep)
        if not head_only:
            su, qu = st_ref["su"], st_ref["qu"]
            enc = O.vit_forward if backbone == "ViT-B/16" else O.rn50_forward
            assert rel(enc(w, ep["context_images"]).reshape(-1, T, D), su) < 2e-4
        st = O.cpm2c_head_forward(w, text, su, qu, ep["context_labels"], ep["real_support_labels"], ep["real_target_labels"],
                                  O.CPM2C_PARAMS, motion_coeff=1.0, normal_coeff=0.7, single_direct=single)
    st_ref["su_real_motion"], st_ref["qu_fake_motion"] = eh_calls[0][0], eh_calls[0][1]
    st_ref["su_real"], st_ref["qu_fake"], st_ref["su_pro"] = eh_calls[1][0], eh_calls[1][1], eh_calls[1][2]
    for k in ("class_logits", "logits_local", "logits_global"):
        st_ref[k] = out[k]
    st_ref["target_consist_distance"] = out["target_consist_distance"].reshape(())
    st["target_consist_distance"] = st["target_consist_distance"].reshape(())
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import utils.utils as U
    P = O.CPM2C_PARAMS
    real = torch.cat([ep["real_support_labels"], ep["real_target_labels"]], 0).long()
    total_ref = P["lambdas1"] * out["logits_local"] + P["lambdas2"] * out["logits_global"]      # run/main_run.py:373
    loss_ref = (P["lambdas0"] * U.loss(out["class_logits"], real, "cpu") + P["lambdas1"] * U.loss(out["logits_local"], ep["target_labels"].long(), "cpu")
                + P["lambdas2"] * U.loss(out["logits_global"], ep["target_labels"].long(), "cpu")) / CPM2C_TASKS_PER_BATCH
    acc_ref = U.aggregate_accuracy(total_ref, ep["target_labels"])
    loss, acc, pred, total = O.cpm2c_loss_and_acc(st, ep["target_labels"], ep["real_support_labels"], ep["real_target_labels"],
                                                  P, CPM2C_TASKS_PER_BATCH)
    st_ref["loss"], st_ref["acc"], st_ref["logits_total"] = loss_ref.reshape(()), acc_ref.reshape(()), total_ref
    st["loss"], st["acc"], st["logits_total"] = loss.reshape(()), acc.reshape(()), total
    st.update(su=su, qu=qu)
    worst = 0.0
    for k, v in st_ref.items():
        r = rel(st[k].reshape(v.shape), v)
        worst = max(worst, r)
        assert r < 2e-4, "oracle disagrees with the reference on %s/%s: rel err %.3e" % (name, k, r)
    lg = total_ref[0]
    top2 = lg.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    print("%-28s oracle==reference (CLIP_CPMMC_FSAR), worst stage rel err %.2e | min top1-top2 margin %.4f, ref acc %.2f"
          % (name, worst, float(margin.min()), float(acc_ref)))
    gold = {k: v.detach().float().numpy() for k, v in st_ref.items()}
    gold["pred"] = lg.argmax(-1).numpy()
    gold["margin"] = margin.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **gold)


SOFTDTW_CASES = {
    # name: (B, N, M, d, gamma, bandwidth, seed)
    "softdtw_8x8_g01": (6, 8, 8, 64, 0.1, 0.0, 3101),        # TA2N's setting: SoftDTW(gamma=0.1), models/model_ta2n.py:87
    "softdtw_17x15_g1": (4, 17, 15, 8, 1.0, 0.0, 3102),      # the reference's own profile() shape, models/OTAM.py:517
    "softdtw_40x38_bw5": (3, 40, 38, 8, 0.5, 5.0, 3103),     # more rows than a warp, Sakoe-Chiba band
}


def run_softdtw_case(name):
    """models/OTAM.py executed here through its CPU path (numba @jit compute_softdtw / compute_softdtw_backward and the
    SoftDTW module with use_cuda=False) -- the path its own test (profile(), :461-505) holds the CUDA kernels against."""
    import models.OTAM as ref
    B, N, M, d, gamma, bw, seed = SOFTDTW_CASES[name]
    X, Y, D = O.make_softdtw_inputs(B, N, M, d, seed)
    R_ref = ref.compute_softdtw(D.numpy(), gamma, bw)
    E_ref = ref.compute_softdtw_backward(D.numpy(), R_ref.copy(), gamma, bw)
    R = O.softdtw_forward_np(D.numpy(), gamma, bw)
    E = O.softdtw_backward_np(D.numpy(), R, gamma, bw)
    fin = np.isfinite(R_ref)
    assert (np.isfinite(R) == fin).all() and np.abs(R[fin] - R_ref[fin]).max() < 1e-9 and np.abs(E - E_ref).max() < 1e-9
    mod = ref.SoftDTW(use_cuda=False, gamma=gamma, bandwidth=bw if bw > 0 else None)
    x = X.clone().requires_grad_(True)
    val = mod(x, Y)
    val.sum().backward()
    mine = O.softdtw_module(X, Y, gamma, False, bw)
    nrm = ref.SoftDTW(use_cuda=False, gamma=gamma, normalize=True, bandwidth=bw if bw > 0 else None)(X[:, :min(N, M)], Y[:, :min(N, M)])
    mine_n = O.softdtw_module(X[:, :min(N, M)], Y[:, :min(N, M)], gamma, True, bw)
    r = max(rel(mine, val.detach()), rel(mine_n, nrm))
    assert r < 1e-5, r
    print("%-28s oracle==reference (models/OTAM.py CPU path): table/gradient exact to 1e-9, module rel err %.2e" % (name, r))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), R=R_ref.astype(np.float32),
                        E=E_ref.astype(np.float32), module=val.detach().numpy(), module_norm=nrm.numpy(),
                        grad_x=x.grad.numpy())


if __name__ == "__main__":
    m = import_reference()
    names = sys.argv[1:] or (list(CASES) + ["text", "otam_grad"] + list(FSAR_CASES) + ["fsar_dead_branches"] + list(STEN_CASES) + list(CPM2C_CASES) + list(SOFTDTW_CASES) + list(HEAD_GRAD_CASES) + list(TRAIN_CASES) + list(FSAR_GRAD_CASES))
    for n in names:
        if n == "text":
            run_text_case(m)
        elif n == "fsar_dead_branches":
            run_fsar_dead_branches(m)
        elif n == "otam_grad":
            run_otam_grad_case(m)
        elif n in FSAR_CASES:
            run_fsar_case(m, n)
        elif n in STEN_CASES:
            run_sten_case(m, n)
        elif n in CPM2C_CASES:
            run_cpm2c_case(m, n)
        elif n in SOFTDTW_CASES:
            run_softdtw_case(n)
        elif n in HEAD_GRAD_CASES:
            run_head_grad_case(m, n)
        elif n in TRAIN_CASES:
            run_train_case(m, n)
        elif n in FSAR_GRAD_CASES:
            run_fsar_grad_case(m, n)
        else:
            run_case(m, n)
