"""Training step of the head, CPU side: the oracle's autograd against the goldens written from the REFERENCE's own
`CNN.forward(...)` + `.backward()` in train mode (oracle/pin_against_reference.py head_grad_*), and the host composition
of clip_spm_b200/train.py (which dense node is fed what) with its CUDA nodes swapped for torch stand-ins."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import clipspm_oracle as O
from tests.helpers import golden

# must match oracle/pin_against_reference.py::HEAD_GRAD_CASES
HEAD_GRAD_CASES = {
    "head_grad_5w2s_t8": ("ViT-B/16", 5, 2, 1, 8, 24, False, 4001),
    "head_grad_3w1s_t4_d1024_single": ("RN50", 3, 1, 2, 4, 10, True, 4002),
}


def head_grad_inputs(name):
    backbone, way, shot, qpc, T, ncls, single, seed = HEAD_GRAD_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    w = O.make_weights(backbone, seed=0, protocol="P1", head_only=True)
    text_train = O.make_text_features(ncls, D, seed=1)
    ep = O.make_episode(seed, way, shot, qpc, T, ncls, "P1", images=False)
    su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    return dict(backbone=backbone, D=D, T=T, way=way, single=single, w=w, text=text_train, ep=ep, su=su, qu=qu)


def check_against_golden(grads, loss, gold, tol, l2=False):
    """every gradient the reference produced: its stored sample (or full tensor) and its L2 norm.  l2: the sample is judged by
    its relative L2 error (tf32 products: a pre-activation within rounding of zero takes the other LeakyReLU slope, which
    moves single entries by far more than the rest) with a loose cap on single entries"""
    names = sorted(k[2:] for k in gold if k.startswith("g:"))
    assert sorted(grads) == names
    assert abs(float(loss) - float(gold["loss"])) < tol * max(1.0, abs(float(gold["loss"])))
    for k in names:
        flat = grads[k].detach().cpu().double().reshape(-1)
        ref = gold["g:" + k].double()
        mine = flat[O.grad_sample_index(flat.numel())]
        scale = float(ref.abs().max().clamp_min(1e-30))
        if l2:
            assert float((mine - ref).norm()) < tol * float(ref.norm()) + 1e-30, (k, float((mine - ref).norm()), float(ref.norm()))
            assert float((mine - ref).abs().max()) < 10 * tol * scale, (k, float((mine - ref).abs().max()), scale)
        else:
            assert float((mine - ref).abs().max()) < tol * scale, (k, float((mine - ref).abs().max()), scale)
        assert abs(float(flat.norm()) - float(gold["n:" + k])) < tol * float(gold["n:" + k]) + 1e-30, k


@pytest.mark.parametrize("name", list(HEAD_GRAD_CASES))
def test_oracle_autograd_matches_reference_golden(name):
    ci = head_grad_inputs(name)
    ep = ci["ep"]
    loss, grads = O.head_loss_and_grads(ci["w"], ci["text"], ci["su"], ci["qu"], ep["context_labels"],
                                        ep["real_support_labels"], ep["real_target_labels"], ep["target_labels"],
                                        O.DEFAULT_PARAMS, ci["single"])
    check_against_golden(grads, loss, golden(name), 2e-4)


class _Block:
    """stand-in for train.TransformerV1: the oracle's restatement of ONE layer of the block"""

    def __init__(self, dim_head=256):
        self.dim_head = dim_head

    def __call__(self, x, w, prefix, dropout_seed=None):
        assert dropout_seed is None
        layer = {k.replace(prefix, "L.layers.0."): v for k, v in w.items() if k.startswith(prefix)}
        return O.transformer_v1(x, layer, "L.", dim_head=self.dim_head)


def _linear(x, W, b=None, act="none", slope=0.0, exact=False):
    y = F.linear(x, W, b)
    return {"none": lambda t: t, "gelu": F.gelu, "sigmoid": torch.sigmoid,
            "leaky_relu": lambda t: F.leaky_relu(t, slope)}[act](y)


def _otam(support, target, single_direct=False):
    return torch.stack([O.otam_distance(support[p], target[p], single_direct) for p in range(support.shape[0])])


@pytest.mark.parametrize("name", list(HEAD_GRAD_CASES))
def test_train_head_composition_matches_oracle(name, monkeypatch):
    """train.spm_head_forward batches the four se_te calls, the motion passes and the class means differently from the
    reference; with torch stand-ins for its CUDA nodes it must still be the oracle's head, value and gradients"""
    from clip_spm_b200 import train
    monkeypatch.setattr(train, "linear", _linear)
    monkeypatch.setattr(train, "otam_distance", _otam)
    ci = head_grad_inputs(name)
    ep = ci["ep"]
    w = {k: v.clone().requires_grad_(True) for k, v in ci["w"].items() if v.dtype.is_floating_point}
    su, qu = ci["su"].clone().requires_grad_(True), ci["qu"].clone().requires_grad_(True)
    out = train.spm_head_forward(w, ci["text"], su, qu, ep["context_labels"], ep["real_support_labels"],
                                 ep["real_target_labels"], O.DEFAULT_PARAMS, _Block(), _Block(), ci["single"])
    loss = train.spm_loss(out, ep["target_labels"], 16.0)
    loss.backward()
    grads = {k: v.grad for k, v in w.items() if v.grad is not None}
    grads["su"], grads["qu"] = su.grad, qu.grad
    gold = golden(name)
    assert torch.allclose(out["logits"].detach(), gold["logits"], atol=2e-4, rtol=1e-4)
    check_against_golden(grads, loss.detach(), gold, 5e-4)


# must match oracle/pin_against_reference.py::FSAR_GRAD_CASES (+ FSAR_TASKS_PER_BATCH, FSAR_CLS_VALUE)
FSAR_GRAD_CASES = {
    "fsar_grad_5w2s_t8": ("ViT-B/16", 5, 2, 1, 8, 30, False, {}, 4201),
    "fsar_grad_3w2s_t4_d1024_depth2": ("RN50", 3, 2, 2, 4, 12, True, dict(depth=2), 4202),
}
FSAR_TPB, FSAR_CLS = 4, 3.0


def fsar_grad_inputs(name):
    backbone, way, shot, qpc, T, ntrain, single, opt, seed = FSAR_GRAD_CASES[name]
    D = 512 if backbone == "ViT-B/16" else 1024
    w = O.make_fsar_weights(D, seed=0, depth=opt.get("depth", 1))
    text = O.make_text_features(ntrain, D, seed=1)
    ep = O.make_episode(seed, way, shot, qpc, T, ntrain, "P1", images=False)
    su, qu = O.make_features(seed, way * shot, way * qpc, T, D, ep["context_labels"], ep["target_labels"].float())
    return dict(backbone=backbone, D=D, T=T, way=way, single=single, opt=opt, w=w, text=text, ep=ep, su=su, qu=qu)


@pytest.mark.parametrize("name", list(FSAR_GRAD_CASES))
def test_fsar_oracle_autograd_and_train_composition_match_reference_golden(name, monkeypatch):
    """CLIP-FSAR's training branch: the oracle's autograd, and train.fsar_head_forward with torch stand-ins for its CUDA
    nodes, against the gradients of the reference's own backward"""
    from clip_spm_b200 import train
    ci = fsar_grad_inputs(name)
    ep, gold = ci["ep"], golden(name)
    loss, grads = O.fsar_head_loss_and_grads(ci["w"], ci["text"], ci["su"], ci["qu"], ep["context_labels"],
                                             ep["real_support_labels"], ep["real_target_labels"], ep["target_labels"],
                                             FSAR_TPB, FSAR_CLS, ci["single"], **ci["opt"])
    check_against_golden(grads, loss, gold, 2e-4)
    monkeypatch.setattr(train, "linear", _linear)
    monkeypatch.setattr(train, "otam_distance", _otam)
    w = {k: v.clone().requires_grad_(True) for k, v in ci["w"].items()}
    su, qu = ci["su"].clone().requires_grad_(True), ci["qu"].clone().requires_grad_(True)
    out = train.fsar_head_forward(w, ci["text"], su, qu, ep["context_labels"], ep["real_support_labels"],
                                  ep["real_target_labels"], _Block(ci["D"] // 8), ci["opt"].get("depth", 1), ci["single"])
    loss2 = train.fsar_loss(out, ep["target_labels"], ep["real_support_labels"], ep["real_target_labels"], FSAR_TPB, FSAR_CLS)
    loss2.backward()
    g2 = {k: v.grad for k, v in w.items() if v.grad is not None}
    g2["su"], g2["qu"] = su.grad, qu.grad
    assert torch.allclose(out["class_logits"].detach(), gold["class_logits"], atol=2e-4, rtol=1e-4)
    check_against_golden(g2, loss2.detach(), gold, 5e-4)


def _dp_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    from clip_spm_b200 import train
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    params = [torch.zeros(s, requires_grad=True) for s in [(5,), (300, 7), (3, 4, 5), (1,), (70000,)]]
    task_grads = [[torch.randn(p.shape, generator=g) for p in params] for _ in range(6)]    # the same on every rank
    for t in train.shard_tasks(6, rank, world):
        for p, tg in zip(params, task_grads[t]):
            p.grad = tg.clone() if p.grad is None else p.grad + tg
    train.allreduce_gradients(params, bucket_numel=4096)
    want = [sum(task_grads[t][i] for t in range(6)) for i in range(len(params))]
    out[rank] = max(float((p.grad - w).abs().max()) for p, w in zip(params, want))
    dist.destroy_process_group()


def test_task_sharded_gradients_equal_single_process_accumulation():
    """world_size 2 over gloo (the GPU path is the same code over NCCL): tasks split over the ranks + one bucketed
    all-reduce(SUM) == the gradients one process accumulates over all TASKS_PER_BATCH tasks"""
    import socket
    import torch.multiprocessing as mp
    from clip_spm_b200 import train
    assert [list(train.shard_tasks(16, r, 3)) for r in range(3)] == [list(range(0, 5)), list(range(5, 10)), list(range(10, 16))]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.Manager().dict()
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert len(out) == 2 and max(out.values()) < 1e-5, dict(out)


def test_philox_known_answers_and_mask_statistics():
    """oracle.philox4x32_10 against the Random123 known-answer vectors of Philox4x32-10; the keep rate of dropout_mask"""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert [int(v) for v in O.philox4x32_10(ctr, key)] == want
    m = O.dropout_mask((1000, 101), 0.2, 99, 0)
    assert set(np.unique(m.numpy()).round(4)) == {0.0, 1.25} and abs(float((m > 0).float().mean()) - 0.8) < 5e-3
    assert not torch.equal(m, O.dropout_mask((1000, 101), 0.2, 99, 1)) and not torch.equal(m, O.dropout_mask((1000, 101), 0.2, 98, 0))


def test_training_ops_fail_loudly_without_gpu():
    """the differentiable nodes have no CPU path: CPU tensors (or a box without a GPU) raise instead of falling back"""
    from clip_spm_b200 import CNN, train
    from clip_spm_b200.config import make_cfg
    x, W = torch.randn(4, 64), torch.randn(32, 64)
    for fn in (lambda: train.linear(x, W), lambda: train.layer_norm(x, torch.ones(64), torch.zeros(64)),
               lambda: train.dropout(x, 0.5, 1), lambda: train.TransformerV1(64)(torch.randn(1, 2, 64), {}),
               lambda: train.VitBlock()(torch.randn(1, 197, 768), {}, "b.")):
        with pytest.raises(RuntimeError):
            fn()
    if not torch.cuda.is_available():
        net = CNN(make_cfg("ViT-B/16", 2), text_features_test=torch.zeros(4, 512), text_features_train=torch.zeros(4, 512))
        net.train()
        with pytest.raises(RuntimeError):
            net.head(torch.zeros(1, 2, 2, 512), torch.zeros(1, 2, 2, 512), torch.tensor([0., 1.]), torch.tensor([0., 1.]),
                     torch.tensor([0., 1.]))


def test_training_loop_schedule_on_cpu_stand_ins(monkeypatch):
    """host logic of train.run_listing_training with stand-ins for the CUDA pieces: which iterations step the optimiser
    ((iteration + 1) % TASKS_PER_BATCH == 0 or the last one, run/main_run.py:203-209), MultiStepLR per iteration (:99,210), the
    episodes are the seeded train-mode plans (same for every rank split), frames go through the training transform with the
    plan's draws, and a 2-rank split computes disjoint iterations that together cover all of them"""
    import random
    from clip_spm_b200 import frames as Fr
    from clip_spm_b200 import ops, train
    sp = Fr.Split()
    for vid in range(3):
        for cls in range(3):
            sp.add_vid([(cls, vid, f) for f in range(6 + vid)], cls)
    H_, W_ = 256, 288
    load = lambda h: standin_frame_small(h, H_, W_)    # noqa: E731
    seen_aug = []

    def fake_transform(frames, aug):
        seen_aug.append([tuple(a) for a in aug])
        return torch.zeros(frames.shape[0], 3, 224, 224)
    monkeypatch.setattr(ops, "transform_frames_train", fake_transform)
    monkeypatch.setattr(ops, "frame_geometry", lambda h, w: P_geometry(h, w))
    monkeypatch.setattr(Fr, "frame_geometry", lambda h, w: P_geometry(h, w))

    class FakeNet:
        training, seq_len, tasks_per_batch, _dev = True, 2, 3.0, "cpu"

        def __init__(self):
            self.p = torch.nn.Parameter(torch.zeros(2))
            self.calls = 0

        def trainable_parameters(self):
            return [self.p]

        def __call__(self, inputs):
            self.calls += 1
            q = inputs["target_labels"].numel()
            return {"logits": (self.p.view(1, 1, 2) + torch.zeros(1, q, 2))}

        def loss(self, out, tl, rs, rt):
            return out["logits"].sum() * 0.0 + self.p.sum()

    class FakeOpt:
        def __init__(self):
            self.param_groups, self.steps, self.lrs = [{"lr": 1.0}], [], []

        def zero_grad(self, set_to_none=True):
            pass

    class FakeScaler:
        def __init__(self, opt, it):
            self.opt, self.it = opt, it

        def scale(self, loss):
            return loss

        def step(self, opt):
            opt.steps.append(self.it[0])
            opt.lrs.append(opt.param_groups[0]["lr"])

        def update(self):
            pass

    def run(rank, world):
        net, opt, it = FakeNet(), FakeOpt(), [0]
        log = train.run_listing_training(net, sp, load, 8, 2, 1, 1, opt, FakeScaler(opt, it), seed=5, lr_milestone=6, rank=rank,
                                         world_size=world, on_iteration=lambda i, l, a: it.__setitem__(0, i))
        return net, opt, log
    monkeypatch.setattr(train, "allreduce_gradients", lambda params, group=None: None)
    net, opt, log = run(0, 1)
    assert net.calls == 8 and len(log) == 8
    # optimiser steps: iterations where (i + 1) % 3 == 0 -> 2, 5, 8 (8 is also the last one)
    assert opt.steps == [2, 5, 8]
    assert opt.lrs == [1.0, 1.0, 0.1]                      # MultiStepLR(milestones=[6]) fires after iteration 6
    # the draws handed to the transform are the plan's, one row per frame, support clips then target clips
    plan = Fr.sample_episode_plan(sp, 2, 1, 1, 2, train=True, rng=random.Random(5 + 1), frame_size=(H_, W_), flip=True)
    want = [tuple([a[0], a[1], int(a[2])]) for _, fr, a in plan["support"] for _f in fr]
    assert seen_aug[0] == want
    # two ranks: disjoint iterations (i % 2 == rank), every optimiser step on both
    n0, o0, l0 = run(0, 2)
    n1, o1, l1 = run(1, 2)
    assert n0.calls + n1.calls == 8 and len(l0) == 4 and len(l1) == 4


def standin_frame_small(triple, H, W):
    cls, vid, f = triple
    return np.full((H, W, 3), (cls * 31 + vid * 7 + f) % 256, np.uint8)


def P_geometry(h, w):
    from oracle import preprocess_oracle as P
    return P.geometry(h, w)
