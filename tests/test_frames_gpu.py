"""GPU parity of the frame transform kernel (csrc/frame_transform.cu): bit-exact against the golden hashes written
from the reference's own Resize(256)/CenterCrop(224)/ToTensor chain and against the numpy restatement."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as P
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _gold():
    return {k: v.numpy() for k, v in H.golden("preprocess").items()}


@pytest.mark.parametrize("name", list(P.CASES))
def test_transform_frames_bit_exact_against_reference_golden(name):
    from clip_spm_b200.ops import transform_frames
    g = _gold()
    frames = P.make_frames(name)
    out = transform_frames(torch.from_numpy(frames).cuda()).cpu().numpy()
    assert out.shape == (frames.shape[0], 3, 224, 224) and out.dtype == np.float32
    assert np.array_equal(np.round(out[0, :, 100:116, :] * 255).astype(np.uint8), g[name + "/frame0_band"])
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest() == g[name + "/sha256"].tobytes()


@pytest.mark.parametrize("hw", [(224, 224), (225, 300), (1080, 1920), (2160, 3840), (900, 257), (256, 1024)])
def test_transform_frames_bit_exact_against_oracle_other_sizes(hw):
    """sizes beyond the golden set: no-crop, 4.2x and 8.4x antialiased down-scaling, extreme aspect ratios"""
    from clip_spm_b200.ops import transform_frames
    rng = np.random.RandomState(hw[0] * 7 + hw[1])
    frames = rng.randint(0, 256, size=(2, hw[0], hw[1], 3)).astype(np.uint8)
    frames[1, ::2] = 255
    frames[1, 1::2] = 0          # saturating stripes: exercises the clip to [0, 255] and the rounding offset
    out = transform_frames(torch.from_numpy(frames).cuda()).cpu().numpy()
    ref = P.preprocess_frames(frames)
    assert np.array_equal(out, ref), int((out != ref).sum())


def test_transform_frames_edge_cases():
    from clip_spm_b200.ops import transform_frames
    empty = transform_frames(torch.zeros(0, 256, 340, 3, dtype=torch.uint8, device="cuda"))
    assert empty.shape == (0, 3, 224, 224)
    const = transform_frames(torch.full((3, 300, 500, 3), 255, dtype=torch.uint8, device="cuda"))
    assert float(const.min()) == 1.0 and float(const.max()) == 1.0          # weights sum to one after rounding
    with pytest.raises(RuntimeError):
        transform_frames(torch.zeros(1, 256, 340, 3, dtype=torch.uint8))    # host tensor: no CPU path


def test_encode_frames_u8_equals_transform_then_encode():
    """uint8 frames straight into the encoder (bf16 patch matrix written by the transform kernel) == transforming to
    fp32 images first and encoding those -- the two routes produce the same patch matrix bit for bit"""
    ci = H.case_inputs("vit_2w1s_t2_p0")
    m = H.build_cuda_model(ci, 1, "bf16")
    frames = torch.from_numpy(P.make_frames("up_320x240")).cuda()
    from clip_spm_b200.ops import transform_frames
    a = m.encode_frames(transform_frames(frames))
    b = m.encode_frames_u8(frames)
    assert torch.equal(a, b)
    m32 = H.build_cuda_model(ci, 1, "fp32")
    assert torch.equal(m32.encode_frames(transform_frames(frames)), m32.encode_frames_u8(frames))


def test_eval_host_u8_equals_eval_host_on_transformed_frames():
    """whole path from decoded host frames: same logits/loss/acc/pred as feeding the transformed fp32 images"""
    from clip_spm_b200.ops import transform_frames
    name = "vit_2w1s_t2_p0"
    ci = H.case_inputs(name)
    ep = ci["episode"]
    backbone, way, shot, qpc, T = H.CASES[name][:5]
    S, Q, E = way * shot, way * qpc, 3
    rng = np.random.RandomState(5)
    su_u8 = torch.from_numpy(rng.randint(0, 256, size=(E * S * T, 240, 320, 3)).astype(np.uint8))
    qu_u8 = torch.from_numpy(rng.randint(0, 256, size=(E * Q * T, 240, 320, 3)).astype(np.uint8))
    lab = ep["context_labels"].float().repeat(E).contiguous()
    rs = ep["real_support_labels"].float().repeat(E).contiguous()
    rt = ep["real_target_labels"].float().repeat(E).contiguous()
    tl = ep["target_labels"].long().repeat(E).contiguous()
    m = H.build_cuda_model(ci, 2, "bf16")
    a = m.evaluate_host_u8(su_u8, lab, qu_u8, rs, rt, tl, E, way)
    su_f = transform_frames(su_u8.cuda()).cpu().contiguous()
    qu_f = transform_frames(qu_u8.cuda()).cpu().contiguous()
    b = m.evaluate_host(su_f, lab, qu_f, rs, rt, tl, E, way)
    for k in ("logits", "dists", "loss", "acc", "pred"):
        assert torch.equal(a[k], b[k]), k
    assert torch.isfinite(a["logits"]).all()


def test_rn50_encode_frames_u8():
    ci = H.case_inputs("rn50_2w1s_t2_p1")
    m = H.build_cuda_model(ci, 1, "bf16")
    from clip_spm_b200.ops import transform_frames
    frames = torch.from_numpy(P.make_frames("portrait_360x480")).cuda()
    assert torch.equal(m.encode_frames(transform_frames(frames)), m.encode_frames_u8(frames))


def test_listing_sweep_matches_per_episode_forward():
    """sampler plan -> decoded uint8 frames -> evaluate_host_u8 batches with next-call prefetch (sweep.run_listing_sweep)
    == the same episodes one at a time through transform_frames + CNN.evaluate (the reference loop's granularity)"""
    import random
    from clip_spm_b200 import frames as F, ops, sweep
    from oracle import clipspm_oracle as O
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci, max_episodes=2)
    g = torch.Generator().manual_seed(3)
    sp = F.Split()
    for vid in range(3):
        for cls in range(4):
            n = 4 + (cls + vid) % 3
            sp.add_vid([torch.randint(0, 256, (120, 160, 3), dtype=torch.uint8, generator=g) for _ in range(n)], cls)
    n_ep, way, shot, nq, T = 5, 2, 1, 1, 2
    res = sweep.run_listing_sweep(net, sp, lambda fr: fr, n_ep, way, shot, nq, seed=50, episodes_per_call=2)
    accs, losses = [], []
    for e in range(n_ep):
        plan = F.sample_episode_plan(sp, way, shot, nq, T, train=False, rng=random.Random(50 + e))
        grab = lambda items: ops.transform_frames(torch.stack([sp.videos[v][f] for v, fr in items for f in fr]).cuda())
        ep = dict(context_images=grab(plan["support"]), target_images=grab(plan["target"]),
                  context_labels=torch.tensor(plan["support_labels"]).cuda(),
                  real_support_labels=torch.tensor(plan["real_support_labels"]).cuda(),
                  real_target_labels=torch.tensor(plan["real_target_labels"]).cuda(),
                  target_labels=torch.tensor([int(x) for x in plan["target_labels"]]).cuda())
        loss, acc = net.evaluate(ep)
        accs.append(float(acc)); losses.append(float(loss))
    assert res["n"] == n_ep
    assert abs(res["accuracy"] - 100.0 * sum(accs) / n_ep) < 1e-4
    assert abs(res["loss"] - sum(losses) / n_ep) < 2e-3 * max(1.0, abs(sum(losses) / n_ep))


def test_training_transform_bit_exact_against_oracle():
    """ops.transform_frames_train (Resize -> mirror -> crop at the clip's random origin -> ToTensor) against the numpy
    restatement that is pinned on the reference's own train-mode __getitem__ pixels; per-frame origins, every flip state,
    corner origins, and geometries with real resampling"""
    import random
    from clip_spm_b200.ops import frame_geometry, transform_frames_train
    rng = random.Random(5)
    for name in ("k100_340x256", "up_320x240", "down_1280x720", "portrait_360x480", "odd_427x241", "square_256"):
        frames = P.make_frames(name)
        F, Hh, Ww = frames.shape[:3]
        oh, ow, _, _ = frame_geometry(Hh, Ww)
        augs = [(0, 0, False), (oh - 224, ow - 224, True), (0, ow - 224, True), (rng.randint(0, oh - 224), rng.randint(0, ow - 224), False),
                (rng.randint(0, oh - 224), rng.randint(0, ow - 224), True)]
        for y1, x1, fl in augs:
            ref = P.preprocess_frames_train(frames, y1, x1, fl)
            out = transform_frames_train(torch.from_numpy(frames).cuda(), (y1, x1, int(fl))).cpu().numpy()
            assert (out != ref).sum() == 0, (name, y1, x1, fl, int((out != ref).sum()))
    # one row per frame: two frames of a clip with different draws in ONE call
    frames = P.make_frames("k100_340x256")
    out = transform_frames_train(torch.from_numpy(frames).cuda(), [[3, 77, 1], [30, 5, 0]]).cpu().numpy()
    assert np.array_equal(out[0], P.preprocess_frames_train(frames[:1], 3, 77, True)[0])
    assert np.array_equal(out[1], P.preprocess_frames_train(frames[1:], 30, 5, False)[0])
    with pytest.raises(RuntimeError):
        transform_frames_train(torch.from_numpy(frames).cuda(), (40, 0, 0))      # 256 - 224 = 32 is the largest y origin


def test_training_episode_from_plan_matches_reference_golden():
    """a whole train-mode episode: plan (frames + draws) -> GPU training transform == the pixels the reference's loader produced"""
    import hashlib
    from clip_spm_b200.ops import transform_frames_train
    from tests.test_frames_cpu import AUG_CASES, aug_case, standin_frame
    for name in AUG_CASES:
        plan, sp, g, fh, fw = aug_case(name)
        for key, items in (("support_set", plan["support"]), ("target_set", plan["target"])):
            frames = np.stack([standin_frame(sp.videos[v][f], fh, fw) for v, fr, _ in items for f in fr])
            aug = [[a[0], a[1], int(a[2])] for _, fr, a in items for _f in fr]
            px = transform_frames_train(torch.from_numpy(frames).cuda(), aug).cpu().numpy()
            assert hashlib.sha256(np.ascontiguousarray(px).tobytes()).digest() == g[key + "_sha256"].tobytes(), (name, key)
