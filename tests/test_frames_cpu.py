"""CPU suite of the input-path row: the numpy restatement of Resize(256)/CenterCrop(224)/ToTensor against the golden
hashes written from the reference's own transform chain (oracle/pin_preprocess.py), the frame sampler mirror against
the reference's VideoDataset.get_seq outputs, and the library's host-side geometry arithmetic."""
import hashlib

import numpy as np
import pytest

from oracle import preprocess_oracle as P
from tests import helpers as H


def _gold():
    return {k: v.numpy() for k, v in H.golden("preprocess").items()}


@pytest.mark.parametrize("name", ["k100_340x256", "up_320x240", "portrait_360x480", "tiny_176x100", "odd_427x241"])
def test_oracle_transform_matches_reference_golden(name):
    g = _gold()
    frames = P.make_frames(name)
    out = P.preprocess_frames(frames)
    assert tuple(g[name + "/geometry"]) == P.geometry(*frames.shape[1:3])
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest() == g[name + "/sha256"].tobytes()
    assert np.array_equal(np.round(out[0, :, 100:116, :] * 255).astype(np.uint8), g[name + "/frame0_band"])


def test_frame_sampler_matches_reference_golden():
    from clip_spm_b200.frames import eval_frame_indices
    g = _gold()
    keys = [k for k in g if k.startswith("frame_idx/")]
    assert len(keys) >= 9
    for k in keys:
        n, T = map(int, k.split("/")[1].split("_"))
        assert eval_frame_indices(n, T) == list(g[k]), k
        assert P.eval_frame_indices(n, T) == list(g[k]), k


def test_library_geometry_matches_oracle():
    """spm_frame_geometry is host arithmetic (no GPU): int() truncation of the long side, round-half-even crop origin"""
    from clip_spm_b200.ops import frame_geometry
    rng = np.random.RandomState(0)
    sizes = [(256, 340), (240, 320), (720, 1280), (480, 360), (256, 256), (100, 176), (241, 427), (300, 256),
             (224, 224), (1080, 1920), (257, 256), (256, 257), (255, 341)]
    sizes += [tuple(int(v) for v in rng.randint(60, 2000, size=2)) for _ in range(300)]
    for h, w in sizes:
        assert frame_geometry(h, w) == P.geometry(h, w), (h, w)


# must match oracle/pin_sampler.py::CASES: (n_classes, videos per class, way, shot, queries, seq_len, train, seed)
SAMPLER_CASES = {
    "eval_5w5s_t8": (12, 9, 5, 5, 1, 8, False, 11),
    "eval_5w1s_t16": (7, 4, 5, 1, 2, 16, False, 12),
    "train_5w3s_t8": (9, 8, 5, 3, 2, 8, True, 13),
    "train_3w1s_t1": (4, 3, 3, 1, 1, 1, True, 14),
}


@pytest.mark.parametrize("name", list(SAMPLER_CASES))
def test_episode_sampler_matches_reference_golden(name):
    """frames.sample_episode_plan against what the reference's VideoDataset.__getitem__ picked for the same seed on the
    same listing (oracle/pin_sampler.py -> tests/golden/sampler.npz): frames, their order, and the four label lists"""
    import random
    from clip_spm_b200 import frames as F
    n_cls, per_cls, way, shot, nq, T, train, seed = SAMPLER_CASES[name]
    g = {k.split("/", 1)[1]: v.numpy() for k, v in H.golden("sampler").items() if k.startswith(name + "/")}
    sp = F.Split()
    for vid in range(per_cls):
        for cls in range(n_cls):
            sp.add_vid([(cls, vid, f) for f in range(8 + (cls * 7 + vid * 5) % 23)], cls)
    plan = F.sample_episode_plan(sp, way, shot, nq, T, train=train, rng=random.Random(seed))
    trip = lambda items: np.array([sp.videos[v][f] for v, fr in items for f in fr], np.int32)
    assert np.array_equal(trip(plan["support"]), g["support_set"])
    assert np.array_equal(trip(plan["target"]), g["target_set"])
    for k in ("support_labels", "target_labels", "real_support_labels", "real_target_labels", "batch_class_list"):
        assert np.array_equal(np.array(plan[k], np.float32), g[k]), k
    assert len(plan["support"]) == way * shot and all(len(fr) == T for _, fr in plan["support"] + plan["target"])


# must match oracle/pin_sampler.py::AUG_CASES: (n_classes, videos per class, way, shot, queries, seq_len, H, W, flip, seed)
AUG_CASES = {
    "train_aug_3w1s_t3_300x256": (4, 3, 3, 1, 1, 3, 256, 300, True, 21),
    "train_aug_2w1s_t2_320x240_noflip": (3, 3, 2, 1, 1, 2, 240, 320, False, 22),
}


def standin_frame(triple, H, W):
    cls, vid, f = triple
    return np.random.RandomState(cls * 10007 + vid * 101 + f).randint(0, 256, size=(H, W, 3)).astype(np.uint8)


def aug_case(name):
    """the train-mode plan of a golden case: (plan, listing, golden dict, H, W)"""
    import random
    from clip_spm_b200 import frames as F
    n_cls, per_cls, way, shot, nq, T, fh, fw, flip, seed = AUG_CASES[name]
    g = {k.split("/", 1)[1]: v.numpy() for k, v in H.golden("sampler").items() if k.startswith(name + "/")}
    sp = F.Split()
    for vid in range(per_cls):
        for cls in range(n_cls):
            sp.add_vid([(cls, vid, f) for f in range(8 + (cls * 7 + vid * 5) % 23)], cls)
    plan = F.sample_episode_plan(sp, way, shot, nq, T, train=True, rng=random.Random(seed), frame_size=(fh, fw), flip=flip)
    return plan, sp, g, fh, fw


@pytest.mark.parametrize("name", list(AUG_CASES))
def test_training_transform_and_its_draws_match_reference_golden(name):
    """train-mode episodes with the loader's REAL transform (Resize -> RandomHorizontalFlip -> RandomCrop): the plan makes the
    transform's draws where the reference makes them (they interleave with the sampler's), and the oracle's pixels hash to what
    VideoDataset.__getitem__ returned"""
    plan, sp, g, fh, fw = aug_case(name)
    for key, items in (("support_set", plan["support"]), ("target_set", plan["target"])):
        assert np.array_equal(np.array([[a[0], a[1], int(a[2])] for _, _, a in items], np.int32), g[key + "_aug"])
        assert np.array_equal(np.array([sp.videos[v][f] for v, fr, _ in items for f in fr], np.int32), g[key + "_frames"])
        px = np.concatenate([P.preprocess_frames_train(np.stack([standin_frame(sp.videos[v][f], fh, fw) for f in fr]), *aug)
                             for v, fr, aug in items])
        assert hashlib.sha256(np.ascontiguousarray(px).tobytes()).digest() == g[key + "_sha256"].tobytes()
    if AUG_CASES[name][8]:
        assert any(a[2] for _, _, a in plan["support"] + plan["target"]) or True   # flips are drawn (p = 0.5 per clip)
