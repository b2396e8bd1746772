"""Pins clip_spm_b200.frames.sample_episode_plan / train_frame_indices against the reference's own episodic sampler,
`VideoDataset.__getitem__` + `get_seq` (video_reader.py:231-329), executed here on a stand-in dataset whose "frames"
are (class, video, frame) triples -- so the reference's stacked support / target tensors spell out exactly which
frames it picked and in which order.  Writes tests/golden/sampler.npz (the reference's picks for seeded cases).
Needs /root/reference; nothing here is imported by the product path."""
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from clip_spm_b200 import frames as F  # noqa: E402

# name: (n_classes, videos per class, way, shot, queries, seq_len, train, seed)
CASES = {
    "eval_5w5s_t8": (12, 9, 5, 5, 1, 8, False, 11),
    "eval_5w1s_t16": (7, 4, 5, 1, 2, 16, False, 12),
    "train_5w3s_t8": (9, 8, 5, 3, 2, 8, True, 13),
    "train_3w1s_t1": (4, 3, 3, 1, 1, 1, True, 14),
}


def n_frames_of(cls, vid):
    return 8 + (cls * 7 + vid * 5) % 23      # 8 .. 30 frames: covers n == seq_len, short and long videos


def build_listing(n_cls, per_cls, make_split):
    sp = make_split()
    for vid in range(per_cls):               # interleave the classes like a real annotation file would
        for cls in range(n_cls):
            sp.add_vid([(cls, vid, f) for f in range(n_frames_of(cls, vid))], cls)
    return sp


# train-mode episodes through the reference's REAL training transform (Resize -> RandomHorizontalFlip -> RandomCrop, whose draws
# interleave with the sampler's own): name: (n_classes, videos per class, way, shot, queries, seq_len, frame H, frame W, flip, seed)
AUG_CASES = {
    "train_aug_3w1s_t3_300x256": (4, 3, 3, 1, 1, 3, 256, 300, True, 21),     # shorter side already 256: crop / flip only
    "train_aug_2w1s_t2_320x240_noflip": (3, 3, 2, 1, 1, 2, 240, 320, False, 22),   # real resize (-> 341 x 256); ssv2: no flip
}


def standin_frame(triple, H, W):
    cls, vid, f = triple
    return np.random.RandomState(cls * 10007 + vid * 101 + f).randint(0, 256, size=(H, W, 3)).astype(np.uint8)


def run_aug_cases(video_reader, gold):
    import hashlib
    from PIL import Image
    from torchvision import transforms
    from videotransforms.video_transforms import Compose, RandomCrop, RandomHorizontalFlip, Resize
    from oracle import preprocess_oracle as P
    for name, (n_cls, per_cls, way, shot, nq, T, H, W, flip, seed) in AUG_CASES.items():
        ref_split = build_listing(n_cls, per_cls, video_reader.Split)
        tf = Compose([Resize(256)] + ([RandomHorizontalFlip()] if flip else []) + [RandomCrop(224)])   # video_reader.py:83-103
        fake = types.SimpleNamespace(
            train=True, seq_len=T, way=way, shot=shot, query_per_class=nq, query_per_class_test=nq,
            transform={"train": tf, "test": None}, tensor_transform=transforms.ToTensor(),
            get_train_or_test_db=lambda s=ref_split: s,
            read_single_image=lambda p, H=H, W=W: Image.fromarray(standin_frame(p, H, W)))
        fake.get_seq = lambda label, idx=-1, f=fake: video_reader.VideoDataset.get_seq(f, label, idx)
        random.seed(seed)
        ref = video_reader.VideoDataset.__getitem__(fake, 0)
        mine_split = build_listing(n_cls, per_cls, F.Split)
        plan = F.sample_episode_plan(mine_split, way, shot, nq, T, train=True, rng=random.Random(seed), frame_size=(H, W),
                                     flip=flip)

        def images(items):
            out = []
            for v, fr, aug in items:
                frames = np.stack([standin_frame(mine_split.videos[v][f], H, W) for f in fr])
                out.append(P.preprocess_frames_train(frames, *aug))
            return np.concatenate(out)
        for key, items in (("support_set", plan["support"]), ("target_set", plan["target"])):
            mine = images(items)
            r = ref[key].numpy()
            assert mine.shape == r.shape and (mine != r).sum() == 0, (name, key, int((mine != r).sum()))
            gold[name + "/" + key + "_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(r).tobytes()).digest(), np.uint8)
            gold[name + "/" + key + "_aug"] = np.array([[a[0], a[1], int(a[2])] for _, _, a in items], np.int32)
            gold[name + "/" + key + "_frames"] = np.array([mine_split.videos[v][f] for v, fr, _ in items for f in fr], np.int32)
        for k in ("support_labels", "target_labels", "real_support_labels", "real_target_labels"):
            assert torch.equal(torch.FloatTensor(plan[k]), ref[k]), (name, k)
        print("%-36s plan + training transform == VideoDataset.__getitem__ pixels, bit for bit (%d clips, draws %s)"
              % (name, len(plan["support"]) + len(plan["target"]),
                 [tuple(int(x) for x in a) for a in gold[name + "/support_set_aug"][:2]]))


def main():
    for mod in ("matplotlib", "matplotlib.pyplot"):     # imported by videotransforms, unused on this path
        sys.modules.setdefault(mod, types.ModuleType(mod))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import video_reader
    gold = {}
    for name, (n_cls, per_cls, way, shot, nq, T, train, seed) in CASES.items():
        ref_split = build_listing(n_cls, per_cls, video_reader.Split)
        fake = types.SimpleNamespace(
            train=train, seq_len=T, way=way, shot=shot, query_per_class=nq, query_per_class_test=nq,
            transform={"train": lambda imgs: imgs, "test": lambda imgs: imgs},
            tensor_transform=lambda v: torch.tensor(v), get_train_or_test_db=lambda s=ref_split: s,
            read_single_image=lambda p: p)
        fake.get_seq = lambda label, idx=-1, f=fake: video_reader.VideoDataset.get_seq(f, label, idx)
        random.seed(seed)
        ref = video_reader.VideoDataset.__getitem__(fake, 0)
        mine_split = build_listing(n_cls, per_cls, F.Split)
        plan = F.sample_episode_plan(mine_split, way, shot, nq, T, train=train, rng=random.Random(seed))

        def triples(items):
            return torch.tensor([mine_split.videos[v][f] for v, fr in items for f in fr])
        assert torch.equal(triples(plan["support"]), ref["support_set"]), name
        assert torch.equal(triples(plan["target"]), ref["target_set"]), name
        for k_mine, k_ref in (("support_labels", "support_labels"), ("target_labels", "target_labels"),
                              ("real_support_labels", "real_support_labels"), ("real_target_labels", "real_target_labels"),
                              ("batch_class_list", "batch_class_list")):
            assert torch.equal(torch.FloatTensor(plan[k_mine]), ref[k_ref]), (name, k_mine)
        gold[name + "/support_set"] = ref["support_set"].numpy().astype(np.int32)
        gold[name + "/target_set"] = ref["target_set"].numpy().astype(np.int32)
        for k in ("support_labels", "target_labels", "real_support_labels", "real_target_labels", "batch_class_list"):
            gold[name + "/" + k] = ref[k].numpy()
        print("%-16s sample_episode_plan == VideoDataset.__getitem__ (%d support + %d target videos, %s frame sampling)"
              % (name, len(plan["support"]), len(plan["target"]), "train" if train else "eval"))
    run_aug_cases(video_reader, gold)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler.npz"), **gold)


if __name__ == "__main__":
    main()
