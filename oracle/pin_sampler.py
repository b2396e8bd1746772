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
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler.npz"), **gold)


if __name__ == "__main__":
    main()
