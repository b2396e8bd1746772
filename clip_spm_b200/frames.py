"""Host side of the episode input path: which frames of a video the evaluation sampler reads,
and the GPU frame transform that replaces the PIL Resize/CenterCrop/ToTensor chain of the data loader."""
import numpy as np

from .ops import frame_geometry, transform_frames, transform_frames_train  # noqa: F401  (re-exported)


def eval_frame_indices(n_frames, seq_len):
    """video_reader.py:231-260 with self.train False: the seq_len frame numbers read from a video of n_frames
    (uniform over [1, n-2]; over [0, n-1] for short videos; all of them when n_frames == seq_len)."""
    n_frames, seq_len = int(n_frames), int(seq_len)
    if n_frames == seq_len:
        return list(range(n_frames))
    start, end = 1, n_frames - 2
    if end - start < seq_len:
        start, end = 0, n_frames - 1
    return [int(f) for f in np.linspace(start, end, num=seq_len)]


def train_frame_indices(n_frames, seq_len, rng):
    """video_reader.py:236-262 with self.train True: like eval_frame_indices but the first / last frame are jittered by
    up to min(5, excess/2) frames -- consumes `rng.randint` exactly as the reference does (two draws, or none)."""
    n_frames, seq_len = int(n_frames), int(seq_len)
    if n_frames == seq_len:
        return list(range(n_frames))
    excess_pad = int(min(5, (n_frames - seq_len) / 2))
    if excess_pad < 1:
        start, end = 0, n_frames - 1
    else:
        start = rng.randint(0, excess_pad)
        end = rng.randint(n_frames - 1 - excess_pad, n_frames - 1)
    if end - start < seq_len:
        start, end = 0, n_frames - 1
    idxs = [int(f) for f in np.linspace(start, end, num=seq_len)]
    if seq_len == 1:
        idxs = [rng.randint(start, end - 1)]
    return idxs


def train_augmentation(rng, frame_h, frame_w, flip=True):
    """The random draws of the loader's TRAINING transform for one clip, in the reference's order (video_reader.py:97-103 ->
    videotransforms/video_transforms.py:46 RandomHorizontalFlip, absent for ssv2; :152-153 RandomCrop on the resized clip):
    -> (crop y1, crop x1, flip), the `aug` row of ops.transform_frames_train."""
    oh, ow, _, _ = frame_geometry(frame_h, frame_w)
    fl = bool(rng.random() < 0.5) if flip else False
    x1 = rng.randint(0, ow - 224)
    y1 = rng.randint(0, oh - 224)
    return y1, x1, fl


class Split:
    """The listing the sampler draws from (video_reader.py:14-50 `Split`): videos[i] = list of frame paths (or any
    per-frame handles), gt_a_list[i] = its class id."""

    def __init__(self):
        self.videos, self.gt_a_list = [], []

    def add_vid(self, paths, gt_a):
        self.videos.append(paths)
        self.gt_a_list.append(gt_a)

    def videos_of(self, label):
        return [i for i, g in enumerate(self.gt_a_list) if g == label]

    def get_unique_classes(self):
        return list(set(self.gt_a_list))


def sample_episode_plan(split, way, shot, n_queries, seq_len, train=False, rng=None, frame_size=None, flip=True):
    """The episode a `VideoDataset.__getitem__` call builds (video_reader.py:275-329), as a PLAN: which frames of which
    videos form the support / target sets, in the order the reference stacks them, plus the four label lists --
    everything but the pixel work (decode + Resize/CenterCrop/ToTensor), which `CNN.evaluate_host_u8` does on the GPU.
    `rng` is a `random.Random` (default: the global `random` module, the reference's own source); the draws are made in
    the reference's order (classes; per class the videos; per video the frame jitter when training; the two shuffles),
    so the same seed yields the same episode.
    Training with the loader's real transform: pass frame_size = (H, W) of the decoded frames (or a callable video_index ->
    (H, W)); the per-clip draws of RandomHorizontalFlip / RandomCrop (flip=False for ssv2, video_reader.py:95-100) are then
    made where the reference makes them -- right after the clip's frame jitter -- and every entry becomes
    (video_index, [frame indices], (crop y1, crop x1, flip)).
    Returns dict(support=[(video_index, [frame indices])...], target=[...], support_labels, target_labels,
    real_support_labels, real_target_labels, batch_class_list)."""
    import random as _random
    rng = rng or _random
    classes = split.get_unique_classes()
    batch_classes = rng.sample(classes, way)
    support, target = [], []
    for bl, bc in enumerate(batch_classes):
        vids = split.videos_of(bc)
        idxs = rng.sample([i for i in range(len(vids))], shot + n_queries)
        for k, idx in enumerate(idxs):
            v = vids[idx]
            n = len(split.videos[v])
            fr = train_frame_indices(n, seq_len, rng) if train else eval_frame_indices(n, seq_len)
            entry = (v, fr)
            if train and frame_size is not None:
                fh, fw = frame_size(v) if callable(frame_size) else frame_size
                entry = (v, fr, train_augmentation(rng, fh, fw, flip))
            (support if k < shot else target).append((entry, bl, bc))
    rng.shuffle(support)
    rng.shuffle(target)
    return dict(support=[s[0] for s in support], target=[t[0] for t in target],
                support_labels=[float(s[1]) for s in support], target_labels=[float(t[1]) for t in target],
                real_support_labels=[float(s[2]) for s in support], real_target_labels=[float(t[2]) for t in target],
                batch_class_list=[float(c) for c in batch_classes])
