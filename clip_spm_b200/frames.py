"""Host side of the episode input path: which frames of a video the evaluation sampler reads,
and the GPU frame transform that replaces the PIL Resize/CenterCrop/ToTensor chain of the data loader."""
import numpy as np

from .ops import frame_geometry, transform_frames  # noqa: F401  (re-exported)


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
