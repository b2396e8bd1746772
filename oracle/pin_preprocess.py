"""Pins oracle/preprocess_oracle.py against the REAL reference transform and writes tests/golden/preprocess.npz.

Runs only in the build container (needs /root/reference and Pillow):   python oracle/pin_preprocess.py
It feeds seeded uint8 frames through the reference's own `Compose([Resize(256), CenterCrop(224)])`
(videotransforms/video_transforms.py, exactly as video_reader.py:83-111 builds the test transform) followed by
torchvision's ToTensor (video_reader.py:65,271), asserts the numpy restatement is bit-identical, and stores per case
the geometry, a SHA-256 of the float32 result and (for small cases) the cropped uint8 image.  Frames are not stored:
they regenerate from the seed (make_frames)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import preprocess_oracle as P  # noqa: E402

CASES, make_frames = P.CASES, P.make_frames


def reference_transform(frames):
    import types
    sys.path.insert(0, REF)
    for mod in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.transform"):   # imported there, unused on this path
        try:
            __import__(mod)
        except ImportError:
            sys.modules.setdefault(mod, types.ModuleType(mod))
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from PIL import Image
    from torchvision import transforms
    from videotransforms.video_transforms import CenterCrop, Compose, Resize
    tf = Compose([Resize(256), CenterCrop(224)])
    to_tensor = transforms.ToTensor()
    imgs = tf([Image.fromarray(f) for f in frames])
    return np.stack([to_tensor(v).numpy() for v in imgs])


if __name__ == "__main__":
    import PIL
    gold = {}
    for name in CASES:
        frames = make_frames(name)
        ref = reference_transform(frames)
        mine = P.preprocess_frames(frames)
        assert ref.dtype == np.float32 and ref.shape == mine.shape
        nbad = int((ref != mine).sum())
        assert nbad == 0, "%s: %d of %d values differ from the reference transform" % (name, nbad, ref.size)
        H, W = frames.shape[1:3]
        gold[name + "/geometry"] = np.array(P.geometry(H, W), np.int32)
        gold[name + "/sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(ref).tobytes()).digest(), np.uint8)
        gold[name + "/frame0_band"] = np.round(ref[0, :, 100:116, :] * 255).astype(np.uint8)   # 16 rows, for debugging
        print("%-18s %4dx%-4d -> resized %s, crop at %s: oracle == reference transform bit for bit (Pillow %s)"
              % (name, W, H, tuple(gold[name + "/geometry"][:2]), tuple(gold[name + "/geometry"][2:]), PIL.__version__))
    # evaluation-time frame selection: the reference's own VideoDataset.get_seq (video_reader.py:231-272) run on a
    # stand-in dataset object whose "images" are just the frame numbers
    import types
    import video_reader

    class FakeDB:
        def __init__(self, n):
            self.n = n

        def get_rand_vid(self, label, idx):
            return list(range(self.n)), 0

    for n, T in [(8, 8), (9, 8), (10, 8), (30, 8), (17, 16), (300, 8), (5, 8), (18, 16), (20, 16)]:
        fake = types.SimpleNamespace(train=False, seq_len=T, transform=None, get_train_or_test_db=lambda n=n: FakeDB(n),
                                     read_single_image=lambda p: p)
        ref_idx, _ = video_reader.VideoDataset.get_seq(fake, 0, 0)
        assert list(ref_idx) == P.eval_frame_indices(n, T), (n, T, ref_idx, P.eval_frame_indices(n, T))
        gold["frame_idx/%d_%d" % (n, T)] = np.array(ref_idx, np.int32)
    print("eval_frame_indices == VideoDataset.get_seq on %d (n_frames, seq_len) pairs" % 9)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **gold)
