"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's evaluation-time frame transform
(video_reader.py:83-111,265-272): Resize(256) -> CenterCrop(224) -> ToTensor, on decoded RGB uint8 frames.

    videotransforms/video_transforms.py:91-110   Resize(256): interpolation defaults to 'nearest', which
    videotransforms/functional.py:24-63          resize_clip maps to PIL.Image.BILINEAR (the branch is inverted there);
                                                 shorter side -> 256, other side int(256 * long / short) (:66-73);
                                                 a clip whose shorter side already is 256 is returned untouched (:48-50)
    videotransforms/video_transforms.py:204-247  CenterCrop(224): x1 = round((W-224)/2), y1 = round((H-224)/2)
    torchvision.transforms.ToTensor              uint8 HWC -> float32 CHW / 255

The resampling itself lives in a third-party dependency that is not vendored in the reference: Pillow (the reference
pins no version; this image has Pillow 12.2.0), `Image.resize(size, BILINEAR)` = `ImagingResample` of
src/libImaging/Resample.c.  Its published algorithm for 8-bit images, restated below: a separable triangle filter
whose support is widened by the down-scaling factor (antialiasing), coefficients normalised in double precision and
rounded to 22-bit fixed point, a horizontal pass into an 8-bit intermediate, then a vertical pass; each pass starts
the accumulator at 2^21 and clips `acc >> 22` to [0, 255].

Pinned: oracle/pin_preprocess.py runs the reference's own Compose([Resize(256), CenterCrop(224)]) + ToTensor (with the
installed Pillow) on seeded frames of several geometries, asserts bit equality with this restatement, and writes
tests/golden/preprocess_*.npz.  Only tests/, smoke() and bench.py's CPU-baseline legs may import this file."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
RESIZE, CROP = 256, 224


def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter over the
    full input range -> (xmin[out], count[out], k[out, ksize] int32)"""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            a = (x + lo - center + 0.5) * ss
            a = -a if a < 0.0 else a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v
        for x in range(n):
            if ww != 0.0:
                w[x] /= ww
            f = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + f) if w[x] < 0 else int(0.5 + f)
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _pass(img, xmin, cnt, kk, axis):
    """one resampling pass along `axis` of an uint8 [H, W, C] image"""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((len(xmin),) + img.shape[1:], np.uint8)
    for i in range(len(xmin)):
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt[i]):
            acc += img[xmin[i] + x] * int(kk[i, x])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_sizes(im_h, im_w, size=RESIZE):
    """functional.py:66-73 (-> oh, ow); None when the clip is returned untouched (:48-50)"""
    if (im_w <= im_h and im_w == size) or (im_h <= im_w and im_h == size):
        return None
    if im_w < im_h:
        return int(size * im_h / im_w), size
    return size, int(size * im_w / im_h)


def pil_resize_bilinear(img, out_h, out_w):
    """Image.resize((out_w, out_h), BILINEAR) of an uint8 [H, W, C] array: horizontal pass, then vertical pass;
    a pass whose size does not change is skipped (Resample.c need_horizontal / need_vertical)"""
    h, w = img.shape[:2]
    if out_w != w:
        img = _pass(img, *resample_coeffs(w, out_w), axis=1)
    if out_h != h:
        img = _pass(img, *resample_coeffs(h, out_h), axis=0)
    return img


def crop_offsets(im_h, im_w, size=CROP):
    """video_transforms.py:244-245 (Python round: half to even)"""
    if size > im_w or size > im_h:
        raise ValueError("Initial image size should be larger then cropped size")
    return int(round((im_h - size) / 2.)), int(round((im_w - size) / 2.))


def geometry(im_h, im_w):
    """-> (resized h, resized w, crop y1, crop x1) of the test transform"""
    rs = resize_sizes(im_h, im_w)
    oh, ow = (im_h, im_w) if rs is None else rs
    y1, x1 = crop_offsets(oh, ow)
    return oh, ow, y1, x1


def preprocess_frames(frames):
    """frames uint8 [F, H, W, 3] -> float32 [F, 3, 224, 224] in [0, 1]"""
    frames = np.asarray(frames)
    F, H, W, _ = frames.shape
    oh, ow, y1, x1 = geometry(H, W)
    out = np.empty((F, 3, CROP, CROP), np.float32)
    for f in range(F):
        img = frames[f]
        if (oh, ow) != (H, W):
            img = pil_resize_bilinear(img, oh, ow)
        img = img[y1:y1 + CROP, x1:x1 + CROP]
        out[f] = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    return out


def train_augmentation(rng, oh, ow, flip=True):
    """The draws the TRAINING transform makes per clip, in its order (video_reader.py:97-103): RandomHorizontalFlip
    (video_transforms.py:46: `random.random() < 0.5`; absent for ssv2) then RandomCrop (:152-153: x1 = randint(0, w - 224),
    y1 = randint(0, h - 224)) on the resized clip of oh x ow.  rng: a random.Random or the random module.  -> (y1, x1, flip)"""
    fl = bool(rng.random() < 0.5) if flip else False
    x1 = rng.randint(0, ow - CROP)
    y1 = rng.randint(0, oh - CROP)
    return y1, x1, fl


def preprocess_frames_train(frames, y1, x1, flip):
    """Resize(256) -> [mirror] -> crop at (y1, x1) -> ToTensor: frames uint8 [F, H, W, 3] of ONE clip -> float32 [F,3,224,224]"""
    frames = np.asarray(frames)
    F, H, W, _ = frames.shape
    oh, ow, _, _ = geometry(H, W)
    out = np.empty((F, 3, CROP, CROP), np.float32)
    for f in range(F):
        img = frames[f]
        if (oh, ow) != (H, W):
            img = pil_resize_bilinear(img, oh, ow)
        if flip:
            img = img[:, ::-1]
        img = img[y1:y1 + CROP, x1:x1 + CROP]
        out[f] = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    return out


def eval_frame_indices(n_frames, seq_len):
    """video_reader.py:231-260, evaluation branch: which of a video's n_frames are read"""
    if n_frames == seq_len:
        return list(range(n_frames))
    start, end = 1, n_frames - 2
    if end - start < seq_len:
        end, start = n_frames - 1, 0
    return [int(f) for f in np.linspace(start, end, num=seq_len)]


# ---- seeded test frames (shared by oracle/pin_preprocess.py and the tests; nothing large is committed) ----
# name: (H, W, n_frames, seed, kind)
CASES = {
    "k100_340x256": (256, 340, 2, 1, "noise"),        # Kinetics frames as extracted: shorter side already 256 -> crop only
    "up_320x240": (240, 320, 2, 2, "noise"),          # up-scaling (2 taps)
    "down_1280x720": (720, 1280, 1, 3, "noise"),      # down-scaling by 2.8 (antialiased, 7 taps)
    "portrait_360x480": (480, 360, 2, 4, "smooth"),   # width is the shorter side
    "square_256": (256, 256, 1, 5, "noise"),
    "tiny_176x100": (100, 176, 1, 6, "noise"),        # strong up-scaling
    "odd_427x241": (241, 427, 1, 7, "smooth"),        # odd sizes: int() truncation of the long side, round-half-even crop
    "w256_tall": (300, 256, 1, 8, "noise"),
}


def make_frames(name):
    H, W, n, seed, kind = CASES[name]
    rng = np.random.RandomState(seed)
    if kind == "noise":
        return rng.randint(0, 256, size=(n, H, W, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.empty((n, H, W, 3), np.uint8)
    for f in range(n):
        for c in range(3):
            ph = rng.uniform(0, 6.28, 2)
            v = 127.5 + 80 * np.sin(xx / (7.0 + 3 * c) + ph[0]) + 47 * np.cos(yy / (5.0 + f) + ph[1])
            out[f, :, :, c] = np.clip(v + rng.randint(-3, 4, size=(H, W)), 0, 255).astype(np.uint8)
    return out
