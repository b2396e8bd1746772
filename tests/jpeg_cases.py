"""Synthetic JPEG files for the decode tests: encoded with PIL (the encoder is irrelevant to the decoder under test), in
the flavours frame dumps come in -- 4:2:0 / 4:2:2 / 4:4:4, qualities 30..100, optimised Huffman tables, restart
intervals, sizes that are not multiples of the MCU.  The expected pixels are PIL's own decode of the same bytes
(video_reader.py:227-230), computed at test time on whichever box runs the test."""
import io

import numpy as np
from PIL import Image


def _image(h, w, kind, rng):
    if kind == "noise":
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == "smooth":
        y, x = np.mgrid[0:h, 0:w]
        a = np.stack([(np.sin(x / 9.0) + np.cos(y / 7.0)) * 60 + 128, x * 255.0 / w, y * 255.0 / h], -1)
        a = a.clip(0, 255).astype(np.uint8)
    else:   # saturated 8x8 blocks: the IDCT overshoots 0 / 255, exercising the range limiting
        a = (rng.integers(0, 2, (h // 8 + 1, w // 8 + 1, 3)) * 255).astype(np.uint8).repeat(8, 0).repeat(8, 1)[:h, :w]
    return Image.fromarray(a)


def encode(h, w, kind, quality, subsampling, seed=0, **extra):
    buf = io.BytesIO()
    _image(h, w, kind, np.random.default_rng(seed)).save(buf, "JPEG", quality=quality, subsampling=subsampling, **extra)
    return buf.getvalue()


def pil_decode(data):
    with Image.open(io.BytesIO(data)) as im:
        im.load()
        assert im.mode == "RGB"
        return np.asarray(im).copy()


def cases():
    """(name, bytes)"""
    out = []
    for (h, w) in [(256, 340), (240, 320), (17, 23), (100, 99), (64, 64), (225, 401)]:
        for kind in ("noise", "smooth", "blocks"):
            for q, sub, extra in ((75, 2, {}), (95, 1, {}), (30, 0, {}), (100, 2, {"optimize": True}),
                                  (85, 2, {"restart_marker_blocks": 5}), (90, 0, {"restart_marker_rows": 1})):
                try:
                    out.append(("%dx%d_%s_q%d_s%d%s" % (h, w, kind, q, sub, "_" + "_".join(extra) if extra else ""),
                                encode(h, w, kind, q, sub, **extra)))
                except (TypeError, OSError):   # an option this Pillow lacks, or its encoder buffer is too small
                    pass
    return out
