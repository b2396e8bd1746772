"""JPEG decode, CPU suite: the arithmetic the CUDA kernels run (csrc/jpeg_core.cuh + jpeg_parse.h, compiled for the host
as oracle/_ref/libjpeg_check.so) against PIL -- the decoder the reference calls (video_reader.py:227-230) -- bit for bit."""
import ctypes
import os

import numpy as np
import pytest

from tests import jpeg_cases as J

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    path = os.path.join(ROOT, "oracle", "_ref", "libjpeg_check.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return ctypes.CDLL(path)


def _decode(lib, data):
    H, W, hs, vs = (ctypes.c_int() for _ in range(4))
    err = ctypes.create_string_buffer(256)
    if lib.jpeg_check_info(data, ctypes.c_longlong(len(data)), ctypes.byref(H), ctypes.byref(W), ctypes.byref(hs),
                           ctypes.byref(vs), err, 256):
        raise RuntimeError(err.value.decode())
    out = np.zeros((H.value, W.value, 3), np.uint8)
    if lib.jpeg_check_decode(data, ctypes.c_longlong(len(data)), out.ctypes.data_as(ctypes.c_void_p), err, 256):
        raise RuntimeError(err.value.decode())
    return out


def test_jpeg_arithmetic_is_bit_exact_with_pil(lib):
    cs = J.cases()
    assert len(cs) >= 60
    for name, data in cs:
        assert np.array_equal(_decode(lib, data), J.pil_decode(data)), name


def test_unsupported_files_are_rejected_with_a_reason(lib):
    from PIL import Image
    import io
    buf = io.BytesIO()
    Image.fromarray(np.zeros((32, 32, 3), np.uint8)).save(buf, "JPEG", progressive=True)
    with pytest.raises(RuntimeError, match="progressive"):
        _decode(lib, buf.getvalue())
    buf = io.BytesIO()
    Image.fromarray(np.zeros((32, 32), np.uint8)).save(buf, "JPEG")
    with pytest.raises(RuntimeError, match="3-component"):
        _decode(lib, buf.getvalue())
    with pytest.raises(RuntimeError, match="SOI"):
        _decode(lib, b"not a jpeg at all")
    good = J.encode(64, 64, "smooth", 80, 2)
    with pytest.raises(RuntimeError):
        _decode(lib, good[:200])      # truncated inside the headers
