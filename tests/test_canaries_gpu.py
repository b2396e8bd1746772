"""Out-of-bounds write checks without a sanitizer: outputs are carved out of a larger NaN-filled buffer and the guard
zones on both sides must still be NaN after the call (and every output element must have been written)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import clipspm_oracle as O
from oracle import preprocess_oracle as P

pytestmark = pytest.mark.gpu
PAD = 4096


def _guarded(n, dtype=torch.float32):
    buf = torch.full((n + 2 * PAD,), float("nan"), device="cuda", dtype=dtype)
    return buf, buf[PAD:PAD + n]


def _check(buf, n):
    assert torch.isnan(buf[:PAD]).all() and torch.isnan(buf[PAD + n:]).all(), "write outside the output buffer"
    assert not torch.isnan(buf[PAD:PAD + n]).any(), "output elements left unwritten"


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("name", ["up_320x240", "down_1280x720", "odd_427x241"])
def test_transform_frames_stays_in_bounds(name):
    from clip_spm_b200 import _lib
    lib = _lib.load()
    frames = torch.from_numpy(P.make_frames(name)).cuda()
    F, H, W, _ = frames.shape
    n = F * 3 * 224 * 224
    buf, out = _guarded(n)
    _lib.check(lib.spm_transform_frames(_st(), ctypes.c_void_p(frames.data_ptr()), F, H, W, ctypes.c_void_p(out.data_ptr())))
    torch.cuda.synchronize()
    _check(buf, n)


@pytest.mark.parametrize("W,Q,T,D", [(5, 5, 8, 512), (3, 2, 16, 1024), (2, 3, 5, 512)])
def test_otam_forward_and_backward_stay_in_bounds(W, Q, T, D):
    from clip_spm_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(W * 100 + T)
    sup = torch.randn(2, W, T, D, generator=g).cuda()
    tgt = torch.randn(2, Q, T, D, generator=g).cuda()
    go = torch.randn(2, Q, W, generator=g).cuda()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    bo, out = _guarded(2 * Q * W)
    _lib.check(lib.spm_otam_distance(_st(), 2, W, Q, T, D, p(sup), p(tgt), 0, 1.0, 0.0, p(out)))
    bs, gs = _guarded(sup.numel())
    bt, gt = _guarded(tgt.numel())
    _lib.check(lib.spm_otam_distance_backward(_st(), 2, W, Q, T, D, p(sup), p(tgt), 0, 1.0, p(go), p(gs), p(gt)))
    torch.cuda.synchronize()
    _check(bo, 2 * Q * W); _check(bs, sup.numel()); _check(bt, tgt.numel())


def test_text_encode_stays_in_bounds():
    from clip_spm_b200 import TextTower
    from tests import helpers as H
    g = H.golden("text_tower_4cls")
    tok = g["tokens"].reshape(-1, 77)[:5].int().cuda().contiguous()
    tw = TextTower(O.make_text_weights(512, seed=0))
    from clip_spm_b200 import _lib
    buf, out = _guarded(5 * 512)
    _lib.check(_lib.load().spm_text_encode(tw._h, _st(), ctypes.c_void_p(tok.data_ptr()), 5, ctypes.c_void_p(out.data_ptr())))
    torch.cuda.synchronize()
    _check(buf, 5 * 512)


@pytest.mark.parametrize("M,N,K", [(19999, 768, 768), (1000, 96, 64), (130, 32, 32)])
def test_gemm_outputs_stay_in_bounds(M, N, K):
    """ragged M (TMA zero-fill on loads, masked stores), 2-CTA residual ring and 1-CTA paths"""
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(M)
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    b = torch.randn(N, K, generator=g).cuda().bfloat16()
    res = torch.randn(M, N, generator=g).cuda()
    buf, flat = _guarded(M * N)
    out = flat.view(M, N)
    ops.gemm(a, b, residual=res, out=out)
    torch.cuda.synchronize()
    _check(buf, M * N)
    buf2, flat2 = _guarded(M * N, torch.bfloat16)
    ops.gemm(a, b, act="relu", out=flat2.view(M, N))
    torch.cuda.synchronize()
    _check(buf2, M * N)
