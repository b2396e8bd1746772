"""Per-stage entry points of the C-ABI library on torch CUDA tensors (torch is used for device memory and
streams only).  These are what the `-m gpu` parity tests call."""
import ctypes

import torch

from . import _lib

ACT = {"none": 0, "quickgelu": 1, "gelu": 2, "leaky_relu": 3, "sigmoid": 4, "relu": 5}


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("clip_spm_b200 ops take CUDA tensors (no CPU fallback exists)")


def gemm(a, b, bias=None, act="none", slope=0.0, residual=None, res_row_mod=0, res_row_off=0, out_row_group=0,
         out_group_stride=0, out_row_off=0, out=None, out_dtype=torch.float32):
    """out[orow(m), :] = act(a @ b.T + bias) (+ residual[rrow(m), :]).  a [M,K], b [N,K]: both bf16 (tensor-core
    bf16 path) or both fp32 (tf32 path), last dim contiguous."""
    lib = _lib.load()
    _need_cuda(a, b, bias, residual, out)
    assert a.dtype == b.dtype and a.dtype in (torch.bfloat16, torch.float32)
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        assert out_row_group == 0, "row-mapped output needs an explicit `out`"
        rows = M
        out = torch.empty(rows, N, device=a.device, dtype=out_dtype)
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
    kind = 0 if a.dtype == torch.bfloat16 else 1
    _lib.check(lib.spm_gemm(_stream(), kind, _ptr(a), a.stride(0), _ptr(b), b.stride(0), M, N, K, _ptr(bias),
                            ACT[act], float(slope), _ptr(residual),
                            0 if residual is None else residual.stride(0), res_row_mod, res_row_off, out_row_group,
                            out_group_stride, out_row_off, _ptr(out), out.stride(0), 1 if out.dtype == torch.bfloat16 else 0))
    return out


class _OtamDistance(torch.autograd.Function):
    """otam_distance with gradients: forward = spm_otam_distance, backward = spm_otam_distance_backward."""

    @staticmethod
    def forward(ctx, support, target, single_direct, alpha):
        lib = _lib.load()
        P, W, T, D = support.shape
        Q = target.shape[1]
        out = torch.zeros(P, Q, W, device=support.device)
        _lib.check(lib.spm_otam_distance(_stream(), P, W, Q, T, D, _ptr(support), _ptr(target), int(single_direct),
                                         float(alpha), 0.0, _ptr(out)))
        ctx.save_for_backward(support, target)
        ctx.cfg = (bool(single_direct), float(alpha))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        support, target = ctx.saved_tensors
        single_direct, alpha = ctx.cfg
        P, W, T, D = support.shape
        Q = target.shape[1]
        go = grad_out.contiguous().float()
        gs, gt = torch.empty_like(support), torch.empty_like(target)
        _lib.check(lib.spm_otam_distance_backward(_stream(), P, W, Q, T, D, _ptr(support), _ptr(target),
                                                  int(single_direct), alpha, _ptr(go), _ptr(gs), _ptr(gt)))
        return gs, gt, None, None


def otam_distance(support, target, single_direct=False, alpha=1.0, beta=0.0, out=None):
    """cos_sim + (bi)directional OTAM (models/model_clipspm.py:348-362): support [P,W,T,D], target [P,Q,T,D] fp32
    -> [P,Q,W].  Differentiable: when an input requires grad the call goes through an autograd Function whose
    backward is the library's own kernel (the accumulate form `out = beta*out + ...` is inference-only)."""
    lib = _lib.load()
    _need_cuda(support, target, out)
    support, target = support.contiguous().float(), target.contiguous().float()
    if torch.is_grad_enabled() and (support.requires_grad or target.requires_grad):
        if out is not None or beta != 0.0:
            raise RuntimeError("otam_distance: `out` / `beta` cannot be combined with autograd")
        return _OtamDistance.apply(support, target, bool(single_direct), float(alpha))
    P, W, T, D = support.shape
    Q = target.shape[1]
    if out is None:
        out = torch.zeros(P, Q, W, device=support.device)
    _lib.check(lib.spm_otam_distance(_stream(), P, W, Q, T, D, _ptr(support), _ptr(target), int(single_direct),
                                     float(alpha), float(beta), _ptr(out)))
    return out


def vit_attention(qkv, n_frames, impl="tcgen05"):
    """softmax(q k^T / 8) v per (frame, head) (models/clip_fsar.py:626,638): qkv [F*197, 2304] bf16 -> [F*197, 768]."""
    lib = _lib.load()
    _need_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (n_frames * 197, 2304)
    out = torch.empty(n_frames * 197, 768, device=qkv.device, dtype=torch.bfloat16)
    _lib.check(lib.spm_vit_attention(_stream(), _ptr(qkv), _ptr(out), n_frames, 1 if impl == "mma" else 0))
    return out


def frame_geometry(H, W):
    """(resized_h, resized_w, crop_y, crop_x) the evaluation transform uses for an H x W frame (host arithmetic)."""
    lib = _lib.load()
    v = [ctypes.c_int() for _ in range(4)]
    _lib.check(lib.spm_frame_geometry(int(H), int(W), *[ctypes.byref(x) for x in v]))
    return tuple(x.value for x in v)


def transform_frames(frames):
    """Resize(256) -> CenterCrop(224) -> ToTensor of the reference's test pipeline (video_reader.py:83-111,265-272),
    bit-exact: frames uint8 [F, H, W, 3] (decoded RGB) -> fp32 [F, 3, 224, 224] in [0, 1]."""
    lib = _lib.load()
    _need_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
    frames = frames.contiguous()
    F, H, W, _ = frames.shape
    out = torch.empty(F, 3, 224, 224, device=frames.device)
    _lib.check(lib.spm_transform_frames(_stream(), _ptr(frames), F, H, W, _ptr(out)))
    return out


def transform_frames_train(frames, aug):
    """The loader's TRAINING transform (video_reader.py:83-103: Resize(256) -> RandomHorizontalFlip -> RandomCrop(224) ->
    ToTensor), bit-exact given the draws: frames uint8 [F, H, W, 3], aug = per frame (crop y1, crop x1, flip) -- an
    [F, 3] int tensor / list, or ONE triple for the whole clip (frames.train_augmentation draws it the reference's way)."""
    lib = _lib.load()
    _need_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
    frames = frames.contiguous()
    F, H, W, _ = frames.shape
    a = torch.as_tensor(aug, dtype=torch.int32).reshape(-1, 3)
    if a.shape[0] == 1:
        a = a.expand(F, 3)
    oh, ow, _, _ = frame_geometry(H, W)
    if a.shape[0] != F or int(a[:, 0].min()) < 0 or int(a[:, 1].min()) < 0 or int(a[:, 0].max()) > oh - 224 or \
            int(a[:, 1].max()) > ow - 224:
        raise RuntimeError("transform_frames_train: one (y1, x1, flip) per frame with 0 <= y1 <= %d, 0 <= x1 <= %d" % (oh - 224, ow - 224))
    a = a.contiguous().to(frames.device)
    out = torch.empty(F, 3, 224, 224, device=frames.device)
    _lib.check(lib.spm_transform_frames_train(_stream(), _ptr(frames), F, H, W, _ptr(a), _ptr(out)))
    return out


def jpeg_info(data):
    """Header of one JPEG file (bytes): (height, width, luma h sampling, luma v sampling)."""
    lib = _lib.load()
    buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
    v = [ctypes.c_int() for _ in range(4)]
    _lib.check(lib.spm_jpeg_info(ctypes.cast(buf, ctypes.c_void_p), len(data), *[ctypes.byref(x) for x in v]))
    return tuple(x.value for x in v)


def decode_jpegs(files, device="cuda"):
    """What the reference's data loader does with PIL for every frame (video_reader.py:227-230 `Image.open(p).load()`),
    on the GPU and bit-identical to it: `files` = list of JPEG file contents (bytes), all of one size and chroma
    subsampling -> uint8 CUDA tensor [n, H, W, 3] (RGB), ready for transform_frames / CNN.encode_frames_u8."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("clip_spm_b200.ops.decode_jpegs needs a CUDA device; there is no CPU fallback")
    n = len(files)
    if n == 0:
        return torch.empty(0, 0, 0, 3, dtype=torch.uint8, device=device)
    H, W, _, _ = jpeg_info(files[0])
    bufs = [(ctypes.c_ubyte * len(f)).from_buffer_copy(f) for f in files]
    ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
    sizes = (ctypes.c_int64 * n)(*[len(f) for f in files])
    out = torch.empty(n, H, W, 3, dtype=torch.uint8, device=device)
    with torch.cuda.device(out.device):
        _lib.check(lib.spm_jpeg_decode(_stream(), n, ptrs, sizes, H, W, _ptr(out)))
    return out


class _SoftDTWFunction(torch.autograd.Function):
    """models/OTAM.py:134-203 `_SoftDTWCUDA`: D [B,N,M] -> R[:, N, M]; backward = E * grad (spm_softdtw_backward)."""

    @staticmethod
    def forward(ctx, D, gamma, bandwidth):
        lib = _lib.load()
        _need_cuda(D)
        D = D.contiguous().float()
        B, N, M = D.shape
        R = torch.empty(B, N + 2, M + 2, device=D.device)
        out = torch.empty(B, device=D.device)
        _lib.check(lib.spm_softdtw_forward(_stream(), B, N, M, _ptr(D), float(gamma), float(bandwidth), _ptr(R), _ptr(out)))
        ctx.save_for_backward(D, R)
        ctx.cfg = (float(gamma), float(bandwidth))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        lib = _lib.load()
        D, R = ctx.saved_tensors
        gamma, bandwidth = ctx.cfg
        B, N, M = D.shape
        E = torch.empty_like(D)
        _lib.check(lib.spm_softdtw_backward(_stream(), B, N, M, _ptr(D), _ptr(R), gamma, bandwidth, _ptr(E)))
        return grad_output.reshape(-1, 1, 1) * E, None, None


def softdtw(D, gamma=1.0, bandwidth=0.0):
    """soft-DTW value of every [N, M] distance matrix of D [B,N,M] (differentiable in D)."""
    return _SoftDTWFunction.apply(D, gamma, bandwidth)


class SoftDTW(torch.nn.Module):
    """Mirror of models/OTAM.py:318-424 `SoftDTW` (TA2N's metric, models/model_ta2n.py:87) with the DP on this
    library's kernels.  Same constructor (`use_cuda` is accepted for signature compatibility: there is no CPU path
    here), same forward contract: X [B,n,d], Y [B,m,d] -> [B,1] (bi-directional mean, :414-424), or the normalised
    variant (:405-413).  The point-wise distance is the caller's `dist_func` (default: the reference's
    1 - cosine_similarity, :381-388, plain torch ops as there)."""

    def __init__(self, use_cuda=True, gamma=1.0, normalize=False, bandwidth=None, dist_func=None):
        super().__init__()
        self.normalize = normalize
        self.gamma = gamma
        self.bandwidth = 0 if bandwidth is None else float(bandwidth)
        self.use_cuda = use_cuda
        self.dist_func = dist_func if dist_func is not None else SoftDTW._similarity_dist_func

    @staticmethod
    def _euclidean_dist_func(x, y):
        return torch.pow(x.unsqueeze(2) - y.unsqueeze(1), 2).sum(3)

    @staticmethod
    def _similarity_dist_func(x, y):
        n, m, d = x.size(1), y.size(1), x.size(2)
        return 1 - torch.cosine_similarity(x.unsqueeze(2).expand(-1, n, m, d), y.unsqueeze(1).expand(-1, n, m, d), dim=3)

    def forward(self, X, Y):
        if X.shape[2] != Y.shape[2]:
            raise RuntimeError("SoftDTW: feature dimensions differ")
        if self.normalize:
            x, y = torch.cat([X, X, Y]), torch.cat([Y, X, Y])
            out = softdtw(self.dist_func(x, y), self.gamma, self.bandwidth)
            out_xy, out_xx, out_yy = torch.split(out, X.shape[0])
            return out_xy - 1 / 2 * (out_xx + out_yy)
        pad = torch.nn.functional.pad
        D_xy = pad(self.dist_func(X, Y), (0, 0, 1, 1), "constant", 0)           # zero first / last row (:416)
        D_yx = pad(self.dist_func(Y, X), (0, 0, 1, 1), "constant", 0)
        a = softdtw(D_xy, self.gamma, self.bandwidth).unsqueeze(-1)
        b = softdtw(D_yx, self.gamma, self.bandwidth).unsqueeze(-1)
        return (a + b) / 2
