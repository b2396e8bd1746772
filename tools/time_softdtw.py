"""Soft-DTW forward / backward at batch scale: this library's warp-wavefront kernels (csrc/softdtw.cu) against the bar
the reference ships for this op -- its numba.cuda kernels (models/OTAM.py:34-130: one block per pair, one thread per
row, a block barrier per anti-diagonal, R / E re-read from global memory every step).  The reference cannot travel to
the GPU box, so `numba_bar` below RESTATES that kernel design with numba.cuda for timing only (baseline harness, not
product code; values are cross-checked against this library's output).
    python tools/time_softdtw.py            # B = 4096, shapes 8x8 .. 40x38"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import _lib
import ctypes


def ours(D, gamma, bw):
    lib = _lib.load()
    B, N, M = D.shape
    R = torch.empty(B, N + 2, M + 2, device=D.device); out = torch.empty(B, device=D.device)
    E = torch.empty(B, N, M, device=D.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    fwd = lambda: _lib.check(lib.spm_softdtw_forward(st, B, N, M, p(D), gamma, bw, p(R), p(out)))
    bwd = lambda: _lib.check(lib.spm_softdtw_backward(st, B, N, M, p(D), p(R), gamma, bw, p(E)))
    return fwd, bwd, out, E


def numba_bar(D, gamma, bw):
    """block per pair / thread per row / syncthreads per diagonal, tables in global memory (the reference's design)"""
    from numba import cuda

    @cuda.jit
    def fwd_k(Dm, g, band, n, m, passes, Rm):
        b = cuda.blockIdx.x; t = cuda.threadIdx.x
        ig = 1.0 / g
        for p in range(passes):
            col = max(0, min(p - t, m - 1))
            if t + col == p and t < n and col < m:
                i = t + 1; j = col + 1
                if not (abs(i - j) > band > 0):
                    x0 = -Rm[b, i - 1, j - 1] * ig; x1 = -Rm[b, i - 1, j] * ig; x2 = -Rm[b, i, j - 1] * ig
                    mx = max(max(x0, x1), x2)
                    sm = math.exp(x0 - mx) + math.exp(x1 - mx) + math.exp(x2 - mx)
                    Rm[b, i, j] = Dm[b, i - 1, j - 1] - g * (math.log(sm) + mx)
            cuda.syncthreads()

    @cuda.jit
    def bwd_k(Dp, Rm, ig, band, n, m, passes, Em):
        b = cuda.blockIdx.x; t = cuda.threadIdx.x
        for p in range(passes):
            q = passes - p - 1
            col = max(0, min(q - t, m - 1))
            if t + col == q and t < n and col < m:
                i = t + 1; j = col + 1
                if math.isinf(Rm[b, i, j]):
                    Rm[b, i, j] = -math.inf
                if not (abs(i - j) > band > 0):
                    a = math.exp((Rm[b, i + 1, j] - Rm[b, i, j] - Dp[b, i + 1, j]) * ig)
                    bb = math.exp((Rm[b, i, j + 1] - Rm[b, i, j] - Dp[b, i, j + 1]) * ig)
                    c = math.exp((Rm[b, i + 1, j + 1] - Rm[b, i, j] - Dp[b, i + 1, j + 1]) * ig)
                    Em[b, i, j] = Em[b, i + 1, j] * a + Em[b, i, j + 1] * bb + Em[b, i + 1, j + 1] * c
            cuda.syncthreads()

    B, N, M = D.shape
    tpb = max(N, M); passes = 2 * tpb - 1
    state = {}

    def fwd():   # includes the table initialisation the reference's Function.forward performs (:147-149)
        R = torch.ones((B, N + 2, M + 2), device=D.device) * math.inf
        R[:, 0, 0] = 0
        fwd_k[B, tpb](cuda.as_cuda_array(D), gamma, bw, N, M, passes, cuda.as_cuda_array(R))
        state["R"] = R

    def bwd():   # includes the padding / edits of Function.backward (:160-167)
        R = state["R"].clone()
        Dp = torch.zeros((B, N + 2, M + 2), device=D.device); Dp[:, 1:N + 1, 1:M + 1] = D
        R[:, :, -1] = -math.inf; R[:, -1, :] = -math.inf; R[:, -1, -1] = R[:, -2, -2]
        E = torch.zeros((B, N + 2, M + 2), device=D.device); E[:, -1, -1] = 1
        bwd_k[B, tpb](cuda.as_cuda_array(Dp), cuda.as_cuda_array(R), 1.0 / gamma, bw, N, M, passes, cuda.as_cuda_array(E))
        state["E"] = E[:, 1:N + 1, 1:M + 1]
    return fwd, bwd, state


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B = 4096
for N, M, gamma in ((8, 8, 0.1), (16, 16, 0.1), (17, 15, 1.0), (32, 32, 0.5), (40, 38, 0.5)):
    D = torch.rand(B, N, M, device="cuda", generator=torch.Generator(device="cuda").manual_seed(N * 100 + M))
    f, b, out, E = ours(D, gamma, 0.0)
    t_f, t_b = timeit(f), timeit(b)
    byts_f = B * (N * M + (N + 2) * (M + 2) + 1) * 4; byts_b = B * (2 * N * M + (N + 2) * (M + 2)) * 4
    line = "softdtw B=%d %dx%d: forward %.1f us (%.0f GB/s), backward %.1f us (%.0f GB/s)" % (
        B, N, M, t_f, byts_f / t_f / 1e3, t_b, byts_b / t_b / 1e3)
    try:
        nf, nb, st = numba_bar(D, gamma, 0.0)
        n_f = timeit(nf); n_b = timeit(nb)
        err_f = float((st["R"][:, N, M] - out).abs().max()); err_b = float((st["E"] - E).abs().max())
        line += " | numba bar (reference kernel design): forward %.1f us, backward %.1f us -> %.1fx / %.1fx; max diff %.1e / %.1e" % (
            n_f, n_b, n_f / t_f, n_b / t_b, err_f, err_b)
    except Exception as ex:
        line += " | numba.cuda unavailable here: %r" % (ex,)
    print(line)
