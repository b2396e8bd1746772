"""GPU parity of the soft-DTW kernels (csrc/softdtw.cu) through the C ABI against the goldens of the reference's CPU
path (what the reference's own profile() test, models/OTAM.py:461-505, holds its numba.cuda kernels against) and the
fp64 oracle.  fp32 kernels: 1e-4 relative on values and gradients (the reference's own tolerances: allclose / 1e-5..1e-3)."""
import numpy as np
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("name", list(H.SOFTDTW_CASES))
def test_softdtw_tables_match_reference_golden(name):
    from clip_spm_b200 import _lib, ops
    import ctypes
    B, N, M, d, gamma, bw, seed = H.SOFTDTW_CASES[name]
    g = H.golden(name)
    X, Y, D = O.make_softdtw_inputs(B, N, M, d, seed)
    Dc = D.cuda().requires_grad_(True)
    out = ops.softdtw(Dc, gamma, bw)
    assert H.rel_err(out, g["R"][:, -2, -2]) < TOL
    out.sum().backward()
    assert H.rel_err(Dc.grad, g["E"]) < TOL
    # the whole cumulative table, +inf pattern included
    lib = _lib.load()
    R = torch.empty(B, N + 2, M + 2, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.spm_softdtw_forward(st, B, N, M, ctypes.c_void_p(D.cuda().data_ptr()), gamma, bw,
                                       ctypes.c_void_p(R.data_ptr()), None))
    R, Rg = R.cpu(), g["R"]
    fin = torch.isfinite(Rg)
    assert torch.equal(torch.isfinite(R), fin)
    assert float((R[fin] - Rg[fin]).abs().max()) < TOL * float(Rg[fin].abs().max())


@pytest.mark.parametrize("name", list(H.SOFTDTW_CASES))
def test_softdtw_module_matches_reference_golden(name):
    """ops.SoftDTW (mirror of models/OTAM.py:318-424): value, gradient w.r.t. X through the distance function, and
    the normalised variant"""
    from clip_spm_b200 import ops
    B, N, M, d, gamma, bw, seed = H.SOFTDTW_CASES[name]
    g = H.golden(name)
    X, Y, _ = O.make_softdtw_inputs(B, N, M, d, seed)
    mod = ops.SoftDTW(use_cuda=True, gamma=gamma, bandwidth=bw if bw > 0 else None)
    x = X.cuda().requires_grad_(True)
    val = mod(x, Y.cuda())
    assert tuple(val.shape) == tuple(g["module"].shape)
    assert H.rel_err(val, g["module"]) < TOL
    val.sum().backward()
    assert H.rel_err(x.grad, g["grad_x"]) < 1e-3
    k = min(N, M)
    nrm = ops.SoftDTW(use_cuda=True, gamma=gamma, normalize=True, bandwidth=bw if bw > 0 else None)(X[:, :k].cuda(), Y[:, :k].cuda())
    assert float((nrm.cpu() - g["module_norm"]).abs().max()) < 1e-3 * max(1.0, float(g["module_norm"].abs().max()))


def test_softdtw_large_batch_and_long_sequences():
    """B = 4096 problems of 8x8 (TA2N's shape at sweep scale) and one 300x257 problem, against the fp64 oracle"""
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(9)
    D = torch.rand(4096, 8, 8, generator=g)
    out = ops.softdtw(D.cuda(), 0.1)
    ref = torch.from_numpy(O.softdtw_forward_np(D.numpy(), 0.1)[:, -2, -2]).float()
    assert H.rel_err(out, ref) < TOL
    D = torch.rand(1, 300, 257, generator=g)
    Dc = D.cuda().requires_grad_(True)
    out = ops.softdtw(Dc, 1.0)
    R = O.softdtw_forward_np(D.numpy(), 1.0)
    assert abs(float(out.detach()) - R[0, -2, -2]) < 2e-4 * abs(R[0, -2, -2])
    out.backward()
    E = O.softdtw_backward_np(D.numpy(), R, 1.0)
    assert float((Dc.grad.cpu() - torch.from_numpy(E).float()).abs().max()) < 2e-4
    # gradient of a soft-min path cost: the entries of E along any monotone path sum to ... at least E is a flow:
    assert abs(float(Dc.grad[0, 0, 0]) - 1.0) < 1e-4 and abs(float(Dc.grad[0, -1, -1]) - 1.0) < 1e-4
