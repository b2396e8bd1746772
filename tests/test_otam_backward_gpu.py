"""Gradients of the metric tail (spm_otam_distance_backward) against autograd through the oracle restatement of
otam_distance / cos_sim / OTAM_cum_dist_v2 (the expressions the reference differentiates when it trains)."""
import pytest
import torch

from oracle import clipspm_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_grads(sup, tgt, single, go, dtype=torch.float64):
    s = sup.detach().cpu().to(dtype).requires_grad_(True)
    t = tgt.detach().cpu().to(dtype).requires_grad_(True)
    outs = torch.stack([O.otam_distance(s[p], t[p], single) for p in range(s.shape[0])])
    (outs * go.detach().cpu().to(dtype)).sum().backward()
    return outs.detach(), s.grad, t.grad


@pytest.mark.parametrize("P,W,Q,T,D,single", [(1, 5, 5, 8, 512, False), (2, 5, 3, 8, 512, True), (2, 3, 4, 16, 512, False),
                                              (1, 5, 5, 8, 1024, False), (1, 2, 2, 2, 512, False), (1, 4, 2, 11, 512, False)])
def test_otam_backward_matches_autograd_through_oracle(P, W, Q, T, D, single):
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(P * 1000 + T * 10 + W)
    base = torch.randn(1, 1, 1, D, generator=g)                  # shared component: cosine similarities well above 0
    sup = (torch.randn(P, W, T, D, generator=g) + 0.7 * base).cuda().requires_grad_(True)
    tgt = (torch.randn(P, Q, T, D, generator=g) + 0.7 * base).cuda().requires_grad_(True)
    go = torch.randn(P, Q, W, generator=g)
    out = ops.otam_distance(sup, tgt, single)
    (out * go.cuda()).sum().backward()
    ref_out, gs, gt = _oracle_grads(sup, tgt, single, go)
    assert torch.allclose(out.detach().cpu().double(), ref_out, atol=1e-4, rtol=1e-5)
    for mine, ref in ((sup.grad, gs), (tgt.grad, gt)):
        err = (mine.cpu().double() - ref).abs().max().item()
        assert err < 2e-4 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_otam_backward_directional_derivative():
    """finite differences of the library's own forward along a random direction (independent of the oracle)"""
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(3)
    sup = torch.randn(1, 5, 8, 512, generator=g).cuda()
    tgt = torch.randn(1, 5, 8, 512, generator=g).cuda()
    ds, dt = torch.randn(sup.shape, generator=g).cuda(), torch.randn(tgt.shape, generator=g).cuda()
    s, t = sup.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    ops.otam_distance(s, t).sum().backward()
    analytic = float((s.grad * ds).sum() + (t.grad * dt).sum())
    h = 1e-2
    with torch.no_grad():
        fp = ops.otam_distance(sup + h * ds, tgt + h * dt).double().sum()
        fm = ops.otam_distance(sup - h * ds, tgt - h * dt).double().sum()
    numeric = float((fp - fm) / (2 * h))
    assert abs(analytic - numeric) < 2e-2 * max(1.0, abs(numeric)), (analytic, numeric)


def test_otam_backward_scaling_and_errors():
    from clip_spm_b200 import ops
    g = torch.Generator().manual_seed(4)
    sup = torch.randn(1, 3, 8, 512, generator=g).cuda().requires_grad_(True)
    tgt = torch.randn(1, 2, 8, 512, generator=g).cuda().requires_grad_(True)
    ops.otam_distance(sup, tgt, alpha=1.0).sum().backward()
    g1 = sup.grad.clone()
    sup.grad = None
    ops.otam_distance(sup, tgt, alpha=-2.5).sum().backward()
    assert torch.allclose(sup.grad, -2.5 * g1, rtol=1e-5, atol=1e-7)
    with pytest.raises(RuntimeError):
        ops.otam_distance(sup, tgt, beta=1.0, out=torch.zeros(1, 2, 3, device="cuda"))
    with torch.no_grad():
        assert ops.otam_distance(sup, tgt).shape == (1, 2, 3)     # inference path unchanged


def test_otam_backward_matches_reference_golden():
    """golden = autograd through the reference's own CNN.otam_distance (oracle/pin_against_reference.py otam_grad)"""
    from clip_spm_b200 import ops
    from tests import helpers as H
    g = H.golden("otam_grad_3w2q_t8")
    W, Q, T, D, seed = [int(v) for v in g["shape"]]
    sup, tgt, go = O.make_otam_grad_inputs(W, Q, T, D, seed)
    s, t = sup.cuda()[None].requires_grad_(True), tgt.cuda()[None].requires_grad_(True)
    out = ops.otam_distance(s, t)
    (out[0] * go.cuda()).sum().backward()
    assert torch.allclose(out[0].detach().cpu(), g["out"], atol=1e-4, rtol=1e-5)
    for mine, ref in ((s.grad[0].cpu(), g["grad_support"]), (t.grad[0].cpu(), g["grad_target"])):
        assert (mine - ref).abs().max().item() < 2e-4 * ref.abs().max().item()
