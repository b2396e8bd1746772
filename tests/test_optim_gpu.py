"""GPU parity of the optimiser kernels (csrc/optimizer.cu) through the C ABI: clip_spm_b200.optim.Adam / GradScaler against the
numpy oracle (itself pinned to torch.optim.Adam / torch.amp.GradScaler on the CPU) and against torch's own CUDA objects."""
import numpy as np
import pytest
import torch

from oracle.optim_oracle import AdamOracle, GradScalerOracle, SGDOracle
from tests.test_optim_cpu import _problem

pytestmark = pytest.mark.gpu


def test_adam_matches_oracle_and_torch():
    from clip_spm_b200 import optim
    params, grads = _problem()
    mine = [p.clone().cuda().requires_grad_(True) for p in params]
    theirs = [p.clone().cuda().requires_grad_(True) for p in params]
    opt = optim.Adam(mine, lr=1e-3, betas=(0.5, 0.999), weight_decay=5e-4)
    ref = torch.optim.Adam(theirs, lr=1e-3, betas=(0.5, 0.999), weight_decay=5e-4)
    orc = AdamOracle([p.numpy() for p in params], lr=1e-3, betas=(0.5, 0.999), weight_decay=5e-4)
    for k, gs in enumerate(grads):
        for p, q, g in zip(mine, theirs, gs):
            p.grad = g.clone().cuda()
            q.grad = g.clone().cuda()
        if k == 3:
            mine[1].grad = None          # a parameter without a gradient is left alone (torch skips it too)
            theirs[1].grad = None
            gs = [g if i != 1 else None for i, g in enumerate(gs)]
        opt.step()
        ref.step()
        orc.step([None if g is None else g.numpy() for g in gs])
    for p, q, o in zip(mine, theirs, orc.p):
        assert np.allclose(p.detach().cpu().numpy(), o, rtol=2e-6, atol=5e-7)
        assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=5e-7)
    m, v, t = opt.state(4)
    assert t == len(grads) and opt.state(1)[2] == len(grads) - 1 and np.allclose(m.cpu().numpy(), orc.m[4], rtol=2e-6, atol=1e-6)
    assert np.allclose(v.cpu().numpy(), orc.v[4], rtol=1e-5, atol=1e-7)
    opt.zero_grad()
    assert all(p.grad is None for p in mine)


@pytest.mark.parametrize("momentum", [0.0, 0.9])
def test_sgd_matches_oracle_and_torch(momentum):
    from clip_spm_b200 import optim
    params, grads = _problem(1)
    mine = [p.clone().cuda().requires_grad_(True) for p in params]
    theirs = [p.clone().cuda().requires_grad_(True) for p in params]
    opt = optim.SGD(mine, lr=1e-2, momentum=momentum, weight_decay=5e-4)
    ref = torch.optim.SGD(theirs, lr=1e-2, momentum=momentum, weight_decay=5e-4)
    orc = SGDOracle([p.numpy() for p in params], lr=1e-2, momentum=momentum, weight_decay=5e-4)
    for k, gs in enumerate(grads):
        for p, q, g in zip(mine, theirs, gs):
            p.grad = g.clone().cuda()
            q.grad = g.clone().cuda()
        if k == 0:
            mine[2].grad = None
            theirs[2].grad = None
            gs = [g if i != 2 else None for i, g in enumerate(gs)]
        opt.step()
        ref.step()
        orc.step([None if g is None else g.numpy() for g in gs])
    for p, q, o in zip(mine, theirs, orc.p):
        assert np.allclose(p.detach().cpu().numpy(), o, rtol=2e-6, atol=2e-6)
        assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=2e-6)


def test_grad_scaler_skips_overflowed_steps_on_the_device():
    from clip_spm_b200 import optim
    params, grads = _problem(1)
    mine = [p.clone().cuda().requires_grad_(True) for p in params]
    opt = optim.Adam(mine, lr=1e-2, betas=(0.5, 0.999))
    scaler = optim.GradScaler(init_scale=1024.0, growth_interval=3)
    orc = AdamOracle([p.numpy() for p in params], lr=1e-2, betas=(0.5, 0.999))
    osc = GradScalerOracle(init_scale=1024.0, growth_interval=3)
    for k, gs in enumerate(grads):
        scaled = [g * float(osc.scale) for g in gs]
        if k == 1:
            scaled[2][0, 0, 0] = float("inf")
        if k == 4:
            scaled[0][3] = float("nan")
        for p, g in zip(mine, scaled):
            p.grad = g.clone().cuda()
        scaler.step(opt)
        scaler.update()
        osc.step(orc, [g.numpy() for g in scaled])
        osc.update()
        assert scaler.get_scale() == float(osc.scale), k
    for p, o in zip(mine, orc.p):
        assert np.allclose(p.detach().cpu().numpy(), o, rtol=2e-6, atol=5e-7)
    assert opt.state(0)[2] == 4      # two of the six steps were skipped, decided without a host round trip
    loss = torch.ones((), device="cuda")
    assert float(scaler.scale(loss)) == scaler.get_scale()


def test_disabled_scaler_is_a_plain_step():
    from clip_spm_b200 import optim
    p = torch.ones(10, device="cuda", requires_grad=True)
    opt = optim.Adam([p], lr=0.1, betas=(0.5, 0.999))
    scaler = optim.GradScaler(enabled=False)
    p.grad = torch.full((10,), 2.0, device="cuda")
    scaler.step(opt)
    scaler.update()
    assert torch.allclose(p.detach(), torch.full((10,), 0.9, device="cuda"), atol=1e-6)   # first Adam step moves by lr
