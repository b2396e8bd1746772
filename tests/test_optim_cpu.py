"""CPU suite of the optimiser half of the training step: the numpy oracle against the objects the
reference itself uses (run/main_run.py:84-88 torch.optim.Adam(betas=(0.5, 0.999), weight_decay); :76,207-209 GradScaler)."""
import numpy as np
import pytest
import torch

from oracle.optim_oracle import AdamOracle, GradScalerOracle, SGDOracle


def _problem(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(7,), (33, 5), (3, 4, 5), (1,), (70000,)]
    params = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) * (10.0 ** (k % 3 - 1)) for s in shapes] for k in range(6)]
    return params, grads


def test_adam_oracle_matches_torch_adam():
    params, grads = _problem()
    tp = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam(tp, lr=1e-3, betas=(0.5, 0.999), weight_decay=5e-4)
    orc = AdamOracle([p.numpy() for p in params], lr=1e-3, betas=(0.5, 0.999), weight_decay=5e-4)
    for k, gs in enumerate(grads):
        for p, g in zip(tp, gs):
            p.grad = g.clone()
        if k == 3:
            tp[1].grad = None            # a parameter without a gradient: skipped, its own step count does not advance
            gs = [g if i != 1 else None for i, g in enumerate(gs)]
        opt.step()
        orc.step([None if g is None else g.numpy() for g in gs])
    for p, q in zip(tp, orc.p):
        assert np.allclose(p.detach().numpy(), q, rtol=2e-6, atol=5e-7)
    for p, m, v in zip(tp, orc.m, orc.v):
        assert np.allclose(opt.state[p]["exp_avg"].numpy(), m, rtol=2e-6, atol=1e-6)
        assert np.allclose(opt.state[p]["exp_avg_sq"].numpy(), v, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("momentum", [0.0, 0.9])
def test_sgd_oracle_matches_torch_sgd(momentum):
    params, grads = _problem(1)
    tp = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.SGD(tp, lr=1e-2, momentum=momentum, weight_decay=5e-4)
    orc = SGDOracle([p.numpy() for p in params], lr=1e-2, momentum=momentum, weight_decay=5e-4)
    for k, gs in enumerate(grads):
        for p, g in zip(tp, gs):
            p.grad = g.clone()
        if k == 0:
            tp[2].grad = None            # no gradient on the first step: its momentum buffer starts one step later
            gs = [g if i != 2 else None for i, g in enumerate(gs)]
        opt.step()
        orc.step([None if g is None else g.numpy() for g in gs])
    for p, q in zip(tp, orc.p):
        assert np.allclose(p.detach().numpy(), q, rtol=2e-6, atol=2e-6)


def test_grad_scaler_oracle_matches_torch_grad_scaler():
    """scale, skip-on-overflow and growth / backoff bookkeeping against torch.amp.GradScaler on CPU tensors"""
    try:
        scaler = torch.amp.GradScaler("cpu", init_scale=1024.0, growth_interval=3)
    except Exception as e:   # an older torch without the CPU scaler
        pytest.skip("torch.amp.GradScaler('cpu') unavailable: %r" % (e,))
    params, grads = _problem(1)
    tp = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam(tp, lr=1e-2, betas=(0.5, 0.999))
    orc = AdamOracle([p.numpy() for p in params], lr=1e-2, betas=(0.5, 0.999))
    osc = GradScalerOracle(init_scale=1024.0, growth_interval=3)
    scaler.scale(torch.ones(()))                   # torch creates its scale tensor lazily, on the first scale() call
    for k, gs in enumerate(grads):
        scaled = [g * float(scaler.get_scale()) for g in gs]
        if k == 1:
            scaled[2][0, 0, 0] = float("inf")      # an overflowed step: skipped, scale halved
        if k == 4:
            scaled[0][3] = float("nan")
        for p, g in zip(tp, scaled):
            p.grad = g.clone()
        scaler.step(opt)
        scaler.update()
        osc.step(orc, [g.numpy() for g in scaled])
        osc.update()
        assert float(scaler.get_scale()) == float(osc.scale), k
    for p, q in zip(tp, orc.p):
        assert np.allclose(p.detach().numpy(), q, rtol=2e-6, atol=5e-7)
    assert orc.t == [4] * len(params)   # two of the six steps were skipped


def test_optim_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clip_spm_b200 import optim
    with pytest.raises(RuntimeError):
        optim.Adam([torch.zeros(4, requires_grad=True)])
