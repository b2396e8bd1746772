"""TEST INFRASTRUCTURE (oracle/): numpy restatement of what run/main_run.py does with the gradients -- torch.optim.Adam
(:84-88, betas=(0.5, 0.999), weight_decay as classic L2) stepped through a GradScaler (:76, :207-209: unscale, skip the step
on inf / nan, grow / back off the scale).  Pinned against torch.optim.Adam and torch.amp.GradScaler themselves -- the
reference's own dependency -- in tests/test_optim_cpu.py; the CUDA kernels (csrc/optimizer.cu) are checked against it.
Only tests/ import this module."""
import numpy as np


class AdamOracle:
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) -- _single_tensor_adam, amsgrad=False."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.p = [np.array(p, dtype=np.float32) for p in params]
        self.m = [np.zeros_like(p) for p in self.p]
        self.v = [np.zeros_like(p) for p in self.p]
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.t = [0] * len(self.p)   # torch keeps state['step'] per parameter: one without a gradient does not advance

    def step(self, grads):
        for i, g in enumerate(grads):
            if g is None:
                continue
            self.t[i] += 1
            bc1, bc2 = 1.0 - self.b1 ** self.t[i], 1.0 - self.b2 ** self.t[i]
            step_size, sq = np.float32(self.lr / bc1), np.float32(bc2 ** 0.5)
            g = np.asarray(g, dtype=np.float32) + np.float32(self.wd) * self.p[i]
            self.m[i] = self.m[i] + np.float32(1.0 - self.b1) * (g - self.m[i])
            self.v[i] = np.float32(self.b2) * self.v[i] + np.float32(1.0 - self.b2) * (g * g)
            denom = np.sqrt(self.v[i]) / sq + np.float32(self.eps)
            self.p[i] = self.p[i] - step_size * (self.m[i] / denom)


class SGDOracle:
    """torch.optim.SGD(params, lr, momentum, weight_decay), dampening 0, nesterov False (run/main_run.py:92-96)."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        self.p = [np.array(p, dtype=np.float32) for p in params]
        self.buf = [None] * len(self.p)
        self.lr, self.mom, self.wd = lr, momentum, weight_decay

    def step(self, grads):
        for i, g in enumerate(grads):
            if g is None:
                continue
            g = np.asarray(g, dtype=np.float32) + np.float32(self.wd) * self.p[i]
            if self.mom != 0:
                self.buf[i] = g.copy() if self.buf[i] is None else np.float32(self.mom) * self.buf[i] + g
                g = self.buf[i]
            self.p[i] = self.p[i] - np.float32(self.lr) * g


class GradScalerOracle:
    """torch.amp.GradScaler(init_scale=65536, growth_factor=2, backoff_factor=0.5, growth_interval=2000):
    step(opt, grads) unscales, skips the optimiser step when a gradient is inf / nan; update() adapts the scale."""

    def __init__(self, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.scale, self.growth, self.backoff, self.interval = np.float32(init_scale), growth_factor, backoff_factor, growth_interval
        self.tracker, self.found_inf = 0, False

    def step(self, opt, grads):
        inv = np.float32(1.0 / np.float64(self.scale))
        un = [None if g is None else np.asarray(g, dtype=np.float32) * inv for g in grads]
        self.found_inf = any(g is not None and not np.isfinite(g).all() for g in un)
        if not self.found_inf:
            opt.step(un)
        return un

    def update(self):
        if self.found_inf:
            self.scale, self.tracker = np.float32(self.scale * self.backoff), 0
        else:
            self.tracker += 1
            if self.tracker >= self.interval:
                grown = np.float32(self.scale * np.float32(self.growth))
                if np.isfinite(grown):
                    self.scale = grown
                self.tracker = 0
        self.found_inf = False
