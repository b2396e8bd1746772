"""Optimiser half of the training step on the C-ABI library -- host-side mirrors of the two torch
objects run/main_run.py builds:

    reference                                                          here
    :84-88  torch.optim.Adam(model.parameters(), lr, betas=(0.5, 0.999),   optim.Adam(params, lr, betas, eps, weight_decay)
            weight_decay)
    :76     GradScaler(device, enabled=USE_AMP)                            optim.GradScaler(init_scale, growth_factor, ...)
    :252    scaler.scale(loss).backward()                                  scaler.scale(loss)
    :207-9  scaler.step(optimizer); scaler.update(); optimizer.zero_grad() same three calls

All parameters are stepped by ONE multi-tensor launch per kernel (spm_adam_step), and whether an overflowed step is skipped
is decided on the device: no `.item()` / host synchronisation anywhere (torch's GradScaler.step reads found_inf back on the
host).  fp32 CUDA parameters only; no CPU fallback."""
import ctypes

import torch

from . import _lib


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("clip_spm_b200.optim.Adam steps contiguous fp32 CUDA parameters (no CPU fallback)")
        self.param_groups = [dict(params=self.params, lr=float(lr), betas=tuple(betas), eps=float(eps),
                                  weight_decay=float(weight_decay))]   # lr schedulers edit param_groups[0]["lr"]
        h = ctypes.c_void_p()
        numel = (ctypes.c_longlong * len(self.params))(*[p.numel() for p in self.params])
        with torch.cuda.device(self.params[0].device):
            _lib.check(_lib.load().spm_adam_create(len(self.params), _ptr_array(self.params), numel, ctypes.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().spm_adam_destroy(h)
            except Exception:
                pass

    def step(self, _scaler_state=None):
        g = self.param_groups[0]
        grads = [None if p.grad is None else p.grad for p in self.params]
        for t in grads:
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError("gradients must be contiguous fp32")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().spm_adam_step(self._h, st, _ptr_array(grads), g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                                             g["weight_decay"],
                                             None if _scaler_state is None else ctypes.c_void_p(_scaler_state.data_ptr())))

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def state(self, i):
        """(exp_avg, exp_avg_sq, step) of parameter i, copied out of the library's state."""
        a, b, s = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(_lib.load().spm_adam_state(self._h, i, ctypes.byref(a), ctypes.byref(b), ctypes.byref(s)))
        n, dev = self.params[i].numel(), self.params[i].device
        torch.cuda.current_stream().synchronize()
        m, v, t = (torch.as_tensor(_DeviceArray(ptr.value, cnt), device=dev).clone() for ptr, cnt in ((a, n), (b, n), (s, 1)))
        return m.view_as(self.params[i]), v.view_as(self.params[i]), float(t)


class SGD(Adam):
    """torch.optim.SGD(params, lr, momentum, weight_decay) (run/main_run.py:92-96, SOLVER.OPTIM_METHOD == "sgd") on the same
    multi-tensor machinery; `state(i)[0]` is the momentum buffer."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        super().__init__(params, lr=lr, weight_decay=weight_decay)
        self.param_groups[0].update(momentum=float(momentum))

    def step(self, _scaler_state=None):
        g = self.param_groups[0]
        grads = [None if p.grad is None else p.grad for p in self.params]
        for t in grads:
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError("gradients must be contiguous fp32")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().spm_sgd_step(self._h, st, _ptr_array(grads), g["lr"], g["momentum"], g["weight_decay"],
                                            None if _scaler_state is None else ctypes.c_void_p(_scaler_state.data_ptr())))


class _DeviceArray:
    """A raw fp32 device buffer of the library, viewed through __cuda_array_interface__ (read-only use: state_dict)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class GradScaler:
    """torch.amp.GradScaler's scale / step / update / get_scale on a device-resident state [scale, growth_tracker, found_inf]."""

    def __init__(self, device="cuda", init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000,
                 enabled=True):
        self.enabled = bool(enabled)
        self.growth_factor, self.backoff_factor, self.growth_interval = float(growth_factor), float(backoff_factor), int(growth_interval)
        self._state = torch.tensor([float(init_scale), 0.0, 0.0], device=device) if self.enabled else None

    def scale(self, loss):
        return loss * self._state[0] if self.enabled else loss   # a device-side multiply: no host read of the scale

    def step(self, optimizer):
        optimizer.step(self._state if self.enabled else None)

    def update(self):
        if self.enabled:
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.load().spm_scaler_update(st, ctypes.c_void_p(self._state.data_ptr()), self.growth_factor,
                                                     self.backoff_factor, self.growth_interval))

    def get_scale(self):
        return float(self._state[0]) if self.enabled else 1.0
