"""Fused optimizers of the training step (reference: train.py:319-336, step at
train.py:529 / 651).

`FusedAdam` / `FusedSGD` are `torch.optim.Optimizer` subclasses with torch's own state layout
(`exp_avg`, `exp_avg_sq`, `step` / `momentum_buffer`), so `state_dict()` round-trips with the stock
optimizers the reference checkpoints (train.py:362, 436) and LR schedulers (train.py:322, 336) drive them
unchanged.  `step()` is ONE launch per 48 parameter tensors of the multi-tensor kernels in
`csrc/step_ops.cu`; `grad_scale` folds the data-parallel 1/world averaging into the same pass
(`bdbnn_b200.ddp.GradAllReduce(..., scale=False)`).  CUDA only: CPU parameters raise."""
import ctypes

import torch

from . import _lib


def _tables(entries):
    """entries: list of (p, g, m, v, wd, lr) -> ctypes host tables."""
    n = len(entries)
    PA, FA, IA = ctypes.c_void_p * n, ctypes.c_float * n, ctypes.c_int64 * n
    ptr = lambda t: t.data_ptr() if t is not None else 0
    return (PA(*[ptr(e[0]) for e in entries]), PA(*[ptr(e[1]) for e in entries]), PA(*[ptr(e[2]) for e in entries]),
            PA(*[ptr(e[3]) for e in entries]), IA(*[e[0].numel() for e in entries]),
            FA(*[float(e[4]) for e in entries]), FA(*[float(e[5]) for e in entries]))


def _dense_like(p, g):
    """Gradient with the parameter's own (dense) layout, so the kernels can walk raw storage."""
    if g.stride() != p.stride():
        g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
    return g


def _device_of(param_groups):
    """Device of the first parameter that has a gradient (the launches go to ITS current stream)."""
    for group in param_groups:
        for p in group["params"]:
            if p.grad is not None and p.is_cuda:
                return p.device
    return None


def _check(p):
    if not p.is_cuda:
        raise RuntimeError("bdbnn_b200.optim: parameters must live on a CUDA device (no CPU path)")
    if p.dtype != torch.float32:
        raise RuntimeError("bdbnn_b200.optim: fp32 parameters only")
    if not (p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)):
        raise RuntimeError("bdbnn_b200.optim: parameters must be dense (contiguous or channels_last)")


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (L2 weight decay, no amsgrad) — train.py:331-335."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = float(grad_scale)

    @torch.no_grad()
    def step(self, closure=None):
        dev = _device_of(self.param_groups)
        if dev is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):                 # device guard (parameters on a non-current device)
                return self._step(closure)
        return self._step(closure)

    def _step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    _check(p)
        L = _lib.lib()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for group in self.param_groups:                 # groups may differ in betas / eps: one call each
            entries, step_no = [], None
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p)
                state = self.state[p]
                if not state:
                    state["step"] = torch.tensor(0.0)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                s = int(state["step"])
                if step_no is not None and s != step_no:      # parameters added later: separate launch
                    self._launch(L, entries, group, step_no, st)
                    entries = []
                step_no = s
                entries.append((p, _dense_like(p, p.grad), state["exp_avg"], state["exp_avg_sq"],
                                group["weight_decay"], group["lr"]))
            if entries:
                self._launch(L, entries, group, step_no, st)
        return loss

    def _launch(self, L, entries, group, step_no, st):
        P, G, M, V, N, WD, LR = _tables(entries)
        b1, b2 = group["betas"]
        _lib.check(L.bdbnn_optim_adam_multi(P, G, M, V, N, WD, LR, len(entries), float(b1), float(b2),
                                            float(group["eps"]), int(step_no), self.grad_scale, st), "optim_adam_multi")
        _lib.count((len(entries) + 47) // 48)


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD(momentum, weight_decay) semantics (dampening 0, no nesterov) — train.py:319-321."""

    def __init__(self, params, lr=0.1, momentum=0.0, weight_decay=0.0, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0,
                                      nesterov=False))
        self.grad_scale = float(grad_scale)

    @torch.no_grad()
    def step(self, closure=None):
        dev = _device_of(self.param_groups)
        if dev is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):                 # device guard (parameters on a non-current device)
                return self._step(closure)
        return self._step(closure)

    def _step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    _check(p)
        L = _lib.lib()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for group in self.param_groups:
            mu = float(group["momentum"])
            fresh, warm = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p)
                state = self.state[p]
                buf = None
                first = False
                if mu != 0.0:
                    if state.get("momentum_buffer") is None:
                        state["momentum_buffer"] = torch.empty_like(p, memory_format=torch.preserve_format)
                        first = True
                    buf = state["momentum_buffer"]
                (fresh if first else warm).append((p, _dense_like(p, p.grad), buf, None, group["weight_decay"],
                                                   group["lr"]))
            for entries, first in ((fresh, 1), (warm, 0)):
                if entries:
                    P, G, M, _, N, WD, LR = _tables(entries)
                    _lib.check(L.bdbnn_optim_sgd_multi(P, G, M, N, WD, LR, len(entries), mu, first, self.grad_scale,
                                                       st), "optim_sgd_multi")
                    _lib.count((len(entries) + 47) // 48)
        return loss
