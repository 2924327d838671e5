"""Fused Adam over the optimisable tensors: one CUDA launch per step instead of torch's ~30 foreach launches.

Same update rule, hyper-parameters and state-dict keys ('step', 'exp_avg', 'exp_avg_sq') as ``torch.optim.Adam`` with its
defaults (the optimiser the reference builds, reconstruction.py:285-368, model_params.py:11), so checkpoints that carry
``optim_state_dict`` stay interchangeable.  Tensors whose ``.grad`` is None are skipped, like torch does
(frozen by start_iter, reconstruction.py:783-790), and -- like torch -- every tensor keeps its OWN ``step``: the kernel reads
the per-tensor device counters (``state['step']``, float32 scalars), so a tensor that joins at a later iteration starts its
bias correction at 1, and a loaded state dict is honoured without any extra bookkeeping.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def _state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not (torch.is_tensor(st["step"]) and st["step"].is_cuda and st["step"].dtype == torch.float32):
            # a state dict written by torch.optim.Adam on the CPU / with a python number: move the counter to the device once
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=p.device)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        todo = []
        betas = eps = None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam runs on CUDA tensors only")
                if betas is None:
                    betas, eps = group["betas"], group["eps"]
                elif (betas, eps) != (group["betas"], group["eps"]):
                    raise ValueError("FusedAdam needs the same betas/eps in every param group")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous parameters and gradients")
                todo.append((p, self._state(p), float(group["lr"])))
        if not todo:
            return loss
        lib = _lib.lib()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(0, len(todo), 8):
            chunk = todo[i:i + 8]
            n = len(chunk)
            arr = lambda xs: (C.c_void_p * n)(*xs)
            _lib.check(lib.ptyb200_adam_step(
                n, arr([p.data_ptr() for p, _, _ in chunk]), arr([p.grad.data_ptr() for p, _, _ in chunk]),
                arr([s["exp_avg"].data_ptr() for _, s, _ in chunk]), arr([s["exp_avg_sq"].data_ptr() for _, s, _ in chunk]),
                arr([s["step"].data_ptr() for _, s, _ in chunk]),
                (C.c_float * n)(*[lr for _, _, lr in chunk]), (C.c_int64 * n)(*[p.numel() for p, _, _ in chunk]),
                betas[0], betas[1], eps, stream))
        return loss
