"""Fused Adam over the optimisable tensors: one CUDA launch per step instead of torch's ~30 foreach launches.

Same update rule, hyper-parameters and state-dict keys ('step', 'exp_avg', 'exp_avg_sq') as ``torch.optim.Adam`` with its
defaults (the optimiser the reference builds, reconstruction.py:285-368, model_params.py:11), so checkpoints that carry
``optim_state_dict`` stay interchangeable.  Tensors whose ``.grad`` is None are skipped, like torch does
(frozen by start_iter, reconstruction.py:783-790).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._step_dev = None

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
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous parameters and gradients")
                todo.append((p, st, float(group["lr"])))
        if not todo:
            return loss
        dev = todo[0][0].device
        if self._step_dev is None or self._step_dev.device != dev:
            first = todo[0][1]["step"]
            self._step_dev = first.to(torch.int64).reshape(1).clone()
        lib = _lib.lib()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if len(todo) > 8:
            raise RuntimeError("FusedAdam handles at most 8 tensors (PtychoAD has 6)")
        for i in range(0, len(todo), 8):
            chunk = todo[i:i + 8]
            n = len(chunk)
            arr = lambda xs: (C.c_void_p * n)(*xs)
            counter = self._step_dev          # device int64, incremented inside the call
            _lib.check(lib.ptyb200_adam_step(
                n, arr([p.data_ptr() for p, _, _ in chunk]), arr([p.grad.data_ptr() for p, _, _ in chunk]),
                arr([s["exp_avg"].data_ptr() for _, s, _ in chunk]), arr([s["exp_avg_sq"].data_ptr() for _, s, _ in chunk]),
                (C.c_float * n)(*[lr for _, _, lr in chunk]), (C.c_int64 * n)(*[p.numel() for p, _, _ in chunk]),
                betas[0], betas[1], eps, C.c_void_p(counter.data_ptr()), stream))
        for _, st, _ in todo:
            st["step"] += 1          # kept for state_dict compatibility (device tensor, no sync)
        return loss
