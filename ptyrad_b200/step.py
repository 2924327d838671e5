"""Per-batch reconstruction step and the data-parallel gradient exchange.

``recon_batch`` is the body of the reference's hot loop (``recon_step``, reconstruction.py:741-770: forward, measurements,
loss, backward, optimizer step, clear cache) with the three host synchronisations per batch removed: losses stay on the
device and are returned as one (5,) tensor; nothing is copied to the host here.

``GradArena`` replaces DDP's bucketed reducer (reconstruction.py:132,753 via accelerate): the ``.grad`` of every
optimisable tensor is a view into ONE contiguous float32 buffer, so a step needs one memset and - on several GPUs - one
NCCL all-reduce (sum, then 1/world) over NVLink.  Scan positions shard across ranks (each rank runs its slice of the
batch and builds its OWN loss on it, exactly what DDP with split_batches=True computes); measurements are immutable and
are never broadcast.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .losses import MeasurementView


class GradArena:
    def __init__(self, model):
        self.params = [p for g in model.optimizable_params for p in g["params"]]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.attach()

    def attach(self):
        """(re)bind .grad views; tensors whose requires_grad is off keep grad=None so the optimiser skips them, as
        zero_grad(set_to_none=True) does in the reference (reconstruction.py:739,760,783-790)."""
        for p, v in zip(self.params, self.views):
            if p.requires_grad:
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v
            else:
                p.grad = None

    def zero(self):
        self.flat.zero_()

    def allreduce(self, world: int, group=None):
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / world)


def shard_indices(indices, rank: int, world: int):
    """Rank `rank`'s slice of one global batch: contiguous split, sizes differing by at most one
    (accelerate's split_batches=True dispatch, utils/common.py:61-65, reconstruction.py:134-137)."""
    return np.array_split(np.asarray(indices), world)[rank]


def recon_batch(model, loss_fn, optimizer, indices, arena: GradArena | None = None, world: int = 1,
                grad_accumulation: int = 1, do_step: bool = True, measurements=None):
    """One batch: zero grads, forward, loss, backward, (all-reduce), optimizer step.  Returns the 5 loss terms as a device
    tensor (no host sync)."""
    if arena is not None:
        arena.attach()
        arena.zero()
    else:
        optimizer.zero_grad()
    dp = model(indices)
    idx = model._current_object_patches.idx
    if measurements is not None:
        meas = measurements
    elif model.meas_padded is None and model.meas_scale_factors is None:
        meas = MeasurementView(model.measurements, idx)          # read rows in place, no gathered copy
    else:
        meas = model.get_measurements(idx)
    total, losses = loss_fn(dp, meas, model._current_object_patches, model.omode_occu)
    (total / grad_accumulation if grad_accumulation != 1 else total).backward()
    if world > 1:
        if arena is None:
            raise RuntimeError("multi-GPU steps need a GradArena")
        arena.allreduce(world)
    if do_step:
        optimizer.step()
    model.clear_cache()
    return torch.stack([l.detach().reshape(()) for l in losses])
