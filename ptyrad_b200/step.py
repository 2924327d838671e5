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

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine
from ._lib import ptr
from .losses import CombinedLoss, MeasurementView


class GradArena:
    def __init__(self, model):
        self.params = [p for g in model.optimizable_params for p in g["params"]]
        # every view starts on a 256-byte boundary: the kernels write gradients with 8- and 16-byte vector accesses
        A = 64
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + A - 1) // A * A
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
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
        """Mean over the ranks of every gradient: ONE collective on the flat buffer.  NCCL averages inside the reduction
        (ReduceOp.AVG: no separate scaling launch after the last adjoint CTA); gloo (CPU tests) has no AVG, so sum then scale."""
        if world > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.mul_(1.0 / world)


def shard_indices(indices, rank: int, world: int):
    """Rank `rank`'s slice of one global batch: contiguous split, sizes differing by at most one
    (accelerate's split_batches=True dispatch, utils/common.py:61-65, reconstruction.py:134-137)."""
    indices = np.asarray(indices)
    if len(indices) < world:
        # np.array_split would hand some rank an empty shard: that rank would fail its kernel-configuration check while the others
        # block in the all-reduce.  Every rank sees the same global batch, so every rank raises here.
        raise ValueError(f"a global batch of {len(indices)} positions cannot be split over {world} ranks")
    return np.array_split(indices, world)[rank]


_SIDE = {}


def _side_stream(dev, which: str = "aux"):
    """Side streams per device (forked from / joined into the caller's stream; a stream capture records them as parallel branches):
    "aux" for work that is independent of the multislice kernels, "half" for the second half of a split step."""
    key = (dev.type, dev.index, which)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def direct_step_eligible(model, loss_fn, arena, grad_accumulation, do_step, measurements) -> bool:
    """The autograd-free step covers: native loss terms (single / poissn / pacbed / sparse), detector blur, on-the-fly measurement
    pad / resample; gradients written once into the arena (no accumulation over batches).  Object pre-blur and loss_simlar take
    the autograd route."""
    if arena is None or grad_accumulation != 1 or not do_step or not torch.is_grad_enabled():
        return False
    if type(loss_fn) is not CombinedLoss or loss_fn.loss_params["loss_simlar"]["state"]:
        return False
    lp = loss_fn.loss_params
    if not (lp["loss_single"]["state"] or lp["loss_poissn"]["state"] or lp["loss_pacbed"]["state"]):
        return False
    if model.obj_preblur_std:
        return False
    if measurements is not None and not isinstance(measurements, MeasurementView):
        return False
    return True


def chunked_step_eligible(model, loss_fn) -> bool:
    """A chunked step needs dL/dI = (scalar of the batch sums) x (per-pixel term): exactly one of loss_single / loss_poissn, no
    loss_pacbed, no detector blur; tilt / thickness gradients are not covered."""
    lp = loss_fn.loss_params
    one = bool(lp["loss_single"]["state"]) != bool(lp["loss_poissn"]["state"])
    prop = (model.opt_obj_tilts.requires_grad and model.tilt_obj) or (model.opt_slice_thickness.requires_grad and model.change_thickness)
    return one and not lp["loss_pacbed"]["state"] and not model.detector_blur_std and not prop


def split_step_eligible(model, loss_fn, B: int) -> bool:
    """What `recon_batch(split=True)` covers: the fused 128^2 kernels (one tile per SM, ragged last wave) and what a chunked step
    needs (one separable data term, no pacbed / detector blur / tilt-thickness gradients)."""
    if not chunked_step_eligible(model, loss_fn) or B < 4:
        return False
    N = model.opt_probe.shape[1]
    return N == 128 and model.kernel_path != _lib.PATH_GENERAL


class _StepIO:
    """What one autograd-free step hands to the C ABI, gathered once: parameter storages, the need mask, the gradient output buffers
    (arena views of the live parameters, scratch for frozen ones), loss configuration, and the ABI calls of a step as methods so that
    the whole-batch, chunked and split variants below differ only in how they sequence them."""

    def __init__(self, model, loss_fn, idx, meas: MeasurementView, cap: int, prop_grads: bool = True):
        self.model, self.lib, self.idx, self.meas = model, _lib.lib(), idx, meas
        self.dev = dev = model.opt_obja.device
        self.B = idx.numel()
        self.obja, self.objp = model.opt_obja.data.contiguous(), model.opt_objp.data.contiguous()
        self.probe = model.opt_probe.data.contiguous()
        tilts, self.dz, shifts = model.opt_obj_tilts.data.contiguous(), model.opt_slice_thickness.data, model.opt_probe_pos_shifts.data.contiguous()
        Z = self.obja.shape[1]
        n_obj = model.opt_obja.requires_grad or model.opt_objp.requires_grad
        n_probe = model.opt_probe.requires_grad
        n_shifts = model.opt_probe_pos_shifts.requires_grad and model.shift_probes
        n_tilts = prop_grads and model.opt_obj_tilts.requires_grad and model.tilt_obj and Z > 1
        n_dz = prop_grads and model.opt_slice_thickness.requires_grad and model.change_thickness and Z > 1
        need_prop = prop_grads and ((model.opt_obj_tilts.requires_grad and model.tilt_obj) or
                                    (model.opt_slice_thickness.requires_grad and model.change_thickness))
        self.need = ((_lib.NEED_OBJ if n_obj else 0) | (_lib.NEED_PROBE if n_probe else 0) | (_lib.NEED_SHIFTS if n_shifts else 0) |
                     (_lib.NEED_TILTS if n_tilts else 0) | (_lib.NEED_DZ if n_dz else 0))
        self.cfg = cfg = model._cfg(stash_fourier=bool(need_prop))
        self.cap = int(cap)                                     # samples per ABI call; the workspaces are laid out for it
        if self.cap != self.B:
            cfg.reserved[0] = self.cap
        self.Hbase = engine.propagator(cfg, self.dz) if model.change_thickness else model.H
        self.tl = tilts if cfg.tilt_mode else None
        self.sh = shifts if cfg.shift_probes else None
        self.lcfg = loss_fn.lcfg()
        self.sparse = bool(self.lcfg.sparse_state)
        if self.sparse and not self.need and model.opt_objp.requires_grad:
            raise RuntimeError("unreachable: objp.requires_grad implies NEED_OBJ")
        ones = getattr(model, "_ones3", None)                   # d(total)/d(term) = 1 for every term: total is their plain sum
        if ones is None or ones.device != dev:
            ones = model._ones3 = torch.ones(3, dtype=torch.float32, device=dev)
        self.ones = ones

        def out(p, wanted):                                     # the arena view of a live parameter, scratch for a frozen one
            if not wanted:
                return None
            return p.grad if p.requires_grad else torch.empty_like(p.data)

        self.g_obja, self.g_objp = out(model.opt_obja, n_obj), out(model.opt_objp, n_obj)
        self.g_probe, self.g_shifts = out(model.opt_probe, n_probe), out(model.opt_probe_pos_shifts, n_shifts)
        self.g_tilts, self.g_dz = out(model.opt_obj_tilts, n_tilts), out(model.opt_slice_thickness, n_dz)
        # losses: [single, poissn, pacbed] and [sparse] land in one (5,) tensor; simlar (slot 4) is off on these paths
        self.losses = torch.zeros(5, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self.pac = torch.empty(2 * cfg.N * cfg.N, dtype=torch.float32, device=dev) if self.lcfg.pacbed_state else None
        if self.sparse:
            self.Ssum = torch.empty(cfg.M, dtype=torch.float64, device=dev)
            self.cover = torch.empty(cfg.Noy * cfg.Nox, dtype=torch.int32, device=dev)

    def workspace(self):
        n = self.lib.ptyb200_workspace_bytes(C.byref(self.cfg), self.cap)
        if n == 0:
            _lib.check(1)
        return torch.empty(n, dtype=torch.uint8, device=self.dev)

    def flagged(self, flags: int):
        c = type(self.cfg).from_buffer_copy(self.cfg)
        c.reserved[4] = flags
        return c

    def _params(self):
        m = self.model
        return (ptr(self.obja), ptr(self.objp), ptr(m.crop_pos), ptr(self.probe), ptr(self.sh), ptr(self.Hbase), ptr(self.tl), ptr(self.dz),
                ptr(m.omode_occu))

    # -- the ABI calls of a step; lo:hi = the samples of this call, everything on the CURRENT stream
    def forward_loss(self, flags, lo, hi, dp, ws):
        """forward with the mode reduction fused with the data losses (the kernel that completes a pattern adds its loss sums)"""
        me = self.meas
        _lib.check(self.lib.ptyb200_forward_loss(C.byref(self.flagged(flags)), ptr(self.idx[lo:hi]), hi - lo, *self._params(), ptr(dp), ptr(ws),
                                                 C.byref(self.lcfg), ptr(me.all), ptr(me.idx[lo:hi]), engine.mref(me.mcfg), ptr(me.padded),
                                                 ptr(self.losses), ptr(self.stats), ptr(self.pac), engine._stream()))

    def loss_grad(self, lo, hi, dp, G, unscaled: bool):
        me = self.meas
        _lib.check(self.lib.ptyb200_loss_grad(C.byref(self.cfg), C.byref(self.lcfg), ptr(dp), ptr(me.all), ptr(me.idx[lo:hi]), hi - lo,
                                              None if unscaled else ptr(self.stats), None if unscaled else ptr(self.pac),
                                              None if unscaled else ptr(self.ones), ptr(G), engine.mref(me.mcfg), ptr(me.padded), engine._stream()))

    def backward(self, flags, lo, hi, G, ws):
        _lib.check(self.lib.ptyb200_backward(C.byref(self.flagged(flags)), ptr(self.idx[lo:hi]), hi - lo, *self._params(), ptr(G), ptr(ws),
                                             ptr(self.g_obja), ptr(self.g_objp), ptr(self.g_probe), ptr(self.g_shifts), ptr(self.g_tilts),
                                             ptr(self.g_dz), self.need, engine._stream()))

    def backward_zero(self, ws):
        _lib.check(self.lib.ptyb200_backward_zero(C.byref(self.cfg), self.cap, ptr(ws), ptr(self.g_probe), ptr(self.g_shifts), self.need,
                                                  engine._stream()))

    def loss_finalize(self):
        _lib.check(self.lib.ptyb200_loss_finalize(C.byref(self.cfg), C.byref(self.lcfg), self.B, ptr(self.stats), ptr(self.pac), ptr(self.losses),
                                                  engine._stream()))

    def finish_scaled(self, flags, ws):
        """accumulators of `ws` -> gradients, times the batch-level factor of the (unscaled) loss gradient"""
        scale = torch.empty(1, dtype=torch.float32, device=self.dev)
        st = engine._stream()
        _lib.check(self.lib.ptyb200_loss_scale(C.byref(self.cfg), C.byref(self.lcfg), self.B, ptr(self.stats), ptr(self.ones), ptr(scale), st))
        _lib.check(self.lib.ptyb200_backward_finish(C.byref(self.flagged(flags)), self.cap, ptr(self.obja), ptr(self.objp), ptr(ws), ptr(self.g_obja),
                                                    ptr(self.g_objp), ptr(self.g_probe), ptr(self.g_shifts), self.need, ptr(scale), st))

    def sparse_forward(self):
        m = self.model
        _lib.check(self.lib.ptyb200_sparse_forward(C.byref(self.cfg), C.byref(self.lcfg), ptr(self.objp), ptr(m.crop_pos), ptr(self.idx), self.B,
                                                   ptr(m.omode_occu), C.c_void_p(self.losses.data_ptr() + 12), ptr(self.Ssum), ptr(self.cover),
                                                   engine._stream()))

    def sparse_grad(self):
        m = self.model
        if self.need and m.opt_objp.requires_grad:
            _lib.check(self.lib.ptyb200_sparse_grad(C.byref(self.cfg), C.byref(self.lcfg), ptr(self.objp), ptr(m.crop_pos), ptr(self.idx), self.B,
                                                    ptr(m.omode_occu), ptr(self.Ssum), ptr(self.ones), ptr(self.cover), ptr(m.opt_objp.grad),
                                                    engine._stream()))

    def aux_work(self, workspaces, arena, zero_arena: bool):
        """Everything that does not depend on the multislice kernels, for a side stream: zeroing of the gradient arena and of the
        adjoint's accumulators, loss_sparse (object and batch indices only) and its gradient (ADDED into the zeroed arena; the
        adjoint's completion then adds the object gradients on top: PTYB200_ACC_ADD_OBJ)."""
        if zero_arena:
            arena.zero()
        if self.need:
            for w in workspaces:
                self.backward_zero(w)
        if self.sparse:
            self.sparse_forward()
            self.sparse_grad()


def _direct_grads(model, loss_fn, idx, meas: MeasurementView, arena: GradArena, zero_arena: bool = False):
    """Forward, loss, loss gradient and adjoint through the C ABI with the gradient tensors of the arena as the kernels' output
    buffers -- what ``model(idx)`` -> ``loss_fn`` -> ``backward()`` computes (engine.MultisliceFunction / DataLossFunction /
    SparseLossFunction), minus the autograd graph, its ~30 small elementwise launches per step and the accumulate-into-.grad copies.
    Returns the five loss terms as one device tensor.

    A step is a chain of ~30 launches of which two matter.  Everything that does not depend on the multislice kernels
    (`_StepIO.aux_work`, and -- once the forward is through -- the loss scalars) runs on a side stream while the forward holds the
    caller's stream: the fused kernels use the whole register file, so these short kernels execute in its last, partial wave."""
    io = _StepIO(model, loss_fn, idx, meas, cap=idx.numel())
    B, cfg, dev = io.B, io.cfg, io.dev
    ws = io.workspace()
    dp = torch.empty((B, cfg.N, cfg.N), dtype=torch.float32, device=dev)
    blur = model.detector_blur_std
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        io.aux_work([ws], arena, zero_arena)
    if not blur:
        # (loss_pacbed: the kernel that forms the loss scalars also completes a sum the loss gradient reads, so it stays in line)
        defer_final = not io.lcfg.pacbed_state
        io.forward_loss(_lib.ACC_NO_LOSS_FINAL if defer_final else 0, 0, B, dp, ws)
        if defer_final:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                io.loss_finalize()
    else:
        st = engine._stream()
        me = io.meas
        _lib.check(io.lib.ptyb200_forward(C.byref(cfg), ptr(idx), B, *io._params(), ptr(dp), ptr(ws), st))
        dp = engine._blur5(dp, float(blur), 0)                # detector blur (models.py:379-380): native 5x5 kernel on the intensities
        _lib.check(io.lib.ptyb200_loss_forward(C.byref(cfg), C.byref(io.lcfg), ptr(dp), ptr(me.all), ptr(me.idx), B, ptr(io.losses), ptr(io.stats),
                                               ptr(io.pac), engine.mref(me.mcfg), ptr(me.padded), st))
    if io.need:
        G = torch.empty_like(dp)
        io.loss_grad(0, B, dp, G, unscaled=False)
        if blur:
            G = engine._blur5(G, float(blur), 1)              # adjoint of the blur
        cur.wait_stream(side)                                 # accumulators zeroed, sparse term in place
        io.backward(_lib.ACC_KEEP_GRADS | _lib.ACC_ADD_OBJ, 0, B, G, ws)
    else:
        cur.wait_stream(side)
    return io.losses


def _direct_grads_chunked(model, loss_fn, idx, meas: MeasurementView, arena: GradArena, chunk: int, zero_arena: bool = False):
    """`_direct_grads` for a batch that is processed `chunk` samples at a time so that only one chunk's wave stash is alive
    (memory must not grow with the per-GPU batch; the strong-scaling runs put 2 048 patterns of C4 on one GPU, whose
    stash alone would be 206 GB).  The result is the gradient of the loss of the WHOLE batch, not a sum of per-chunk losses: every
    chunk adds its loss sums and runs its adjoint on the unscaled loss gradient into shared accumulators; the batch-level factor is
    applied once, when the accumulators are turned into gradients (include/ptyrad_b200.h, PTYB200_ACC_*)."""
    if not chunked_step_eligible(model, loss_fn):
        raise ValueError("a chunked step needs exactly one of loss_single / loss_poissn, no loss_pacbed, no detector blur and no "
                         "tilt / thickness gradients")
    io = _StepIO(model, loss_fn, idx, meas, cap=chunk, prop_grads=False)
    B, cfg, dev = io.B, io.cfg, io.dev
    ws = io.workspace()
    dp = torch.empty((chunk, cfg.N, cfg.N), dtype=torch.float32, device=dev)
    G = torch.empty_like(dp)
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        io.aux_work([ws], arena, zero_arena)
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        io.forward_loss(_lib.ACC_KEEP_STATS | _lib.ACC_NO_LOSS_FINAL, lo, hi, dp, ws)      # `stats` starts zeroed (_StepIO)
        if io.need:
            io.loss_grad(lo, hi, dp, G, unscaled=True)
            if lo == 0:
                cur.wait_stream(side)                         # accumulators zeroed
            io.backward(_lib.ACC_KEEP_GRADS | _lib.ACC_NO_FINISH, lo, hi, G, ws)
    cur.wait_stream(side)
    io.loss_finalize()
    if io.need:
        io.finish_scaled(_lib.ACC_ADD_OBJ, ws)                # the arena was zeroed and holds the loss_sparse term
    return io.losses


def _direct_grads_split(model, loss_fn, idx, meas: MeasurementView, arena: GradArena, zero_arena: bool = False):
    """`_direct_grads` with the batch cut in two halves that run forward -> unscaled loss gradient -> adjoint on TWO streams, each into
    its own workspace; the accumulators are added and completed once, with the batch-level factor (the machinery of the chunked
    step).  Why: the fused kernels run one tile per SM, so a kernel over T tiles takes ceil(T / 148) waves -- C2's 1536 tiles are
    10.38 waves, and the last, 38 %-full wave of the forward and of the adjoint costs 0.14 ms of a 2.54 ms step.  Four half-size
    kernels on two streams could fill each other's partial waves.  MEASURED (C2, B200): 2.819 ms against 2.534 ms unsplit --
    concurrent forward (stash writes) and adjoint (stash reads) kernels cost more than the tails; kept as an option
    (`recon_batch(split=True)`), off by default.  Loss and gradients are those of the whole batch."""
    if not chunked_step_eligible(model, loss_fn):
        raise ValueError("a split step needs exactly one of loss_single / loss_poissn, no loss_pacbed, no detector blur and no "
                         "tilt / thickness gradients")
    B = idx.numel()
    cuts = [(0, B // 2), (B // 2, B)]
    io = _StepIO(model, loss_fn, idx, meas, cap=max(hi - lo for lo, hi in cuts), prop_grads=False)
    cfg, dev = io.cfg, io.dev
    ws = [io.workspace() for _ in cuts]
    dp = torch.empty((B, cfg.N, cfg.N), dtype=torch.float32, device=dev)
    G = torch.empty_like(dp)
    cur = torch.cuda.current_stream(dev)
    aux, half = _side_stream(dev), _side_stream(dev, "half")
    aux.wait_stream(cur)
    half.wait_stream(cur)
    with torch.cuda.stream(aux):
        io.aux_work(ws, arena, zero_arena)
    for (lo, hi), stream, w in zip(cuts, (cur, half), ws):
        with torch.cuda.stream(stream):
            io.forward_loss(_lib.ACC_KEEP_STATS | _lib.ACC_NO_LOSS_FINAL, lo, hi, dp[lo:hi], w)
            if io.need:
                io.loss_grad(lo, hi, dp[lo:hi], G[lo:hi], unscaled=True)
                stream.wait_stream(aux)                       # accumulators zeroed
                io.backward(_lib.ACC_KEEP_GRADS | _lib.ACC_NO_FINISH, lo, hi, G[lo:hi], w)
    cur.wait_stream(half)
    cur.wait_stream(aux)
    io.loss_finalize()
    if io.need:
        _lib.check(io.lib.ptyb200_accumulators_add(C.byref(cfg), io.cap, ptr(ws[0]), ptr(ws[1]), io.need, engine._stream()))
        io.finish_scaled(_lib.ACC_ADD_OBJ, ws[0])
    return io.losses


def recon_batch(model, loss_fn, optimizer, indices, arena: GradArena | None = None, world: int = 1,
                grad_accumulation: int = 1, do_step: bool = True, measurements=None, direct: bool | None = None,
                first_of_group: bool = True, chunk: int = 0, split: bool | None = None):
    """One batch: (zero grads), forward, loss, backward, (all-reduce), optimizer step.  Returns the 5 loss terms as a device
    tensor (no host sync).  `direct` (default: whenever eligible) takes the autograd-free route of `_direct_grads`.

    Gradient accumulation (reconstruction.py:750-760): pass `grad_accumulation` = group size, `first_of_group` = True only for the
    first batch of a group (gradients are zeroed there and nowhere else) and `do_step` = True only for the last one.

    `split=True` runs the batch as two halves on two streams (`_direct_grads_split`; same loss and gradients).  Off by default: it
    was built to let the partial last waves of the forward and adjoint kernels fill each other, and measured SLOWER at C2 (2.534 ->
    2.819 ms, profiles/r02/ab_split_step.txt): the write-heavy forward and the read-heavy adjoint slow each other down by more than
    the two 38 %-full waves cost.
    `chunk` > 0 bounds the memory of the step: the batch runs through the kernels `chunk` samples at a time (one chunk's wave stash
    alive) and still yields the loss and gradients of the whole batch (`_direct_grads_chunked`; direct route only)."""
    if world > 1 and arena is None:
        raise RuntimeError("multi-GPU steps need a GradArena (the gradient exchange is one all-reduce over its flat buffer)")
    if direct is None:
        direct = direct_step_eligible(model, loss_fn, arena, grad_accumulation, do_step, measurements)
    elif direct and not direct_step_eligible(model, loss_fn, arena, grad_accumulation, do_step, measurements):
        raise ValueError("this configuration needs the autograd path (direct=False)")
    if direct:
        arena.attach()
        idx = model._index_tensor(indices)
        meas = measurements if measurements is not None else MeasurementView(model.measurements, idx, model)
        if chunk and idx.numel() > chunk:
            losses = _direct_grads_chunked(model, loss_fn, idx, meas, arena, int(chunk), zero_arena=True)
        elif split:
            losses = _direct_grads_split(model, loss_fn, idx, meas, arena, zero_arena=True)
        else:
            losses = _direct_grads(model, loss_fn, idx, meas, arena, zero_arena=True)     # zeroes the arena itself, on its side stream
        if world > 1:
            arena.allreduce(world)
        optimizer.step()
        return losses
    if chunk and len(indices) > chunk:
        raise ValueError("chunked steps exist on the direct (autograd-free) route only")
    if arena is not None:
        arena.attach()
        if first_of_group:
            arena.zero()
    elif first_of_group:
        optimizer.zero_grad()
    dp = model(indices)
    idx = model._index_tensor(indices)
    if measurements is not None:
        meas = measurements
    elif type(loss_fn) is CombinedLoss:
        meas = MeasurementView(model.measurements, idx, model)   # read rows in place (pad / resample inside the loss kernels)
    else:
        meas = model.get_measurements(idx)                       # a foreign loss (e.g. the reference's) gets a real tensor
    total, losses = loss_fn(dp, meas, model._current_object_patches, model.omode_occu)
    (total / grad_accumulation if grad_accumulation != 1 else total).backward()
    if do_step:
        if world > 1:
            arena.allreduce(world)
        optimizer.step()
    model.clear_cache()
    return torch.stack([l.detach().reshape(()) for l in losses])


class GraphedStep:
    """`recon_batch` captured into a CUDA graph and replayed per batch (sync-free, launch-free step).

    At the reference's default batch size (32) the per-batch work on a B200 is a few tens of microseconds while ~40 kernel
    launches plus autograd bookkeeping cost ~1 ms of host time; replaying a graph removes that.  One instance handles one batch
    size (``make_batches`` yields at most two distinct sizes).  The scan indices are the only per-step input: they are copied
    into a static device buffer.  Parameter storage must stay put between replays; constraints that rebind ``opt_*.data``
    (constraints.py:38-224) are handled by ``_rebind`` (values are copied back into the captured storage).

    A captured graph bakes in WHICH tensors are trainable (the adjoint's need mask, the arena views the kernels write, the
    optimiser's tensor list), while the reference re-evaluates ``requires_grad`` every iteration (``start_iter``,
    reconstruction.py:783-790).  Graphs are therefore keyed by the requires_grad signature: a signature seen for the first time
    is captured on the spot, later ones replay.  Capturing runs the step a few times; parameters AND optimiser state are
    snapshotted before and restored afterwards, so building (or re-capturing) a GraphedStep mid-run, or after
    ``optimizer.load_state_dict`` (reconstruction.py:356-364), leaves the Adam moments and step counters exactly as they were.
    """

    def __init__(self, model, loss_fn, optimizer, arena: GradArena, batch_size: int, grad_accumulation: int = 1, warmup: int = 2,
                 world: int = 1, stream_measurements: bool = False, chunk: int = 0, split: bool = False):
        if grad_accumulation != 1:
            raise ValueError("GraphedStep captures a whole step (zero, forward, adjoint, exchange, optimizer): use the eager "
                             "recon_batch / recon_step for gradient accumulation")
        self.model, self.loss_fn, self.opt, self.arena = model, loss_fn, optimizer, arena
        self.B = int(batch_size)
        self.world, self.warmup, self.chunk, self.split = world, warmup, int(chunk), bool(split)
        dev = model.opt_obja.device
        self.idx = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.params = list(model.optimizable_tensors.values())
        # stream_measurements: this batch's patterns are copied into `self.meas` before every replay (dataset in host memory)
        self.meas = None
        self._mv = None
        if stream_measurements:
            Nm = model.measurements.shape[-1]
            self.meas = torch.zeros((self.B, Nm, Nm), dtype=torch.float32, device=dev)
            self._mv = MeasurementView(self.meas, torch.arange(self.B, device=dev), model)   # read by address: lives with the graphs
            # double buffering for `prefetch`: the next batch is copied host -> device on a side stream while this one runs
            self._next_meas = torch.zeros_like(self.meas)
            self._next_idx = torch.zeros_like(self.idx)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_done = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        self._graphs = {}
        self._ptrs = [p.data_ptr() for p in self.params]
        self._store = [p.data for p in self.params]
        self._capture(self.signature())

    def signature(self):
        return tuple(bool(p.requires_grad) for p in self.params)

    # -- optimiser state: values are saved and written back INTO the same tensors (the graphs hold their addresses)
    def _snapshot_opt(self):
        snap = {}
        for group in self.opt.param_groups:
            for p in group["params"]:
                st = self.opt.state.get(p)
                snap[p] = None if not st else {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return snap

    def _restore_opt(self, snap):
        with torch.no_grad():
            for p, old in snap.items():
                st = self.opt.state.get(p)
                if not st:
                    continue
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is None or k not in old:
                            v.zero_()                     # created by the warm-up steps: back to its initial value
                        elif v.data_ptr() != old[k].data_ptr():
                            v.copy_(old[k])
                    elif old is not None and k in old:
                        st[k] = old[k]

    def _capture(self, sig):
        model, dev = self.model, self.idx.device
        self._rebind()
        snap_p = [p.detach().clone() for p in self.params]
        snap_o = self._snapshot_opt()
        run = lambda: recon_batch(model, self.loss_fn, self.opt, self.idx, self.arena, self.world, measurements=self._mv, chunk=self.chunk, split=self.split)
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for i in range(self.warmup):
                run()
        torch.cuda.current_stream(dev).wait_stream(stream)
        torch.cuda.synchronize(dev)
        # the warm-up may have created optimiser state; restore values before capture so that nothing captured depends on them
        self._restore(snap_p, snap_o)
        # the graph allocates from its own pool: hand the warm-up's cached blocks (a wave stash can be > 100 GB) back first
        torch.cuda.empty_cache()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            losses = run()
        self._restore(snap_p, snap_o)
        self._graphs[sig] = (graph, losses)

    def _restore(self, snap_p, snap_o):
        with torch.no_grad():
            for p, s in zip(self.params, snap_p):
                p.data.copy_(s)
        self._restore_opt(snap_o)

    def _rebind(self):
        for i, p in enumerate(self.params):
            if p.data_ptr() != self._ptrs[i]:
                with torch.no_grad():
                    self._store[i].copy_(p.data)
                p.data = self._store[i]

    def _replay(self):
        sig = self.signature()
        if sig not in self._graphs:
            self._capture(sig)                            # first time this set of trainable tensors is seen (start_iter)
        graph, losses = self._graphs[sig]
        self.arena.attach()
        graph.replay()
        return losses

    @property
    def losses(self):
        return self._graphs[self.signature()][1]

    def prefetch(self, indices, measurements):
        """Start the host -> device copy of the NEXT batch (pinned tensors) on a side stream; pair with `step_prefetched`."""
        self._copy_stream.wait_event(self._consumed)
        with torch.cuda.stream(self._copy_stream):
            self._next_idx.copy_(indices, non_blocking=True)
            self._next_meas.copy_(measurements, non_blocking=True)
            self._copy_done.record(self._copy_stream)

    def step_prefetched(self):
        """Run the step on the batch handed to the last `prefetch` call."""
        self._rebind()
        cur = torch.cuda.current_stream(self.idx.device)
        cur.wait_event(self._copy_done)
        self.idx.copy_(self._next_idx, non_blocking=True)
        self.meas.copy_(self._next_meas, non_blocking=True)
        self._consumed.record(cur)
        return self._replay()

    def __call__(self, indices, measurements=None):
        self._rebind()
        if self.meas is not None:
            self.meas.copy_(measurements, non_blocking=True)
        if not torch.is_tensor(indices):
            indices = torch.as_tensor(np.asarray(indices, dtype=np.int64))
        if indices.numel() != self.B:
            raise ValueError(f"this graph was captured for batch size {self.B}, got {indices.numel()}")
        self.idx.copy_(indices, non_blocking=True)
        return self._replay()


def toggle_grad_requires(model, niter: int):
    """requires_grad per optimisable tensor from its start_iter (reference reconstruction.py:783-790)."""
    for name, start in model.start_iter.items():
        model.optimizable_tensors[name].requires_grad = start is not None and niter >= start


def _lbfgs_iteration(batches, grad_accumulation, model, optimizer, loss_fn, arena, world, rank):
    """The LBFGS branch of the reference's ``recon_step`` (reconstruction.py:697-735): the batches are shuffled and cut into groups
    of `grad_accumulation`; ``optimizer.step(closure)`` runs once per group, the closure evaluating forward + loss over the whole
    group (several forwards before one backward: every forward keeps its own workspace) and the mean loss of the group driving
    the line search; one extra closure evaluation at the end yields the logged loss terms.  With `world` > 1 the gradients AND
    the closure's loss are averaged over the ranks, so that every rank's line search takes the same decisions (the reference
    leaves the loss rank-local).  Returns the five loss terms of that last evaluation as a device tensor."""
    nb = len(batches)
    order = np.arange(nb)
    np.random.shuffle(order)
    groups = np.array_split(order, max(nb // max(int(grad_accumulation), 1), 1))
    state = {}

    def closure(group):
        if arena is not None:
            arena.attach()
            arena.zero()
        else:
            optimizer.zero_grad()
        total = 0
        for bi in group:
            mine = shard_indices(batches[bi], rank, world) if world > 1 else batches[bi]
            dp = model(mine)
            idx = model._index_tensor(mine)
            meas = MeasurementView(model.measurements, idx, model) if type(loss_fn) is CombinedLoss else model.get_measurements(idx)
            loss_batch, losses = loss_fn(dp, meas, model._current_object_patches, model.omode_occu)
            total = total + loss_batch
        total = total / len(group)
        total.backward()
        if world > 1:
            arena.allreduce(world)
            total = total.detach().clone()
            dist.all_reduce(total)
            total /= world
        state["losses"] = losses
        return total

    for group in groups:
        optimizer.step(lambda: closure(group))
    with torch.enable_grad():
        closure(groups[-1])                                    # logging only, like the reference's extra evaluation
    if arena is not None:
        arena.zero()
    else:
        optimizer.zero_grad()
    model.clear_cache()
    return torch.stack([l.detach().reshape(()) for l in state["losses"]])


def recon_step(batches, grad_accumulation, model, optimizer, loss_fn, constraint_fn, niter, verbose=True, arena: GradArena | None = None,
               world: int = 1, rank: int = 0, graphed: dict | None = None):
    """One iteration over all batches: the non-LBFGS branch of the reference's ``recon_step`` (reconstruction.py:658-781) with the
    same bookkeeping (``batch_losses`` dict in ``loss_params`` key order, ``model.loss_iters / iter_times / dz_iters /
    avg_tilt_iters``, constraints once per iteration) but ONE host synchronisation per iteration instead of three per batch:
    the per-batch losses stay on the device until the end of the iteration.

    `graphed`: optional {batch_size: GraphedStep}; batches of a captured size are replayed as CUDA graphs.  With `world` > 1
    every rank passes the same global batches and takes its own slice (split_batches=True semantics).  A ``torch.optim.LBFGS``
    optimizer takes the closure loop of the reference (``_lbfgs_iteration``, reconstruction.py:697-735).
    """
    import time
    dev = model.opt_obja.device
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    if world > 1 and arena is None:
        raise RuntimeError("multi-GPU iterations need a GradArena (the gradient exchange is one all-reduce over its flat buffer)")
    toggle_grad_requires(model, niter)
    if isinstance(optimizer, torch.optim.LBFGS):
        per_batch = [_lbfgs_iteration(batches, grad_accumulation, model, optimizer, loss_fn, arena, world, rank)]
        batches = []                                           # the mini-batch loop below has nothing left to do
    elif arena is not None:
        arena.attach()
        arena.zero()
        per_batch = []
    else:
        optimizer.zero_grad()
        per_batch = []
    nb = len(batches)
    for bi, batch in enumerate(batches):
        mine = shard_indices(batch, rank, world) if world > 1 else batch
        last_of_group = (bi + 1) % grad_accumulation == 0 or (bi + 1) == nb
        g = graphed.get(len(mine)) if (graphed and grad_accumulation == 1) else None
        if g is not None:
            per_batch.append(g(mine).clone())
            continue
        if grad_accumulation == 1 and direct_step_eligible(model, loss_fn, arena, 1, True, None):
            per_batch.append(recon_batch(model, loss_fn, optimizer, mine, arena, world))      # autograd-free step
            continue
        dp = model(mine)
        idx = model._index_tensor(mine)
        meas = MeasurementView(model.measurements, idx, model) if type(loss_fn) is CombinedLoss else model.get_measurements(idx)
        total, losses = loss_fn(dp, meas, model._current_object_patches, model.omode_occu)
        (total / grad_accumulation).backward()
        if last_of_group:
            if world > 1:
                arena.allreduce(world)
            optimizer.step()
            if arena is not None:
                arena.zero()
            else:
                optimizer.zero_grad()
        model.clear_cache()
        per_batch.append(torch.stack([l.detach().reshape(()) for l in losses]))
    if constraint_fn is not None:
        constraint_fn(model, niter)
    host = torch.stack(per_batch).cpu().numpy()                  # the one synchronisation of the iteration
    iter_t = time.perf_counter() - t0
    names = list(loss_fn.loss_params.keys())
    batch_losses = {n: [host[i, j] for i in range(host.shape[0])] for j, n in enumerate(names)}
    loss_iter = float(sum(np.mean(v) for v in batch_losses.values()))
    if verbose and rank == 0:
        print(f"Iter: {niter}, Total Loss: {loss_iter:.4f}, " + ", ".join(f"{n}: {np.mean(v):.4f}" for n, v in batch_losses.items()) +
              f", in {iter_t:.3f} sec")
    model.loss_iters.append((niter, loss_iter))
    model.iter_times.append(iter_t)
    model.dz_iters.append((niter, model.opt_slice_thickness.detach().cpu().numpy()))
    model.avg_tilt_iters.append((niter, model.opt_obj_tilts.detach().mean(0).cpu().numpy()))
    return batch_losses
