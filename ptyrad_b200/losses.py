"""Drop-in ``CombinedLoss`` (reference ``src/ptyrad/losses.py:17-155``) backed by the CUDA library.

Same constructor, same call signature ``loss_fn(model_DP, measured_DP, object_patches, omode_occu) -> (total, [single,
poissn, pacbed, sparse, simlar])``, same key order of ``loss_params`` (it defines the log order,
reconstruction.py:687,769); inactive terms are 0-dim zeros.  The three data terms are one native reduction +
one native gradient kernel (``engine.DataLossFunction``); ``loss_sparse`` is evaluated on the batch ROIs natively when
``object_patches`` is the ``LazyPatches`` handle our model hands out, and with plain tensor ops when a caller passes a
materialised patch tensor.  ``loss_simlar`` (mixed-object regulariser) is native too on those handles (ROI planes
blurred straight from the dense object, area interpolation + std + mean in one kernel).
"""
from __future__ import annotations

import torch

from . import engine
import math

from .models import LazyPatches, PatchPlanes, gaussian_blur5


class CombinedLoss(torch.nn.Module):
    def __init__(self, loss_params, device="cuda"):
        super().__init__()
        self.device = device
        self.loss_params = loss_params

    def lcfg(self):
        """Kernel-side view of `loss_params`, rebuilt on every call: the reference reads the dict live (losses.py:41-47), so a
        caller that edits weights / states between iterations (hypertune, notebooks) is honoured."""
        return engine.make_loss_cfg(self.loss_params)

    # data terms -------------------------------------------------------------------------------------
    def _data_losses(self, model_DP, measured_DP):
        lp = self.loss_params
        if not (lp["loss_single"]["state"] or lp["loss_poissn"]["state"] or lp["loss_pacbed"]["state"]):
            z = torch.zeros((), dtype=torch.float32, device=model_DP.device)
            return z, z.clone(), z.clone()
        B, N = model_DP.shape[0], model_DP.shape[-1]
        mcfg = padded = None
        if isinstance(measured_DP, MeasurementView):
            meas_all, idx, mcfg, padded = measured_DP.all, measured_DP.idx, measured_DP.mcfg, measured_DP.padded
        else:
            meas_all = measured_DP.contiguous().float()
            idx = torch.arange(B, dtype=torch.int64, device=model_DP.device)
        cfg = engine.make_cfg(N, 1, 1, 1, N, N, meas_all.shape[0], 0, 0, 0, 1.0, 1.0)
        l3 = engine.DataLossFunction.apply(model_DP, meas_all, idx, cfg, self.lcfg(), mcfg, padded)
        return l3[0], l3[1], l3[2]

    def get_loss_single(self, model_DP, measured_DP):
        return self._data_losses(model_DP, measured_DP)[0]

    def get_loss_poissn(self, model_DP, measured_DP):
        return self._data_losses(model_DP, measured_DP)[1]

    def get_loss_pacbed(self, model_DP, measured_DP):
        return self._data_losses(model_DP, measured_DP)[2]

    # object regularisers ----------------------------------------------------------------------------
    def get_loss_sparse(self, objp_patches, omode_occu):
        sp = self.loss_params["loss_sparse"]
        if not sp["state"]:
            return torch.zeros((), dtype=torch.float32, device=omode_occu.device)
        if isinstance(objp_patches, LazyPatches):
            m = objp_patches.model
            return engine.SparseLossFunction.apply(m.opt_objp, m.crop_pos, objp_patches.idx, omode_occu, m._cfg(False), self.lcfg())
        n = sp["ln_order"]
        return sp["weight"] * (objp_patches.abs().pow(n).mean(dim=(0, 2, 3, 4)).pow(1.0 / n) * omode_occu).sum()

    def get_loss_simlar(self, object_patches, omode_occu):
        """losses.py:106-141.  With the handles our model hands out (LazyPatches / PatchPlanes) everything is native: ROI planes
        gathered and 5x5-blurred straight from the dense object (engine.RoiBlurFunction), then area interpolation + std over the
        object modes + mean in one kernel (engine.SimlarFunction).  A foreign, materialised patch tensor takes plain tensor ops."""
        s = self.loss_params["loss_simlar"]
        if not s["state"]:
            return torch.zeros((), dtype=torch.float32, device=omode_occu.device)
        sf = s.get("scale_factor")
        blur = s.get("blur_std") or 0
        native = isinstance(object_patches, (LazyPatches, PatchPlanes)) and omode_occu.is_cuda and 2 <= omode_occu.numel() <= 8
        if native:
            if isinstance(object_patches, LazyPatches):
                m = object_patches.model
                cfg = m._cfg(False)
                a, p = engine.RoiBlurFunction.apply(m.opt_obja, m.opt_objp, object_patches.idx, m.crop_pos, cfg, float(blur))
            else:                                          # pre-blurred planes (obj_preblur_std): loss_simlar's own blur on top
                a, p = object_patches.a, object_patches.p
                if blur:
                    a, p = gaussian_blur5(a, blur), gaussian_blur5(p, blur)
                B_, M_, Z_, N_ = a.shape[0], a.shape[1], a.shape[2], a.shape[-1]
                cfg = engine.make_cfg(N_, 1, M_, Z_, N_, N_, B_, 0, 0, 0, 1.0, 1.0)
            f = tuple(sf) if sf is not None else (1, 1, 1)
            dims = (int(math.floor(a.shape[2] * f[0])), int(math.floor(a.shape[3] * f[1])), int(math.floor(a.shape[4] * f[2])))
            tot = torch.zeros((), dtype=torch.float32, device=omode_occu.device)
            for name, x in (("amplitude", a), ("phase", p)):
                if s["obj_type"] in (name, "both"):
                    tot = tot + engine.SimlarFunction.apply(x, omode_occu, cfg, dims, float(s["weight"]))
            return tot
        if isinstance(object_patches, (LazyPatches, PatchPlanes)):
            object_patches = object_patches.materialize()
        tot = torch.zeros((), dtype=torch.float32, device=omode_occu.device)
        for name, x in (("amplitude", object_patches[..., 0]), ("phase", object_patches[..., 1])):
            if s["obj_type"] not in (name, "both"):
                continue
            if blur:
                x = gaussian_blur5(x, blur)
            if sf is not None and any(f != 1 for f in sf):
                x = torch.nn.functional.interpolate(x, scale_factor=tuple(sf), mode="area")
            tot = tot + (x * omode_occu[:, None, None, None]).std(1).mean()
        return s["weight"] * tot

    def forward(self, model_DP, measured_DP, object_patches, omode_occu):
        single, poissn, pacbed = self._data_losses(model_DP, measured_DP)
        objp = object_patches if isinstance(object_patches, LazyPatches) else object_patches[..., 1]      # PatchPlanes: the phase plane
        losses = [single, poissn, pacbed, self.get_loss_sparse(objp, omode_occu), self.get_loss_simlar(object_patches, omode_occu)]
        return sum(losses), losses


class MeasurementView:
    """(all measurements, batch indices) pair: lets the native loss read rows ``idx`` of the stored (Ntot,H,W) array in place
    instead of a gathered copy.  With `model` given, the model's "on-the-fly" padding / bilinear resampling options
    (models.py:392-409) travel along and are evaluated INSIDE the loss kernels (nothing is materialised).
    ``PtychoAD.get_measurements`` keeps returning a real tensor for compatibility; the fast step
    (``ptyrad_b200.step.recon_batch``) passes this view."""

    def __init__(self, all_meas, idx, model=None):
        self.all, self.idx = all_meas, idx
        self.mcfg, self.padded = (None, None) if model is None else model._meas_cfg(all_meas)
        if model is not None and all_meas is model.measurements:
            self.idx = model.meas_rows(idx)                  # a rank that holds only its shard: scan index -> local row
