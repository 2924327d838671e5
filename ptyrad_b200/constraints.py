"""Per-iteration object constraints on the B200 (reference ``src/ptyrad/constraints.py:83-114,165-208``).

``CombinedConstraint(constraint_params, device, verbose, fallback=None)`` has the reference's constructor / call signature
(``constraint_fn(model, niter)``, called once per iteration at reconstruction.py:774) and the reference's ORDER of operations
(constraints.py:227-246).  The constraints that are pure streaming passes over the object run in the CUDA library, IN PLACE:

* ``obj_rblur``   separable Gaussian, reflect padding (torchvision ``gaussian_blur``)        -> ``ptyb200_blur_axis`` x then y
* ``obj_zblur``   Gaussian along z, replicate padding (``gaussian_blur_1d``)                 -> ``ptyb200_blur_axis`` z
* ``mirrored_amp`` -> ``obja_thresh`` -> ``objp_postiv``                                     -> ONE launch, ``ptyb200_object_constraints``

In place means no ``.data`` re-binding: parameter storage stays where a captured CUDA graph (``GraphedStep``) and the optimiser
state expect it, and an iteration's worth of constraints costs two to four passes over the object instead of ~15 temporaries.
(Folding these passes into the last Adam launch of the iteration was considered and dropped: one pass over the object per
ITERATION is ~1e-4 of the iteration's memory traffic.)

Everything else -- ``ortho_pmode``, ``probe_mask_k``, ``fix_probe_int``, ``kr_filter``, ``kz_filter``, ``complex_ratio``,
``tilt_smooth`` (FFT / eigen-decomposition based, off the hot path) -- is delegated, at its place in the sequence, to `fallback`,
an instance of the reference's own ``CombinedConstraint`` built from the same ``constraint_params``; requesting one of them without
a fallback raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ptr

_NATIVE = ("obj_rblur", "obj_zblur", "mirrored_amp", "obja_thresh", "objp_postiv")
_ORDER = ("ortho_pmode", "probe_mask_k", "fix_probe_int", "obj_rblur", "obj_zblur", "kr_filter", "kz_filter", "complex_ratio",
          "mirrored_amp", "obja_thresh", "objp_postiv", "tilt_smooth")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class CombinedConstraint(torch.nn.Module):
    def __init__(self, constraint_params, device="cuda", verbose=True, fallback=None):
        super().__init__()
        self.device, self.constraint_params, self.verbose, self.fallback = device, constraint_params, verbose, fallback
        self._tmp = None

    # ------------------------------------------------------------------------------------------------
    def _active(self, name, niter):
        p = self.constraint_params.get(name)
        f = None if p is None else p.get("freq")
        return f is not None and niter % f == 0

    def _scratch(self, like):
        if self._tmp is None or self._tmp.shape != like.shape or self._tmp.device != like.device:
            self._tmp = torch.empty_like(like)
        return self._tmp

    def _blur(self, t, axis, ksize, sigma, pad_mode):
        """In-place 1-D Gaussian along `axis` of the contiguous (M,Z,Y,X) tensor `t` (via one scratch copy)."""
        M, Z, Y, X = t.shape
        outer, L, inner = {"x": (M * Z * Y, X, 1), "y": (M * Z, Y, X), "z": (M, Z, Y * X)}[axis]
        tmp = self._scratch(t)
        _lib.check(_lib.lib().ptyb200_blur_axis(ptr(t), ptr(tmp), outer, L, inner, int(ksize), float(sigma), pad_mode, _stream()))
        return tmp

    def _targets(self, model, obj_type):
        out = []
        if obj_type in ("amplitude", "both"):
            out.append(model.opt_obja)
        if obj_type in ("phase", "both"):
            out.append(model.opt_objp)
        return out

    def apply_obj_rblur(self, model, niter):
        p = self.constraint_params["obj_rblur"]
        if not self._active("obj_rblur", niter) or p["std"] == 0:
            return
        ks = p["kernel_size"]
        ky, kx = (ks, ks) if isinstance(ks, int) else (ks[0], ks[1])                   # torchvision: [kx, ky] = kernel_size
        for t in self._targets(model, p["obj_type"]):
            d = self._contig(t)
            tmp = self._blur(d, "x", kx, p["std"], 0)
            M, Z, Y, X = d.shape
            _lib.check(_lib.lib().ptyb200_blur_axis(ptr(tmp), ptr(d), M * Z, Y, X, int(ky), float(p["std"]), 0, _stream()))

    def apply_obj_zblur(self, model, niter):
        p = self.constraint_params["obj_zblur"]
        if not self._active("obj_zblur", niter) or p["std"] == 0:
            return
        for t in self._targets(model, p["obj_type"]):
            d = self._contig(t)
            d.copy_(self._blur(d, "z", p["kernel_size"], p["std"], 1))

    def apply_voxel_chain(self, model, niter):
        cp = self.constraint_params
        oc = _lib.ObjConstraints()
        if self._active("mirrored_amp", niter):
            m = cp["mirrored_amp"]
            oc.mirrored_on, oc.mirrored_relax, oc.mirrored_scale, oc.mirrored_power = 1, float(m["relax"]), float(m["scale"]), float(m["power"])
        if self._active("obja_thresh", niter):
            m = cp["obja_thresh"]
            oc.thresh_on, oc.thresh_relax, oc.thresh_lo, oc.thresh_hi = 1, float(m["relax"]), float(m["thresh"][0]), float(m["thresh"][1])
        if self._active("objp_postiv", niter):
            m = cp["objp_postiv"]
            oc.postiv_on, oc.postiv_relax = 1, float(m["relax"])
            oc.postiv_subtract_min = 1 if m.get("mode", "clip_neg") == "subtract_min" else 0
        if not (oc.mirrored_on or oc.thresh_on or oc.postiv_on):
            return
        a, p = self._contig(model.opt_obja), self._contig(model.opt_objp)
        scratch = torch.empty(1, dtype=torch.float32, device=a.device) if oc.postiv_subtract_min else None
        _lib.check(_lib.lib().ptyb200_object_constraints(C.byref(oc), ptr(a), ptr(p), a.numel(), ptr(scratch), _stream()))

    @staticmethod
    def _contig(param):
        if not param.data.is_cuda:
            raise RuntimeError("ptyrad_b200 constraints run on CUDA tensors only")
        if not param.data.is_contiguous():
            param.data = param.data.contiguous()
        return param.data

    # ------------------------------------------------------------------------------------------------
    def forward(self, model, niter):
        others = [n for n in _ORDER if n not in _NATIVE and self._active(n, niter)]
        if others and self.fallback is None:
            raise NotImplementedError(f"constraints {others} are not native: pass fallback=<the reference's CombinedConstraint>")
        fb = self.fallback
        with torch.no_grad():
            if fb is not None:
                fb.apply_ortho_pmode(model, niter)
                fb.apply_probe_mask_k(model, niter)
                fb.apply_fix_probe_int(model, niter)
            self.apply_obj_rblur(model, niter)
            self.apply_obj_zblur(model, niter)
            if fb is not None:
                fb.apply_kr_filter(model, niter)
                fb.apply_kz_filter(model, niter)
                fb.apply_complex_ratio(model, niter)
            self.apply_voxel_chain(model, niter)
            if fb is not None:
                fb.apply_tilt_smooth(model, niter)
