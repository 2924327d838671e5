"""ctypes binding of ``libptyrad_b200.so`` (C ABI declared in ``include/ptyrad_b200.h``).

The library is the product: there is no Python/torch fallback for the hot path.  ``lib()`` raises if the
shared object is missing, and every call raises ``RuntimeError`` with ``ptyb200_last_error()`` on failure.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTYB_LIB") or os.path.join(_HERE, "lib", "libptyrad_b200.so")   # PTYB_LIB: kernel-variant experiments
ABI_VERSION = 3
ACC_KEEP_STATS, ACC_KEEP_GRADS, ACC_NO_FINISH, ACC_NO_LOSS_FINAL, ACC_ADD_OBJ = 1, 2, 4, 8, 16      # cfg.reserved[4] of a chunked step (include/ptyrad_b200.h)

NEED_OBJ, NEED_PROBE, NEED_SHIFTS, NEED_TILTS, NEED_DZ = 1, 2, 4, 8, 16
PATH_AUTO, PATH_GENERAL, PATH_FUSED = 0, 1, 2
SUPPORTED_N = (16, 32, 48, 64, 96, 128, 192, 256)


class Cfg(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("P", C.c_int32), ("M", C.c_int32), ("Z", C.c_int32),
        ("Noy", C.c_int32), ("Nox", C.c_int32), ("Ntot", C.c_int32),
        ("shift_probes", C.c_int32), ("tilt_mode", C.c_int32), ("stash_fourier", C.c_int32),
        ("path", C.c_int32), ("reserved", C.c_int32 * 5),
        ("dx", C.c_float), ("lambd", C.c_float), ("eps", C.c_float), ("reserved_f", C.c_float),
    ]


class LossCfg(C.Structure):
    _fields_ = [
        ("single_state", C.c_int32), ("single_weight", C.c_float), ("single_pow", C.c_float),
        ("poissn_state", C.c_int32), ("poissn_weight", C.c_float), ("poissn_pow", C.c_float), ("poissn_eps", C.c_float),
        ("pacbed_state", C.c_int32), ("pacbed_weight", C.c_float), ("pacbed_pow", C.c_float),
        ("sparse_state", C.c_int32), ("sparse_weight", C.c_float), ("sparse_order", C.c_float),
    ]


class MeasCfg(C.Structure):
    _fields_ = [
        ("Hs", C.c_int32), ("Ws", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32),
        ("h1", C.c_int32), ("h2", C.c_int32), ("w1", C.c_int32), ("w2", C.c_int32),
        ("scale_y", C.c_float), ("scale_x", C.c_float),
    ]


class ObjConstraints(C.Structure):
    _fields_ = [
        ("mirrored_on", C.c_int32), ("mirrored_relax", C.c_float), ("mirrored_scale", C.c_float), ("mirrored_power", C.c_float),
        ("thresh_on", C.c_int32), ("thresh_relax", C.c_float), ("thresh_lo", C.c_float), ("thresh_hi", C.c_float),
        ("postiv_on", C.c_int32), ("postiv_relax", C.c_float), ("postiv_subtract_min", C.c_int32),
    ]


_P = C.c_void_p
_SIGNATURES = {
    "ptyb200_abi_version": (C.c_int, []),
    "ptyb200_last_error": (C.c_char_p, []),
    "ptyb200_launch_count": (C.c_longlong, []),
    "ptyb200_timing_enable": (None, [C.c_int]),
    "ptyb200_timing_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ptyb200_workspace_bytes": (C.c_size_t, [C.POINTER(Cfg), C.c_int32]),
    "ptyb200_propagator": (C.c_int, [C.POINTER(Cfg), _P, _P, _P]),
    "ptyb200_gather_patches": (C.c_int, [C.POINTER(Cfg), _P, C.c_int32, _P, _P, _P, _P, _P]),
    "ptyb200_forward": (C.c_int, [C.POINTER(Cfg), _P, C.c_int32] + [_P] * 12),
    "ptyb200_forward_loss": (C.c_int, [C.POINTER(Cfg), _P, C.c_int32] + [_P] * 11 + [C.POINTER(LossCfg), _P, _P, C.POINTER(MeasCfg), _P, _P, _P, _P, _P]),
    "ptyb200_backward": (C.c_int, [C.POINTER(Cfg), _P, C.c_int32] + [_P] * 17 + [C.c_uint32, _P]),
    "ptyb200_loss_forward": (C.c_int, [C.POINTER(Cfg), C.POINTER(LossCfg), _P, _P, _P, C.c_int32, _P, _P, _P, C.POINTER(MeasCfg), _P, _P]),
    "ptyb200_loss_grad": (C.c_int, [C.POINTER(Cfg), C.POINTER(LossCfg), _P, _P, _P, C.c_int32, _P, _P, _P, _P, C.POINTER(MeasCfg), _P, _P]),
    "ptyb200_gather_measurements": (C.c_int, [C.POINTER(Cfg), C.POINTER(MeasCfg), _P, _P, _P, C.c_int32, _P, _P]),
    "ptyb200_sparse_forward": (C.c_int, [C.POINTER(Cfg), C.POINTER(LossCfg), _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P]),
    "ptyb200_sparse_grad": (C.c_int, [C.POINTER(Cfg), C.POINTER(LossCfg), _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "ptyb200_gaussian_blur5": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P]),
    "ptyb200_roi_blur": (C.c_int, [C.POINTER(Cfg), _P, C.c_int32, _P, _P, _P, C.c_float, _P, _P, _P, _P]),
    "ptyb200_roi_blur_adjoint": (C.c_int, [C.POINTER(Cfg), _P, C.c_int32, _P, C.c_float, _P, _P, _P, _P, _P, _P]),
    "ptyb200_simlar_forward": (C.c_int, [C.POINTER(Cfg), C.c_int32, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P]),
    "ptyb200_simlar_backward": (C.c_int, [C.POINTER(Cfg), C.c_int32, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, _P]),
    "ptyb200_blur_axis": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_float, C.c_int32, _P]),
    "ptyb200_object_constraints": (C.c_int, [C.POINTER(ObjConstraints), _P, _P, C.c_int64, _P, _P]),
    "ptyb200_backward_zero": (C.c_int, [C.POINTER(Cfg), C.c_int32, _P, _P, _P, C.c_uint32, _P]),
    "ptyb200_accumulators_add": (C.c_int, [C.POINTER(Cfg), C.c_int32, _P, _P, C.c_uint32, _P]),
    "ptyb200_backward_finish": (C.c_int, [C.POINTER(Cfg), C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_uint32, _P, _P]),
    "ptyb200_loss_finalize": (C.c_int, [C.POINTER(Cfg), C.POINTER(LossCfg), C.c_int32, _P, _P, _P, _P]),
    "ptyb200_loss_scale": (C.c_int, [C.POINTER(Cfg), C.POINTER(LossCfg), C.c_int32, _P, _P, _P, _P]),
    "ptyb200_sparse_groups": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "ptyb200_adam_step": (C.c_int, [C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load the shared library once; raise loudly if it is not built (run ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with ptyrad_b200.build.build_library(); "
                               "there is no CPU or torch fallback for the multislice hot path")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        v = h.ptyb200_abi_version()
        if v != ABI_VERSION:
            raise RuntimeError(f"libptyrad_b200 ABI {v} != expected {ABI_VERSION}; rebuild the library")
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("ptyrad_b200: " + lib().ptyb200_last_error().decode())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
