"""torch.autograd glue over the C ABI: tensors stay torch-owned, the arithmetic runs in libptyrad_b200.so.

``MultisliceFunction`` replaces the autograd graph the reference builds for PtychoAD.forward
(models.py:422-436 -> forward.py:20-80); ``DataLossFunction`` / ``SparseLossFunction`` replace the graph of
CombinedLoss.forward (losses.py:36-104).  Saved state is per call (``ctx``), so the LBFGS closure, which runs
several forwards before one backward (reconstruction.py:705-718), works.  All calls are pinned to float32
(the reference may wrap them in autocast, reconstruction.py:794-799).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import Cfg, LossCfg, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ptyrad_b200 runs on CUDA tensors only: the multislice hot path has no CPU fallback")


def make_cfg(N, P, M, Z, Noy, Nox, Ntot, shift_probes, tilt_mode, stash_fourier, dx, lambd, eps=1e-10, path=_lib.PATH_AUTO):
    if N not in _lib.SUPPORTED_N:
        raise ValueError(f"pattern size N={N} is not supported by the CUDA kernels (supported: {_lib.SUPPORTED_N})")
    c = Cfg()
    c.N, c.P, c.M, c.Z, c.Noy, c.Nox, c.Ntot = N, P, M, Z, Noy, Nox, Ntot
    c.shift_probes, c.tilt_mode, c.stash_fourier, c.path = int(shift_probes), int(tilt_mode), int(stash_fourier), int(path)
    c.dx, c.lambd, c.eps = float(dx), float(lambd), float(eps)
    return c


def make_loss_cfg(lp: dict) -> LossCfg:
    l = LossCfg()
    s, p, b, sp = lp["loss_single"], lp["loss_poissn"], lp["loss_pacbed"], lp["loss_sparse"]
    l.single_state, l.single_weight, l.single_pow = int(bool(s["state"])), float(s.get("weight", 1.0)), float(s.get("dp_pow", 0.5))
    l.poissn_state, l.poissn_weight, l.poissn_pow, l.poissn_eps = int(bool(p["state"])), float(p.get("weight", 1.0)), float(p.get("dp_pow", 1.0)), float(p.get("eps", 1e-6))
    l.pacbed_state, l.pacbed_weight, l.pacbed_pow = int(bool(b["state"])), float(b.get("weight", 1.0)), float(b.get("dp_pow", 0.2))
    l.sparse_state, l.sparse_weight, l.sparse_order = int(bool(sp["state"])), float(sp.get("weight", 1.0)), float(sp.get("ln_order", 1))
    return l


def mref(mcfg):
    """ctypes reference to an optional MeasCfg (None -> NULL)."""
    return None if mcfg is None else C.byref(mcfg)


def gather_measurements(cfg: Cfg, mcfg, meas_all, padded, idx):
    """Rows idx of the measurements with the on-the-fly pad / resample applied, as a (B,N,N) tensor (models.py:384-416)."""
    _require_cuda(meas_all, idx)
    out = torch.empty((idx.numel(), cfg.N, cfg.N), dtype=torch.float32, device=meas_all.device)
    _lib.check(_lib.lib().ptyb200_gather_measurements(C.byref(cfg), mref(mcfg), ptr(meas_all), ptr(padded), ptr(idx), idx.numel(), ptr(out), _stream()))
    return out


def propagator(cfg: Cfg, dz: torch.Tensor) -> torch.Tensor:
    """exp(i*dz*Kz) (N,N) complex64, evaluated in float64 on the device (models.py:222-223,341,355)."""
    _require_cuda(dz)
    H = torch.empty((cfg.N, cfg.N), dtype=torch.complex64, device=dz.device)
    _lib.check(_lib.lib().ptyb200_propagator(C.byref(cfg), ptr(dz), ptr(H), _stream()))
    return H


def gather_patches(cfg: Cfg, idx, obja, objp, crop_pos):
    _require_cuda(idx, obja, objp, crop_pos)
    B = idx.numel()
    out = torch.empty((B, cfg.M, cfg.Z, cfg.N, cfg.N, 2), dtype=torch.float32, device=obja.device)
    _lib.check(_lib.lib().ptyb200_gather_patches(C.byref(cfg), ptr(idx), B, ptr(obja), ptr(objp), ptr(crop_pos), ptr(out), _stream()))
    return out


class MultisliceFunction(torch.autograd.Function):
    """dp = forward(obja, objp, tilts, dz, probe(real view), shifts); backward = hand-derived adjoint kernels."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, obja, objp, tilts, dz, probe, shifts, st):
        # st: dict(cfg=Cfg, idx, crop_pos, H, occu, change_thickness)
        _require_cuda(obja, objp, probe, st["idx"])
        cfg = st["cfg"]
        idx = st["idx"]
        B = idx.numel()
        obja, objp, probe = obja.contiguous(), objp.contiguous(), probe.contiguous()
        tilts_c = tilts.contiguous() if cfg.tilt_mode else None
        shifts_c = shifts.contiguous() if cfg.shift_probes else None
        Hbase = propagator(cfg, dz) if st["change_thickness"] else st["H"]
        ws_bytes = _lib.lib().ptyb200_workspace_bytes(C.byref(cfg), B)
        if ws_bytes == 0:
            _lib.check(1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=obja.device)
        dp = torch.empty((B, cfg.N, cfg.N), dtype=torch.float32, device=obja.device)
        _lib.check(_lib.lib().ptyb200_forward(
            C.byref(cfg), ptr(idx), B, ptr(obja), ptr(objp), ptr(st["crop_pos"]), ptr(probe), ptr(shifts_c), ptr(Hbase),
            ptr(tilts_c), ptr(dz), ptr(st["occu"]), ptr(dp), ptr(ws), _stream()))
        ctx.st, ctx.ws, ctx.Hbase = st, ws, Hbase
        # the contiguous tensors the kernels actually read (constraints may leave strided .data behind)
        ctx.save_for_backward(obja, objp, tilts_c if tilts_c is not None else tilts, dz, probe,
                              shifts_c if shifts_c is not None else shifts)
        return dp

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, G):
        obja, objp, tilts, dz, probe, shifts = ctx.saved_tensors
        st, cfg = ctx.st, ctx.st["cfg"]
        idx = st["idx"]
        B = idx.numel()
        n_obja, n_objp, n_tilts, n_dz, n_probe, n_shifts, _ = ctx.needs_input_grad
        need = 0
        g_obja = g_objp = g_probe = g_shifts = g_tilts = g_dz = None
        if n_obja or n_objp:
            need |= _lib.NEED_OBJ
            g_obja, g_objp = torch.empty_like(obja), torch.empty_like(objp)
        if n_probe:
            need |= _lib.NEED_PROBE
            g_probe = torch.empty_like(probe)
        if n_shifts and cfg.shift_probes:
            need |= _lib.NEED_SHIFTS
            g_shifts = torch.empty_like(shifts)
        if n_tilts and cfg.tilt_mode and cfg.Z > 1:
            need |= _lib.NEED_TILTS
            g_tilts = torch.empty_like(tilts)
        if n_dz and st["change_thickness"] and cfg.Z > 1:
            need |= _lib.NEED_DZ
            g_dz = torch.empty_like(dz)
        if need:
            G = G.contiguous().float()
            _lib.check(_lib.lib().ptyb200_backward(
                C.byref(cfg), ptr(idx), B, ptr(obja), ptr(objp), ptr(st["crop_pos"]), ptr(probe),
                ptr(shifts if cfg.shift_probes else None), ptr(ctx.Hbase), ptr(tilts if cfg.tilt_mode else None), ptr(dz),
                ptr(st["occu"]), ptr(G), ptr(ctx.ws), ptr(g_obja), ptr(g_objp), ptr(g_probe), ptr(g_shifts), ptr(g_tilts),
                ptr(g_dz), need, _stream()))
        ctx.ws = None
        # tensors that took no part in the forward get zero gradients, like unused leaves would get None
        return (g_obja if n_obja else None, g_objp if n_objp else None, g_tilts if n_tilts else None,
                g_dz if n_dz else None, g_probe if n_probe else None, g_shifts if n_shifts else None, None)


class DataLossFunction(torch.autograd.Function):
    """(single, poissn, pacbed) = f(dp, measurements[idx]); losses.py:36-89."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, dp, meas_all, idx, cfg, lcfg, mcfg=None, padded=None):
        _require_cuda(dp, meas_all, idx)
        dp = dp.contiguous()
        B = dp.shape[0]
        dev = dp.device
        losses3 = torch.empty(3, dtype=torch.float32, device=dev)
        stats = torch.empty(8, dtype=torch.float64, device=dev)
        pac = torch.empty(2 * cfg.N * cfg.N, dtype=torch.float32, device=dev) if lcfg.pacbed_state else None
        _lib.check(_lib.lib().ptyb200_loss_forward(C.byref(cfg), C.byref(lcfg), ptr(dp), ptr(meas_all), ptr(idx), B,
                                                   ptr(losses3), ptr(stats), ptr(pac), mref(mcfg), ptr(padded), _stream()))
        ctx.save_for_backward(dp, meas_all, idx, stats)
        ctx.pac, ctx.cfg, ctx.lcfg, ctx.mcfg, ctx.padded = pac, cfg, lcfg, mcfg, padded
        return losses3

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, up):
        dp, meas_all, idx, stats = ctx.saved_tensors
        G = torch.empty_like(dp)
        up = up.contiguous().float()
        _lib.check(_lib.lib().ptyb200_loss_grad(C.byref(ctx.cfg), C.byref(ctx.lcfg), ptr(dp), ptr(meas_all), ptr(idx), dp.shape[0],
                                                ptr(stats), ptr(ctx.pac), ptr(up), ptr(G), mref(ctx.mcfg), ptr(ctx.padded), _stream()))
        return G, None, None, None, None, None, None


class SparseLossFunction(torch.autograd.Function):
    """loss_sparse on the batch ROIs of objp without materialising the patches; losses.py:91-104."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, objp, crop_pos, idx, occu, cfg, lcfg):
        _require_cuda(objp, crop_pos, idx, occu)
        objp = objp.contiguous()
        dev = objp.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        Ssum = torch.empty(cfg.M, dtype=torch.float64, device=dev)
        cover = torch.empty(cfg.Noy * cfg.Nox, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().ptyb200_sparse_forward(C.byref(cfg), C.byref(lcfg), ptr(objp), ptr(crop_pos), ptr(idx), idx.numel(),
                                                     ptr(occu), ptr(loss), ptr(Ssum), ptr(cover), _stream()))
        ctx.save_for_backward(objp, crop_pos, idx, occu, Ssum, cover)
        ctx.cfg, ctx.lcfg = cfg, lcfg
        return loss

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, up):
        objp, crop_pos, idx, occu, Ssum, cover = ctx.saved_tensors
        cfg = ctx.cfg
        g = torch.zeros_like(objp)
        up = up.reshape(1).contiguous().float()
        _lib.check(_lib.lib().ptyb200_sparse_grad(C.byref(cfg), C.byref(ctx.lcfg), ptr(objp), ptr(crop_pos), ptr(idx), idx.numel(),
                                                  ptr(occu), ptr(Ssum), ptr(up), ptr(cover), ptr(g), _stream()))
        return g, None, None, None, None, None


class RoiBlurFunction(torch.autograd.Function):
    """(a, phi) ROI planes (B,M,Z,N,N) of the batch, 5x5-blurred inside the patch when sigma > 0: get_obj_ROI + the Gaussian pre-blur
    of models.py:251-284 in two native launches, no gather tensor; backward = adjoint blur + scatter-add into the dense gradients."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, obja, objp, idx, crop_pos, cfg, sigma):
        _require_cuda(obja, objp, idx, crop_pos)
        obja, objp = obja.contiguous(), objp.contiguous()
        B = idx.numel()
        shape = (B, cfg.M, cfg.Z, cfg.N, cfg.N)
        out_a = torch.empty(shape, dtype=torch.float32, device=obja.device)
        out_p = torch.empty_like(out_a)
        tmp = torch.empty((2,) + shape, dtype=torch.float32, device=obja.device) if sigma > 0 else None
        _lib.check(_lib.lib().ptyb200_roi_blur(C.byref(cfg), ptr(idx), B, ptr(obja), ptr(objp), ptr(crop_pos), float(sigma), ptr(tmp),
                                               ptr(out_a), ptr(out_p), _stream()))
        ctx.cfg, ctx.sigma, ctx.idx, ctx.crop_pos, ctx.shape_obj = cfg, float(sigma), idx, crop_pos, obja.shape
        return out_a, out_p

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, ga, gp):
        cfg, B = ctx.cfg, ctx.idx.numel()
        need_a, need_p = ctx.needs_input_grad[0] and ga is not None, ctx.needs_input_grad[1] and gp is not None
        dev = ctx.idx.device
        g_obja = torch.zeros(ctx.shape_obj, dtype=torch.float32, device=dev) if need_a else None
        g_objp = torch.zeros(ctx.shape_obj, dtype=torch.float32, device=dev) if need_p else None
        ga = ga.contiguous().float() if need_a else None
        gp = gp.contiguous().float() if need_p else None
        tmp = torch.empty((2, B, cfg.M, cfg.Z, cfg.N, cfg.N), dtype=torch.float32, device=dev) if ctx.sigma > 0 else None
        if need_a or need_p:
            _lib.check(_lib.lib().ptyb200_roi_blur_adjoint(C.byref(cfg), ptr(ctx.idx), B, ptr(ctx.crop_pos), ctx.sigma, ptr(ga), ptr(gp), ptr(tmp),
                                                           ptr(g_obja), ptr(g_objp), _stream()))
        return g_obja, g_objp, None, None, None, None


class SimlarFunction(torch.autograd.Function):
    """weight * mean over the pooled batch volume of std_m(occu_m * area_pool(plane)) (losses.py:113-138) for one set of ROI planes
    (B,M,Z,N,N); area interpolation, std and mean in one native kernel, backward in another."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, plane, occu, cfg, out_dims, weight):
        _require_cuda(plane, occu)
        plane = plane.contiguous()
        B = plane.shape[0]
        acc = torch.zeros(1, dtype=torch.float64, device=plane.device)
        _lib.check(_lib.lib().ptyb200_simlar_forward(C.byref(cfg), B, ptr(plane), ptr(occu), int(out_dims[0]), int(out_dims[1]), int(out_dims[2]),
                                                     float(weight), ptr(acc), _stream()))
        ctx.save_for_backward(plane, occu)
        ctx.cfg, ctx.out_dims, ctx.weight = cfg, tuple(int(v) for v in out_dims), float(weight)
        return acc[0].to(torch.float32)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        plane, occu = ctx.saved_tensors
        gp = torch.zeros_like(plane)
        up = g.reshape(1).to(torch.float32).contiguous()
        _lib.check(_lib.lib().ptyb200_simlar_backward(C.byref(ctx.cfg), plane.shape[0], ptr(plane), ptr(occu), *ctx.out_dims, ctx.weight, ptr(up),
                                                      ptr(gp), _stream()))
        return gp, None, None, None, None


class GaussianBlur5Function(torch.autograd.Function):
    """5x5 Gaussian blur, reflect padding, last two dims (torchvision gaussian_blur(kernel_size=5); models.py:275-284,379-380,
    losses.py:125,134); backward = the adjoint kernel."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, sigma):
        _require_cuda(x)
        x = x.contiguous()
        ctx.sigma = float(sigma)
        return _blur5(x, ctx.sigma, 0)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        return _blur5(g.contiguous().float(), ctx.sigma, 1), None


def _blur5(x, sigma, transpose):
    H, W = x.shape[-2], x.shape[-1]
    out, tmp = torch.empty_like(x), torch.empty_like(x)
    _lib.check(_lib.lib().ptyb200_gaussian_blur5(ptr(x), ptr(tmp), ptr(out), x.numel() // (H * W), H, W, float(sigma), transpose, _stream()))
    return out
