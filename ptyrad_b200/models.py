"""Drop-in ``PtychoAD`` whose forward/backward run in the B200 CUDA library.

Mirrors the Python surface the rest of PtyRAD consumes (reference ``src/ptyrad/models.py:30-436``; the list of
attributes and methods that ``recon_step``, ``CombinedConstraint``, ``save_results`` and ``plot_forward_pass`` touch is in
SURVEY.md section 8b): same constructor, same ``opt_*`` ``nn.Parameter`` names (``.data`` re-bindable, ``requires_grad``
toggled per iteration), same buffers, same behaviour flags, ``model(indices) -> dp_fwd (B,N,N)`` with autograd
connectivity, ``_current_object_patches`` / ``clear_cache()``.  The arithmetic of ``forward`` is NOT torch: it is one
``torch.autograd.Function`` over the C ABI (``engine.MultisliceFunction``).

Helper getters used by saving / plotting (``get_probes``, ``get_propagators``, ``get_propagated_probe``,
``get_obj_patches``) are small torch expressions of the same formulas; they are not on the per-batch path.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, engine


def vprint(*args, verbose=True, **kw):
    if verbose:
        print(*args, **kw)


class LazyPatches:
    """Stand-in for the (B,omode,Nz,Ny,Nx,2) ROI tensor the reference materialises every forward
    (models.py:260-264, 268 MB per batch at the C2 shape).  ``ptyrad_b200.losses.CombinedLoss`` recognises it and
    evaluates the object regularisers natively on the ROIs; anything else that indexes it (e.g. the reference's own
    CombinedLoss doing ``object_patches[...,1]``) transparently gets the materialised, autograd-connected tensor."""

    def __init__(self, model, idx):
        self.model, self.idx = model, idx
        self._t = None

    def materialize(self):
        if self._t is None:
            self._t = self.model._roi_tensor(self.idx)
        return self._t

    def __getitem__(self, key):
        return self.materialize()[key]

    @property
    def shape(self):
        m = self.model
        N = m.opt_probe.shape[-2]
        return torch.Size((self.idx.numel(), m.opt_obja.shape[0], m.opt_obja.shape[1], N, N, 2))

    def __getattr__(self, name):            # .permute, .detach, ... on demand
        return getattr(self.materialize(), name)


class PatchPlanes:
    """The pre-blurred ROI patches as two planes (amplitude, phase), each (B,omode,Nz,Ny,Nx): ``patches[..., 0]`` / ``patches[..., 1]``
    (how the losses read ``_current_object_patches``, losses.py:152-153) return the planes themselves; anything else gets the stacked
    (B,omode,Nz,Ny,Nx,2) tensor of the reference (models.py:264,284)."""

    def __init__(self, a, p):
        self.a, self.p = a, p
        self._t = None

    def materialize(self):
        if self._t is None:
            self._t = torch.stack([self.a, self.p], -1)
        return self._t

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and key[0] is Ellipsis and key[1] in (0, 1):
            return self.p if key[1] else self.a
        return self.materialize()[key]

    @property
    def shape(self):
        return torch.Size(tuple(self.a.shape) + (2,))

    def __getattr__(self, name):
        return getattr(self.materialize(), name)


class PtychoAD(nn.Module):
    """See module docstring.  Constructor signature = reference ``models.py:70``."""

    def __init__(self, init_variables, model_params, device="cuda", verbose=True):
        super().__init__()
        with torch.no_grad():
            vprint("### Initializing PtychoAD model (ptyrad_b200) ###", verbose=verbose)
            self.device = device
            self.verbose = verbose
            self.detector_blur_std = model_params["detector_blur_std"]
            self.obj_preblur_std = model_params["obj_preblur_std"]
            if init_variables.get("on_the_fly_meas_padded", None) is not None:
                self.meas_padded = torch.tensor(init_variables["on_the_fly_meas_padded"], dtype=torch.float32, device=device)
                self.meas_padded_idx = torch.tensor(init_variables["on_the_fly_meas_padded_idx"], dtype=torch.int32, device=device)
                self._meas_pad_idx = tuple(int(v) for v in init_variables["on_the_fly_meas_padded_idx"])   # host copy: no sync per batch
                self.meas_padded = self.meas_padded.contiguous()
            else:
                self.meas_padded = None
            self.meas_scale_factors = init_variables.get("on_the_fly_meas_scale_factors", None)

            self.start_iter = {k: v["start_iter"] for k, v in model_params["update_params"].items()}
            self.lr_params = {k: v["lr"] for k, v in model_params["update_params"].items()}
            self.optimizer_params = model_params["optimizer_params"]

            def t(x, dtype):
                return torch.as_tensor(np.asarray(x), dtype=dtype).to(device)

            obj = torch.as_tensor(np.asarray(init_variables["obj"])).to(device)
            self.opt_obja = nn.Parameter(torch.abs(obj).to(torch.float32))
            self.opt_objp = nn.Parameter(torch.angle(obj).to(torch.float32))
            self.opt_obj_tilts = nn.Parameter(t(init_variables["obj_tilts"], torch.float32))
            self.opt_slice_thickness = nn.Parameter(t(init_variables["slice_thickness"], torch.float32))
            # real view (P,N,N,2): NCCL/optimisers handle it as plain float32 (reference models.py:103,147-150)
            self.opt_probe = nn.Parameter(torch.view_as_real(t(init_variables["probe"], torch.complex64)).clone())
            self.opt_probe_pos_shifts = nn.Parameter(t(init_variables["probe_pos_shifts"], torch.float32))

            self.register_buffer("omode_occu", t(init_variables["omode_occu"], torch.float32))
            self.register_buffer("H", t(init_variables["H"], torch.complex64))
            self.register_buffer("measurements", t(init_variables["measurements"], torch.float32))
            # extension: a data-parallel rank may hold only ITS rows of the measurements.  `measurements_positions` lists
            # the scan indices of the rows given; the loss looks rows up through `_meas_row_of` (scan index -> row, -1 = not held).
            # The reference replicates the whole array on every rank (models.py:109).
            self._meas_row_of = None
            if init_variables.get("measurements_positions", None) is not None:
                pos = torch.as_tensor(np.asarray(init_variables["measurements_positions"], dtype=np.int64), device=device)
                if pos.numel() != self.measurements.shape[0]:
                    raise ValueError("measurements_positions must list one scan index per measurement row")
                Ntot_ = np.asarray(init_variables["crop_pos"]).shape[0]
                self._meas_row_of = torch.full((Ntot_,), -1, dtype=torch.int64, device=device)
                self._meas_row_of[pos] = torch.arange(pos.numel(), dtype=torch.int64, device=device)
            self.register_buffer("N_scan_slow", t(init_variables["N_scan_slow"], torch.int32))
            self.register_buffer("N_scan_fast", t(init_variables["N_scan_fast"], torch.int32))
            self.register_buffer("crop_pos", t(np.asarray(init_variables["crop_pos"]).astype(np.int32), torch.int32))
            self.register_buffer("slice_thickness", t(init_variables["slice_thickness"], torch.float32))
            self.register_buffer("dx", t(init_variables["dx"], torch.float32))
            self.register_buffer("dk", t(init_variables["dk"], torch.float32))
            self.register_buffer("lambd", t(init_variables["lambd"], torch.float32))

            self.scan_affine = init_variables["scan_affine"]
            self.tilt_obj = bool(self.lr_params["obj_tilts"] != 0 or torch.any(self.opt_obj_tilts))
            self.shift_probes = bool(self.lr_params["probe_pos_shifts"] != 0)
            self.change_thickness = bool(self.lr_params["slice_thickness"] != 0)
            self.probe_int_sum = self.get_complex_probe_view().abs().pow(2).sum()
            self.loss_iters, self.iter_times, self.dz_iters, self.avg_tilt_iters = [], [], [], []
            self._current_object_patches = None
            self.kernel_path = _lib.PATH_AUTO
            self.kernel_chunk = 0          # general path: samples per L2-resident chunk (0 = library heuristic; cfg.reserved[2])
            self.kernel_pmodes_per_cta = 0  # general path: probe modes looped over by one CTA (0 = heuristic; cfg.reserved[3])

            self._validate(init_variables)
            self.create_grids()
            self.optimizable_tensors = {
                "obja": self.opt_obja, "objp": self.opt_objp, "obj_tilts": self.opt_obj_tilts,
                "slice_thickness": self.opt_slice_thickness, "probe": self.opt_probe,
                "probe_pos_shifts": self.opt_probe_pos_shifts}
            self.create_optimizable_params_dict(self.lr_params, self.verbose)
            self.init_propagator_vars()
            # host copies of the scalars the kernel configuration needs (no device sync on the hot path)
            self._dx_host = float(np.float32(init_variables["dx"]))
            self._lambd_host = float(np.float32(init_variables["lambd"]))
            vprint("### Done initializing PtychoAD model ###", verbose=verbose)

    # ------------------------------------------------------------------------------------------------
    def _validate(self, iv):
        """'custom' sources bypass Initializer.init_check (initialization.py:508-588); re-check what the kernels assume."""
        P, Ny, Nx, _ = self.opt_probe.shape
        if Ny != Nx:
            raise ValueError(f"probe must be square, got {Ny}x{Nx}")
        if Ny not in _lib.SUPPORTED_N:
            raise ValueError(f"pattern size N={Ny} is not supported by the CUDA kernels (supported: {_lib.SUPPORTED_N})")
        if tuple(self.H.shape) != (Ny, Nx):
            raise ValueError(f"H has shape {tuple(self.H.shape)}, expected {(Ny, Nx)}")
        M, Z, Noy, Nox = self.opt_obja.shape
        if self.omode_occu.numel() != M:
            raise ValueError("omode_occu length does not match the number of object modes")
        Ntot = self.crop_pos.shape[0]
        if self._meas_row_of is not None:
            if tuple(self.measurements.shape[1:]) != (Ny, Nx) and self.meas_padded is None and self.meas_scale_factors is None:
                raise ValueError(f"measurement rows have shape {tuple(self.measurements.shape[1:])}, expected {(Ny, Nx)}")
        elif self.meas_padded is None and self.meas_scale_factors is None and tuple(self.measurements.shape) != (Ntot, Ny, Nx):
            raise ValueError(f"measurements have shape {tuple(self.measurements.shape)}, expected {(Ntot, Ny, Nx)}")
        if self.opt_probe_pos_shifts.shape != (Ntot, 2):
            raise ValueError("probe_pos_shifts must be (Ntot,2)")
        if self.opt_obj_tilts.shape[0] not in (1, Ntot) or self.opt_obj_tilts.shape[1] != 2:
            raise ValueError("obj_tilts must be (1,2) or (Ntot,2)")
        cp = np.asarray(iv["crop_pos"]).astype(np.int64)
        if cp.min() < 0 or (cp[:, 0].max() + Ny > Noy) or (cp[:, 1].max() + Nx > Nox):
            raise ValueError("crop_pos + probe size exceeds the object canvas")

    def get_complex_probe_view(self):
        return torch.view_as_complex(self.opt_probe)

    def create_grids(self):
        """Grids kept for API parity (models.py:152-185); the kernels regenerate them on the fly."""
        dev = self.device
        N = self.opt_probe.shape[-2]
        Noy, Nox = self.opt_objp.shape[-2:]
        g = (torch.arange(-(N // 2), N - N // 2, device=dev) + 0.5) / N
        k1 = torch.fft.ifftshift(2 * math.pi * g / self.dx)
        Ky, Kx = torch.meshgrid(k1, k1, indexing="ij")
        self.propagator_grid = torch.stack([Ky, Kx], 0)
        ar = torch.arange(N, dtype=torch.int32, device=dev)
        self.rpy_grid, self.rpx_grid = torch.meshgrid(ar, ar, indexing="ij")
        self.roy_grid, self.rox_grid = torch.meshgrid(torch.arange(Noy, dtype=torch.int32, device=dev),
                                                      torch.arange(Nox, dtype=torch.int32, device=dev), indexing="ij")
        self.shift_probes_grid = torch.stack([self.rpy_grid / N, self.rpx_grid / N], 0)
        self.shift_object_grid = torch.stack([self.roy_grid / Noy, self.rox_grid / Nox], 0)

    def create_optimizable_params_dict(self, lr_params, verbose=True):
        self.lr_params = lr_params
        self.optimizable_params = []
        for name, lr in lr_params.items():
            if name not in self.optimizable_tensors:
                raise ValueError(f"WARNING: '{name}' is not a valid parameter name, check your `update_params` and choose from "
                                 "'obja', 'objp', 'obj_tilts', 'slice_thickness', 'probe', and 'probe_pos_shifts'")
            self.optimizable_tensors[name].requires_grad = (lr != 0)
            if lr != 0:
                self.optimizable_params.append({"params": [self.optimizable_tensors[name]], "lr": lr})
        if verbose:
            self.print_model_summary()

    def init_propagator_vars(self):
        dz = self.opt_slice_thickness.detach()
        Ky, Kx = self.propagator_grid
        ty = self.opt_obj_tilts[:, 0, None, None] / 1e3
        tx = self.opt_obj_tilts[:, 1, None, None] / 1e3
        self.H_fixed_tilts_full = self.H * torch.exp(1j * dz * (Ky * torch.tan(ty) + Kx * torch.tan(tx)))
        self.k = 2 * math.pi / self.lambd
        self.Kz = torch.sqrt(self.k ** 2 - Kx ** 2 - Ky ** 2)

    def print_model_summary(self):
        v = self.verbose
        vprint("### PtychoAD optimizable variables ###", verbose=v)
        for name, tensor in self.optimizable_tensors.items():
            vprint(f"{name.ljust(16)}: {str(tuple(tensor.shape)).ljust(32)}, {str(tensor.dtype).ljust(16)}, device:{tensor.device}, "
                   f"grad:{str(tensor.requires_grad).ljust(5)}, lr:{self.lr_params[name]:.0e}", verbose=v)
        total_var = sum(t.numel() for t in self.optimizable_tensors.values() if t.requires_grad)
        vprint(f"Total measurement values  : {self.measurements.numel():,d}", verbose=v)
        vprint(f"Total optimizing variables: {total_var:,d}", verbose=v)
        vprint(f"Sub-px probe shift        : {self.shift_probes}; tilt propagator: {self.tilt_obj}; change thickness: {self.change_thickness}", verbose=v)

    # ------------------------------------------------------------------------------------------------
    def _index_tensor(self, indices):
        """indices: numpy int64 array (make_batches) or a torch tensor (DataLoader); reconstruction.py:519-522,130."""
        if isinstance(indices, torch.Tensor):
            return indices.to(device=self.opt_obja.device, dtype=torch.int64).contiguous()
        return torch.as_tensor(np.asarray(indices, dtype=np.int64), device=self.opt_obja.device)

    def _tilt_mode(self):
        if not self.tilt_obj:
            return 0
        return 1 if self.opt_obj_tilts.shape[0] == 1 else 2

    def _cfg(self, stash_fourier, patch_mode=False):
        M, Z, Noy, Nox = self.opt_obja.shape
        P, N = self.opt_probe.shape[0], self.opt_probe.shape[1]
        if patch_mode:          # the "object" handed to the kernels is the (pre-blurred) per-sample ROI stack (B,M,Z,N,N)
            Noy = Nox = N
        cfg = engine.make_cfg(N, P, M, Z, Noy, Nox, self.crop_pos.shape[0], self.shift_probes, self._tilt_mode(), stash_fourier,
                              self._dx_host, self._lambd_host, 1e-10, self.kernel_path)
        cfg.reserved[1] = 1 if patch_mode else 0
        cfg.reserved[2] = int(self.kernel_chunk)
        cfg.reserved[3] = int(self.kernel_pmodes_per_cta)
        return cfg

    def _roi_tensor(self, idx):
        """Materialised, autograd-connected ROI tensor (B,M,Z,N,N,2): the gather of models.py:251-265."""
        N = self.opt_probe.shape[1]
        gy = (self.rpy_grid[None] + self.crop_pos[idx, None, None, 0]).long()
        gx = (self.rpx_grid[None] + self.crop_pos[idx, None, None, 1]).long()
        a = self.opt_obja[:, :, gy, gx]
        p = self.opt_objp[:, :, gy, gx]
        return torch.stack([a, p], -1).permute(2, 0, 1, 3, 4, 5)

    def get_obj_ROI(self, indices):
        return self._roi_tensor(self._index_tensor(indices))

    def get_obj_patches(self, indices):
        """ROI tensor, Gaussian pre-blurred (5x5, reflect) on amplitude and phase if obj_preblur_std is set (models.py:267-284)."""
        patches = self.get_obj_ROI(indices)
        if self.obj_preblur_std is None or self.obj_preblur_std == 0:
            return patches
        return gaussian_blur5(patches.permute(5, 0, 1, 2, 3, 4), self.obj_preblur_std).permute(1, 2, 3, 4, 5, 0)

    def get_probes(self, indices):
        probe = self.get_complex_probe_view()
        idx = self._index_tensor(indices)
        if not self.shift_probes:
            return torch.broadcast_to(probe, (idx.numel(), *probe.shape))
        s = self.opt_probe_pos_shifts[idx]
        ky, kx = self.shift_probes_grid
        w = torch.exp(-2j * math.pi * (s[:, 1, None, None, None] * kx + s[:, 0, None, None, None] * ky))
        spec = torch.fft.fftshift(torch.fft.fft2(probe), dim=(-2, -1))
        return torch.fft.ifft2(torch.fft.ifftshift(spec[None] * w, dim=(-2, -1)))

    def get_propagators(self, indices):
        """The 4+1 cases of models.py:339-360 (helper for saving/plotting; the kernels form H_n on the fly)."""
        idx = self._index_tensor(indices)
        glob = self.opt_obj_tilts.shape[0] == 1
        change_tilt = self.lr_params["obj_tilts"] != 0
        dz, Kz = self.opt_slice_thickness, self.Kz
        Ky, Kx = self.propagator_grid
        t = self.opt_obj_tilts if glob else self.opt_obj_tilts[idx]
        ty, tx = t[:, 0, None, None] / 1e3, t[:, 1, None, None] / 1e3
        if self.tilt_obj and self.change_thickness:
            return torch.exp(1j * dz * Kz) * torch.exp(1j * dz * (Ky * torch.tan(ty) + Kx * torch.tan(tx)))
        if self.tilt_obj:
            if change_tilt:
                return self.H * torch.exp(1j * dz * (Ky * torch.tan(ty) + Kx * torch.tan(tx)))
            return self.H_fixed_tilts_full if glob else self.H_fixed_tilts_full[idx]
        if self.change_thickness:
            return torch.exp(1j * dz * Kz)[None]
        return self.H[None]

    def get_propagated_probe(self, index):
        probe = self.get_probes(index)[0].detach()
        H = self.get_propagators(index)[[0]].detach()
        Z = self.opt_objp.shape[1]
        out = torch.zeros((Z, *probe.shape), dtype=probe.dtype, device=probe.device)
        psi = probe
        for n in range(Z):
            out[n] = psi
            psi = torch.fft.ifft2(H[None] * torch.fft.fft2(psi))
        return out

    def _meas_cfg(self, meas_all=None):
        """(MeasCfg, padded canvas) describing the on-the-fly pad / resample of the measurements for the kernels, or (None, None)
        when the stored patterns are used as they are (models.py:392-409)."""
        sf = self.meas_scale_factors
        resample = sf is not None and any(f != 1 for f in sf)
        if self.meas_padded is None and not resample:
            return None, None
        meas_all = self.measurements if meas_all is None else meas_all
        m = _lib.MeasCfg()
        m.Hs, m.Ws = int(meas_all.shape[-2]), int(meas_all.shape[-1])
        padded = None
        if self.meas_padded is not None:
            padded = self.meas_padded
            m.Hp, m.Wp = int(padded.shape[-2]), int(padded.shape[-1])
            m.h1, m.h2, m.w1, m.w2 = self._meas_pad_idx
        if resample:
            m.scale_y, m.scale_x = float(sf[0]), float(sf[1])
        return m, padded

    def meas_rows(self, idx):
        """Rows of `self.measurements` that hold the patterns of scan indices `idx` (identity unless this rank holds a shard)."""
        return idx if self._meas_row_of is None else self._meas_row_of[idx]

    def get_measurements(self, indices=None):
        """measurements[indices] with the optional on-the-fly pad / bilinear resample (models.py:384-416); without indices the
        stored array is returned as is, like the reference does (models.py:411-414)."""
        if indices is None:
            return self.measurements
        idx = self.meas_rows(self._index_tensor(indices))
        mcfg, padded = self._meas_cfg()
        if mcfg is None:
            return self.measurements[idx]
        N = self.opt_probe.shape[1]
        cfg = engine.make_cfg(N, 1, 1, 1, N, N, self.measurements.shape[0], 0, 0, 0, 1.0, 1.0)
        return engine.gather_measurements(cfg, mcfg, self.measurements, padded, idx)

    def clear_cache(self):
        self._current_object_patches = None

    # ------------------------------------------------------------------------------------------------
    def forward(self, indices):
        """dp_fwd (B,N,N) float32 for the scan indices of one batch (reference models.py:422-436)."""
        idx = self._index_tensor(indices)
        need_prop = (self.opt_obj_tilts.requires_grad and self.tilt_obj) or (self.opt_slice_thickness.requires_grad and self.change_thickness)
        preblur = self.obj_preblur_std is not None and self.obj_preblur_std != 0
        cfg = self._cfg(stash_fourier=bool(need_prop) and torch.is_grad_enabled(), patch_mode=preblur)
        st = dict(cfg=cfg, idx=idx, crop_pos=self.crop_pos, H=self.H, occu=self.omode_occu, change_thickness=self.change_thickness)
        if preblur:
            # pre-blurred ROIs (models.py:267-284): gather + 5x5 blur in two native launches straight from the dense object (no gather
            # tensor; backward = adjoint blur + scatter-add); the multislice kernels take the per-sample planes as their "object" (patch
            # mode).  Blurring once per (sample, object mode, slice) instead of inside the wave kernels: DESIGN 3.6.
            obja, objp = engine.RoiBlurFunction.apply(self.opt_obja, self.opt_objp, idx, self.crop_pos, self._cfg(False), float(self.obj_preblur_std))
            self._current_object_patches = PatchPlanes(obja, objp)
        else:
            obja, objp = self.opt_obja, self.opt_objp
            self._current_object_patches = LazyPatches(self, idx)
        dp = engine.MultisliceFunction.apply(obja, objp, self.opt_obj_tilts, self.opt_slice_thickness,
                                             self.opt_probe, self.opt_probe_pos_shifts, st)
        if self.detector_blur_std is not None and self.detector_blur_std != 0:
            dp = gaussian_blur5(dp, self.detector_blur_std)
        return dp


def gaussian_blur5(x, sigma):
    """5x5 Gaussian, reflect padding, on the last two dims (what torchvision's gaussian_blur(kernel_size=5) computes;
    reference models.py:275-284,379-380).  CUDA float32 tensors go through the native blur / adjoint-blur kernels
    (ptyb200_gaussian_blur5); anything else (the CPU helper getters, float64 checks) uses separable shifted sums -- not conv2d:
    cuDNN convolutions default to TF32 (1e-3 errors in the gradient).  Off the default path."""
    if x.is_cuda and x.dtype == torch.float32:
        return engine.GaussianBlur5Function.apply(x, float(sigma))       # native kernels: blur and its adjoint
    t = torch.arange(-2, 3, dtype=x.dtype, device=x.device)
    k = torch.exp(-0.5 * (t / sigma) ** 2)
    k = k / k.sum()
    sh = x.shape
    H, W = sh[-2], sh[-1]
    y = torch.nn.functional.pad(x.reshape(-1, 1, H, W), (2, 2, 2, 2), mode="reflect")
    y = sum(k[i] * y[..., :, i:i + W] for i in range(5))
    y = sum(k[i] * y[..., i:i + H, :] for i in range(5))
    return y.reshape(sh)
