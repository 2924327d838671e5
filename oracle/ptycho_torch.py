"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the PtyRAD hot path in torch.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product package ``ptyrad_b200`` never does.

What it restates (all citations relative to /root/reference):
  * object ROI gather with int32 index arithmetic          src/ptyrad/models.py:251-265
  * sub-pixel Fourier probe shift (non-centred j/N grid)    src/ptyrad/utils/image_proc.py:531-532,
                                                           src/ptyrad/models.py:179,286-298
  * the 4+1 propagator cases                                src/ptyrad/models.py:300-360, :164-171, :210-223
  * mixed-state multislice forward model                    src/ptyrad/forward.py:53-79
  * the five loss terms                                     src/ptyrad/losses.py:36-155
  * one step of the non-LBFGS branch of recon_step          src/ptyrad/reconstruction.py:738-772
Gradients come from torch autograd exactly as in the reference (``loss.backward()``,
reconstruction.py:753); ``oracle/adjoint_np.py`` holds the independent hand-derived adjoint.

Parity pin: the reference has no tests or golden vectors of its own, so this
restatement is pinned against outputs of the reference itself run in the build container --
``tests/golden/make_golden.py`` imports the unmodified reference modules and stores inputs +
outputs; ``tests/test_oracle_golden.py`` checks this file against them.

The arithmetic lives in third-party torch (pocketfft/MKL FFT + ATen), not vendored in the
reference; pinned here: torch 2.11.0+cu128.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

C64, C128 = torch.complex64, torch.complex128


def _cdtype(rdtype):
    return C64 if rdtype == torch.float32 else C128


class OracleModel:
    """Holds the six optimisable tensors and the fixed buffers exactly as the reference model does
    (models.py:99-122), in `dtype` (float32 = like-for-like, float64 = arbiter)."""

    def __init__(self, iv: dict, model_params: dict, dtype=torch.float32, device="cpu"):
        rd, cd = dtype, _cdtype(dtype)
        self.rd, self.cd = rd, cd
        self.device = torch.device(device)
        obj = torch.as_tensor(np.asarray(iv["obj"]))
        # amplitude / phase are taken from the complex64 object in float32 first (models.py:99-100)
        self.obja = torch.abs(obj).to(torch.float32).to(rd).requires_grad_(True)
        self.objp = torch.angle(obj).to(torch.float32).to(rd).requires_grad_(True)
        self.tilts = torch.as_tensor(np.asarray(iv["obj_tilts"]), dtype=torch.float32).to(rd).clone().requires_grad_(True)   # clone: never alias the caller's arrays
        self.dz = torch.as_tensor(np.asarray(iv["slice_thickness"]), dtype=torch.float32).to(rd).clone().requires_grad_(True)
        pr = torch.as_tensor(np.asarray(iv["probe"])).to(C64)
        self.probe = torch.view_as_real(pr).to(rd).clone().requires_grad_(True)      # (P,N,N,2) real view
        self.shifts = torch.as_tensor(np.asarray(iv["probe_pos_shifts"]), dtype=torch.float32).to(rd).clone().requires_grad_(True)
        self.occu = torch.as_tensor(np.asarray(iv["omode_occu"]), dtype=torch.float32).to(rd)
        self.H = torch.as_tensor(np.asarray(iv["H"])).to(C64).to(cd)
        self.meas = torch.as_tensor(np.asarray(iv["measurements"]), dtype=torch.float32).to(rd)
        self.crop = torch.as_tensor(np.asarray(iv["crop_pos"]).astype(np.int32))
        self.dx = torch.as_tensor(np.asarray(iv["dx"]), dtype=torch.float32).to(rd)
        self.lambd = torch.as_tensor(np.asarray(iv["lambd"]), dtype=torch.float32).to(rd)
        lr = {k: v["lr"] for k, v in model_params["update_params"].items()}
        self.lr = lr
        self.tilt_obj = bool(lr["obj_tilts"] != 0 or torch.any(self.tilts.detach() != 0))
        self.shift_probes = bool(lr["probe_pos_shifts"] != 0)
        self.change_thickness = bool(lr["slice_thickness"] != 0)
        self.change_tilt = bool(lr["obj_tilts"] != 0)
        self.N = int(pr.shape[-1])
        if self.device.type != "cpu":          # same restatement on another device (bench.py: torch-eager GPU baseline)
            for name in ("obja", "objp", "tilts", "dz", "probe", "shifts"):
                setattr(self, name, getattr(self, name).detach().to(self.device).requires_grad_(True))
            for name in ("occu", "H", "meas", "crop", "dx", "lambd"):
                setattr(self, name, getattr(self, name).to(self.device))
        self._grids()

    # --- grids (models.py:152-185, 210-223) ------------------------------------------------
    def _grids(self):
        N, rd = self.N, self.rd
        g = (torch.arange(-(N // 2), N - N // 2, device=self.device) + 0.5).to(rd) / N
        k1 = torch.fft.ifftshift(2 * math.pi * g / self.dx)
        self.Ky, self.Kx = torch.meshgrid(k1, k1, indexing="ij")
        self.k0 = 2 * math.pi / self.lambd
        self.Kz = torch.sqrt(self.k0 ** 2 - self.Kx ** 2 - self.Ky ** 2)
        ar = torch.arange(N, dtype=torch.int32, device=self.device)
        self.ry, self.rx = torch.meshgrid(ar, ar, indexing="ij")
        self.sy = self.ry.to(rd) / N        # shift grid: j/N, j = 0..N-1 (not centred)
        self.sx = self.rx.to(rd) / N
        with torch.no_grad():
            ty = self.tilts[:, 0, None, None] / 1e3
            tx = self.tilts[:, 1, None, None] / 1e3
            self.H_fixed_tilts = self.H * torch.exp(1j * self.dz * (self.Ky * torch.tan(ty) + self.Kx * torch.tan(tx)))

    def params(self):
        return dict(obja=self.obja, objp=self.objp, obj_tilts=self.tilts, slice_thickness=self.dz,
                    probe=self.probe, probe_pos_shifts=self.shifts)

    # --- pieces ---------------------------------------------------------------------------
    def _idx(self, idx):
        if torch.is_tensor(idx):
            return idx.to(device=self.device, dtype=torch.int64)
        return torch.as_tensor(np.asarray(idx), dtype=torch.int64, device=self.device)

    def roi_index(self, idx):
        """int32 ROI addresses: gy = y + crop[n,0], gx = x + crop[n,1]  (models.py:261-262)."""
        idx = self._idx(idx)
        gy = self.ry[None] + self.crop[idx, None, None, 0]
        gx = self.rx[None] + self.crop[idx, None, None, 1]
        return gy.long(), gx.long()

    def patches(self, idx):
        gy, gx = self.roi_index(idx)
        a = self.obja[:, :, gy, gx].permute(2, 0, 1, 3, 4)       # (B,M,Z,N,N)
        p = self.objp[:, :, gy, gx].permute(2, 0, 1, 3, 4)
        return a, p

    def probes(self, idx):
        pc = torch.view_as_complex(self.probe)
        idx = self._idx(idx)
        if not self.shift_probes:
            return pc[None].expand(len(idx), *pc.shape)
        s = self.shifts[idx]
        ramp = torch.exp(-2j * math.pi * (s[:, 1, None, None, None] * self.sx + s[:, 0, None, None, None] * self.sy))
        spec = torch.fft.fftshift(torch.fft.fft2(pc), dim=(-2, -1))
        return torch.fft.ifft2(torch.fft.ifftshift(spec[None] * ramp, dim=(-2, -1)))

    def propagators(self, idx):
        idx = self._idx(idx)
        glob = self.tilts.shape[0] == 1
        t = self.tilts if glob else self.tilts[idx]
        ty, tx = t[:, 0, None, None] / 1e3, t[:, 1, None, None] / 1e3
        ramp = lambda: torch.exp(1j * self.dz * (self.Ky * torch.tan(ty) + self.Kx * torch.tan(tx)))
        if self.tilt_obj and self.change_thickness:
            return torch.exp(1j * self.dz * self.Kz) * ramp()
        if self.tilt_obj:
            if self.change_tilt:
                return self.H * ramp()
            return self.H_fixed_tilts if glob else self.H_fixed_tilts[idx]
        if self.change_thickness:
            return torch.exp(1j * self.dz * self.Kz)[None]
        return self.H[None]

    def forward(self, idx, eps=1e-10):
        a, p = self.patches(idx)
        O = torch.polar(a, p).to(self.cd)                        # (B,M,Z,N,N)
        psi = self.probes(idx)[:, :, None]                       # (B,P,1,N,N)
        Hn = self.propagators(idx)[:, None, None]                # (B|1,1,1,N,N)
        Z = O.shape[2]
        for z in range(Z - 1):
            psi = torch.fft.ifft2(Hn * torch.fft.fft2(psi * O[:, None, :, z]))
        psi = psi * O[:, None, :, Z - 1]
        far = torch.fft.fftshift(torch.fft.fft2(psi, norm="ortho"), dim=(-2, -1))
        dp = (far.abs().square() * self.occu[:, None, None]).sum(dim=(1, 2)) + eps
        return dp, (a, p)


# --- losses (losses.py:36-155) --------------------------------------------------------------

def loss_terms(dp, meas, objp_patches, occu, lp: dict, obja_patches=None):
    zero = lambda: torch.zeros((), dtype=dp.dtype, device=dp.device)
    out = []
    s = lp["loss_single"]
    if s["state"]:
        pw = s.get("dp_pow", 0.5)
        mp = meas.pow(pw)
        out.append(s["weight"] * torch.sqrt(torch.mean((dp.pow(pw) - mp) ** 2)) / mp.mean())
    else:
        out.append(zero())
    s = lp["loss_poissn"]
    if s["state"]:
        pw, e = s.get("dp_pow", 1), s.get("eps", 1e-6)
        mp, ip = meas.pow(pw), dp.pow(pw)
        out.append(-s["weight"] * torch.mean(mp * torch.log(ip + e) - ip) / mp.mean())
    else:
        out.append(zero())
    s = lp["loss_pacbed"]
    if s["state"]:
        pw = s.get("dp_pow", 0.2)
        out.append(s["weight"] * torch.sqrt(torch.mean((dp.mean(0).pow(pw) - meas.mean(0).pow(pw)) ** 2)) / meas.pow(pw).mean())
    else:
        out.append(zero())
    s = lp["loss_sparse"]
    if s["state"]:
        n = s["ln_order"]
        out.append(s["weight"] * (objp_patches.abs().pow(n).mean(dim=(0, 2, 3, 4)).pow(1.0 / n) * occu).sum())
    else:
        out.append(zero())
    s = lp["loss_simlar"]
    if s["state"]:
        out.append(_loss_simlar(obja_patches, objp_patches, occu, s))
    else:
        out.append(zero())
    return sum(out), out


def _gauss5(x, sigma):
    """5-tap separable Gaussian with reflect padding on the last two dims
    (torchvision.transforms.functional.gaussian_blur(kernel_size=5) semantics, losses.py:125,134)."""
    t = torch.arange(-2, 3, dtype=x.dtype)
    k = torch.exp(-0.5 * (t / sigma) ** 2)
    k = k / k.sum()
    sh = x.shape
    y = x.reshape(-1, 1, sh[-2], sh[-1])
    y = torch.nn.functional.pad(y, (2, 2, 2, 2), mode="reflect")
    y = torch.nn.functional.conv2d(y, k.view(1, 1, 1, 5))
    y = torch.nn.functional.conv2d(y, k.view(1, 1, 5, 1))
    return y.reshape(sh)


def _loss_simlar(obja_patches, objp_patches, occu, s):
    tot = torch.zeros((), dtype=objp_patches.dtype)
    sf = s.get("scale_factor")
    for name, x in (("amplitude", obja_patches), ("phase", objp_patches)):
        if s["obj_type"] not in (name, "both"):
            continue
        if s.get("blur_std"):
            x = _gauss5(x, s["blur_std"])
        if sf is not None and any(f != 1 for f in sf):
            x = torch.nn.functional.interpolate(x, scale_factor=tuple(sf), mode="area")
        tot = tot + (x * occu[:, None, None, None]).std(1).mean()
    return s["weight"] * tot


# --- one step: forward + loss + backward -----------------------------------------------------

def oracle_step(iv, model_params, loss_params, idx, dtype=torch.float32, grad_names: Optional[Sequence[str]] = None,
                model: Optional[OracleModel] = None):
    """Returns dict(dp, losses(5), total, grads{name: dense ndarray}) for batch `idx`."""
    m = model or OracleModel(iv, model_params, dtype)
    ps = m.params()
    if grad_names is None:
        grad_names = [k for k, v in m.lr.items() if v != 0]
    for k, t in ps.items():
        t.requires_grad_(k in grad_names)
        t.grad = None
    dp, (a, p) = m.forward(idx)
    meas = m.meas[torch.as_tensor(np.asarray(idx), dtype=torch.int64)]
    total, terms = loss_terms(dp, meas, p, m.occu, loss_params, obja_patches=a)
    total.backward()
    grads = {}
    for k in grad_names:
        g = ps[k].grad
        grads[k] = (torch.zeros_like(ps[k]) if g is None else g).detach().numpy().copy()
    return dict(dp=dp.detach().numpy(), total=float(total.detach()), losses=np.array([float(t.detach()) for t in terms]), grads=grads, model=m)


def ddp_emulated_grads(iv, model_params, loss_params, idx, world: int, dtype=torch.float32):
    """What the reference computes under DDP with split_batches=True (utils/common.py:61-65,
    reconstruction.py:134-137,753): each rank takes a contiguous 1/world slice of the batch, builds its
    OWN loss on it, and the gradients are averaged.  Used as the multi-GPU oracle."""
    chunks = np.array_split(np.asarray(idx), world)
    acc = None
    for c in chunks:
        r = oracle_step(iv, model_params, loss_params, c, dtype)
        if acc is None:
            acc = {k: v / world for k, v in r["grads"].items()}
        else:
            for k, v in r["grads"].items():
                acc[k] += v / world
    return acc


# --- timing harness for bench.py's CPU legs --------------------------------------------------

class OracleTrainer:
    """zero_grad -> forward -> measurements -> loss -> backward -> Adam.step on the CPU, the same
    sequence as the non-LBFGS branch of recon_step (reconstruction.py:738-772)."""

    def __init__(self, iv, model_params, loss_params, threads: Optional[int] = None, device="cpu"):
        if threads:
            torch.set_num_threads(threads)
        self.m = OracleModel(iv, model_params, torch.float32, device=device)
        self.lp = loss_params
        groups = [dict(params=[t], lr=self.m.lr[k]) for k, t in self.m.params().items() if self.m.lr[k] != 0]
        for k, t in self.m.params().items():
            t.requires_grad_(self.m.lr[k] != 0)
        self.opt = torch.optim.Adam(groups)

    def step(self, idx):
        self.opt.zero_grad()
        dp, (a, p) = self.m.forward(idx)
        meas = self.m.meas[self.m._idx(idx)]
        total, terms = loss_terms(dp, meas, p, self.m.occu, self.lp, obja_patches=a)
        total.backward()
        self.opt.step()
        return float(total.detach()) if self.m.device.type == "cpu" else total.detach()
