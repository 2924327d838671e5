"""ORACLE #2 (test infrastructure, NOT product code) -- NumPy complex128 forward model and the
HAND-DERIVED adjoint, with no autograd anywhere.

This is the explicit statement of what the CUDA adjoint kernels compute; the
reference never writes its adjoint down (it is whatever torch autograd derives from
src/ptyrad/forward.py:53-79, src/ptyrad/utils/image_proc.py:531-532, src/ptyrad/models.py:251-360 and
src/ptyrad/losses.py:36-104).  ``tests/test_oracle_golden.py`` pins it against the reference's own
float64 autograd gradients (golden vectors).  loss_simlar is not restated here (autograd oracle only).

Gradient convention = torch's: for complex x, g_x = dL/dRe(x) + i dL/dIm(x).
"""
from __future__ import annotations

import numpy as np
import torch

fft2 = lambda x: np.fft.fft2(x, axes=(-2, -1))
ifft2 = lambda x: np.fft.ifft2(x, axes=(-2, -1))


def _setup(iv, mp):
    lr = {k: v["lr"] for k, v in mp["update_params"].items()}
    obj = np.asarray(iv["obj"])
    st = dict(
        # torch.abs / torch.angle on complex64, as models.py:99-100 does (numpy's differ in the last ulp)
        a=torch.abs(torch.as_tensor(obj)).to(torch.float32).numpy().astype(np.float64),
        ph=torch.angle(torch.as_tensor(obj)).to(torch.float32).numpy().astype(np.float64),
        probe=np.asarray(iv["probe"]).astype(np.complex64).astype(np.complex128),
        shifts=np.asarray(iv["probe_pos_shifts"], np.float32).astype(np.float64),
        tilts=np.asarray(iv["obj_tilts"], np.float32).astype(np.float64),
        dz=float(np.float32(iv["slice_thickness"])),
        H=np.asarray(iv["H"]).astype(np.complex64).astype(np.complex128),
        occu=np.asarray(iv["omode_occu"], np.float32).astype(np.float64),
        meas=np.asarray(iv["measurements"], np.float32).astype(np.float64),
        crop=np.asarray(iv["crop_pos"]).astype(np.int64),
        dx=float(np.float32(iv["dx"])), lambd=float(np.float32(iv["lambd"])), lr=lr,
    )
    N = st["probe"].shape[-1]
    g = (np.arange(-(N // 2), N - N // 2) + 0.5) / N
    k1 = np.fft.ifftshift(2 * np.pi * g / st["dx"])
    st["Ky"], st["Kx"] = np.meshgrid(k1, k1, indexing="ij")
    st["Kz"] = np.sqrt((2 * np.pi / st["lambd"]) ** 2 - st["Kx"] ** 2 - st["Ky"] ** 2)
    st["tilt_obj"] = bool(lr["obj_tilts"] != 0 or np.any(st["tilts"] != 0))
    st["shift_probes"] = bool(lr["probe_pos_shifts"] != 0)
    st["change_thickness"] = bool(lr["slice_thickness"] != 0)
    st["change_tilt"] = bool(lr["obj_tilts"] != 0)
    st["N"] = N
    return st


def propagators(st, idx):
    """H_n (B|1,N,N) for the 4+1 cases of models.py:339-360 and the per-sample tilt angles."""
    glob = st["tilts"].shape[0] == 1
    t = st["tilts"] if glob else st["tilts"][idx]
    ty, tx = t[:, 0, None, None] / 1e3, t[:, 1, None, None] / 1e3
    ramp = np.exp(1j * st["dz"] * (st["Ky"] * np.tan(ty) + st["Kx"] * np.tan(tx)))
    if st["tilt_obj"] and st["change_thickness"]:
        return np.exp(1j * st["dz"] * st["Kz"]) * ramp, ty, tx
    if st["tilt_obj"]:
        return st["H"] * ramp, ty, tx          # 2A and 2B are numerically the same function of the stored tilts
    if st["change_thickness"]:
        return np.exp(1j * st["dz"] * st["Kz"])[None], ty, tx
    return st["H"][None], ty, tx


def step(iv, mp, lp, idx, eps=1e-10):
    st = _setup(iv, mp)
    idx = np.asarray(idx, np.int64)
    N, B = st["N"], len(idx)
    P = st["probe"].shape[0]
    M, Z = st["a"].shape[:2]
    ar = np.arange(N)
    gy = st["crop"][idx, 0, None, None] + ar[None, :, None]
    gx = st["crop"][idx, 1, None, None] + ar[None, None, :]
    a = st["a"][:, :, gy, gx].transpose(2, 0, 1, 3, 4)          # (B,M,Z,N,N)
    ph = st["ph"][:, :, gy, gx].transpose(2, 0, 1, 3, 4)
    O = a * np.exp(1j * ph)
    occu = st["occu"]

    # ---- forward ----
    Phat = fft2(st["probe"])
    kap = ((ar + N // 2) % N) / N
    if st["shift_probes"]:
        s = st["shifts"][idx]
        w = np.exp(-2j * np.pi * (s[:, 0, None, None] * kap[None, :, None] + s[:, 1, None, None] * kap[None, None, :]))
        psi0 = ifft2(Phat[None] * w[:, None])                    # (B,P,N,N)
    else:
        psi0 = np.broadcast_to(st["probe"][None], (B, P, N, N))
    Hn, ty, tx = propagators(st, idx)
    Hb = Hn[:, None, None]
    psi = np.broadcast_to(psi0[:, :, None], (B, P, M, N, N))
    psis, Phis = [], []
    for z in range(Z - 1):
        psis.append(psi)
        Phi = fft2(psi * O[:, None, :, z])
        Phis.append(Phi)
        psi = ifft2(Hb * Phi)
    psis.append(psi)
    Psi = fft2(psi * O[:, None, :, Z - 1]) / N
    I = np.fft.fftshift((np.abs(Psi) ** 2 * occu[None, None, :, None, None]).sum((1, 2)), axes=(-2, -1)) + eps
    meas = st["meas"][idx]

    # ---- losses and G = dL/dI ----
    Nel = B * N * N
    losses = np.zeros(5)
    G = np.zeros_like(I)
    s_ = lp["loss_single"]
    if s_["state"]:
        pw = s_.get("dp_pow", 0.5)
        Ip, Mp = I ** pw, meas ** pw
        rmse = np.sqrt(np.mean((Ip - Mp) ** 2))
        losses[0] = s_["weight"] * rmse / Mp.mean()
        G += s_["weight"] * (Ip - Mp) * pw * I ** (pw - 1) / (Nel * rmse * Mp.mean())
    s_ = lp["loss_poissn"]
    if s_["state"]:
        pw, e = s_.get("dp_pow", 1), s_.get("eps", 1e-6)
        Ip, Mp = I ** pw, meas ** pw
        losses[1] = -s_["weight"] * np.mean(Mp * np.log(Ip + e) - Ip) / Mp.mean()
        G += -s_["weight"] * (Mp / (Ip + e) - 1.0) * pw * I ** (pw - 1) / (Nel * Mp.mean())
    s_ = lp["loss_pacbed"]
    if s_["state"]:
        pw = s_.get("dp_pow", 0.2)
        Ib, Mb = I.mean(0), meas.mean(0)
        rmse = np.sqrt(np.mean((Ib ** pw - Mb ** pw) ** 2))
        dm = (meas ** pw).mean()
        losses[2] = s_["weight"] * rmse / dm
        G += (s_["weight"] * (Ib ** pw - Mb ** pw) * pw * Ib ** (pw - 1) / (B * N * N * rmse * dm))[None]
    g_ph_extra = np.zeros_like(ph)
    s_ = lp["loss_sparse"]
    if s_["state"]:
        n = s_["ln_order"]
        S = (np.abs(ph) ** n).mean(axis=(0, 2, 3, 4))           # (M,)
        losses[3] = s_["weight"] * np.sum(occu * S ** (1.0 / n))
        g_ph_extra = (s_["weight"] * occu * S ** (1.0 / n - 1.0))[None, :, None, None, None] * \
            np.abs(ph) ** (n - 1) * np.sign(ph) / (B * Z * N * N)

    # ---- adjoint ----
    gPsi = 2.0 * occu[None, None, :, None, None] * np.fft.ifftshift(G, axes=(-2, -1))[:, None, None] * Psi
    gphi = N * ifft2(gPsi)
    gO = np.zeros_like(O)
    S_n = np.zeros((B, N, N))                                   # Im(conj(H_n) gH_n), only if tilts/dz are optimised
    need_H = st["change_tilt"] or st["change_thickness"]
    for z in range(Z - 1, -1, -1):
        gO[:, :, z] = (np.conj(psis[z]) * gphi).sum(1)
        gpsi = np.conj(O[:, None, :, z]) * gphi
        if z > 0:
            T = fft2(gpsi) / (N * N)
            cHT = np.conj(Hb) * T
            if need_H:
                S_n += np.imag(np.conj(Phis[z - 1]) * cHT).sum((1, 2))
            gphi = (N * N) * ifft2(cHT)
    g_a_p = np.real(gO * np.exp(-1j * ph))
    g_ph_p = np.imag(gO * np.conj(O)) + g_ph_extra
    grads = {}
    g_a = np.zeros_like(st["a"])
    g_ph = np.zeros_like(st["ph"])
    for b in range(B):                                          # scatter-add (models.py:264 backward)
        cy, cx = st["crop"][idx[b]]
        g_a[:, :, cy:cy + N, cx:cx + N] += g_a_p[b]
        g_ph[:, :, cy:cy + N, cx:cx + N] += g_ph_p[b]
    grads["obja"], grads["objp"] = g_a, g_ph

    gpsi0 = gpsi.sum(2)                                         # (B,P,N,N): sum over object modes
    if st["shift_probes"]:
        T0 = fft2(gpsi0) / (N * N)
        gPhat = (np.conj(w)[:, None] * T0).sum(0)
        gP = (N * N) * ifft2(gPhat)
        gw = (np.conj(Phat)[None] * T0).sum(1)                  # (B,N,N)
        q = np.imag(np.conj(w) * gw)
        gs = np.zeros_like(st["shifts"])
        np.add.at(gs, (idx, 0), -2 * np.pi * (q * kap[None, :, None]).sum((1, 2)))
        np.add.at(gs, (idx, 1), -2 * np.pi * (q * kap[None, None, :]).sum((1, 2)))
        grads["probe_pos_shifts"] = gs
    else:
        gP = gpsi0.sum(0)
    grads["probe"] = np.stack([gP.real, gP.imag], -1)

    if st["change_tilt"]:
        gt = np.zeros_like(st["tilts"])
        gty = st["dz"] * (st["Ky"][None] * S_n).sum((1, 2)) / np.cos(ty[:, 0, 0]) ** 2 / 1e3 * np.ones(B)
        gtx = st["dz"] * (st["Kx"][None] * S_n).sum((1, 2)) / np.cos(tx[:, 0, 0]) ** 2 / 1e3 * np.ones(B)
        if st["tilts"].shape[0] == 1:
            gt[0, 0], gt[0, 1] = gty.sum(), gtx.sum()
        else:
            np.add.at(gt, (idx, 0), gty)
            np.add.at(gt, (idx, 1), gtx)
        grads["obj_tilts"] = gt
    if st["change_thickness"]:
        Kt = st["Kz"][None] + (st["Ky"][None] * np.tan(ty) + st["Kx"][None] * np.tan(tx) if st["tilt_obj"] else 0.0)
        grads["slice_thickness"] = np.array((Kt * S_n).sum())
    return dict(dp=I, losses=losses, total=losses.sum(), grads=grads)
