"""Deterministic synthetic 4D-STEM inputs for the multislice hot path (bench / test input generation -- NOT part of the
product package ``ptyrad_b200``; lives at the repo root so that ``bench.py``, ``tests/`` and ``__graft_entry__`` share it).

Produces the ``init_variables`` / ``model_params`` / ``loss_params`` dictionaries the
reference feeds to ``PtychoAD`` / ``CombinedLoss`` (layouts: reference
``src/ptyrad/models.py:99-118``; geometry rules: ``src/ptyrad/initialization.py:342-367``
(canvas, crop_pos, sub-pixel shifts) and ``:1455-1474`` (raster scan)).  Everything here is
plain NumPy so that the same inputs can be rebuilt bit-for-bit on a box that has neither
the reference tree nor a GPU.  Nothing in this module touches ``oracle/``.

The physics helpers are written from the published formulas (Kirkland, *Advanced Computing
in Electron Microscopy*, eq. 2.10 for the aberration function; angular-spectrum propagator),
matching the conventions the reference uses (``src/ptyrad/utils/physics.py:92-118,219-305,
382-489``): relativistic wavelength, aperture-limited defocused probe, Hermite-like
orthogonalised probe modes, half-bin-shifted propagator grid.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Optional

import numpy as np

SEED = 20260101

# ----------------------------------------------------------------------------------------
# configs (BASELINE.json "configs", completed per SURVEY.md section 8)
# ----------------------------------------------------------------------------------------


@dataclass
class ScanConfig:
    name: str
    N: int                 # pattern / probe size in px (square)
    scan: int              # scan is scan x scan positions
    P: int                 # probe modes
    M: int                 # object modes
    Z: int                 # slices
    batch: int
    kv: float = 80.0
    conv_angle: float = 24.9
    dx: float = 0.1494
    step: float = 0.429
    dz: float = 2.0
    defocus: float = 50.0
    jitter_px: float = 0.15
    dose: float = 1.0e4
    tilt_each: bool = False          # per-position tilts, optimised (C3)
    lr_shifts: float = 0.0           # probe_pos_shifts lr (0 -> shift_probes False, models.py:120)
    lr_tilts: float = 0.0
    lr_dz: float = 0.0
    loss: str = "single"             # "single" | "poissn"
    notes: str = ""

    def as_dict(self):
        return asdict(self)


CONFIGS = {
    # C1: reference-runnable CPU case
    "C1": ScanConfig("C1", N=128, scan=32, P=1, M=1, Z=1, batch=32),
    # C2: the configuration the metric is quoted on (tBL_WSe2-demo shape)
    "C2": ScanConfig("C2", N=128, scan=64, P=6, M=1, Z=8, batch=256, lr_shifts=1e-4),
    # C3: thick sample with position correction and per-position object tilt
    "C3": ScanConfig("C3", N=256, scan=128, P=8, M=1, Z=32, batch=64, kv=300.0, conv_angle=21.4,
                     dz=10.0, defocus=-200.0, tilt_each=True, lr_shifts=1e-4, lr_tilts=1e-4),
    # C4: large field (multi-GPU config)
    "C4": ScanConfig("C4", N=256, scan=256, P=12, M=1, Z=16, batch=256, lr_shifts=1e-4),
    # C5: mixed-state object + Poisson loss
    "C5": ScanConfig("C5", N=192, scan=96, P=1, M=2, Z=10, batch=512, loss="poissn"),
    # tuning stand-ins: C3 / C4 kernels shapes on a smaller scan (seconds instead of a minute of input generation); not bench lines
    "C3s": ScanConfig("C3s", N=256, scan=64, P=8, M=1, Z=32, batch=64, kv=300.0, conv_angle=21.4,
                      dz=10.0, defocus=-200.0, tilt_each=True, lr_shifts=1e-4, lr_tilts=1e-4),
    "C4s": ScanConfig("C4s", N=256, scan=96, P=12, M=1, Z=16, batch=256, lr_shifts=1e-4),
    # BASELINE configs at their OWN depth (same N, P, M, Z, loss and tilt / shift options) on a tiny scan, so the float64 CPU
    # oracle finishes in seconds: the parity cases for C2..C5 (tests/test_gpu_parity.py::test_baseline_configs_at_depth)
    # the C2 physics on 64^2 patterns (the reference's small-pattern use: demo batch 32; here a saturating batch): fused64 vs general
    "S64": ScanConfig("S64", N=64, scan=64, P=6, M=1, Z=8, batch=1024, lr_shifts=1e-4, dx=0.2988),
    "C2d": ScanConfig("C2d", N=128, scan=4, P=6, M=1, Z=8, batch=8, lr_shifts=1e-4),
    "C3d": ScanConfig("C3d", N=256, scan=3, P=8, M=1, Z=32, batch=4, kv=300.0, conv_angle=21.4,
                      dz=10.0, defocus=-200.0, tilt_each=True, lr_shifts=1e-4, lr_tilts=1e-4),
    "C4d": ScanConfig("C4d", N=256, scan=3, P=12, M=1, Z=16, batch=4, lr_shifts=1e-4),
    "C5d": ScanConfig("C5d", N=192, scan=3, P=1, M=2, Z=10, batch=6, loss="poissn"),
    # tiny cases for parity tests (oracle finishes in well under a second)
    "T32": ScanConfig("T32", N=32, scan=6, P=2, M=2, Z=3, batch=5, lr_shifts=1e-4, step=0.6),
    "T64": ScanConfig("T64", N=64, scan=5, P=3, M=1, Z=4, batch=7, lr_shifts=1e-4),
    "T128": ScanConfig("T128", N=128, scan=4, P=2, M=1, Z=3, batch=6, lr_shifts=1e-4),
    "T128m": ScanConfig("T128m", N=128, scan=4, P=3, M=2, Z=2, batch=5, lr_shifts=1e-4, loss="poissn"),
    "T48": ScanConfig("T48", N=48, scan=4, P=2, M=1, Z=3, batch=4, lr_shifts=1e-4),
    "T256": ScanConfig("T256", N=256, scan=3, P=2, M=1, Z=2, batch=3, lr_shifts=1e-4),
    "T192": ScanConfig("T192", N=192, scan=3, P=1, M=2, Z=2, batch=4, loss="poissn"),
}


# ----------------------------------------------------------------------------------------
# physics helpers
# ----------------------------------------------------------------------------------------

def electron_wavelength(kv: float) -> float:
    """Relativistic electron wavelength in Angstrom from CODATA constants
    (same definition as reference ``physics.py:92-118``)."""
    h = 6.62607015e-34
    m0 = 9.1093837015e-31
    e = 1.602176634e-19
    c = 299792458.0
    hc_kev_ang = h * c / e * 1e-3 * 1e10
    rest_kev = m0 * c * c / e * 1e-3
    return hc_kev_ang / np.sqrt((2.0 * rest_kev + kv) * kv)


def fresnel_propagator(N: int, dx: float, dz: float, lambd: float) -> np.ndarray:
    """Angular-spectrum propagator with the half-bin-shifted frequency grid, zero frequency
    at the corner (reference ``physics.py:475-489``)."""
    g = (np.arange(-(N // 2), N - N // 2, dtype=np.float64) + 0.5) / N
    kk = 2.0 * np.pi * g / dx
    Ky, Kx = np.meshgrid(kk, kk, indexing="ij")
    k0 = 2.0 * np.pi / lambd
    H = np.exp(1j * dz * np.sqrt(k0 * k0 - Kx * Kx - Ky * Ky))
    return np.fft.ifftshift(H)


def stem_probe(N: int, dx: float, kv: float, conv_angle_mrad: float, defocus: float) -> np.ndarray:
    """Aperture-limited defocused probe at the sample plane, unit total intensity.
    chi(k) = -pi*lambda*k^2*df (Kirkland eq. 2.10, defocus term only)."""
    lam = 12.398 / np.sqrt((2 * 511.0 + kv) * kv)
    dk = 1.0 / (dx * N)
    f = (np.arange(N) - N // 2) * dk
    kX, kY = np.meshgrid(f, f, indexing="xy")
    kR2 = kX * kX + kY * kY
    aperture = kR2 <= (conv_angle_mrad * 1e-3 / lam) ** 2
    wave_k = aperture * np.exp(1j * np.pi * lam * kR2 * defocus)
    probe = np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(wave_k)))
    return probe / np.sqrt(np.sum(np.abs(probe) ** 2))


def mixed_probe(base: np.ndarray, P: int, minor_power: float = 0.02) -> np.ndarray:
    """P mutually orthogonal probe modes built from polynomial x Gaussian envelopes of the
    fundamental and Gram-Schmidt orthogonalisation (the construction the reference takes from
    PtychoShelves, ``physics.py:382-472``).  Mode 0 keeps 1-(P-1)*minor_power of the power."""
    if P == 1:
        return base[None].copy()
    N = base.shape[-1]
    mx = int(np.ceil(np.sqrt(P))) - 1
    my = int(np.ceil(P / (mx + 1))) - 1
    c = np.arange(N) - N / 2
    X, Y = np.meshgrid(c, c)
    w = np.abs(base) ** 2
    w = w / w.sum()
    cx, cy = (X * w).sum(), (Y * w).sum()
    vx, vy = (((X - cx) ** 2) * w).sum(), (((Y - cy) ** 2) * w).sum()
    env = np.exp(-((X - cx) ** 2) / (2 * vx) - ((Y - cy) ** 2) / (2 * vy))
    modes = []
    for iy in range(my + 1):
        for ix in range(mx + 1):
            f = ((X - cx) ** ix) * ((Y - cy) ** iy) * base
            if modes:
                f = f * env
            f = f / np.linalg.norm(f)
            for g in modes:
                f = f - np.sum(g * np.conj(f)) * g
            modes.append(f / np.linalg.norm(f))
    modes = np.stack(modes[:P])
    pw = np.full(P, minor_power)
    pw[0] = 1.0 - minor_power * (P - 1)
    return modes * np.sqrt(pw)[:, None, None]


def raster_positions(scan: int, step_px: float, N: int):
    """Raster scan centred in a canvas of int(1.2*ceil(range+N)) px
    (reference ``initialization.py:352,1471-1474``). Returns float positions (Ntot,2) as (y,x)."""
    yy, xx = np.meshgrid(np.arange(scan), np.arange(scan), indexing="ij")
    pos = step_px * np.stack([yy.ravel(), xx.ravel()], -1).astype(np.float64)
    pos = pos - pos.mean(0)
    extent = 1.2 * np.ceil(pos.max(0) - pos.min(0) + N)
    pos = pos + np.ceil(extent / 2 - N / 2)
    return pos


def smooth_field(rng, shape, sigma):
    """Band-limited random field in [0,1] (periodic Gaussian low-pass of white noise)."""
    ny, nx = shape[-2:]
    fy = np.fft.fftfreq(ny)[:, None]
    fx = np.fft.fftfreq(nx)[None, :]
    lp = np.exp(-2 * (np.pi * sigma) ** 2 * (fy * fy + fx * fx))
    f = np.fft.ifft2(np.fft.fft2(rng.standard_normal(shape)) * lp).real
    f = f - f.min(axis=(-2, -1), keepdims=True)
    return f / f.max(axis=(-2, -1), keepdims=True)


# ----------------------------------------------------------------------------------------
# measurement synthesis (independent of the oracle: strong-phase approximation)
# ----------------------------------------------------------------------------------------

def _projected_measurements(obj_true, probe, crop, N, rng, dose, chunk=256):
    """Cheap stand-in for an experiment: I_n = sum_{p,m} occu_m |F_ortho(P_p * prod_z O_z[roi_n])|^2,
    Poisson-sampled at `dose` counts per pattern and normalised so the brightest pixel is 1
    ('max_at_one', reference ``initialization.py:928-930``).  It only fixes the value
    distribution (many zeros, heavy tail); it is not used as a correctness reference."""
    M = obj_true.shape[0]
    proj = np.prod(obj_true, axis=1)                        # (M,Noy,Nox) complex
    Ntot = crop.shape[0]
    meas = np.empty((Ntot, N, N), np.float32)
    ar = np.arange(N)
    for s in range(0, Ntot, chunk):
        c = crop[s:s + chunk]
        gy = c[:, 0, None, None] + ar[None, :, None]
        gx = c[:, 1, None, None] + ar[None, None, :]
        acc = 0.0
        for m in range(M):
            ex = probe[None] * proj[m][gy, gx][:, None]      # (b,P,N,N)
            acc = acc + (np.abs(np.fft.fft2(ex, norm="ortho")) ** 2).sum(1) / M
        acc = np.fft.fftshift(acc, axes=(-2, -1))
        acc = acc / acc.sum(axis=(-2, -1), keepdims=True) * dose
        meas[s:s + chunk] = rng.poisson(acc).astype(np.float32)
    return meas / meas.max()


# ----------------------------------------------------------------------------------------
# public entry
# ----------------------------------------------------------------------------------------

def make_inputs(cfg: ScanConfig | str, seed: int = SEED, measurements: Optional[np.ndarray] = None,
                simulate_measurements: bool = True, positions: Optional[np.ndarray] = None):
    """Build (init_variables, model_params, loss_params) for `cfg`.

    init_variables keys follow reference ``models.py:99-118``; model_params / loss_params follow
    SURVEY appendix B.  If `measurements` is given it is used as is; else if
    `simulate_measurements` they come from `_projected_measurements`; else smooth random
    non-negative patterns (fast, for very large configs).  `positions` (sorted scan indices): only those rows of the measurements
    are produced / kept (a data-parallel rank holds its own block of the scan), announced to the model through the extension key
    `measurements_positions`."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    rng = np.random.default_rng(seed)
    N, P, M, Z = cfg.N, cfg.P, cfg.M, cfg.Z
    lam = electron_wavelength(cfg.kv)
    step_px = cfg.step / cfg.dx

    pos = raster_positions(cfg.scan, step_px, N)
    pos = pos + rng.normal(0.0, cfg.jitter_px, pos.shape)
    Ntot = pos.shape[0]
    canvas = (1.2 * np.ceil(pos.max(0) - pos.min(0) + N)).astype(int)
    crop = np.round(pos).astype(np.int16)
    shifts = (pos - crop).astype(np.float32)
    assert crop.min() >= 0 and (crop.max(0) + N <= canvas).all()
    Noy, Nox = int(canvas[0]), int(canvas[1])

    # current estimate of the object (what the solver holds) and a hidden "true" object
    amp = 0.98 + 0.02 * rng.random((M, Z, Noy, Nox))
    phs = 0.3 * smooth_field(rng, (M, Z, Noy, Nox), 1.5) * rng.random((M, Z, 1, 1)) + 0.02 * rng.random((M, Z, Noy, Nox))
    obj = (amp * np.exp(1j * phs)).astype(np.complex64)
    phs_true = 0.35 * smooth_field(rng, (M, Z, Noy, Nox), 1.2)
    obj_true = np.exp(1j * phs_true)

    base = stem_probe(N, cfg.dx, cfg.kv, cfg.conv_angle, cfg.defocus)
    probe = mixed_probe(base, P)

    if measurements is None:
        if simulate_measurements:
            measurements = _projected_measurements(obj_true, probe, crop.astype(np.int64), N, rng, cfg.dose)
        else:
            env = np.abs(np.fft.fftshift(np.fft.fft2(base))) ** 2
            env = (env / env.max()).astype(np.float32)
            # every 512-row chunk has its own seeded stream, so that any subset of the rows can be produced on its own
            rows = np.arange(Ntot) if positions is None else np.asarray(positions)
            measurements = np.empty((len(rows), N, N), np.float32)
            for c in np.unique(rows // 512):
                sel = np.nonzero(rows // 512 == c)[0]
                chunk = env[None] * np.random.default_rng([seed, 77, int(c)]).random((512, N, N), dtype=np.float32)
                measurements[sel] = chunk[rows[sel] - c * 512]
        if positions is not None and simulate_measurements:
            measurements = measurements[np.asarray(positions)]
    measurements = np.ascontiguousarray(measurements, dtype=np.float32)

    # probe power matches the mean pattern sum (reference initialization.py:1365-1366); with a block of the rows, the block's mean
    # (identical probes on all ranks are restored by the caller when it matters: the benchmark only needs the value scale)
    probe = probe * np.sqrt(measurements[: min(len(measurements), 4096)].sum(axis=(-2, -1)).mean() / np.sum(np.abs(probe) ** 2))
    probe = probe.astype(np.complex64)

    if cfg.tilt_each:
        tilts = rng.normal(0.0, 1.0, (Ntot, 2)).astype(np.float32)
    else:
        tilts = np.zeros((1, 2), np.float32)

    H = fresnel_propagator(N, cfg.dx, cfg.dz, lam).astype(np.complex64)
    iv = dict(
        obj=obj, probe=probe, probe_pos_shifts=shifts, crop_pos=crop, H=H,
        measurements=measurements, omode_occu=(np.ones(M) / M).astype(np.float32),
        obj_tilts=tilts, N_scan_slow=cfg.scan, N_scan_fast=cfg.scan,
        slice_thickness=np.float32(cfg.dz), dx=np.float32(cfg.dx),
        dk=np.float32(1.0 / (cfg.dx * N)), lambd=np.float32(lam), scan_affine=None,
    )
    if positions is not None:
        iv["measurements_positions"] = np.asarray(positions, dtype=np.int64)
    lr = dict(obja=5e-4, objp=5e-4, obj_tilts=cfg.lr_tilts, slice_thickness=cfg.lr_dz,
              probe=1e-4, probe_pos_shifts=cfg.lr_shifts)
    model_params = dict(
        obj_preblur_std=None, detector_blur_std=None,
        optimizer_params=dict(name="Adam", configs={}, load_state=None),
        update_params={k: dict(start_iter=(1 if v != 0 else None), lr=v) for k, v in lr.items()},
    )
    loss_params = default_loss_params(cfg.loss)
    return iv, model_params, loss_params


def default_loss_params(kind: str = "single"):
    """loss_params in the reference's key order (``losses.py:143-155``; order defines the log order)."""
    return dict(
        loss_single=dict(state=(kind == "single"), weight=1.0, dp_pow=0.5),
        loss_poissn=dict(state=(kind == "poissn"), weight=1.0, dp_pow=1.0, eps=1e-6),
        loss_pacbed=dict(state=False, weight=0.5, dp_pow=0.2),
        loss_sparse=dict(state=True, weight=0.1, ln_order=1),
        loss_simlar=dict(state=False, weight=0.1, obj_type="both", scale_factor=[1, 1, 1], blur_std=1),
    )


def random_batches(Ntot: int, batch: int, seed: int = SEED):
    """'random' grouping: a seeded permutation split with np.array_split
    (reference ``reconstruction.py:515-522``; the reference leaves the rng unseeded)."""
    rng = np.random.default_rng(seed + 1)
    perm = rng.permutation(Ntot)
    nb = max(1, Ntot // batch)
    return [b.astype(np.int64) for b in np.array_split(perm, nb)]
