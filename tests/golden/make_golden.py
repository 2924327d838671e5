"""Generate golden input/output vectors by running the UNMODIFIED reference hot path.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case it builds synthetic inputs (workloads), instantiates the reference's own
``ptyrad.models.PtychoAD`` and ``ptyrad.losses.CombinedLoss`` on the CPU in float32, runs
forward -> get_measurements -> loss -> backward (the sequence of reconstruction.py:792-806,753) and stores
inputs, intensities, the five loss terms and the dense gradients in ``tests/golden/<case>.npz``.
The float64 twin of the same run (model.double()) is stored too, as the arbiter for tolerances.
"""
import os
import sys
from dataclasses import replace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from ptyrad.models import PtychoAD          # noqa: E402  (reference, unmodified)
from ptyrad.losses import CombinedLoss      # noqa: E402

from workloads import CONFIGS, ScanConfig, make_inputs, default_loss_params  # noqa: E402

BASE = ScanConfig("G", N=32, scan=4, P=2, M=2, Z=3, batch=5, step=0.6, lr_shifts=1e-4)


def cases():
    lp_single = default_loss_params("single")
    lp_poissn = default_loss_params("poissn")
    lp_all = default_loss_params("single")
    lp_all["loss_poissn"]["state"] = True
    lp_all["loss_pacbed"]["state"] = True
    lp_all["loss_sparse"]["ln_order"] = 2
    lp_sim = default_loss_params("single")
    lp_sim["loss_simlar"]["state"] = True
    out = {
        "g_base": (BASE, lp_single, {}),
        "g_noshift_poissn": (replace(BASE, lr_shifts=0.0), lp_poissn, {}),
        "g_all_losses": (replace(BASE, M=1), lp_all, {}),
        "g_single_slice": (replace(BASE, Z=1, P=1, M=1), lp_single, {}),
        "g_tilt_each": (replace(BASE, M=1, tilt_each=True, lr_tilts=1e-4), lp_single, {}),
        "g_tilt_fixed": (replace(BASE, M=1, tilt_each=True, lr_tilts=0.0), lp_single, {}),
        "g_tilt_global_dz": (replace(BASE, M=1, lr_tilts=1e-4, lr_dz=1e-4), lp_single, {"global_tilt": (0.7, -0.4)}),
        "g_dz_only": (replace(BASE, M=1, lr_dz=1e-4), lp_single, {}),
        "g_simlar": (BASE, lp_sim, {}),
        "g_n48": (replace(BASE, N=48, scan=3, M=1, P=2, Z=2, batch=4), lp_single, {}),
        # depth: 8 slices = 15 chained FFT pairs per wave (the C2 depth), shifts + per-position tilts optimised
        "g_z8": (replace(BASE, M=1, P=3, Z=8, tilt_each=True, lr_tilts=1e-4), lp_single, {}),
        # depth: 16 slices, two object modes, Poisson + sparse (the C4 / C5 depth)
        "g_z16_poissn": (replace(BASE, P=2, M=2, Z=16, batch=3), lp_poissn, {}),
    }
    return out


def run_reference(iv, mp, lp, idx, dtype):
    torch.manual_seed(0)
    model = PtychoAD(iv, mp, device="cpu", verbose=False)
    loss_fn = CombinedLoss(lp, device="cpu")
    if dtype == torch.float64:
        # promote every tensor the forward touches; the code path is unchanged.  The default dtype is
        # switched too, because the reference builds its k-space grids from python scalars / 0-dim
        # tensors, which would otherwise stay float32 (0-dim tensors do not take part in promotion).
        torch.set_default_dtype(torch.float64)
        for name in ["opt_obja", "opt_objp", "opt_obj_tilts", "opt_slice_thickness", "opt_probe", "opt_probe_pos_shifts"]:
            p = getattr(model, name)
            p.data = p.data.double()
        model.optimizable_tensors = {k: getattr(model, "opt_" + ("obj_tilts" if k == "obj_tilts" else k)) for k in model.optimizable_tensors}
        model.omode_occu = model.omode_occu.double()
        model.H = model.H.to(torch.complex128)
        model.measurements = model.measurements.double()
        model.dx = model.dx.double()
        model.lambd = model.lambd.double()
        model.create_grids()
        model.shift_probes_grid = model.shift_probes_grid.double()
        model.init_propagator_vars()
    for k, t in model.optimizable_tensors.items():
        t.requires_grad = model.lr_params[k] != 0
    dp = model(idx)
    meas = model.get_measurements(idx)
    total, terms = loss_fn(dp, meas, model._current_object_patches, model.omode_occu)
    total.backward()
    grads = {}
    for k, t in model.optimizable_tensors.items():
        if model.lr_params[k] != 0:
            grads[k] = (t.grad if t.grad is not None else torch.zeros_like(t)).detach().numpy().astype(np.float64 if dtype == torch.float64 else np.float32)
    torch.set_default_dtype(torch.float32)
    return dict(dp=dp.detach().numpy(), losses=np.array([float(t.detach()) for t in terms]), total=float(total.detach()), grads=grads,
                roi=model.get_obj_ROI(idx).detach().numpy() if dtype == torch.float32 else None,
                probes=model.get_probes(idx).detach().numpy(), props=model.get_propagators(idx).detach().numpy())


def main():
    only = sys.argv[1:]                       # optional: names of the cases to (re)generate
    for name, (cfg, lp, extra) in cases().items():
        if only and name not in only:
            continue
        iv, mp, _ = make_inputs(cfg, seed=1234)
        if "global_tilt" in extra:
            iv["obj_tilts"] = np.array([extra["global_tilt"]], np.float32)
        rng = np.random.default_rng(7)
        Ntot = iv["crop_pos"].shape[0]
        idx = np.sort(rng.choice(Ntot, cfg.batch, replace=False)).astype(np.int64)
        r32 = run_reference(iv, mp, lp, idx, torch.float32)
        r64 = run_reference(iv, mp, lp, idx, torch.float64)
        save = dict(idx=idx, dp32=r32["dp"], dp64=r64["dp"], losses32=r32["losses"], losses64=r64["losses"],
                    roi32=r32["roi"], probes32=r32["probes"], props32=r32["props"],
                    probes64=r64["probes"], props64=r64["props"])
        for k, v in r32["grads"].items():
            save["g32_" + k] = v
        for k, v in r64["grads"].items():
            save["g64_" + k] = v
        for k, v in iv.items():
            if v is not None:
                save["iv_" + k] = np.asarray(v)
        import json
        save["model_params"] = np.array(json.dumps(mp))
        save["loss_params"] = np.array(json.dumps(lp))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
        print(f"{name:20s} dp f32-vs-f64 {rel(r32['dp'], r64['dp']):.2e}  loss {r32['total']:.6f}/{r64['total']:.6f}  " +
              " ".join(f"{k}:{rel(r32['grads'][k], r64['grads'][k]):.1e}" for k in r32["grads"]))


if __name__ == "__main__":
    main()
