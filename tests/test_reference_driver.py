"""The drop-in claim, exercised against the reference's OWN driver code: the unmodified ``recon_step``
(reconstruction.py:658-781), ``CombinedConstraint`` (constraints.py:227-246) and ``CombinedLoss`` (losses.py:143-155) are run on
``ptyrad_b200.PtychoAD`` (CUDA) and, side by side, on the reference's own ``PtychoAD`` (CPU, float32) from the same inputs, for
several iterations including constraints that rebind ``opt_*.data``, ``start_iter`` toggling and the LBFGS closure.

The reference package comes from ``baseline/_ref`` (or /root/reference/src in the build container); see oracle/ref_import.py.
"""
import copy

import numpy as np
import pytest
import torch

from helpers import rel
from oracle.ref_import import import_reference

REF = import_reference(with_driver=True)
needs_ref = pytest.mark.skipif(REF is None, reason="reference package not importable (no baseline/_ref and no /root/reference)")


def constraint_params(Z):
    """Every key CombinedConstraint.forward reads (constraints.py:231-246); the ones exercised here rebind .data."""
    off = dict(freq=None)
    return dict(
        ortho_pmode=dict(freq=None),
        probe_mask_k=dict(freq=None, radius=0.22, width=0.05, power_thresh=0.95),
        fix_probe_int=dict(freq=1),
        obj_rblur=dict(freq=1, obj_type="both", kernel_size=5, std=0.4),
        obj_zblur=dict(freq=1, obj_type="both", kernel_size=5, std=1.0),
        kr_filter=dict(freq=None, obj_type="both", radius=0.15, width=0.05),
        kz_filter=dict(freq=None, obj_type="both", beta=1, alpha=1),
        complex_ratio=dict(freq=None, obj_type="both", alpha1=1, alpha2=0),
        mirrored_amp=dict(freq=None, relax=0.1, scale=0.03, power=4),
        obja_thresh=dict(freq=1, relax=0, thresh=[0.98 ** (1 / Z), 1.02 ** (1 / Z)]),
        objp_postiv=dict(freq=1, relax=0, mode="clip_neg"),
        tilt_smooth=dict(freq=None, std=2),
    )


def _inputs(cfg_name="T64", seed=51, **over):
    from dataclasses import replace
    from workloads import make_inputs, CONFIGS
    cfg = replace(CONFIGS[cfg_name], **over)
    iv, mp, lp = make_inputs(cfg, seed=seed)
    mp = copy.deepcopy(mp)
    mp["update_params"]["probe"]["start_iter"] = 2               # exercised by toggle_grad_requires (reconstruction.py:783-790)
    return cfg, iv, mp, lp


def _drive(model_cls, loss_cls, device, iv, mp, lp, batches, niters, optimizer="Adam", grad_accumulation=1, constraints=True, step_fn=None):
    model = model_cls(iv, mp, device=device, verbose=False)
    loss_fn = loss_cls(lp, device=device)
    Z = model.opt_obja.shape[1]
    cp = constraint_params(Z)
    if not constraints:
        cp = {k: dict(v, freq=None) for k, v in cp.items()}
    constraint_fn = REF.CombinedConstraint(cp, device=device, verbose=False)
    if optimizer == "Adam":
        opt = torch.optim.Adam(model.optimizable_params)
    else:
        opt = torch.optim.LBFGS([p for g in model.optimizable_params for p in g["params"]], lr=1.0, max_iter=3, history_size=5)
    hist = []
    for it in range(1, niters + 1):
        np.random.seed(100 + it)                                  # the LBFGS branch shuffles the batch order with np.random
        bl = (step_fn or REF.recon_step)(batches, grad_accumulation, model, opt, loss_fn, constraint_fn, it, verbose=False)
        hist.append({k: [float(x) for x in v] for k, v in bl.items()})
    params = {k: t.detach().cpu().numpy().astype(np.float64) for k, t in model.optimizable_tensors.items()}
    return hist, params, model


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("loss_side", ["reference_loss", "native_loss"])
def test_reference_recon_step_and_constraints_drive_the_cuda_model(loss_side):
    """3 iterations of the reference's recon_step (Adam, grad accumulation 2, probe starting at iteration 2, five constraints that
    rebind .data) on our CUDA model -- once with the REFERENCE's CombinedLoss (it indexes `_current_object_patches[..., 1]`, which
    LazyPatches materialises with autograd connectivity) and once with ours -- against the reference model on the CPU."""
    import ptyrad_b200
    cfg, iv, mp, lp = _inputs()
    batches = [np.arange(0, 6), np.arange(6, 12), np.arange(12, 19), np.arange(19, 25)]
    h_ref, p_ref, m_ref = _drive(REF.PtychoAD, REF.CombinedLoss, "cpu", iv, mp, lp, batches, 3, grad_accumulation=2)
    loss_cls = REF.CombinedLoss if loss_side == "reference_loss" else ptyrad_b200.CombinedLoss
    h, p, m = _drive(ptyrad_b200.PtychoAD, loss_cls, "cuda", iv, mp, lp, batches, 3, grad_accumulation=2)
    assert list(h[0].keys()) == list(lp.keys())
    for it in range(3):
        for k in lp:
            np.testing.assert_allclose(h[it][k], h_ref[it][k], rtol=3e-4, atol=1e-7, err_msg=f"iter {it + 1} {k}")
    p0 = {"obja": np.abs(iv["obj"]), "objp": np.angle(iv["obj"]), "probe": np.stack([iv["probe"].real, iv["probe"].imag], -1),
          "probe_pos_shifts": iv["probe_pos_shifts"]}
    for k in ("obja", "objp", "probe", "probe_pos_shifts"):
        assert rel(p[k], p_ref[k]) < 2e-3, (k, rel(p[k], p_ref[k]))
        assert rel(p[k] - p0[k], p_ref[k] - p0[k]) < 8e-2, (k, rel(p[k] - p0[k], p_ref[k] - p0[k]))    # the CHANGE, norm-wise
    # bookkeeping the reference's save / plot code reads (save.py:112-133)
    assert len(m.loss_iters) == 3 and len(m.iter_times) == 3 and len(m.dz_iters) == 3 and len(m.avg_tilt_iters) == 3
    assert m.opt_probe.requires_grad and m.opt_obja.is_contiguous() and m.opt_objp.is_contiguous()


@needs_ref
@pytest.mark.gpu
def test_reference_lbfgs_closure_drives_the_cuda_model():
    """The LBFGS branch of the reference's recon_step (reconstruction.py:697-735): several forwards inside one closure, one backward,
    line-search re-evaluations -- unchanged on our model; losses and object against the reference model on the CPU."""
    import ptyrad_b200
    cfg, iv, mp, lp = _inputs(seed=52)
    mp["update_params"]["probe"]["start_iter"] = 1
    batches = [np.arange(0, 8), np.arange(8, 16), np.arange(16, 25)]
    h_ref, p_ref, _ = _drive(REF.PtychoAD, REF.CombinedLoss, "cpu", iv, mp, lp, batches, 2, optimizer="LBFGS", grad_accumulation=3, constraints=False)
    h, p, _ = _drive(ptyrad_b200.PtychoAD, ptyrad_b200.CombinedLoss, "cuda", iv, mp, lp, batches, 2, optimizer="LBFGS", grad_accumulation=3, constraints=False)
    for it in range(2):
        for k in lp:
            np.testing.assert_allclose(h[it][k], h_ref[it][k], rtol=2e-3, atol=1e-7, err_msg=f"iter {it + 1} {k}")
    for k in ("obja", "objp"):
        assert rel(p[k], p_ref[k]) < 5e-3, (k, rel(p[k], p_ref[k]))


@needs_ref
@pytest.mark.gpu
def test_native_recon_step_lbfgs_branch_matches_the_reference_driver():
    """ptyrad_b200.step.recon_step with a torch.optim.LBFGS optimizer (its own closure loop, mirroring reconstruction.py:697-735)
    against the REFERENCE's recon_step on the reference model (CPU): same shuffled groups (np.random seeded alike), same losses."""
    import ptyrad_b200
    from ptyrad_b200.step import recon_step
    cfg, iv, mp, lp = _inputs(seed=54)
    mp["update_params"]["probe"]["start_iter"] = 1
    batches = [np.arange(0, 6), np.arange(6, 12), np.arange(12, 18), np.arange(18, 25)]
    h_ref, p_ref, _ = _drive(REF.PtychoAD, REF.CombinedLoss, "cpu", iv, mp, lp, batches, 2, optimizer="LBFGS", grad_accumulation=2, constraints=False)
    h, p, m = _drive(ptyrad_b200.PtychoAD, ptyrad_b200.CombinedLoss, "cuda", iv, mp, lp, batches, 2, optimizer="LBFGS", grad_accumulation=2,
                     constraints=False, step_fn=recon_step)
    for it in range(2):
        for k in lp:
            np.testing.assert_allclose(h[it][k], h_ref[it][k], rtol=2e-3, atol=1e-7, err_msg=f"iter {it + 1} {k}")
    for k in ("obja", "objp"):
        assert rel(p[k], p_ref[k]) < 5e-3, (k, rel(p[k], p_ref[k]))
    assert len(m.loss_iters) == 2 and len(m.iter_times) == 2


@needs_ref
@pytest.mark.gpu
def test_reference_ortho_pmode_and_probe_mask_on_the_cuda_model():
    """Probe constraints that go through get_complex_probe_view() and rebind opt_probe.data (constraints.py:34-81): run on our
    model, then one more step must still work (the kernels read the re-bound, possibly re-strided storage)."""
    import ptyrad_b200
    cfg, iv, mp, lp = _inputs(seed=53)
    model = ptyrad_b200.PtychoAD(iv, mp, device="cuda", verbose=False)
    cp = {k: dict(v, freq=None) for k, v in constraint_params(cfg.Z).items()}
    cp["ortho_pmode"]["freq"] = 1
    cp["probe_mask_k"]["freq"] = 1
    cp["fix_probe_int"]["freq"] = 1
    REF.CombinedConstraint(cp, device="cuda", verbose=False)(model, 1)
    pr = model.get_complex_probe_view().detach().reshape(cfg.P, -1).to(torch.complex128)
    gram = (pr @ pr.conj().T).abs().cpu().numpy()
    off = gram - np.diag(np.diag(gram))
    assert off.max() < 2e-2 * gram.max()                                   # modes stay (nearly) orthogonal after the k-space mask
    np.testing.assert_allclose(float(model.get_complex_probe_view().detach().abs().pow(2).sum()), float(model.probe_int_sum), rtol=1e-5)
    loss_fn = ptyrad_b200.CombinedLoss(lp, device="cuda")
    idx = np.arange(7)
    total, _ = loss_fn(model(idx), model.get_measurements(idx), model._current_object_patches, model.omode_occu)
    total.backward()
    assert torch.isfinite(model.opt_probe.grad).all() and float(model.opt_probe.grad.abs().sum()) > 0


@needs_ref
def test_reference_driver_runs_its_own_model_on_cpu():
    """Sanity of the harness itself (no GPU): the stubbed-import reference recon_step drives the reference model."""
    cfg, iv, mp, lp = _inputs("T32", seed=54)
    batches = [np.arange(0, 5), np.arange(5, 11)]
    h, p, m = _drive(REF.PtychoAD, REF.CombinedLoss, "cpu", iv, mp, lp, batches, 2)
    assert len(h) == 2 and all(np.isfinite(h[1][k]).all() for k in lp) and len(m.loss_iters) == 2


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("postiv_mode", ["clip_neg", "subtract_min"])
def test_native_object_constraints_match_the_reference(postiv_mode):
    """ptyrad_b200.constraints.CombinedConstraint (obj_rblur / obj_zblur / mirrored_amp / obja_thresh / objp_postiv in the CUDA library,
    in place; the rest delegated to the reference at its place in the sequence) against the reference's CombinedConstraint on a twin
    model: same parameters afterwards, and no .data re-binding for the native ones."""
    import ptyrad_b200
    from ptyrad_b200.constraints import CombinedConstraint
    cfg, iv, mp, lp = _inputs("T64", seed=55, M=2)
    rng = np.random.default_rng(3)
    iv = dict(iv)
    iv["obj"] = (iv["obj"] * np.exp(1j * rng.normal(0, 0.2, iv["obj"].shape))).astype(np.complex64)      # negative phases too
    cp = constraint_params(cfg.Z)
    cp["obj_rblur"] = dict(freq=1, obj_type="both", kernel_size=5, std=0.7)
    cp["obj_zblur"] = dict(freq=1, obj_type="phase", kernel_size=3, std=0.8)
    cp["mirrored_amp"] = dict(freq=1, relax=0.1, scale=0.03, power=4)
    cp["obja_thresh"] = dict(freq=1, relax=0.2, thresh=[0.985, 1.0])
    cp["objp_postiv"] = dict(freq=1, relax=0.1, mode=postiv_mode)
    cp["kr_filter"]["freq"] = 2                                   # a non-native one in the middle of the sequence (iteration 2)
    a = ptyrad_b200.PtychoAD(iv, mp, device="cuda", verbose=False)
    b = ptyrad_b200.PtychoAD(iv, mp, device="cuda", verbose=False)
    ref_c = REF.CombinedConstraint(cp, device="cuda", verbose=False)
    ours = CombinedConstraint(cp, device="cuda", verbose=False, fallback=REF.CombinedConstraint(cp, device="cuda", verbose=False))
    ptr_a, ptr_p = b.opt_obja.data_ptr(), b.opt_objp.data_ptr()
    for it in (1, 2):
        ref_c(a, it)
        ours(b, it)
        for name in ("opt_obja", "opt_objp", "opt_probe"):
            x, y = getattr(a, name).detach().cpu().numpy(), getattr(b, name).detach().cpu().numpy()
            assert rel(y, x) < 2e-6, (it, name, rel(y, x))
        if it == 1:
            assert (b.opt_obja.data_ptr(), b.opt_objp.data_ptr()) == (ptr_a, ptr_p)       # in place
    with pytest.raises(NotImplementedError):
        CombinedConstraint(cp, device="cuda", verbose=False)(b, 2)                       # kr_filter without a fallback
