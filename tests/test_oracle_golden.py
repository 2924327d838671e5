"""CPU: the oracle restatement must reproduce the reference's own outputs (golden vectors made by
tests/golden/make_golden.py from the unmodified reference) before anything is compared against it."""
import numpy as np
import pytest
import torch

from helpers import golden_cases, load_golden, rel
from oracle.ptycho_torch import OracleModel, oracle_step, ddp_emulated_grads
from oracle import adjoint_np


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_f32_matches_reference_f32(name):
    z, iv, mp, lp = load_golden(name)
    r = oracle_step(iv, mp, lp, z["idx"], torch.float32)
    assert rel(r["dp"], z["dp32"]) < 2e-6
    np.testing.assert_allclose(r["losses"], z["losses32"], rtol=2e-6, atol=1e-9)
    for k, g in r["grads"].items():
        # fp32-vs-fp32: two float32 evaluation orders of the same formula; the reference's own fp32-vs-fp64
        # spread (printed by make_golden.py) is 1e-5 on object/probe, 7e-5 on shifts, 2e-4 on tilts and
        # 8e-2 on dz (k*dz*Kz loses ~5 digits in float32), so these are noise floors, not tolerances to tighten.
        tol = {"probe_pos_shifts": 5e-4, "obj_tilts": 1e-3, "slice_thickness": 0.5}.get(k, 5e-5)
        assert rel(g, z["g32_" + k]) < tol, k


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_f64_matches_reference_f64(name):
    z, iv, mp, lp = load_golden(name)
    r = oracle_step(iv, mp, lp, z["idx"], torch.float64)
    assert rel(r["dp"], z["dp64"]) < 1e-12
    # loss_simlar goes through torchvision's 5x5 blur (2-D kernel, float32 sigma handling): 1e-8 agreement
    np.testing.assert_allclose(r["losses"][:4], z["losses64"][:4], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(r["losses"][4], z["losses64"][4], rtol=1e-7, atol=1e-14)
    for k, g in r["grads"].items():
        assert rel(g, z["g64_" + k]) < (1e-9 if name != "g_simlar" else 1e-7), k


@pytest.mark.parametrize("name", golden_cases())
def test_roi_gather_bit_exact(name):
    """integer ROI addressing + float copy: bit-exact (models.py:261-264)."""
    z, iv, mp, lp = load_golden(name)
    m = OracleModel(iv, mp, torch.float32)
    a, p = m.patches(z["idx"])
    roi = z["roi32"]
    assert np.array_equal(a.detach().numpy(), roi[..., 0])
    assert np.array_equal(p.detach().numpy(), roi[..., 1])


@pytest.mark.parametrize("name", golden_cases())
def test_probes_and_propagators(name):
    z, iv, mp, lp = load_golden(name)
    m = OracleModel(iv, mp, torch.float64)
    assert rel(m.probes(z["idx"]).detach().numpy(), z["probes64"]) < 1e-12
    assert rel(m.propagators(z["idx"]).detach().numpy(), z["props64"]) < 1e-12


@pytest.mark.parametrize("name", [c for c in golden_cases() if c != "g_simlar"])
def test_hand_adjoint_matches_autograd_f64(name):
    """Appendix-A adjoint (numpy complex128, no autograd) against the reference's float64 autograd."""
    z, iv, mp, lp = load_golden(name)
    r = adjoint_np.step(iv, mp, lp, z["idx"])
    assert rel(r["dp"], z["dp64"]) < 1e-12
    np.testing.assert_allclose(r["losses"], z["losses64"], rtol=1e-11, atol=1e-14)
    for k in r["grads"]:
        if "g64_" + k in z.files:
            assert rel(r["grads"][k], z["g64_" + k]) < 1e-9, k


def test_invariants():
    """Docstring-level invariants of the reference: zero tilt => propagator == H;
    ortho norm => sum(dp) == sum|probe|^2 for a unit-amplitude object."""
    z, iv, mp, lp = load_golden("g_base")
    m = OracleModel(iv, mp, torch.float64)
    assert rel(m.propagators([0]).numpy(), iv["H"][None]) < 1e-7
    iv2 = dict(iv)
    iv2["obj"] = np.exp(1j * np.angle(iv["obj"])).astype(np.complex64)
    m2 = OracleModel(iv2, mp, torch.float64)
    dp, _ = m2.forward(z["idx"])
    tot = dp.sum(dim=(-2, -1)).detach().numpy()
    np.testing.assert_allclose(tot, np.sum(np.abs(iv["probe"].astype(np.complex128)) ** 2), rtol=2e-6)


def test_ddp_emulation_differs_from_full_batch():
    """Per-rank loss normalisation makes k-GPU gradients != 1-GPU gradients."""
    z, iv, mp, lp = load_golden("g_base")
    full = oracle_step(iv, mp, lp, z["idx"], torch.float64)["grads"]
    ddp = ddp_emulated_grads(iv, mp, lp, z["idx"], 2, torch.float64)
    assert rel(ddp["objp"], full["objp"]) > 1e-3
