"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host mirror reproduces the reference's
flag / parameter-group logic, the product path refuses to run without CUDA, and the data-parallel exchange works over gloo."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from ptyrad_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "ptyrad_b200.h")).read()
    declared = set(re.findall(r"\b(ptyb200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ptyb200_cfg", "ptyb200_loss_cfg", "ptyb200_stream"}
    h = built.lib()
    for name in sorted(declared):
        assert hasattr(h, name), f"{name} declared in include/ptyrad_b200.h but not exported"
    assert set(built.EXPORTED_SYMBOLS) == declared
    assert h.ptyb200_abi_version() == built.ABI_VERSION


def test_struct_layout_matches_header(built):
    import ctypes as C
    assert C.sizeof(built.Cfg) == 16 * 4 + 4 * 4
    assert C.sizeof(built.LossCfg) == 13 * 4


def test_workspace_query_and_errors_without_gpu(built):
    import ctypes as C
    from ptyrad_b200 import engine
    cfg = engine.make_cfg(128, 6, 1, 8, 370, 370, 4096, 1, 0, 0, 0.1494, 0.0418)
    n = built.lib().ptyb200_workspace_bytes(C.byref(cfg), 256)
    stash = 256 * 6 * 8 * 128 * 128 * 8
    assert stash < n < 2 * stash
    bad = engine.make_cfg(128, 6, 1, 8, 370, 370, 4096, 1, 0, 0, 0.1494, 0.0418)
    bad.N = 100
    assert built.lib().ptyb200_workspace_bytes(C.byref(bad), 4) == 0
    assert b"unsupported N" in built.lib().ptyb200_last_error()
    with pytest.raises(ValueError):
        engine.make_cfg(100, 1, 1, 1, 200, 200, 10, 0, 0, 0, 0.1, 0.02)


def test_general_path_cut_changes_the_workspace_and_entry_points_refuse_bad_arguments(built):
    """cfg.reserved[2] (samples per chunk) sizes the pass buffers of the general path; the blur and backward entry points validate
    their arguments before touching the device (no GPU needed)."""
    import ctypes as C
    from ptyrad_b200 import engine
    lib = built.lib()
    cfg = engine.make_cfg(256, 12, 1, 16, 1186, 1186, 65536, 1, 0, 0, 0.1494, 0.0418)
    whole = lib.ptyb200_workspace_bytes(C.byref(cfg), 256)
    cfg.reserved[2] = 4
    cut = lib.ptyb200_workspace_bytes(C.byref(cfg), 256)
    tile = 12 * 256 * 256 * 8
    assert whole - cut == 2 * (256 - 4) * tile                 # G1 and G2 shrink from the batch to one chunk
    assert lib.ptyb200_gaussian_blur5(None, None, None, 1, 8, 8, 1.0, 0, None) != 0
    assert b"NULL" in lib.ptyb200_last_error()
    one = C.c_void_p(16); two = C.c_void_p(32); three = C.c_void_p(48)
    assert lib.ptyb200_gaussian_blur5(one, two, three, 1, 2, 8, 1.0, 0, None) != 0      # H < 3: reflect padding of 2 impossible
    assert lib.ptyb200_gaussian_blur5(one, two, three, 1, 8, 8, 0.0, 0, None) != 0      # sigma must be positive
    assert lib.ptyb200_gaussian_blur5(one, one, three, 1, 8, 8, 1.0, 0, None) != 0      # aliasing
    assert b"distinct" in lib.ptyb200_last_error()


def test_grad_arena_layout_and_direct_step_eligibility():
    """GradArena: every .grad is a 256-byte-aligned view into one flat buffer (the kernels write gradients with 8/16-byte vector
    accesses), frozen tensors get grad=None; the autograd-free step is chosen only for configurations it covers."""
    from ptyrad_b200 import CombinedLoss
    from ptyrad_b200.step import GradArena, direct_step_eligible
    from workloads import default_loss_params
    m, iv, mp, lp = _model(tilt_each=True, lr_tilts=1e-4, lr_dz=1e-4, lr_shifts=1e-4)
    arena = GradArena(m)
    base = arena.flat.data_ptr()
    seen = 0
    for p, v in zip(arena.params, arena.views):
        assert (v.data_ptr() - base) % 256 == 0 and v.shape == p.shape
        assert p.grad is not None and p.grad.data_ptr() == v.data_ptr()
        seen += p.numel()
    assert seen <= arena.flat.numel() < seen + 64 * len(arena.params) + 64
    m.opt_probe.requires_grad = False
    arena.attach()
    assert m.opt_probe.grad is None and m.opt_obja.grad is not None
    loss = CombinedLoss(lp, device="cpu")
    assert direct_step_eligible(m, loss, arena, 1, True, None)
    assert not direct_step_eligible(m, loss, None, 1, True, None)            # needs the arena as the kernels' output buffers
    assert not direct_step_eligible(m, loss, arena, 2, True, None)           # gradient accumulation adds: autograd path
    simlar = default_loss_params("single"); simlar["loss_simlar"]["state"] = True
    assert not direct_step_eligible(m, CombinedLoss(simlar, device="cpu"), arena, 1, True, None)
    m.detector_blur_std = 1.0
    assert direct_step_eligible(m, loss, arena, 1, True, None)               # detector blur: native blur + adjoint on dp / G
    m.detector_blur_std = None
    m.obj_preblur_std = 1.0
    assert not direct_step_eligible(m, loss, arena, 1, True, None)           # pre-blurred ROIs go through autograd (patch mode)
    m.obj_preblur_std = None
    with torch.no_grad():
        assert not direct_step_eligible(m, loss, arena, 1, True, None)


def _model(name="T32", **kw):
    from dataclasses import replace
    from ptyrad_b200 import PtychoAD
    from workloads import make_inputs, CONFIGS
    iv, mp, lp = make_inputs(replace(CONFIGS[name], **kw), seed=5)
    return PtychoAD(iv, mp, device="cpu", verbose=False), iv, mp, lp


def test_model_surface_matches_reference_contract():
    """Attributes / methods that recon_step, CombinedConstraint, save_results and plot_forward_pass touch."""
    m, iv, mp, lp = _model()
    for name in ["opt_obja", "opt_objp", "opt_obj_tilts", "opt_slice_thickness", "opt_probe", "opt_probe_pos_shifts"]:
        assert isinstance(getattr(m, name), torch.nn.Parameter)
    for name in ["omode_occu", "H", "measurements", "N_scan_slow", "N_scan_fast", "crop_pos", "slice_thickness", "dx", "dk", "lambd"]:
        assert name in dict(m.named_buffers())
    assert m.crop_pos.dtype == torch.int32 and m.opt_probe.shape[-1] == 2
    for name in ["scan_affine", "tilt_obj", "shift_probes", "change_thickness", "probe_int_sum", "detector_blur_std", "obj_preblur_std",
                 "optimizable_tensors", "optimizable_params", "optimizer_params", "start_iter", "lr_params", "loss_iters", "iter_times",
                 "dz_iters", "avg_tilt_iters"]:
        assert hasattr(m, name), name
    assert list(m.optimizable_tensors) == ["obja", "objp", "obj_tilts", "slice_thickness", "probe", "probe_pos_shifts"]
    assert m.get_complex_probe_view().dtype == torch.complex64
    # lr == 0 -> no requires_grad and no param group (models.py:199-206)
    assert not m.opt_obj_tilts.requires_grad and not m.opt_slice_thickness.requires_grad
    assert len(m.optimizable_params) == 4 and {g["lr"] for g in m.optimizable_params} == {5e-4, 1e-4}
    # .data can be rebound (constraints do this) and requires_grad toggled (reconstruction.py:783-790)
    m.opt_objp.data = m.opt_objp.data.clamp(min=0).contiguous()
    m.opt_probe.requires_grad = False
    with pytest.raises(ValueError):
        m.create_optimizable_params_dict({"nonsense": 1e-3}, verbose=False)


def test_flags_follow_reference_rules():
    m, *_ = _model(lr_shifts=0.0)
    assert m.shift_probes is False and m._tilt_mode() == 0           # lr 0 -> stored shifts are not applied (models.py:120)
    m, *_ = _model(tilt_each=True)
    assert m.tilt_obj is True and m._tilt_mode() == 2                # non-zero stored tilts switch the tilted propagator on
    m, *_ = _model(lr_tilts=1e-4)
    assert m.tilt_obj is True and m._tilt_mode() == 1
    m, *_ = _model(lr_dz=1e-4)
    assert m.change_thickness is True


def test_helper_getters_match_oracle_on_cpu():
    from oracle.ptycho_torch import OracleModel
    m, iv, mp, lp = _model(tilt_each=True, lr_tilts=1e-4)
    o = OracleModel(iv, mp, torch.float32)
    idx = np.array([0, 3, 7])
    torch.testing.assert_close(m.get_probes(idx), o.probes(idx).detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(m.get_propagators(idx), o.propagators(idx).detach(), rtol=1e-5, atol=1e-6)
    a, p = o.patches(idx)
    roi = m.get_obj_ROI(idx)
    assert torch.equal(roi[..., 0], a.detach()) and torch.equal(roi[..., 1], p.detach())
    assert torch.equal(m.get_measurements(idx), m.measurements[torch.as_tensor(idx)])
    assert m.get_propagated_probe([0]).shape == (iv["obj"].shape[1], *iv["probe"].shape)


def test_hot_path_refuses_cpu_tensors():
    m, *_ = _model()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(np.array([0, 1]))


def test_validation_rejects_bad_inputs():
    from ptyrad_b200 import PtychoAD
    from workloads import make_inputs
    iv, mp, lp = make_inputs("T32", seed=5)
    bad = dict(iv); bad["crop_pos"] = iv["crop_pos"].copy(); bad["crop_pos"][0] = [10000, 0]
    with pytest.raises(ValueError, match="canvas"):
        PtychoAD(bad, mp, device="cpu", verbose=False)
    bad = dict(iv); bad["probe"] = iv["probe"][:, :30, :30]
    with pytest.raises(ValueError):
        PtychoAD(bad, mp, device="cpu", verbose=False)


def test_loss_cfg_and_shard_indices():
    from ptyrad_b200 import engine
    from ptyrad_b200.step import shard_indices
    from workloads import default_loss_params
    l = engine.make_loss_cfg(default_loss_params("poissn"))
    assert (l.single_state, l.poissn_state, l.pacbed_state, l.sparse_state) == (0, 1, 0, 1)
    assert abs(l.poissn_eps - 1e-6) < 1e-12 and l.sparse_order == 1.0
    idx = np.arange(11)
    parts = [shard_indices(idx, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), idx) and max(map(len, parts)) - min(map(len, parts)) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from ptyrad_b200 import PtychoAD
    from ptyrad_b200.step import GradArena, shard_indices
    from workloads import make_inputs
    from oracle.ptycho_torch import oracle_step, ddp_emulated_grads
    iv, mp, lp = make_inputs("T32", seed=5)
    model = PtychoAD(iv, mp, device="cpu", verbose=False)
    arena = GradArena(model)
    batch = np.array([0, 2, 5, 9, 11, 17, 20])
    mine = shard_indices(batch, rank, world)
    # the device kernels cannot run here: the per-rank gradients come from the oracle; what is under test is the
    # exchange (one flat all-reduce, 1/world) and that it reproduces what DDP computes in the reference
    g = oracle_step(iv, mp, lp, mine, torch.float64)["grads"]
    arena.zero()
    for name, p in model.optimizable_tensors.items():
        if p.requires_grad:
            p.grad.copy_(torch.as_tensor(g[name], dtype=torch.float32))
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(arena.params, arena.views))
    arena.allreduce(world)
    want = ddp_emulated_grads(iv, mp, lp, batch, world, torch.float64)
    err = {n: float(np.linalg.norm(model.optimizable_tensors[n].grad.numpy() - want[n]) / np.linalg.norm(want[n])) for n in want}
    # frozen tensors drop out of the arena views (grad None) like zero_grad(set_to_none=True)
    model.opt_probe.requires_grad = False
    arena.attach()
    frozen_ok = model.opt_probe.grad is None
    q.put((rank, err, frozen_ok))
    dist.destroy_process_group()


def test_gradient_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, frozen_ok in res:
        assert frozen_ok
        assert all(v < 1e-6 for v in err.values()), (rank, err)


def test_fft_index_algebra_on_host(tmp_path):
    """The register DFT templates and the two-stage row FFT (incl. the register-fed outer stages the general kernels use) are
    __host__ __device__: compile them for the host and check them against a double-precision DFT (no GPU needed)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "test_fft_host")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc_host", "test_fft_host.cu")
    subprocess.run([nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "ALL OK" in r.stdout


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under ptyrad_b200/ may import it (a product path through the oracle would void
    every parity claim)."""
    pkg = os.path.join(ROOT, "ptyrad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "adjoint_np" not in txt and "ptycho_torch" not in txt, f


def test_chunked_and_split_step_eligibility():
    """A chunked / split step needs dL/dI = (scalar of the batch sums) x (per-pixel term): exactly one of loss_single / loss_poissn,
    no loss_pacbed, no detector blur, no tilt / thickness gradients (ptyrad_b200/step.py)."""
    from ptyrad_b200.losses import CombinedLoss
    from ptyrad_b200.step import chunked_step_eligible, split_step_eligible
    from workloads import default_loss_params
    m, iv, mp, lp = _model(lr_shifts=1e-4)
    assert chunked_step_eligible(m, CombinedLoss(lp, device="cpu"))
    both = default_loss_params("single"); both["loss_poissn"]["state"] = True
    assert not chunked_step_eligible(m, CombinedLoss(both, device="cpu"))
    pac = default_loss_params("single"); pac["loss_pacbed"]["state"] = True
    assert not chunked_step_eligible(m, CombinedLoss(pac, device="cpu"))
    assert chunked_step_eligible(m, CombinedLoss(default_loss_params("poissn"), device="cpu"))
    m.detector_blur_std = 1.0
    assert not chunked_step_eligible(m, CombinedLoss(lp, device="cpu"))
    m.detector_blur_std = None
    assert not split_step_eligible(m, CombinedLoss(lp, device="cpu"), 64)      # N = 32: not the one-tile-per-SM kernels
    m2, _, _, lp2 = _model(tilt_each=True, lr_tilts=1e-4, lr_dz=1e-4)
    assert not chunked_step_eligible(m2, CombinedLoss(lp2, device="cpu"))      # tilt / thickness gradients


def test_patch_planes_handle():
    """PatchPlanes (the pre-blurred ROI planes handed out as model._current_object_patches): `[..., 0]` / `[..., 1]` are the planes
    themselves, anything else sees the stacked (B,omode,Nz,Ny,Nx,2) tensor of the reference (models.py:264,284)."""
    from ptyrad_b200.models import PatchPlanes
    a, p = torch.rand(2, 1, 3, 4, 4), torch.rand(2, 1, 3, 4, 4)
    h = PatchPlanes(a, p)
    assert h[..., 0] is a and h[..., 1] is p
    assert h.shape == (2, 1, 3, 4, 4, 2)
    assert torch.equal(h[0, 0, 1], torch.stack([a, p], -1)[0, 0, 1])
    assert torch.equal(h.permute(5, 0, 1, 2, 3, 4)[1], p)


def test_round2_entry_points_validate_before_touching_the_device(built):
    """Argument checks of the entry points added in round 2 (chunked steps, grouping, ROI blur, loss_simlar): every refusal happens
    before the first launch, so it can be exercised without a GPU."""
    import ctypes as C
    from ptyrad_b200 import engine
    lib = built.lib()
    cfg = engine.make_cfg(128, 6, 2, 8, 370, 370, 4096, 1, 0, 0, 0.1494, 0.0418)
    lcfg = engine.make_loss_cfg({"loss_single": dict(state=True, weight=1.0, dp_pow=0.5), "loss_poissn": dict(state=True, weight=1.0, dp_pow=1.0, eps=1e-6),
                                 "loss_pacbed": dict(state=False, weight=0.5, dp_pow=0.2), "loss_sparse": dict(state=False, weight=0.1, ln_order=1),
                                 "loss_simlar": dict(state=False, weight=0.1, obj_type="both", scale_factor=[1, 1, 1], blur_std=1)})
    p = lambda v: C.c_void_p(v)
    # unscaled loss gradient / loss_scale: exactly one separable data term
    assert lib.ptyb200_loss_grad(C.byref(cfg), C.byref(lcfg), p(256), p(512), p(768), 4, None, None, None, p(1024), None, None, None) != 0
    assert b"exactly one" in lib.ptyb200_last_error()
    assert lib.ptyb200_loss_scale(C.byref(cfg), C.byref(lcfg), 4, p(256), p(512), p(768), None) != 0
    # chunked completion: no tilt / thickness gradients, no patch mode
    assert lib.ptyb200_backward_finish(C.byref(cfg), 4, p(256), p(512), p(768), None, None, None, None, built.NEED_TILTS, None, None) != 0
    assert b"tilt" in lib.ptyb200_last_error()
    patch = engine.make_cfg(128, 6, 2, 8, 128, 128, 4096, 1, 0, 0, 0.1494, 0.0418)
    patch.reserved[1] = 1
    assert lib.ptyb200_backward_zero(C.byref(patch), 4, p(256), None, None, built.NEED_OBJ, None) != 0
    assert lib.ptyb200_accumulators_add(C.byref(patch), 4, p(256), p(512), built.NEED_OBJ, None) != 0
    assert b"patch mode" in lib.ptyb200_last_error()
    # grouping: one seed per group first
    assert lib.ptyb200_sparse_groups(p(256), 3, 5, p(512), None) != 0
    assert lib.ptyb200_sparse_groups(None, 8, 2, p(512), None) != 0
    # ROI blur: scratch required when blurring; simlar: 2 <= M <= 8 and pooled size within the input
    assert lib.ptyb200_roi_blur(C.byref(cfg), p(256), 4, p(512), p(768), p(1024), 1.0, None, p(2048), p(4096), None) != 0
    assert b"tmp" in lib.ptyb200_last_error()
    assert lib.ptyb200_roi_blur(C.byref(cfg), p(256), 4, p(512), p(768), p(1024), -1.0, p(64), p(2048), p(4096), None) != 0
    one = engine.make_cfg(128, 6, 1, 8, 370, 370, 4096, 1, 0, 0, 0.1494, 0.0418)
    assert lib.ptyb200_simlar_forward(C.byref(one), 4, p(256), p(512), 8, 128, 128, 0.1, p(768), None) != 0
    assert b"object modes" in lib.ptyb200_last_error()
    assert lib.ptyb200_simlar_forward(C.byref(cfg), 4, p(256), p(512), 9, 128, 128, 0.1, p(768), None) != 0
    assert b"pooled size" in lib.ptyb200_last_error()
    # the workspace layout follows the batch capacity of a chunked step
    cfg.reserved[0] = 256
    assert lib.ptyb200_workspace_bytes(C.byref(cfg), 64) == lib.ptyb200_workspace_bytes(C.byref(cfg), 256)


def test_traffic_stamps_are_well_formed_and_the_stamp_maker_reads_a_launch_csv(tmp_path):
    """bench.py quotes `roofline.traffic` from profiles/<round>/ncu_traffic_<cfg>.json only for the build of the kernels the capture
    was taken from (content hash of csrc + the header) and for the workload's batch.  The committed stamps must name their dominant
    kernel, carry the workload's batch, and profiles/make_traffic_json.py must turn an `ncu --page raw --csv` launch list into one
    (per-kernel first launch; general-path launches summed per section)."""
    import glob
    import json
    import subprocess
    import bench
    from workloads import CONFIGS
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stamps = sorted(glob.glob(os.path.join(root, "profiles", bench.PROFILE_ROUND, "ncu_traffic_*.json")))
    assert stamps, "no traffic stamps committed for this round"
    for p in stamps:
        d = json.load(open(p))
        cfg = os.path.basename(p)[len("ncu_traffic_"):-len(".json")]
        assert d["batch"] == CONFIGS[cfg].batch and d["path"] == "auto", p
        k = d["kernels"][d["dominant"]]
        assert k["dram_read_gb"] + k["dram_write_gb"] > 0 and k["duration_ms"] > 0, p
        assert re.fullmatch(r"[0-9a-f]{16}", d["csrc_sha256"]), p
    assert re.fullmatch(r"[0-9a-f]{16}", bench.csrc_hash())
    # the stamp maker on a hand-made launch list: two general-path launches per section + one unrelated kernel
    csv_path = tmp_path / "launches.csv"
    rows = [('"ID"', '"Kernel Name"', '"dram__bytes_read.sum"', '"dram__bytes_write.sum"', '"gpu__time_duration.sum"'),
            ('""', '""', '"Gbyte"', '"Mbyte"', '"us"'),
            ('"0"', '"void ptyb::k_fwd_da<(int)256>(ptyb::Args)"', '"1.5"', '"500"', '"250"'),
            ('"1"', '"void ptyb::k_fwd_bc<(int)256>(ptyb::Args)"', '"0.5"', '"250"', '"150"'),
            ('"2"', '"void ptyb::k_bwd_da<(int)256>(ptyb::Args)"', '"2"', '"1000"', '"400"'),
            ('"3"', '"void ptyb::k_bwd_bc<(int)256>(ptyb::Args)"', '"1"', '"0"', '"100"'),
            ('"4"', '"ptyb::k_adam(ptyb::AdamArgs)"', '"0.01"', '"10"', '"5"')]
    csv_path.write_text("\n".join(",".join(r) for r in rows) + "\n")
    out = subprocess.run([sys.executable, os.path.join(root, "profiles", "make_traffic_json.py"), str(csv_path), "C4", "256", "auto",
                          "adjoint_section", "hand-made"], capture_output=True, text=True, check=True).stdout
    d = json.loads(out)
    assert d["csrc_sha256"] == bench.csrc_hash() and d["dominant"] == "adjoint_section" and d["source"].startswith("hand-made")
    a, f = d["kernels"]["adjoint_section"], d["kernels"]["forward_section"]
    assert a["launches"] == 2 and abs(a["dram_read_gb"] - 3.0) < 1e-9 and abs(a["dram_write_gb"] - 1.0) < 1e-9 and abs(a["duration_ms"] - 0.5) < 1e-9
    assert f["launches"] == 2 and abs(f["dram_read_gb"] - 2.0) < 1e-9 and abs(f["dram_write_gb"] - 0.75) < 1e-9 and abs(f["duration_ms"] - 0.4) < 1e-9
    assert abs(d["kernels"]["k_adam"]["dram_write_gb"] - 0.01) < 1e-9
