#!/usr/bin/env python
"""Benchmark of the multislice reconstruction step (BASELINE.json metric: diffraction patterns / s for a full
forward + loss + backward + optimizer iteration; % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]
                    [--scaling weak|strong --global-batch G]

One "step" = one batch of the workload through the hot path: zero-grad, forward, loss, adjoint, (gradient all-reduce),
Adam step -- the body of the reference's recon_step loop (reconstruction.py:741-770).  N > 1 is launched by torchrun, one
rank per GPU; scan positions shard across ranks: every rank holds ONLY its own block of the measurements and
draws its batches from it.  `--scaling weak` (default): every rank runs a full `batch` of its own; `--scaling strong`: the
global batch (`--global-batch`, default 2048) is fixed and split over the ranks.
The workload can also be chosen with the environment variable PTYB_BENCH_CONFIG (for drivers that pass no --config), e.g.
PTYB_BENCH_CONFIG=C4 for the 256x256-scan configuration.

Printed JSON (rank 0, one line): see README / DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "diffraction patterns/sec (fwd+bwd iter)"
UNIT = "patterns/s"
PROFILE_ROUND = "r02"


def bytes_per_pattern(c):
    """compulsory = N^2(16 M Z + 4): ROI (a,phi) read + ROI gradient write + measurement read;
    stash adds 16 P M Z N^2 (write + re-read of psi_z for the adjoint)."""
    comp = c.N * c.N * (16 * c.M * c.Z + 4)
    return comp, comp + 16 * c.P * c.M * c.Z * c.N * c.N


def flop_per_pattern(c):
    """n_fft = 2 P M (2Z-1) + 2P tile FFTs at 5 N^2 log2(N^2), plus ~ (6 mul-flops) pointwise complex work per FFT input."""
    n_fft = 2 * c.P * c.M * (2 * c.Z - 1) + 2 * c.P
    return n_fft * (5 * c.N * c.N * np.log2(c.N * c.N) + 8 * c.N * c.N)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured burst copy bandwidth (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def csrc_hash():
    """Content hash of the kernel sources: ncu-derived constants (DRAM traffic per launch) are only quoted for the build they
    were captured from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "ptyrad_b200", "csrc")
    for f in sorted(os.listdir(d)) + ["../../include/ptyrad_b200.h"]:
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(iv, mp, lp, batch_cpu, steps, warmup, threads):
    """The reference's CPU torch path timed on the host cores: the UNMODIFIED reference package (PtychoAD(device='cpu') +
    CombinedLoss + backward + Adam.step) when it is importable (baseline/_ref, or /root/reference/src in the
    build container; oracle/ref_import.py), else the oracle port of the same sequence.  Returns (patterns/s, median s/step, kind)."""
    import torch
    from workloads import random_batches
    torch.set_num_threads(threads)
    batches = [b[:batch_cpu] for b in random_batches(iv["crop_pos"].shape[0], batch_cpu, seed=99)]
    kind = "port"
    step = None
    try:
        from oracle.ref_import import import_reference
        ref = import_reference()
        if ref is not None:
            model = ref.PtychoAD(iv, mp, device="cpu", verbose=False)
            loss_fn = ref.CombinedLoss(lp, device="cpu")
            opt = torch.optim.Adam(model.optimizable_params)

            def step(idx):                                       # the non-LBFGS body of recon_step (reconstruction.py:738-772)
                opt.zero_grad()
                dp = model(idx)
                meas = model.get_measurements(idx)
                total, _ = loss_fn(dp, meas, model._current_object_patches, model.omode_occu)
                total.backward()
                opt.step()
                model.clear_cache()
            step(batches[0])                                     # proves the package really runs here before we commit to it
            kind = "reference"
    except Exception as e:                                       # e.g. a torchvision mismatch on the box: fall back to the port
        sys.stderr.write(f"reference package not usable ({type(e).__name__}: {e}); timing the oracle port instead\n")
        step = None
    if step is None:
        from oracle.ptycho_torch import OracleTrainer
        tr = OracleTrainer(iv, mp, lp, threads=threads)
        step = tr.step
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        step(batches[s % len(batches)])
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return batch_cpu / statistics.median(times), statistics.median(times), kind


def rank_block(Ntot, rank, world):
    """Scan positions whose measurements rank `rank` holds: a contiguous block of the scan (sizes differ by at most one)."""
    edges = np.linspace(0, Ntot, world + 1).astype(np.int64)
    return np.arange(edges[rank], edges[rank + 1], dtype=np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default=os.environ.get("PTYB_BENCH_CONFIG", "C2"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch size (weak scaling)")
    ap.add_argument("--scaling", default=os.environ.get("PTYB_BENCH_SCALING", "weak"), choices=["weak", "strong"])
    ap.add_argument("--global-batch", type=int, default=2048, help="strong scaling: the fixed global batch, split over the ranks")
    ap.add_argument("--chunk", type=int, default=-1, help="samples per internal chunk of a step (bounds the wave stash; the step still "
                    "yields the loss / gradients of the whole batch); -1: the config's batch when the per-GPU batch is larger, 0: off")
    ap.add_argument("--path", default="auto", choices=["auto", "general", "fused"])
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"], help="fused: one-launch Adam of this repo; torch: torch.optim.Adam (foreach)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=3.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from workloads import CONFIGS, make_inputs
    cfg = CONFIGS[args.config]
    if args.scaling == "strong":
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        B = args.global_batch // world
    else:
        B = args.batch or cfg.batch
    chunk = args.chunk if args.chunk >= 0 else (cfg.batch if B > cfg.batch else 0)
    threads = os.cpu_count() or 1
    workload = (f"{cfg.name}: {cfg.P} probe modes, {cfg.M} object modes, {cfg.Z} slices, {cfg.N}^2 patterns, "
                f"{cfg.scan}x{cfg.scan} scan, batch {B}/GPU, Adam, loss_{cfg.loss}+sparse")
    config = {"workload": workload, "cfg": cfg.name, "N": cfg.N, "P": cfg.P, "M": cfg.M, "Z": cfg.Z, "scan": cfg.scan,
              "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
              "l2": "per-step working set (wave stash) >> 126 MB L2, no explicit flush"}
    if chunk and B > chunk:
        config["chunk"] = chunk                                   # the step runs `chunk` samples at a time (one chunk's stash alive)
    simulate = cfg.scan <= 64 and cfg.N <= 128

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        iv, mp, lp = make_inputs(cfg, simulate_measurements=simulate)
        b_cpu = min(cfg.batch, 64 if cfg.N <= 128 else 8)
        k, w = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))          # bounded: the whole run ends within minutes
        rate, med, kind = cpu_reference_rate(iv, mp, lp, b_cpu, k, w, threads)
        sample = (f"{b_cpu} patterns/step of the {cfg.name} workload (same model, reduced batch), {w} warm-up + {k} timed steps, "
                  f"median step time; {'unmodified reference package' if kind == 'reference' else 'oracle port of the reference path'}")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": w,
            "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm
    # the library normally arrives built (driver's build()); on a box without it local rank 0 compiles, the others wait for the file
    from ptyrad_b200 import _lib as _libmod
    if not os.path.exists(_libmod.LIB_PATH):
        if local_rank == 0:
            from ptyrad_b200.build import build_library
            build_library()
        else:
            t_wait = time.time()
            while not os.path.exists(_libmod.LIB_PATH):
                if time.time() - t_wait > 900:
                    raise SystemExit(f"{_libmod.LIB_PATH} did not appear within 15 min")
                time.sleep(1.0)
    import torch
    import torch.distributed as dist
    from ptyrad_b200 import PtychoAD, CombinedLoss, MeasurementView, _lib
    from ptyrad_b200.step import GradArena, GraphedStep, recon_batch

    # libraries (NCCL's version banner) may write to fd 1: keep the real stdout for the one JSON line, send the rest to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Ntot = cfg.scan * cfg.scan
    mine_pos = rank_block(Ntot, rank, world)                  # the scan positions whose patterns this rank holds
    if simulate:
        if world > 1:
            # the simulated measurements cost ~30 s of host time with all cores: rank 0 makes them once (torchrun pins every rank to
            # one OpenMP thread) and the other ranks map the node-local file and copy out their own block
            shared = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp",
                                  f"ptyb200_meas_{cfg.name}_{os.environ.get('MASTER_PORT', '0')}.npy")
            if rank == 0:
                torch.set_num_threads(threads)
                iv, mp, lp = make_inputs(cfg, simulate_measurements=True)
                np.save(shared + ".tmp.npy", iv["measurements"])
                os.replace(shared + ".tmp.npy", shared)
            dist.barrier()
            if rank != 0:
                iv, mp, lp = make_inputs(cfg, measurements=np.ascontiguousarray(np.load(shared, mmap_mode="r")[mine_pos]), positions=mine_pos)
            dist.barrier()
            if rank == 0:
                os.remove(shared)
                iv["measurements"] = np.ascontiguousarray(iv["measurements"][mine_pos])
                iv["measurements_positions"] = mine_pos
        else:
            iv, mp, lp = make_inputs(cfg, simulate_measurements=True)
    else:
        iv, mp, lp = make_inputs(cfg, simulate_measurements=False, positions=mine_pos if world > 1 else None)
    model = PtychoAD(iv, mp, device=dev, verbose=False)
    model.kernel_path = {"auto": _lib.PATH_AUTO, "general": _lib.PATH_GENERAL, "fused": _lib.PATH_FUSED}[args.path]
    loss_fn = CombinedLoss(lp, device=dev)
    if args.optimizer == "fused":
        from ptyrad_b200.optim import FusedAdam
        opt = FusedAdam(model.optimizable_params)
    else:
        opt = torch.optim.Adam(model.optimizable_params)
    config["optimizer"] = "Adam (" + args.optimizer + ")"
    config["measurements"] = f"{len(mine_pos)} of {Ntot} patterns resident per GPU (own block of the scan)" if world > 1 else "all resident"
    arena = GradArena(model)
    # every rank draws its batches from a seeded permutation of ITS block of the scan
    perm = np.random.default_rng(8 + rank).permutation(mine_pos)
    nbat = max(1, len(perm) // B)
    if len(perm) < B:
        raise SystemExit(f"rank {rank} holds {len(perm)} positions, fewer than its batch {B}")
    batches = [np.sort(perm[i * B:(i + 1) * B]) for i in range(nbat)]
    my = [torch.as_tensor(batches[i % nbat], device=dev) for i in range(max(4, min(nbat, 64)))]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    lib = _lib.lib()
    eager_fn = lambda ix: recon_batch(model, loss_fn, opt, ix, arena, world, chunk=chunk)

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        out = None
        for s in range(steps):
            out = fn(my[s % len(my)])
        e1.record()
        sync_all()
        return max_over_ranks(e0.elapsed_time(e1)), out

    for s in range(args.warmup):
        eager_fn(my[s % len(my)])
    sync_all()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    # pass 1 (eager launches): per-section CUDA events inside the library + launch count
    import ctypes as C
    lib.ptyb200_timing_enable(1)
    l0 = lib.ptyb200_launch_count()
    ms_eager, last = timed(eager_fn, args.steps)
    launches = lib.ptyb200_launch_count() - l0
    tf, tb, nf, nb = C.c_double(), C.c_double(), C.c_int(), C.c_int()
    lib.ptyb200_timing_read(C.byref(tf), C.byref(tb), C.byref(nf), C.byref(nb))
    lib.ptyb200_timing_enable(0)
    ms = ms_eager
    step_fn = eager_fn
    if not args.no_graph:
        # pass 2 (headline): the identical step replayed as a CUDA graph (same kernels, no per-launch host cost)
        step_fn = GraphedStep(model, loss_fn, opt, arena, B, world=world, chunk=chunk)
        for s in range(args.warmup):
            step_fn(my[s % len(my)])
        ms, last = timed(step_fn, args.steps)
        config["cuda_graph"] = True
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps * B * world / (ms * 1e-3)

    # ------------------------------------------------------------------ sustained: >= 3 s of back-to-back steps with its own clock record
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(math.ceil(args.sustained_seconds * 1e3 / (ms / args.steps))))
        s2 = ClockSampler(local_rank).start() if rank == 0 else None
        ms_sus, _ = timed(step_fn, n_sus)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"value": n_sus * B * world / (ms_sus * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "clocks": c2}

    # ------------------------------------------------------------------ end-to-end: host buffers, H2D + D2H every step
    e2e = None
    if not args.no_e2e:
        k2 = min(args.steps, 40)
        nn = cfg.N * cfg.N
        row_of = {int(p): i for i, p in enumerate(mine_pos)} if world > 1 else None
        host_meas, host_idx = [], []
        for i in range(min(8, len(my))):
            ix = my[i % len(my)].cpu().numpy()
            rows = ix if row_of is None else np.array([row_of[int(p)] for p in ix])
            host_meas.append(torch.from_numpy(np.ascontiguousarray(iv["measurements"][rows])).pin_memory())
            host_idx.append(torch.from_numpy(ix.astype(np.int64)).pin_memory())
        dev_meas = torch.empty((B, cfg.N, cfg.N), dtype=torch.float32, device=dev)
        dev_idx = torch.empty(B, dtype=torch.int64, device=dev)
        ar = torch.arange(B, device=dev)
        host_loss = torch.empty(5, dtype=torch.float32).pin_memory()
        if args.no_graph:
            def e2e_step(j):
                dev_meas.copy_(host_meas[j], non_blocking=True)
                dev_idx.copy_(host_idx[j], non_blocking=True)
                return recon_batch(model, loss_fn, opt, dev_idx, arena, world, measurements=MeasurementView(dev_meas, ar), chunk=chunk)
        else:
            g2 = GraphedStep(model, loss_fn, opt, arena, B, world=world, stream_measurements=True, chunk=chunk)
            e2e_step = lambda j: g2(host_idx[j], host_meas[j])       # H2D of indices + patterns into the graph's static inputs
        for s in range(3):
            e2e_step(s % len(host_meas))
        sync_all()
        nh = len(host_meas)
        if args.no_graph:
            t0 = time.perf_counter()
            for s in range(k2):
                l5 = e2e_step(s % nh)
                host_loss.copy_(l5, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        else:
            # every step still copies its own inputs host -> device and its result device -> host inside the timed region;
            # the copy of batch s+1 runs on a side stream while batch s computes (what a streaming data loader does)
            t0 = time.perf_counter()
            g2.prefetch(host_idx[0], host_meas[0])
            for s in range(k2):
                l5 = g2.step_prefetched()
                g2.prefetch(host_idx[(s + 1) % nh], host_meas[(s + 1) % nh])
                host_loss.copy_(l5, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        sync_all()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": k2 * B * world / dt, "unit": UNIT, "h2d_bytes_per_step": B * nn * 4 + B * 8, "d2h_bytes_per_step": 20,
               "steps": k2, "api": "PtychoAD.forward + CombinedLoss + backward + Adam.step via ptyrad_b200.step.recon_batch"}

    def finish():
        # captured NCCL graphs can dead-lock destroy_process_group(); everything is synchronised and printed by now, so leave
        # through os._exit after a last barrier
        sys.stdout.flush(); sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    peak, peak_src = measured_peaks()
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this workload (per launch); quoted only when
    # the capture was taken from THIS build of the kernels (content hash of ptyrad_b200/csrc) and this batch size
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", PROFILE_ROUND, f"ncu_traffic_{cfg.name}.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
        except ValueError:
            tj = {}
        if tj.get("csrc_sha256") == csrc_hash() and tj.get("batch") == B and tj.get("path", "auto") == args.path:
            kb = tj["kernels"][tj["dominant"]]
            traffic, traffic_src = (kb["dram_read_gb"] + kb["dram_write_gb"]) * 1e9, f"profiles/{PROFILE_ROUND}/" + os.path.basename(tpath)
        else:
            traffic_src = f"profiles/{PROFILE_ROUND}/{os.path.basename(tpath)} is from another build / batch of the kernels: not quoted"
    comp, stash = bytes_per_pattern(cfg)
    # adjoint section (dominant): re-reads the stash, reads the ROIs, read-modify-writes the ROI gradients, reads G
    Bl = chunk if (chunk and B > chunk) else B            # samples per launch of the section (a chunked step launches it per chunk)
    bwd_bytes = Bl * (8 * cfg.P * cfg.M * cfg.Z + 16 * cfg.M * cfg.Z + 4) * cfg.N * cfg.N
    fwd_bytes = Bl * (8 * cfg.P * cfg.M * cfg.Z + 8 * cfg.M * cfg.Z + 4) * cfg.N * cfg.N
    ms_b = tb.value / max(1, nb.value)
    ms_f = tf.value / max(1, nf.value)
    ach_b = bwd_bytes / (ms_b * 1e-3) / 1e9 if ms_b > 0 else 0.0
    flops = flop_per_pattern(cfg)
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config, "clocks": clocks, "sustained": sustained, "e2e": e2e, "gpu_launches": int(launches),
        "eager": {"value": args.steps * B * world / (ms_eager * 1e-3), "ms_per_step": ms_eager / args.steps,
                  "note": "same steps launched kernel by kernel; section timings and gpu_launches come from this pass"},
        "roofline": {"bound": "hbm", "kernel": "multislice adjoint section (ptyb200_backward)", "achieved": ach_b, "peak": peak,
                     "unit": "GB/s", "frac": ach_b / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bwd_bytes, "ms_per_launch": ms_b,
                     "share_of_step": ms_b / (ms_eager / args.steps) if ms_eager > 0 else None},
        "roofline_forward": {"bound": "hbm", "achieved": fwd_bytes / (ms_f * 1e-3) / 1e9 if ms_f > 0 else 0.0, "peak": peak, "unit": "GB/s",
                             "ms_per_launch": ms_f},
        "roofline_step": {"hbm_frac_stash_convention": value / world * stash / (peak * 1e9),
                          "hbm_frac_compulsory": value / world * comp / (peak * 1e9),
                          "fp32_tflops": value / world * flops / 1e12, "fp32_peak_nominal_tflops": fp32_peak,
                          "fp32_frac": value / world * flops / 1e12 / fp32_peak, "bytes_per_pattern_stash": stash,
                          "flop_per_pattern": flops},
        "losses_last_step": [float(x) for x in last.cpu()],
    }
    if world == 1 and not args.no_cpu_baseline:
        # like-for-like GPU baseline: the same torch-eager restatement of the reference (cuFFT + ATen + autograd + Adam) on this
        # B200, full batch, resident inputs (SURVEY 8d asks for it next to the CPU number)
        try:
            from oracle.ptycho_torch import OracleTrainer
            del step_fn
            torch.cuda.empty_cache()
            tg = OracleTrainer(iv, mp, lp, device=dev)
            for s_ in range(3):
                tg.step(my[s_ % len(my)])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ng = 10
            for s_ in range(ng):
                tg.step(my[s_ % len(my)])
            torch.cuda.synchronize()
            dtg = (time.perf_counter() - t0) / ng
            out["torch_cuda_eager_baseline"] = {"value": B / dtg, "unit": UNIT, "ms_per_step": dtg * 1e3, "kind": "port",
                                                "sample": f"{ng} steps of the full {B}-pattern batch, torch eager on the same GPU"}
            del tg
        except Exception as e:           # never let the extra baseline break the contract line
            out["torch_cuda_eager_baseline"] = {"error": str(e)[:200]}
        b_cpu = min(B, 64 if cfg.N <= 128 else 8)
        rate, med, kind = cpu_reference_rate(iv, mp, lp, b_cpu, 6, 1, threads)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": f"{b_cpu} patterns/step of the same workload, 1 warm-up + 6 timed steps, median ({med * 1e3:.0f} ms/step); "
                                         f"{'unmodified reference package' if kind == 'reference' else 'oracle port of the reference path'}"}
    emit(out)
    finish()


if __name__ == "__main__":
    main()
