"""Sweep the general path's chunk size / probe modes per CTA on one config (tuning helper for api.cu: gen_plan).

    python tools/sweep_general.py C4 256 "256:12 8:0 4:0 4:2 3:0 2:0"      # config, batch, list of chunk:pg (0 = heuristic)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ptyrad_b200 import PtychoAD, CombinedLoss, _lib
from ptyrad_b200.optim import FusedAdam
from ptyrad_b200.step import GradArena, recon_batch
from workloads import CONFIGS, make_inputs, random_batches

name, B = sys.argv[1], int(sys.argv[2])
combos = [tuple(int(v) for v in c.split(":")) for c in sys.argv[3].split()]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfg = CONFIGS[name]
t0 = time.time()
iv, mp, lp = make_inputs(cfg, simulate_measurements=(cfg.scan <= 64 and cfg.N <= 128))
dev = torch.device("cuda", 0)
model = PtychoAD(iv, mp, device=dev, verbose=False)
model.kernel_path = _lib.PATH_GENERAL
loss_fn = CombinedLoss(lp, device=dev)
opt = FusedAdam(model.optimizable_params)
arena = GradArena(model)
batches = [torch.as_tensor(b[:B], device=dev) for b in random_batches(iv["crop_pos"].shape[0], B, seed=7)[:8]]
print(f"setup {time.time() - t0:.1f} s", file=sys.stderr)
lib = _lib.lib()
for chunk, pg in combos:
    model.kernel_chunk, model.kernel_pmodes_per_cta = chunk, pg
    for s in range(2):
        recon_batch(model, loss_fn, opt, batches[s % len(batches)], arena, 1)
    torch.cuda.synchronize()
    l0 = lib.ptyb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        recon_batch(model, loss_fn, opt, batches[s % len(batches)], arena, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"cfg": name, "B": B, "chunk": chunk, "pg": pg, "ms_per_step": round(ms, 3), "patterns_per_s": round(B / ms * 1e3, 1),
                      "launches_per_step": (lib.ptyb200_launch_count() - l0) // steps}), flush=True)
