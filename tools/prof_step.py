"""One forward + loss + backward of a synthetic config (profiling target for ncu; no oracle, no timing).

    python tools/prof_step.py C4 general "batch=64, scan=128" [chunk] [pg]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import replace

import numpy as np
import torch

from ptyrad_b200 import PtychoAD, CombinedLoss
from workloads import make_inputs, CONFIGS

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
path = {"auto": 0, "general": 1, "fused": 2}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
kw = eval("dict(%s)" % sys.argv[3]) if len(sys.argv) > 3 else {}
cfg = replace(CONFIGS[name], **kw)
iv, mp, lp = make_inputs(cfg, simulate_measurements=False)
rng = np.random.default_rng(3)
idx = np.sort(rng.choice(cfg.scan ** 2, cfg.batch, replace=False)).astype(np.int64)
model = PtychoAD(iv, mp, device="cuda", verbose=False)
model.kernel_path = path
model.kernel_chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
model.kernel_pmodes_per_cta = int(sys.argv[5]) if len(sys.argv) > 5 else 0
loss_fn = CombinedLoss(lp, device="cuda")
for it in range(2):
    if it == 1 and os.environ.get("PROF_RANGE"):          # `ncu --profile-from-start off`: capture the second iteration only
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    model.zero_grad(set_to_none=True)
    dp = model(idx)
    total, terms = loss_fn(dp, model.get_measurements(idx), model._current_object_patches, model.omode_occu)
    total.backward()
torch.cuda.synchronize()
if os.environ.get("PROF_RANGE"):
    torch.cuda.profiler.stop()
print("ok", float(total))
