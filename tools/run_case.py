"""Run one synthetic case through the CUDA path and compare with the float64 oracle (debug helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import replace
import numpy as np, torch
from ptyrad_b200 import PtychoAD, CombinedLoss
from workloads import make_inputs, CONFIGS
from oracle.ptycho_torch import oracle_step

name = sys.argv[1] if len(sys.argv) > 1 else "T128"
path = {"auto": 0, "general": 1, "fused": 2}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
kw = eval("dict(%s)" % sys.argv[3]) if len(sys.argv) > 3 else {}
cfg = replace(CONFIGS[name], **kw)
iv, mp, lp = make_inputs(cfg, seed=31)
idx = np.arange(min(cfg.batch, cfg.scan ** 2), dtype=np.int64)
model = PtychoAD(iv, mp, device="cuda", verbose=False)
model.kernel_path = path
loss_fn = CombinedLoss(lp, device="cuda")
dp = model(idx)
torch.cuda.synchronize(); print("forward ok")
total, terms = loss_fn(dp, model.get_measurements(idx), model._current_object_patches, model.omode_occu)
total.backward()
torch.cuda.synchronize(); print("backward ok")
ref = oracle_step(iv, mp, lp, idx, torch.float64)
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
print("dp", rel(dp.detach().cpu().numpy(), ref["dp"]), "loss", float(total), ref["total"])
for k, g in ref["grads"].items():
    print(k, rel(model.optimizable_tensors[k].grad.cpu().numpy(), g))
