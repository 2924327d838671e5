#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_batching.py tests/test_reference_driver.py -m gpu -x -q > gpurun_out/gputest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_h.log
tail -15 gpurun_out/gputest_h.log
python - <<'PY' > gpurun_out/grouping_timing.txt 2>&1
import time, numpy as np, torch
from ptyrad_b200.batching import make_batches
for n in (64, 128, 256):
    rng = np.random.default_rng(0)
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    pos = np.stack([yy.ravel(), xx.ravel()], 1) * 7.3 + rng.normal(0, 0.4, (n * n, 2))
    for mode in ("compact", "sparse"):
        torch.cuda.synchronize(); t0 = time.time()
        b = make_batches(np.arange(n * n), pos, 256, mode=mode, verbose=False, seed=1)
        torch.cuda.synchronize()
        print(f"{n}x{n} scan, {len(b)} groups of ~256, mode {mode}: {time.time() - t0:.2f} s")
PY
cat gpurun_out/grouping_timing.txt
