#!/bin/bash
# full GPU test suite + one C2 bench line (quick regression check)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_quick.log
tail -5 gpurun_out/gputest_quick.log
python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("C2", round(d["value"], 1), "step %.4f ms" % d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1), "traffic", d["roofline"]["traffic"], "launches/step", d["gpu_launches"] / d["steps"])
PY
