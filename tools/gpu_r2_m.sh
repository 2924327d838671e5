#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_m.log
tail -4 gpurun_out/gputest_m.log
for nb in 1 0 1 0; do
for cfg in C2 C1; do
PTYB200_NO_BRANCHES=$nb python bench.py --config $cfg --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/bench_${cfg}_nb$nb.json 2> gpurun_out/bench_${cfg}_nb$nb.err || tail -3 gpurun_out/bench_${cfg}_nb$nb.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${cfg}_nb$nb.json")); print("$cfg no_branches=$nb", round(d["value"]), "patterns/s  step %.4f ms (eager %.4f)" % (d["ms_per_step"], d["eager"]["ms_per_step"]))
PY
done
done
