#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_n.log
tail -4 gpurun_out/gputest_n.log
for cfg in C2 C1 C5; do
python bench.py --config $cfg --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/bench_${cfg}_n.json 2> gpurun_out/bench_${cfg}_n.err || tail -3 gpurun_out/bench_${cfg}_n.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${cfg}_n.json")); print("$cfg", round(d["value"]), "patterns/s  step %.4f ms (eager %.4f) launches/step %.1f" % (d["ms_per_step"], d["eager"]["ms_per_step"], d["gpu_launches"]/d["steps"]))
PY
done
