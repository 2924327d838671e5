#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_p.log
tail -6 gpurun_out/gputest_p.log
for ns in 1 "" 1 ""; do
PTYB200_NO_SPLIT=$ns timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-sustained > gpurun_out/bench_C2_split$ns.json 2> gpurun_out/bench_C2_split$ns.err || tail -3 gpurun_out/bench_C2_split$ns.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_C2_split$ns.json")); print("C2 no_split='$ns'", round(d["value"]), "patterns/s  step %.4f ms (eager %.4f) e2e %.0f" % (d["ms_per_step"], d["eager"]["ms_per_step"], d["e2e"]["value"]))
PY
done
