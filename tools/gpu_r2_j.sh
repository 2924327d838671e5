#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gputest_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_j.log
tail -8 gpurun_out/gputest_j.log
for path in general auto; do
python bench.py --config S64 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained --path $path > gpurun_out/bench_S64_$path.json 2> gpurun_out/bench_S64_$path.err || tail -3 gpurun_out/bench_S64_$path.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench_S64_$path.json")); print("S64 $path", round(d["value"]), "patterns/s  step %.3f ms  fwd %.3f bwd %.3f" % (d["ms_per_step"], d["roofline_forward"]["ms_per_launch"], d["roofline"]["ms_per_launch"]))
PY
done
