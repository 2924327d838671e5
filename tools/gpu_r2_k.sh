#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gputest_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_k.log
tail -4 gpurun_out/gputest_k.log
bash tools/ab.sh build/variants/base2.so build/variants/ythoist.so build/variants/base2.so build/variants/ythoist.so > gpurun_out/ab_ythoist.txt 2>&1
cat gpurun_out/ab_ythoist.txt
python bench.py --config S64 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/bench_S64_auto2.json 2> gpurun_out/bench_S64_auto2.err || tail -3 gpurun_out/bench_S64_auto2.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench_S64_auto2.json")); print("S64 auto", round(d["value"]), "patterns/s  step %.3f ms  fwd %.3f bwd %.3f" % (d["ms_per_step"], d["roofline_forward"]["ms_per_launch"], d["roofline"]["ms_per_launch"]))
PY
