#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked" > gpurun_out/gputest_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_l.log
tail -12 gpurun_out/gputest_l.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_l2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_l2.log
tail -4 gpurun_out/gputest_l2.log
python bench.py --config C2 --batch 1024 --chunk 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/bench_C2_b1024_chunk256.json 2> gpurun_out/bench_C2_b1024_chunk256.err || tail -3 gpurun_out/bench_C2_b1024_chunk256.err
python bench.py --config C2 --batch 1024 --chunk 0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/bench_C2_b1024_chunk0.json 2> gpurun_out/bench_C2_b1024_chunk0.err || tail -3 gpurun_out/bench_C2_b1024_chunk0.err
python - <<PY
import json
for n in ("chunk256","chunk0"):
    d=json.load(open("gpurun_out/bench_C2_b1024_%s.json"%n)); print("C2 B=1024", n, round(d["value"]), "patterns/s  step %.3f ms" % d["ms_per_step"])
PY
