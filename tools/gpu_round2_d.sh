#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_d.log
tail -15 gpurun_out/gputest_d.log
python bench.py --steps 60 --warmup 5 > gpurun_out/bench_C2_d.json 2> gpurun_out/bench_C2_d.err || tail -5 gpurun_out/bench_C2_d.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_C2_ref_d.json 2> gpurun_out/bench_C2_ref_d.err || tail -5 gpurun_out/bench_C2_ref_d.err
python bench.py --config C5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C5_d.json 2> gpurun_out/bench_C5_d.err || tail -5 gpurun_out/bench_C5_d.err
cat gpurun_out/bench_C2_d.json | head -c 1500
