#!/bin/bash
# launch list of one C2 step (cold-cache, serialised: compare shares)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/ll_plain.json 2> gpurun_out/ll_plain.err || { tail -3 gpurun_out/ll_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_C2.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/ll_ncu.log 2>&1
tail -2 gpurun_out/ll_ncu.log | cut -c1-300
