#!/bin/bash
# round-2 GPU call A: full GPU test suite, DSMEM microbenchmark, C2 sanity bench, chunked general path at the C4 shape
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/gputest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_a.log
timeout 120 ./build/ubench/dsmem_bw > gpurun_out/dsmem_bw.txt 2>&1; echo "dsmem rc=$?" >> gpurun_out/dsmem_bw.txt
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_C2_a.json 2> gpurun_out/bench_C2_a.err
for v in "0 0" "4 1" "8 1" "8 3" "16 4" "32 12"; do
  set -- $v
  PTYB200_GEN_CHUNK=$1 PTYB200_GEN_PG=$2 python bench.py --config C4s --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/bench_C4s_chunk$1_pg$2.json 2> gpurun_out/bench_C4s_chunk$1_pg$2.err
done
tail -5 gpurun_out/gputest_a.log
