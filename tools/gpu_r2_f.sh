#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_f.log
tail -5 gpurun_out/gputest_f.log
bash tools/ab.sh build/variants/base.so build/variants/tma8.so build/variants/tma4.so build/variants/base.so build/variants/tma8.so build/variants/tma4.so > gpurun_out/ab_tma_stash.txt 2>&1
cat gpurun_out/ab_tma_stash.txt
