#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_e.log
tail -5 gpurun_out/gputest_e.log
bash tools/ab.sh build/variants/opair0.so build/variants/opair1.so build/variants/opair0.so build/variants/opair1.so > gpurun_out/ab_opair.txt 2>&1
cat gpurun_out/ab_opair.txt
timeout 120 build/ubench/l2_bw > gpurun_out/l2_bw.txt 2>&1; cat gpurun_out/l2_bw.txt
