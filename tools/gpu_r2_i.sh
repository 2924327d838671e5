#!/bin/bash
mkdir -p gpurun_out
PTYB_LIB=build/variants/roibulk.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gputest_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_i.log
tail -8 gpurun_out/gputest_i.log
bash tools/ab.sh build/variants/base2.so build/variants/roibulk.so build/variants/base2.so build/variants/roibulk.so > gpurun_out/ab_roibulk.txt 2>&1
cat gpurun_out/ab_roibulk.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k simlar > gpurun_out/gputest_i2.log 2>&1; tail -3 gpurun_out/gputest_i2.log
