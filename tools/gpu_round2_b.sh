#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_b.log
timeout 120 ./build/ubench/dsmem_bw > gpurun_out/dsmem_bw.txt 2>&1; echo "dsmem rc=$?" >> gpurun_out/dsmem_bw.txt
bash tools/ab.sh build/variants/opack.so build/variants/oasync_chs8.so build/variants/oasync_chs4.so build/variants/oasync_chs16.so build/variants/opack.so build/variants/oasync_chs8.so > gpurun_out/ab_oasync.txt 2>&1
tail -5 gpurun_out/gputest_b.log; cat gpurun_out/ab_oasync.txt
