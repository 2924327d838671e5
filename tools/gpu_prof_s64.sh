#!/bin/bash
# one `ncu --set full` capture of the fused64 kernels at the S64 shape (second iteration)
mkdir -p gpurun_out
python tools/prof_step.py S64 auto > gpurun_out/prof_s64_plain.log 2>&1 || { tail -5 gpurun_out/prof_s64_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_forward|k_backward" --launch-skip 2 --launch-count 2 \
    -f -o gpurun_out/fused64_S64 python tools/prof_step.py S64 auto > gpurun_out/prof_s64_ncu.log 2>&1
tail -3 gpurun_out/prof_s64_ncu.log
