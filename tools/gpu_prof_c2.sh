#!/bin/bash
# one `ncu --set full` capture of the two fused kernels at the C2 shape (second iteration) + launch list of one bench step
mkdir -p gpurun_out
python tools/prof_step.py C2 auto > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_forward|k_backward" --launch-skip 2 --launch-count 2 \
    -f -o gpurun_out/fused128_C2 python tools/prof_step.py C2 auto > gpurun_out/prof_ncu.log 2>&1
tail -3 gpurun_out/prof_ncu.log
ls -la gpurun_out/*.ncu-rep
