#!/bin/bash
# DRAM traffic of every general-path kernel of one C4 step (second iteration): light metric set, no replay
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --kernel-name regex:"k_fwd_|k_bwd_|k_init_shift" --launch-skip 66 --launch-count 66 \
    -f -o gpurun_out/general_C4 python tools/prof_step.py C4 auto > gpurun_out/prof_c4_ncu.log 2>&1
tail -3 gpurun_out/prof_c4_ncu.log
ls -la gpurun_out/general_C4.ncu-rep
