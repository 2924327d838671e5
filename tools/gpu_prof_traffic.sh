#!/bin/bash
# DRAM bytes + duration of every kernel of ONE step (the second iteration, by profiler range; light metric set, no replay) for the
# configs whose traffic is not in the --set full captures: C3, C5, S64, C1
mkdir -p gpurun_out
for cfg in "$@"; do
  PROF_RANGE=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
      -f -o gpurun_out/traffic_$cfg python tools/prof_step.py $cfg auto > gpurun_out/prof_traffic_$cfg.log 2>&1
  ncu -i gpurun_out/traffic_$cfg.ncu-rep --page raw --csv > gpurun_out/traffic_$cfg.csv 2>/dev/null; rm -f gpurun_out/traffic_$cfg.ncu-rep    # the reports exceed gpurun's 64 MiB
  tail -1 gpurun_out/prof_traffic_$cfg.log; ls -la gpurun_out/traffic_$cfg.csv
done
