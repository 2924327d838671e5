#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_o.log
tail -12 gpurun_out/gputest_o.log
