#!/bin/bash
# A/B: first CTA barrier of the forward FFT removed (PTYB_FUSED_BARRIER_A=0), stash words loaded before the last register DFT (F128_EARLY_PS)
mkdir -p gpurun_out; out=gpurun_out/ab_barrier.txt; : > $out
V=build/variants
for rep in 1 2; do bash tools/ab.sh $V/base.so $V/noA.so $V/noA_eps2.so $V/noA_eps4.so $V/noA_eps8.so >> $out 2>&1; done
for lib in $V/base.so $V/noA.so $V/base.so $V/noA.so; do
  PTYB_LIB=$lib python bench.py --config S64 --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > /tmp/ab.json 2> /tmp/ab.err || tail -3 /tmp/ab.err
  python - >> $out <<PY
import json; d=json.load(open("/tmp/ab.json")); print("S64 $lib", round(d["value"]), "step %.3f fwd %.3f bwd %.3f" % (d["ms_per_step"], d["roofline_forward"]["ms_per_launch"], d["roofline"]["ms_per_launch"]))
PY
done
PTYB_LIB=$V/noA_eps4.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_noA.log 2>&1; tail -2 gpurun_out/gputest_noA.log >> $out
cat $out
