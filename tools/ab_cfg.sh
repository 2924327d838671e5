#!/bin/bash
# tools/ab_cfg.sh <config> <batch> <lib.so> ... : time one config with several kernel-variant libraries
cfg=$1; batch=$2; shift 2
for lib in "$@"; do
  PTYB_LIB=$lib python bench.py --config $cfg --batch $batch --steps 6 --warmup 2 --no-cpu-baseline --no-e2e --no-graph > /tmp/ab.json 2> /tmp/ab.err || tail -3 /tmp/ab.err
  python - <<PY
import json; d=json.load(open("/tmp/ab.json")); print("$cfg $lib", round(d["value"]), "step %.3f fwd %.3f bwd %.3f" % (d["ms_per_step"], d["roofline_forward"]["ms_per_launch"], d["roofline"]["ms_per_launch"]))
PY
done
