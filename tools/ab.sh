#!/bin/bash
# A/B kernel variants: tools/ab.sh <lib.so> ...   (each built with different -D flags into build/variants/)
for lib in "$@"; do
  PTYB_LIB=$lib python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-graph > /tmp/ab.json 2> /tmp/ab.err || tail -3 /tmp/ab.err
  python - <<PY
import json; d=json.load(open("/tmp/ab.json")); print("$lib", round(d["value"]), "step %.3f fwd %.3f bwd %.3f" % (d["ms_per_step"], d["roofline_forward"]["ms_per_launch"], d["roofline"]["ms_per_launch"]))
PY
done
