#!/bin/bash
# general path: slice sequence on L2-sized chunks of the batch (PTYB200_GEN_CHUNK) -- sweep at C4 / C3 / C5
mkdir -p gpurun_out; out=gpurun_out/chunk_sweep.txt; : > $out
run() {  # config chunk
  PTYB200_GEN_CHUNK=$2 timeout 300 python bench.py --config $1 --steps ${3:-6} --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/cs.json 2> gpurun_out/cs.err || { echo "$1 chunk=$2 FAILED" >> $out; tail -2 gpurun_out/cs.err >> $out; return; }
  python - "$1" "$2" >> $out <<'PY'
import json, sys
d = json.loads(open("gpurun_out/cs.json").read().strip().splitlines()[-1])
print(sys.argv[1], "chunk=%s" % sys.argv[2], "%.1f patterns/s" % d["value"], "step %.3f ms" % d["ms_per_step"], "fwd %.3f bwd %.3f" % (d["roofline_forward"]["ms_per_launch"], d["roofline"]["ms_per_launch"]), d["clocks"]["reasons"])
PY
}
for c in 0 2 4 8 16 32; do run C4 $c; done
for c in 0 8 16; do run C5 $c; done
for c in 0 4; do run C3 $c 3; done
cat $out
