#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_g.log
tail -5 gpurun_out/gputest_g.log
bash tools/ab.sh build/variants/base.so build/variants/fold.so build/variants/base.so build/variants/fold.so > gpurun_out/ab_fold.txt 2>&1
cat gpurun_out/ab_fold.txt
for cfg in C4 C5; do
PTYB_LIB=build/variants/base.so python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_fold_${cfg}_base.json 2>/dev/null
PTYB_LIB=build/variants/fold.so python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_fold_${cfg}_fold.json 2>/dev/null
python - <<PY
import json
for v in ("base","fold"):
    d=json.load(open("gpurun_out/ab_fold_${cfg}_%s.json"%v)); print("$cfg", v, round(d["value"]), d["ms_per_step"])
PY
done
