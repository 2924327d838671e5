#!/bin/bash
# the GPU suite three times with every gradient check's (error, bound) recorded: shows how far each check sits from its bound
# and how much the atomics' summation order moves it from run to run
mkdir -p gpurun_out; rm -f gpurun_out/test_margins.jsonl
for i in 1 2 3; do
  PTYB200_MARGINS_FILE=$PWD/gpurun_out/test_margins.jsonl timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gputest_margins_$i.log 2>&1
  echo "run $i rc=$?"; tail -2 gpurun_out/gputest_margins_$i.log
done
python - <<'PY'
import json, collections
w = collections.defaultdict(list)
for l in open("gpurun_out/test_margins.jsonl"):
    d = json.loads(l); w[(d["test"], d["label"], d["tensor"], d["tol"])].append(d["err"])
rows = sorted(((max(v) / k[3], k, min(v), max(v), len(v)) for k, v in w.items()), reverse=True)
for r, k, lo, hi, n in rows[:25]:
    print("%.2f of bound  %s %s %s  min %.2e max %.2e  n=%d" % (r, k[0].split("::")[-1], k[1], k[2], lo, hi, n))
PY
