#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_final.log
tail -4 gpurun_out/gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
python bench.py --steps 100 --warmup 5 > gpurun_out/final_C2.json 2> gpurun_out/final_C2.err || tail -3 gpurun_out/final_C2.err
python bench.py --config S64 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/final_S64.json 2> gpurun_out/final_S64.err || tail -3 gpurun_out/final_S64.err
python - <<'PY'
import json
for f in ("final_C2", "final_S64"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), "step %.4f ms" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "fwd %.3f bwd %.3f" % (d["roofline_forward"]["ms_per_launch"], d["roofline"]["ms_per_launch"]))
PY
bash tools/gpu_prof_c2.sh
bash tools/gpu_launchlist.sh
