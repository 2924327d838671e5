#!/bin/bash
# final single-GPU evidence: bench lines of every config + ncu captures (one --set full of the fused128 kernels, launch list)
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/final_C2.json 2> gpurun_out/final_C2.err || tail -3 gpurun_out/final_C2.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_C2_reference.json 2> gpurun_out/final_C2_reference.err || tail -3 gpurun_out/final_C2_reference.err
python bench.py --config C1 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/final_C1.json 2> gpurun_out/final_C1.err || tail -3 gpurun_out/final_C1.err
python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final_C3.json 2> gpurun_out/final_C3.err || tail -3 gpurun_out/final_C3.err
python bench.py --config C4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/final_C4.json 2> gpurun_out/final_C4.err || tail -3 gpurun_out/final_C4.err
python bench.py --config C4 --scaling strong --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/final_C4strong_1gpu.json 2> gpurun_out/final_C4strong_1gpu.err || tail -3 gpurun_out/final_C4strong_1gpu.err
python bench.py --config C5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final_C5.json 2> gpurun_out/final_C5.err || tail -3 gpurun_out/final_C5.err
python bench.py --config S64 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/final_S64.json 2> gpurun_out/final_S64.err || tail -3 gpurun_out/final_S64.err
python bench.py --config C2 --path general --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/final_C2_general.json 2> gpurun_out/final_C2_general.err || tail -3 gpurun_out/final_C2_general.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"], 1), d["unit"], "step %.3f ms" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "no line", e)
PY
bash tools/gpu_prof_c2.sh
bash tools/gpu_launchlist.sh
