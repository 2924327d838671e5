#!/bin/bash
# final evidence of the round on ONE build: ncu captures first (traffic JSONs regenerated on the box so that the bench lines quote
# them), then the GPU suite three times with margins, smoke, and a bench line per config
mkdir -p gpurun_out
HOW="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off (second iteration)"
bash tools/gpu_prof_c2.sh
python profiles/make_traffic_json.py gpurun_out/fused128_C2.ncu-rep C2 256 auto k_backward > gpurun_out/ncu_traffic_C2.json
bash tools/gpu_prof_traffic.sh C4 C3 C5 S64 C1
python profiles/make_traffic_json.py gpurun_out/traffic_C4.csv C4 256 auto adjoint_section "$HOW" > gpurun_out/ncu_traffic_C4.json
python profiles/make_traffic_json.py gpurun_out/traffic_C3.csv C3 64 auto adjoint_section "$HOW" > gpurun_out/ncu_traffic_C3.json
python profiles/make_traffic_json.py gpurun_out/traffic_C5.csv C5 512 auto adjoint_section "$HOW" > gpurun_out/ncu_traffic_C5.json
python profiles/make_traffic_json.py gpurun_out/traffic_S64.csv S64 1024 auto k_backward "$HOW" > gpurun_out/ncu_traffic_S64.json
python profiles/make_traffic_json.py gpurun_out/traffic_C1.csv C1 32 auto k_backward "$HOW" > gpurun_out/ncu_traffic_C1.json
cp gpurun_out/ncu_traffic_*.json profiles/r02/
bash tools/gpu_margins.sh > gpurun_out/margins_final.txt 2>&1; grep -E "^run|passed|failed" gpurun_out/margins_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
python bench.py --steps 100 --warmup 5 > gpurun_out/final_C2.json 2> gpurun_out/final_C2.err || tail -3 gpurun_out/final_C2.err
python bench.py --config S64 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/final_S64.json 2> gpurun_out/final_S64.err || tail -3 gpurun_out/final_S64.err
python bench.py --config C1 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/final_C1.json 2> gpurun_out/final_C1.err || tail -3 gpurun_out/final_C1.err
python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final_C3.json 2> gpurun_out/final_C3.err || tail -3 gpurun_out/final_C3.err
python bench.py --config C4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/final_C4.json 2> gpurun_out/final_C4.err || tail -3 gpurun_out/final_C4.err
python bench.py --config C5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final_C5.json 2> gpurun_out/final_C5.err || tail -3 gpurun_out/final_C5.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_C?.json") + glob.glob("gpurun_out/final_S64.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"], 1), d["unit"], "step %.3f ms" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "traffic", d["roofline"].get("traffic"), "frac %.3f" % d["roofline"]["frac"], (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "no line", e)
PY
bash tools/gpu_launchlist.sh
du -sh gpurun_out
