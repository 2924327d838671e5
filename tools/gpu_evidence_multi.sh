#!/bin/bash
# multi-GPU evidence: NCCL numerical test (2 ranks) + weak / strong scaling lines at N = $1 GPUs
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/gputest_nccl2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_nccl2.log
  tail -6 gpurun_out/gputest_nccl2.log
fi
run() { # name, extra args
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" \
      > gpurun_out/bench_${name}_${N}gpu.json 2> gpurun_out/bench_${name}_${N}gpu.err || tail -3 gpurun_out/bench_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${name}_${N}gpu.json").read().strip().splitlines()[-1])
    print("${name} x$N:", round(d["value"]), d["unit"], "step %.3f ms" % d["ms_per_step"], d["scaling"], d.get("clocks", {}).get("reasons"))
except Exception as e:
    print("${name} x$N: no line", e)
PY
}
run C2 --steps 60 --warmup 5 --no-cpu-baseline
run C4 --config C4 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e
run C4strong --config C4 --scaling strong --steps 6 --warmup 3 --no-cpu-baseline --no-e2e
