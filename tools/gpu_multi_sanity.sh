#!/bin/bash
# 2-GPU check of the final build: NCCL numerical test + one weak-scaling C2 line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/gputest_nccl2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_nccl2.log
tail -3 gpurun_out/gputest_nccl2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 60 --warmup 5 --no-cpu-baseline \
    > gpurun_out/bench_C2_2gpu.json 2> gpurun_out/bench_C2_2gpu.err || tail -3 gpurun_out/bench_C2_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_C2_2gpu.json').read().strip().splitlines()[-1]); print('C2 x2', round(d['value']), 'step %.3f ms' % d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks']['reasons'])"
