"""Numerical test of the multi-GPU path proper: CUDA kernels + one NCCL all-reduce (eager and inside a captured CUDA graph) against
the DDP emulation of the float64 oracle.  Needs >= 2 GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py
-m gpu`); skipped otherwise.  The CPU/gloo twin of the arena arithmetic is tests/test_host_cpu.py::test_gradient_exchange_world2_gloo.
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_two_ranks_match_ddp_emulation():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0, "multi-GPU worker reported a mismatch (see output above)"
    assert "FAIL" not in r.stdout
