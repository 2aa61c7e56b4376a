"""Multi-GPU checks (need >= 2 GPUs on the box; skipped otherwise): fused peer-memory optimiser vs the NCCL path."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_data_parallel_optimizer_two_gpus():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_WORKER_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    print([line for line in out.stdout.splitlines() if "MULTI_GPU_WORKER_OK" in line][0])
