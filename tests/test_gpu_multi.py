"""N > 1 on real GPUs (BASELINE configs[4]): two ranks over NCCL — inputs scattered from rank 0, shards compressed
on their own GPUs, packed streams gathered on rank 0 — must reproduce the single-GPU bytes.  Needs two devices."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (run through gpurun --gpus 2)")
def test_two_rank_nccl_gather_equals_single_gpu():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "4099"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_GPU_OK world=2" in r.stdout, (r.stdout + r.stderr)[-3000:]
