"""Data-parallel training step across TWO GPUs (NCCL), run under torchrun as a subprocess: gradients after the bucketed
all-reduce equal the single-GPU gradients of the concatenated batch (rel-RMS <= 1e-2), ranks that were
initialised differently hold identical parameters after the constructor's broadcast and one optimizer step
(tools/ddp_check.py). Skipped on boxes with fewer than two GPUs; the CPU-side bucket logic is covered with gloo in
tests/test_host_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_train_step_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "ddp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    print(res.stdout[-2000:])
    assert res.returncode == 0, res.stderr[-3000:]
    assert "parameters identical on all ranks after the step: True" in res.stdout
