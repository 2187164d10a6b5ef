"""GPU (>= 2 devices): batch-sharded inference and NCCL data-parallel gradients against single-GPU runs
(tests/dist_check.py under torchrun).  Skipped on one-GPU boxes; profiles/r02_dist_check.txt holds a 2-GPU run."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_inference_and_nccl_gradients_match_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0 and "DIST_CHECK ok" in res.stdout, (res.stdout + res.stderr)[-3000:]
