"""GPU (>= 2 devices): the sharded step and the sharded evaluation over real NCCL, one process per GPU
(scripts/multi_parity.py under torchrun): fvx_bpr_step_sharded against the fp64 oracle, both top-k
decompositions against the oracle.  Skipped on a one-GPU box - there the same kernels are covered with emulated
ranks (tests/test_gpu_sharded.py) and the NCCL path by bench.py --gpus N (parity_checked)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_paths_over_nccl_match_the_oracle():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(REPO, "scripts", "multi_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "GREEN" in r.stdout
