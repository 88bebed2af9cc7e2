"""Multi-GPU (>= 2 B200) test of the peer-fused per-Gaussian backward: launches tools/peers_check.py under torchrun and
checks the fused reduction against the sum of the per-view oracle gradients and against backward_gaussians + NCCL all-reduce,
one view and several views per rank.  Skipped on single-GPU boxes (the
host-side sharding logic is covered on CPU by tests/test_distributed_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [[], ["odd"]])
def test_peer_fused_backward_matches_nccl_allreduce(extra):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tools", "peers_check.py")] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all ranks ok = True" in r.stdout
