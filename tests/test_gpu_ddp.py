"""Multi-GPU (>= 2 devices) check of the ray-sharded gradient exchange: the peer-memory all-reduce + Adam kernel against NCCL
and against single-rank training on the concatenated batch (tools/ddp_check.py under torchrun).  Skipped on one GPU."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_memory_allreduce_adam_matches_nccl_and_single_rank():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DDP_CHECK OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
