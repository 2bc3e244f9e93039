"""Multi-GPU test of the m-split transform (needs >= 2 GPUs; skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_msplit_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "dist", "msplit_check.py"), "8", "48", "256"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MSPLIT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_msplit_peer_stores_two_gpus():
    """The fused form (Legendre kernel stores into the peers' receive buffers over NVLink, CUDA IPC
    mappings, one stream-ordered barrier): bit-identical to one GPU, both buffers exercised."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(root, "tests", "dist", "msplit_check.py"), "--p2p", "8", "48", "256"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MSPLIT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_multi_plane_pipeline_two_gpus():
    """Shell-sharded multi-plane recurrence (dist.multi_plane_block): state handed over NCCL,
    every kappa_i bit-identical to the single-GPU recurrence."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(root, "tests", "dist", "multiplane_check.py"), "64", "7"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTIPLANE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
