"""GPU tier, boxes with >= 2 GPUs only (skipped otherwise): the NCCL side of the multi-GPU paths, run under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_gpus() < 2, reason="needs >= 2 GPUs")
def test_observed_grids_fuse_over_nccl():
    """carve mode over 2 ranks: fuse_observed (all-gather + k_or_reduce) == the grid one GPU builds from all views"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "fuse_observed_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "\n".join(l for l in (r.stdout + r.stderr).splitlines() if "rank" in l or "Error" in l or "error" in l)[-3000:]
    assert r.stdout.count("OK") == 2
