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


def _run(cmd, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e)


@pytest.mark.skipif(_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["push", "epilogue", "nccl"])
def test_sweep_single_process_all_gpus(exchange):
    """dmf_comm_init_all: ONE process drives every GPU through the C ABI (the reference drivers' shape): gathered rows, set cover,
    fused observed grids and fused marks equal the single-GPU results; fused peer stores and the NCCL path"""
    r = _run([sys.executable, os.path.join(ROOT, "tools", "comm_check.py"), "all", str(min(_gpus(), 8))], {"DMF_COMM_EXCHANGE": exchange})
    assert r.returncode == 0 and "OK single process" in r.stdout, (r.stdout + r.stderr)[-3000:]
    assert {"push": "push kernel", "epilogue": "epilogue", "nccl": "ncclAllGather"}[exchange] in r.stdout, r.stdout


@pytest.mark.skipif(_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["push", "epilogue", "nccl"])
def test_sweep_one_process_per_gpu(exchange):
    """dmf_comm_init_rank under torchrun: arenas mapped with CUDA IPC, rows pushed by the march kernels (or ncclAllGather)"""
    n = min(_gpus(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "comm_check.py"), "rank"]
    r = _run(cmd, {"DMF_COMM_EXCHANGE": exchange})
    assert r.returncode == 0, "\n".join(l for l in (r.stdout + r.stderr).splitlines() if "rank" in l or "rror" in l or "ssert" in l)[-3000:]
    assert r.stdout.count("OK rank") == n
