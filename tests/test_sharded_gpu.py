"""Multi-GPU path (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multi_gpu_check.py
under torch.distributed.run with one rank per GPU over NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_fill_halo_trace_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi_gpu_check ok" in r.stdout


def test_group_over_real_devices_single_process():
    """sdfgpu_group_* with one device per rank, driven by one host thread (tests/group_devices_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "group_devices_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "group_devices_check ok" in r.stdout
