"""One process per GPU over NCCL: slab-decomposed trajectories equal the single-GPU ones bit for bit.
Needs >= 2 GPUs on the box (skipped otherwise; the in-process multi-slab test in test_gpu_parity.py and the gloo
plumbing test in test_host_logic.py cover the same logic on one GPU / on CPU)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_two_ranks_over_nccl_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "scripts", "mgpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
