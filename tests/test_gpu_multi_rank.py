"""Two ranks, two GPUs, the real library: the cooperative echelon form, RREF and kernel basis must be
byte-identical to the single-GPU ones and the cooperative run must have moved bytes through NCCL.  Skipped on a box
with fewer than two GPUs (the CPU tier covers the slicing rule with gloo, tests/test_multi_rank.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_give_the_single_gpu_results_and_use_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    last = [l for l in out.stdout.splitlines() if l.startswith("MULTI-GPU PARITY")]
    assert last and last[-1].startswith("MULTI-GPU PARITY OK"), out.stdout[-2000:]
    assert "nccl_bytes_rank0=0" not in last[-1]
