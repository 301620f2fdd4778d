"""GPU, needs >= 2 devices (skipped on a 1-GPU box): the fused all-gather decode under torchrun -- several steps with
different latents per step and skewed ranks -- must equal decode + NCCL all_gather_into_tensor bit for bit, with the
double-buffered symmetric allocation and with the single-buffer + leading-barrier variant."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,P", [(8, 4099), (3, 129)])
def test_fused_gather_multi_step_two_ranks(T, P):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "tools", "gather_probe.py"), str(T), str(P), "6"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    assert "mismatches=0" in p.stdout
