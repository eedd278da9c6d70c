"""2-GPU checks of the sharded path on real hardware (skipped on single-GPU boxes):
``gpurun --gpus 2 -- python -m pytest tests -m gpu -k multigpu`` runs them.  The CPU-side (gloo, world_size 2) coverage
of the same orchestration is in tests/test_host_logic_cpu.py."""
import os
import subprocess
import sys

import pytest


@pytest.mark.gpu
def test_multigpu_sharded_lloyd_and_flux_match_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "multigpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert "ALL CHECKS PASSED: True" in out.stdout
