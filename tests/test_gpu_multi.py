"""Runs tests/multigpu_check.py under torchrun when the box has >= 2 GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_fit_is_bitwise_equal_to_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs (run tests/multigpu_check.py under torchrun on a multi-GPU box)')
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr',
           '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tests', 'multigpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and 'MULTIGPU_CHECK PASS' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
