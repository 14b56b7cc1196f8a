"""Coset-sharded proving on >= 2 GPUs of one box (NCCL): every rank's proof equals the C oracle's.  Skipped on a
single-GPU box; the host-side sharding logic is covered on CPU by tests/test_sharding_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_proofs_equal_oracle(world):
    if _gpus() < world:
        pytest.skip(f'needs {world} GPUs')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'shard_check.py'), str(world), '13', '8'],
                       capture_output=True, text=True, timeout=900)
    assert 'SHARD_CHECK OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
