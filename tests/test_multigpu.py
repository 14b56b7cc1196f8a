"""Coset-sharded proving on >= 2 GPUs of one box (NCCL): every rank's proof equals the C oracle's, on the shapes
BASELINE.json names for the sharded runs -- config 4 (MiMC 2^20 steps, E = 16; examples/mimc/mimc128.ts:22-28), config 5
(Poseidon Merkle proof, 128 branches of depth 8 = 2^16 steps, 12 registers, E = 32; examples/poseidon/merkleProof.ts:25-131)
and the north-star shape (MiMC 2^20, E = 8) -- plus two quick shapes.  Skipped on a box with fewer GPUs; the host-side
sharding logic is covered on CPU by tests/test_sharding_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _check(world, *configs):
    if _gpus() < world:
        pytest.skip(f'needs {world} GPUs')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'shard_check.py'), str(world), *configs],
                       capture_output=True, text=True, timeout=1500)
    assert 'SHARD_CHECK OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_proofs_equal_oracle_small(world):
    _check(world, 'small')


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_north_star_shape(world):
    _check(world, 'ns')


@pytest.mark.parametrize('world', [2, 8])
def test_sharded_config4_mimc_2e20_e16(world):
    _check(world, '4')


@pytest.mark.parametrize('world', [2, 8])
def test_sharded_config5_poseidon_2e16(world):
    _check(world, '5')


@pytest.mark.parametrize('world', [2])
def test_sharded_config3_rescue(world):
    _check(world, '3')
