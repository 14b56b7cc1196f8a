"""A short slice of the mutation campaign (tests/fuzz/run.sh: host code built with AddressSanitizer + UBSan, mutated AIR blobs and
proofs in exact-size heap buffers) so that the harness keeps working and a regression in the parsers is seen in the CPU suite."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_short_sanitizer_campaign_is_clean(tmp_path):
    if not shutil.which('g++'):
        pytest.skip('no host compiler')
    asan = subprocess.run(['gcc', '-print-file-name=libasan.so'], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip('libasan is not installed')
    env = dict(os.environ, TMPDIR=str(tmp_path))
    r = subprocess.run(['bash', os.path.join(HERE, 'fuzz', 'run.sh'), '120'], capture_output=True, text=True, timeout=600, env=env)
    out = r.stdout + r.stderr
    assert 'AddressSanitizer' not in out and 'runtime error' not in out, out[-3000:]
    assert r.returncode == 0, out[-3000:]
    done = [l for l in out.splitlines() if ' done 120' in l]
    assert len(done) == 7, out[-3000:]
