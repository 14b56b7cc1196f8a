"""tests/golden/oracle_proofs.json: digests of the proofs the oracle produced when the fixtures were written
(scripts/make_golden.py).  A regression guard for the restated semantics -- NOT reference outputs (parity with genSTARK
itself is unpinned: DESIGN.md §2).  CPU: the C port reproduces every fixture, the Python restatement the small ones;
GPU (-m gpu): so does the device path."""
import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
GOLDEN = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'oracle_proofs.json')))


def _workloads():
    import make_golden
    return make_golden.workloads()


def _check(name, buf):
    g = GOLDEN[name]
    assert len(buf) == g['bytes'], name
    assert buf[:32].hex() == g['evRoot'], name
    assert hashlib.sha256(buf).hexdigest() == g['sha256'], name


def test_c_port_reproduces_every_fixture():
    from oracle import cport
    w = _workloads()
    assert set(w) == set(GOLDEN)
    for name, (air, opts, a, inputs, seed) in w.items():
        _check(name, cport.prove(air, opts, a, inputs, seed))


@pytest.mark.parametrize('name', ['mimc_64_e8_blake2s', 'mimc_64_e16_sha256', 'sponge_asm_b1_w4_e8_blake2s'])
def test_python_restatement_reproduces_the_small_fixtures(name):
    from oracle.stark import Stark as OracleStark
    air, opts, a, inputs, seed = _workloads()[name]
    st = OracleStark(air, opts)
    _check(name, st.serialize(st.prove(a, inputs, seed)))


@pytest.mark.gpu
def test_gpu_path_reproduces_every_fixture():
    from genstark_b200.stark import Stark
    for name, (air, opts, a, inputs, seed) in _workloads().items():
        _check(name, Stark(air, opts).prove_bytes(a, inputs, seed))
