"""Pins the plain-C oracle port to the Python restatement: identical serialized proofs."""
import pytest

from genstark_b200 import airs
from oracle import cport
from oracle.stark import Stark as OracleStark


@pytest.mark.parametrize('steps,e,alg', [(64, 8, 'blake2s256'), (64, 16, 'sha256'), (256, 8, 'sha256'), (1024, 16, 'blake2s256')])
def test_c_port_matches_python_oracle(steps, e, alg):
    opts = dict(hashAlgorithm=alg, extensionFactor=e, exeQueryCount=48, friQueryCount=24)
    air = airs.mimc128(steps)
    ctl = airs.run_mimc(steps, airs.mimc_round_constants(), 3)
    a = [dict(step=0, register=0, value=ctl[0]), dict(step=steps - 1, register=0, value=ctl[-1])]
    ora = OracleStark(air, opts)
    want = ora.serialize(ora.prove(a, [], [3]))
    got = cport.prove(air, opts, a, [], [3])
    assert got == want
    got1 = cport.prove(air, opts, a, [], [3], threads=1)
    assert got1 == want


def test_c_port_rejects_bad_assertion():
    opts = dict(hashAlgorithm='sha256', extensionFactor=8, exeQueryCount=48, friQueryCount=24)
    air = airs.mimc128(64)
    with pytest.raises(RuntimeError, match='conflicts with execution trace'):
        cport.prove(air, opts, [dict(step=0, register=0, value=4)], [], [3])


@pytest.mark.parametrize('log_t,log_n', [(3, 3), (6, 6), (5, 8), (10, 13), (12, 12)])
def test_c_transform_equals_python_restatement(log_t, log_n):
    """oracle_transform (what the large-size K1 parity tests compare with) against oracle/field.py"""
    import random
    from oracle.field import PrimeField
    from genstark_b200.air import P128
    OF = PrimeField(P128)
    r = random.Random(log_t * 100 + log_n)
    v = [r.randrange(P128) for _ in range(1 << log_t)]
    raw = b''.join(x.to_bytes(16, 'little') for x in v)
    got = cport.transform(raw, log_t, log_n)
    dom = OF.get_power_series(OF.get_root_of_unity(1 << log_n), 1 << log_n)
    assert [int.from_bytes(got[i:i + 16], 'little') for i in range(0, len(got), 16)] == OF.eval_polys_at_roots([v], dom)[0]
    if log_t == log_n:
        assert cport.transform(got, log_t, log_t, True) == raw
