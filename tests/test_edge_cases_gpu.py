"""Edge cases of the prover path: the reference's error behaviour and unusual shapes, GPU vs oracle."""
import pytest

import cases
from genstark_b200 import airs
from genstark_b200.air import AirModule, ProgramBuilder, StaticRegister, P128
from genstark_b200.stark import Stark, StarkError
from oracle.stark import Stark as OracleStark

pytestmark = pytest.mark.gpu


def _same_bytes(air, opts, a, inputs, seed):
    gpu = Stark(air, opts)
    got = gpu.prove_bytes(a, inputs, seed)
    ora = OracleStark(air, opts)
    assert got == ora.serialize(ora.prove(a, inputs, seed))
    assert gpu.verify(a, got)
    return got


@pytest.mark.parametrize('e', [2, 4, 8, 16, 32])
def test_extension_factor_sweep_linear_air(e):
    """Fibonacci over p128 (degree-1 constraints: no degree adjustment, composition degree = T, delta = 0)."""
    steps = 128
    air = airs.fibonacci(steps, modulus=P128)
    tr = [1, 1]
    rows = [(1, 1)]
    for _ in range(steps - 1):
        a2 = (tr[0] + tr[1]) % P128
        tr = [a2, (tr[1] + a2) % P128]
        rows.append(tuple(tr))
    a = [dict(step=0, register=0, value=1), dict(step=0, register=1, value=1), dict(step=steps - 1, register=1, value=rows[-1][1])]
    opts = dict(hashAlgorithm='sha256', extensionFactor=e, exeQueryCount=30, friQueryCount=12)
    _same_bytes(air, opts, a, [], [1, 1])


def test_query_count_is_capped_by_the_domain():
    """getExeIndexes caps the count at N - N/E (QueryIndexGenerator.ts:21): T=64, E=2 -> at most 64 positions."""
    air, opts, a, inputs, seed = cases.mimc(64, 8)
    air2 = airs.fibonacci(64, modulus=P128)
    opts2 = dict(hashAlgorithm='blake2s256', extensionFactor=2, exeQueryCount=128, friQueryCount=16)
    a2 = [dict(step=0, register=0, value=1), dict(step=0, register=1, value=1)]
    _same_bytes(air2, opts2, a2, [], [1, 1])


def test_domain_below_128_fails_like_the_reference():
    """getComponentCount (LowDegreeProver.ts:287-291) makes N < 128 throw in the reference."""
    air = airs.fibonacci(16, modulus=P128)
    gpu = Stark(air, dict(extensionFactor=4))
    with pytest.raises(StarkError, match='Low degree proof failed'):
        gpu.prove_bytes([dict(step=0, register=0, value=1)], [], [1, 1])


def test_constraint_violation_is_reported():
    """a transition function that disagrees with the constraints: air-assembly refuses the trace
    ('Failed to evaluate transition constraints', CompositionPolynomial.ts:75-80)."""
    p = P128
    t = ProgramBuilder(p); t.out(0, t.cur(0) * t.cur(0) + 1)
    e = ProgramBuilder(p); e.out(0, e.nxt(0) - (e.cur(0) * e.cur(0) + 2))
    air = AirModule('broken', p, 1, 64, t.build(), e.build(), init=lambda i, s: [3])
    gpu = Stark(air, dict(extensionFactor=8))
    with pytest.raises(StarkError, match="Failed to evaluate transition constraints: Constraint 0 didn't evaluate to 0 at step 0"):
        gpu.prove_bytes([dict(step=0, register=0, value=3)], [], [])


def test_single_assertion_and_many_assertions_on_one_register():
    air, opts, a, inputs, seed = cases.mimc(256, 8, 'sha256')
    ctl = airs.run_mimc(256, airs.mimc_round_constants(), 3)
    _same_bytes(air, opts, [dict(step=17, register=0, value=ctl[17])], inputs, seed)
    many = [dict(step=s, register=0, value=ctl[s]) for s in (0, 1, 2, 100, 101, 200, 254, 255)]
    _same_bytes(air, opts, many, inputs, seed)


def _mixed_air(second_degree):
    p = P128
    t = ProgramBuilder(p)
    e = ProgramBuilder(p)
    t.out(0, t.exp(t.cur(0), 3) + t.static(0))
    e.out(0, e.nxt(0) - (e.exp(e.cur(0), 3) + e.static(0)))
    if second_degree == 2:
        t.out(1, t.cur(1) * t.cur(0) + 1); e.out(1, e.nxt(1) - (e.cur(1) * e.cur(0) + 1))
    else:
        t.out(1, t.cur(1) + t.cur(0)); e.out(1, e.nxt(1) - (e.cur(1) + e.cur(0)))
    air = AirModule('mixed', p, 2, 128, t.build(), e.build(), [StaticRegister('cycle', [5, 7, 11, 13])], init=lambda i, s: [3, 4])
    from oracle.air import ProvingContext
    tr = ProvingContext(air, [], []).generate_execution_trace()
    a = [dict(step=127, register=0, value=tr[0][127]), dict(step=127, register=1, value=tr[1][127]), dict(step=0, register=1, value=4)]
    return air, dict(hashAlgorithm='blake2s256', extensionFactor=8, exeQueryCount=40, friQueryCount=20), a


def test_mixed_degree_constraints_use_degree_adjustment():
    """two constraint groups (degrees 3 and 2): each lower-than-combination group gets its own x^incr copy and
    coefficient, in first-appearance order (CompositionPolynomial.ts:88-100,206-225)."""
    air, opts, a = _mixed_air(2)
    assert air.constraint_degrees == [3, 2]
    _same_bytes(air, opts, a, [], [])


def test_degree_one_constraint_next_to_cubic_overshoots_like_the_restated_protocol():
    """a degree-1 constraint raised by x^(3T) and divided by Z(x) has degree exactly 3T = compositionDegree, one more
    than FRI allows: the restated protocol rejects its own proof, and the GPU prover fails with the same text."""
    air, opts, a = _mixed_air(1)
    assert air.constraint_degrees == [3, 1]
    with pytest.raises(Exception, match='Remainder is not a valid degree 95 polynomial'):
        OracleStark(air, opts).prove(a, [], [])
    with pytest.raises(StarkError, match='Low degree proof failed: Remainder is not a valid degree 95 polynomial'):
        Stark(air, opts).prove_bytes(a, [], [])


def test_repeated_proves_with_different_assertions_reuse_the_instance():
    air, opts, a, inputs, seed = cases.mimc(1024, 8)
    gpu = Stark(air, opts)
    ora = OracleStark(air, opts)
    for sd in (3, 4, 5, 3):
        ctl = airs.run_mimc(1024, airs.mimc_round_constants(), sd)
        aa = [dict(step=0, register=0, value=ctl[0]), dict(step=1023, register=0, value=ctl[-1])]
        assert gpu.prove_bytes(aa, [], [sd]) == ora.serialize(ora.prove(aa, [], [sd]))
