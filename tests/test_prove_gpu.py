"""End-to-end parity of the fused GPU prover against the oracle's restatement of Stark.prove:
serialized proof bytes identical, stage outputs identical, and the GPU proof accepted by the oracle's
restatement of Stark.verify."""
import pytest

from genstark_b200 import airs
from genstark_b200.stark import Stark, StarkError
from oracle.stark import Stark as OracleStark

pytestmark = pytest.mark.gpu

OPTS = dict(hashAlgorithm='blake2s256', exeQueryCount=48, friQueryCount=24)


def _mimc_case(steps, e, hash_alg='blake2s256', seed=3):
    opts = dict(OPTS, extensionFactor=e, hashAlgorithm=hash_alg)
    air = airs.mimc128(steps)
    ctl = airs.run_mimc(steps, airs.mimc_round_constants(), seed)
    a = [dict(step=0, register=0, value=ctl[0]), dict(step=steps - 1, register=0, value=ctl[-1])]
    return air, opts, a, [seed]


@pytest.mark.parametrize('steps,e,alg', [(64, 8, 'blake2s256'), (64, 16, 'sha256'), (256, 16, 'blake2s256'),
                                         (1024, 8, 'sha256'), (2**13, 8, 'blake2s256')])
def test_mimc_proof_bytes_match_oracle(steps, e, alg):
    air, opts, a, seed = _mimc_case(steps, e, alg)
    gpu = Stark(air, opts)
    gpu._set_debug(True)
    got = gpu.prove_bytes(a, [], seed)
    ora = OracleStark(air, opts)
    tr = {}
    want_proof = ora.prove(a, [], seed, trace_out=tr)
    n = steps * e
    # stage-level parity first (sharper failure messages)
    assert gpu._read_intermediate(3) == [v for row in tr['p_polys'] for v in row], 'P(x) coefficients'
    assert gpu._read_intermediate(0) == [v for row in tr['p_evaluations'] for v in row], 'P(x) evaluations'
    assert gpu._read_intermediate(1) == tr['c_evaluations'], 'C(x)'
    assert gpu._read_intermediate(2) == tr['l_evaluations'], 'L(x)'
    want = ora.serialize(want_proof)
    assert got == want
    # the reference verifier logic accepts the GPU proof; round trip through parse/serialize
    proof = gpu.parse(got)
    assert gpu.serialize(proof) == got
    assert gpu.sizeOf(proof) == len(got)
    assert ora.verify(a, ora.parse(got))
    # and so does the library's own verifier (Stark.verify)
    assert gpu.verify(a, proof)
    with pytest.raises(StarkError):
        gpu.verify([dict(a[0]), dict(a[1], value=a[1]['value'] + 1)], proof)


def test_wrong_assertion_is_rejected_like_the_reference():
    air, opts, a, seed = _mimc_case(64, 8)
    gpu = Stark(air, opts)
    bad = [dict(a[0]), dict(a[1], value=(a[1]['value'] + 1))]
    with pytest.raises(StarkError, match='conflicts with execution trace'):
        gpu.prove_bytes(bad, [], seed)
    with pytest.raises(TypeError):
        gpu.prove_bytes([], [], seed)


def test_north_star_shape_proof_verifies_under_oracle_verifier():
    """2^16 steps: too slow for the Python oracle prover, but its verifier is O(queries log N)."""
    air, opts, a, seed = _mimc_case(2**16, 8)
    gpu = Stark(air, opts)
    got = gpu.prove_bytes(a, [], seed)
    ora = OracleStark(air, opts)
    assert ora.verify(a, ora.parse(got))
    assert gpu.securityLevel == ora.security_level


@pytest.mark.parametrize('log_steps,e', [(16, 8), (18, 16), (20, 8), (20, 16)])      # (20, 8) = north star, (20, 16) = config 4 on one GPU
def test_large_proofs_match_c_oracle_port(log_steps, e):
    """BASELINE sizes: the GPU proof is byte-identical to the plain-C oracle port (itself pinned to the Python
    restatement by tests/test_cport.py), which runs the reference's unfused data flow on the host cores."""
    from oracle import cport
    from genstark_b200 import workloads
    air, opts, a, _, _ = workloads.mimc(1 << log_steps, e)
    gpu = Stark(air, opts)
    got = gpu.prove_bytes(a, [], [3])
    want = cport.prove(air, opts, a, [], [3])
    assert len(got) == len(want)
    assert got == want
    # prove again: buffers are reused across calls and the result must not depend on it
    assert gpu.prove_bytes(a, [], [3]) == want


def _check_against(gpu_case, oracle_kind):
    import cases  # noqa: F401
    air, opts, a, inputs, seed = gpu_case
    gpu = Stark(air, opts)
    got = gpu.prove_bytes(a, inputs, seed)
    if oracle_kind == 'python':
        ora = OracleStark(air, opts)
        want = ora.serialize(ora.prove(a, inputs, seed))
        pub = inputs[4:] if air.name == 'poseidon_mp' else None
        assert ora.verify(a, ora.parse(got), pub)
    else:
        from oracle import cport
        want = cport.prove(air, opts, a, inputs, seed)
    assert len(got) == len(want)
    assert got == want


def test_rescue_chain_small_matches_python_oracle():
    import cases
    _check_against(cases.rescue(4), 'python')


def test_rescue_config3_2e12_steps_matches_c_oracle():
    """BASELINE config 3: Rescue 4x128, 2^12 steps (128 chained instances), 4 trace + 4 secret input registers."""
    import cases
    _check_against(cases.rescue(128), 'c')


def test_poseidon_small_matches_python_oracle():
    import cases
    _check_against(cases.poseidon(2, 1, e=16), 'python')


def test_poseidon_config5_2e16_steps_matches_c_oracle():
    """BASELINE config 5: Poseidon Merkle proof, 12 registers, 2^16 steps (128 branches of depth 8), E=32."""
    import cases
    _check_against(cases.poseidon(8, 128, e=32), 'c')
