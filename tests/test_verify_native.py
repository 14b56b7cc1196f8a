"""The library's native Stark.verify (host C++, no device) against proofs produced by the oracle prover:
accepts what the reference protocol accepts, rejects tampering with the reference's error texts."""
import pytest

import cases
from genstark_b200.stark import StarkError, verify_proof
from oracle.stark import Stark as OracleStark


def _proof(case):
    air, opts, a, inputs, seed = case
    ora = OracleStark(air, opts)
    return air, opts, a, inputs, ora.serialize(ora.prove(a, inputs, seed))


@pytest.mark.parametrize('case', [lambda: cases.mimc(64, 8), lambda: cases.mimc(256, 16, 'sha256'), lambda: cases.mimc(1024, 8),
                                  lambda: cases.rescue(4), lambda: cases.poseidon(2, 1, e=16)])
def test_native_verifier_accepts_oracle_proofs(case):
    air, opts, a, inputs, buf = _proof(case())
    pub = inputs[4:] if air.name == 'poseidon_mp' else None
    assert verify_proof(air, opts, a, buf, pub)


def test_native_verifier_rejects_tampering():
    air, opts, a, inputs, buf = _proof(cases.mimc(256, 8))
    assert verify_proof(air, opts, a, buf)
    bad_a = [dict(a[0]), dict(a[1], value=a[1]['value'] + 1)]
    with pytest.raises(StarkError, match='linear combination correctness'):
        verify_proof(air, opts, bad_a, buf)
    # flip a byte in the evaluation root / in a leaf / in the remainder / in a FRI row
    for pos, pattern in ((5, 'Merkle proof failed'), (40, 'evaluation Merkle proof failed'), (len(buf) - 30, 'low degree')):
        t = bytearray(buf); t[pos] ^= 1
        with pytest.raises(StarkError, match=pattern):
            verify_proof(air, opts, a, bytes(t))
    with pytest.raises(StarkError):
        verify_proof(air, opts, a, buf[:len(buf) // 2])
    with pytest.raises(TypeError):
        verify_proof(air, opts, [], buf)


def test_native_verifier_needs_the_right_public_inputs():
    air, opts, a, inputs, buf = _proof(cases.poseidon(2, 1, e=16))
    assert verify_proof(air, opts, a, buf, inputs[4:])
    wrong = [[[1 - b for b in row] for row in inputs[4]]]
    with pytest.raises(StarkError):
        verify_proof(air, opts, a, buf, wrong)


def test_native_verifier_validates_assertions():
    """what the prover checks on its side (register bank, trace length, repeated pairs) is checked here too"""
    air, opts, a, inputs, buf = _proof(cases.mimc(256, 8))
    for bad, pattern in (([dict(a[0], register=1)], 'outside of register bank'),
                         ([dict(a[0], register=2**31)], 'outside of register bank'),
                         ([dict(a[0], step=256)], 'outside of execution trace'),
                         ([a[0], a[1], dict(a[0])], 'repeated assertion')):
        with pytest.raises(StarkError, match=pattern):
            verify_proof(air, opts, bad, buf)


def test_native_verifier_bounds_what_the_proof_claims():
    """component count and batch depths are values read from the proof: both have to match the instance"""
    from genstark_b200.stark import Stark as _S                      # serializer / parser only (no device needed)
    air, opts, a, inputs, buf = _proof(cases.mimc(256, 8))
    es = 16
    from genstark_b200.stark import parse_proof
    st = object.__new__(_S)
    st.air, st.elementSize = air.with_options(8), es
    proof = parse_proof(buf, (air.trace_register_count + air.secret_input_count) * es, 4 * es, es, 32)
    extra = dict(proof, ldProof=dict(proof['ldProof'], components=proof['ldProof']['components'] + proof['ldProof']['components'][-1:]))
    with pytest.raises(StarkError, match='components'):
        verify_proof(air, opts, a, _S.serialize(st, extra))
    fewer = dict(proof, ldProof=dict(proof['ldProof'], components=proof['ldProof']['components'][:-1]))
    with pytest.raises(StarkError):
        verify_proof(air, opts, a, _S.serialize(st, fewer))
    # every depth byte of the proof set to 200 in turn: rejected, no undefined shift
    import copy
    for pick in (lambda p: p['evProof'], lambda p: p['ldProof']['lcProof'],
                 lambda p: p['ldProof']['components'][0]['columnProof'], lambda p: p['ldProof']['components'][0]['polyProof']):
        for depth in (200, 64, 33, 0):
            p2 = copy.deepcopy(proof)
            pick(p2).depth = depth
            with pytest.raises(StarkError):
                verify_proof(air, opts, a, _S.serialize(st, p2))
