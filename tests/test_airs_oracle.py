"""Multi-register AIRs (BASELINE configs 3 and 5) on the oracle: in-tree KATs, independent control values,
prove/verify round trips, and the C port against the Python restatement."""
import pytest

import cases
from genstark_b200 import airs
from oracle import cport
from oracle.air import ProvingContext
from oracle.stark import Stark as OracleStark


def test_rescue_kat_from_reference_example():
    # examples/rescue/hash4x128.ts:114-118: input (42, 43) -> registers 0,1 at step 31
    air = airs.rescue4x128(1)
    row0 = airs.rescue_build_inputs([42, 43])
    tr = ProvingContext(air, [[row0[i]] for i in range(4)], []).generate_execution_trace()
    assert tr[0][31] == 302524937772545017647250309501879538110
    assert tr[1][31] == 205025454306577433144586673939030012640
    assert [tr[r][31] for r in range(4)] == cases.rescue_hash_control([42, 43])


def test_poseidon_branch_reaches_the_control_root():
    air, opts, a, inputs, seed = cases.poseidon(4, 2)
    tr = ProvingContext(air, inputs, seed).generate_execution_trace()
    for x in a:
        assert tr[x['register']][x['step']] == x['value']
    assert set(air.constraint_degrees) == {7} and air.trace_register_count == 12 and air.secret_input_count == 4


@pytest.mark.parametrize('case', [lambda: cases.rescue(4), lambda: cases.poseidon(2, 1, e=16)])
def test_oracle_roundtrip_and_c_port(case):
    air, opts, a, inputs, seed = case()
    ora = OracleStark(air, opts)
    proof = ora.prove(a, inputs, seed)
    buf = ora.serialize(proof)
    assert ora.verify(a, ora.parse(buf), inputs[4:] if air.name == 'poseidon_mp' else None)
    assert cport.prove(air, opts, a, inputs, seed) == buf
    bad = [dict(a[0], value=a[0]['value'] + 1)] + a[1:]
    with pytest.raises(Exception):
        ora.verify(bad, ora.parse(buf), inputs[4:] if air.name == 'poseidon_mp' else None)


def test_numpy_input_expansion_equals_the_list_based_one():
    """AirModule.expand_inputs_blob (prove path) produces byte for byte the columns of expand_inputs"""
    from genstark_b200 import assembly
    from genstark_b200.air import input_blob
    from asm_sources import SPONGE_SOURCE, sponge_inputs
    todo = [cases.poseidon(2, 1, e=16), cases.poseidon(4, 8), cases.rescue(4)]
    inputs = sponge_inputs(4, 8)
    todo.append((assembly.compile(SPONGE_SOURCE).component('sponge').module_for(inputs), None, None, inputs, []))
    for air, _, _, inputs, _ in todo:
        p = air.modulus
        slow = b''.join((int(v) % p).to_bytes(16, 'little') for t in air.expand_inputs(inputs) for v in t)
        assert input_blob(air, inputs) == slow
        assert len(slow) == 16 * air.trace_length * sum(1 for s in air.static_registers if s.kind == 'input')
    # verify path: the public-input columns, numpy against the list-based expansion
    from genstark_b200.air import public_blob
    for air, _, _, inputs, _ in (cases.poseidon(2, 1, e=16), cases.poseidon(4, 8)):
        p, pub = air.modulus, inputs[4:]
        assert air.expand_public_inputs_blob is not None
        assert public_blob(air, pub) == b''.join((int(v) % p).to_bytes(16, 'little') for t in air.expand_public_inputs(pub) for v in t)
