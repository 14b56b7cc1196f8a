"""AirAssembly front-end (genstark_b200/assembly.py): source text -> AirModule.

CPU tests.  The MiMC and sponge modules below are written for these tests; the lib128 checks read the reference's
own library file, kept verbatim as a fixture (tests/golden/lib128.aa = /root/reference/assembly/lib128.aa, so the tests
also run on the GPU box where the reference checkout does not exist) -- they reproduce examples/assembly/lib128.ts:51-118:
the trace the component generates ends in the value an independent plain implementation computes."""
import os

import pytest

from genstark_b200 import airs, assembly
from genstark_b200.air import P128, prng_sha256
from genstark_b200.stark import generate_execution_trace

from asm_sources import MIMC_SOURCE, SPONGE_SOURCE, sponge_control, sponge_inputs

LIB128 = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'lib128.aa')
needs_lib128 = pytest.mark.skipif(not os.path.exists(LIB128), reason='tests/golden/lib128.aa missing')


def test_lib128_fixture_is_the_reference_file():
    ref = '/root/reference/assembly/lib128.aa'
    if not os.path.exists(ref):
        pytest.skip('reference checkout not present (GPU box)')
    assert open(ref, 'rb').read() == open(LIB128, 'rb').read()


def test_sexpr_parser_handles_comments_and_nesting():
    forms = assembly.parse_sexpr('(a (b 1 2) # comment (ignored\n (c))')
    assert forms == [['a', ['b', '1', '2'], ['c']]]
    with pytest.raises(assembly.AssemblyError):
        assembly.parse_sexpr('(a (b)')
    with pytest.raises(assembly.AssemblyError):
        assembly.parse_sexpr('(a))')


def test_mimc_module_equals_the_hand_built_air(monkeypatch):
    monkeypatch.setenv('GS_TRACE_JIT', '0')       # interpreter: tests/test_trace_jit.py owns the compiled MiMC program
    steps = 256
    m = assembly.compile(MIMC_SOURCE.replace('STEPS', str(steps))).component('mimc').module([])
    h = airs.mimc128(steps)
    assert (m.modulus, m.trace_register_count, m.trace_length) == (h.modulus, 1, steps)
    assert m.constraint_degrees == h.constraint_degrees == [3]
    assert m.static_registers[0].values == h.static_registers[0].values
    assert m.extension_factor == h.extension_factor == 8
    assert m.init([], [3]) == [3]
    control = airs.run_mimc(steps, airs.mimc_round_constants(), 3)
    assert generate_execution_trace(m, [], [3])[0] == control


def test_mimc_module_proof_equals_hand_built_proof_under_the_oracle():
    from oracle.stark import Stark as OracleStark
    steps = 64
    opts = dict(hashAlgorithm='blake2s256', extensionFactor=8, exeQueryCount=20, friQueryCount=10)
    m = assembly.compile(MIMC_SOURCE.replace('STEPS', str(steps)).encode()).component('mimc').module([])
    h = airs.mimc128(steps)
    ctl = airs.run_mimc(steps, airs.mimc_round_constants(), 3)
    a = [dict(step=0, register=0, value=3), dict(step=steps - 1, register=0, value=ctl[-1])]
    s1, s2 = OracleStark(m, opts), OracleStark(h, opts)
    p1 = s1.serialize(s1.prove(a, [], [3]))
    assert p1 == s2.serialize(s2.prove(a, [], [3]))
    assert s2.verify(a, s2.parse(p1))


def test_sponge_module_inputs_masks_and_vector_ops():
    comp = assembly.compile(SPONGE_SOURCE).component('sponge')
    blocks, words = 2, 4
    inputs = sponge_inputs(blocks, words)
    m = comp.module_for(inputs)
    assert m.trace_length == blocks * words * 16 and m.trace_register_count == 4
    assert comp.input_shapes(inputs) == [[blocks], [blocks], [blocks, words]]
    assert m.constraint_degrees == [4, 4, 4, 4]
    assert m.secret_input_count == 2
    kinds = [(s.kind, s.secret) if s.kind == 'input' else (s.kind, len(s.values)) for s in m.static_registers]
    assert kinds == [('input', True), ('input', True), ('input', False), ('cycle', 64), ('cycle', 16)] + [('cycle', 16)] * 4
    assert m.static_registers[3].values == [0] * 63 + [1] and m.static_registers[4].values == [0] * 15 + [1]
    trace = generate_execution_trace(m, inputs, [])
    want = sponge_control(inputs, blocks, words)
    for r in range(4):
        assert trace[r] == want[r]
    # the constraint evaluator vanishes on consecutive rows of that trace
    from oracle.air import run_program
    from oracle.field import PrimeField
    f = PrimeField(P128)
    regs = m.expand_inputs(inputs)
    for s in (0, 15, 16, 63, 64, 126):
        row = []
        it = iter(regs)
        for reg in m.static_registers:
            row.append(reg.values[s % len(reg.values)] if reg.kind == 'cycle' else next(it)[s])
        q = run_program(f, m.evaluation, [trace[r][s] for r in range(4)], [trace[r][s + 1] for r in range(4)], row)
        assert q == [0, 0, 0, 0]
    # public inputs alone rebuild the public register
    assert m.expand_public_inputs([inputs[2]]) == [regs[2]]


def test_sponge_proof_round_trip_under_the_oracle():
    from oracle.stark import Stark as OracleStark
    comp = assembly.compile(SPONGE_SOURCE).component('sponge')
    blocks, words = 1, 4
    inputs = sponge_inputs(blocks, words)
    m = comp.module_for(inputs)
    want = sponge_control(inputs, blocks, words)
    T = m.trace_length
    a = [dict(step=T - 1, register=r, value=want[r][T - 1]) for r in range(2)]
    st = OracleStark(m, dict(hashAlgorithm='sha256', extensionFactor=8, exeQueryCount=16, friQueryCount=8))
    proof = st.parse(st.serialize(st.prove(a, inputs, [])))
    assert proof['iShapes'] == [[1], [1], [1, 4]]
    assert st.verify(a, proof, [inputs[2]])
    bad = [dict(a[0], value=(a[0]['value'] + 1) % P128)]
    with pytest.raises(Exception):
        st.verify(bad, proof, [inputs[2]])


def test_front_end_errors():
    with pytest.raises(assembly.AssemblyError, match='not exported'):
        assembly.compile(MIMC_SOURCE.replace('STEPS', '64')).component('nope')
    comp = assembly.compile(SPONGE_SOURCE).component('sponge')
    with pytest.raises(assembly.AssemblyError, match='inputs expected'):
        comp.module_for([[1]])
    with pytest.raises(assembly.AssemblyError, match='ragged'):
        comp.module_for([[1, 2], [3, 4], [[1, 2, 3, 4], [1, 2]]])
    with pytest.raises(assembly.AssemblyError, match='peer'):
        comp.module_for([[1, 2], [3], [[1, 2, 3, 4], [1, 2, 3, 4]]])
    with pytest.raises(assembly.AssemblyError, match='power of 2'):
        comp.module_for([[1], [3], [[1, 2, 3]]])
    with pytest.raises(assembly.AssemblyError, match='yield a vector of 1'):
        assembly.compile(MIMC_SOURCE.replace('STEPS', '64').replace('(registers 1)', '(registers 1)').replace(
            '(sub\n                (load.trace 1)', '(vector (scalar 0)\n                (load.trace 1)')).component('mimc').module([])
    with pytest.raises(assembly.AssemblyError, match='binary'):
        src = SPONGE_SOURCE.replace('(input public (childof 0)', '(input public binary (childof 0)')
        c2 = assembly.compile(src).component('sponge')
        c2.module_for(sponge_inputs(1, 4)).expand_inputs(sponge_inputs(1, 4))


# ------------------------------------------------------------------------------------------------ lib128
def _poseidon_params():
    mds = airs.poseidon_mds(P128)
    ark = [prng_sha256(bytes.fromhex('48616465733%d' % (i + 1)), 64, P128) for i in range(6)]      # lib128.ts:24-31
    return mds, ark


def _poseidon(values, mds, ark, rf=8, rp=55):
    """plain Poseidon as in examples/poseidon/utils.ts:19-49 with the lib128 constants"""
    p, m = P128, 6
    st = [int(v) % p for v in values] + [0] * (m - len(values))
    for i in range(rf + rp):
        st = [(st[j] + ark[j][i]) % p for j in range(m)]
        if i < rf // 2 or i >= rf // 2 + rp:
            st = [pow(x, 5, p) for x in st]
        else:
            st[m - 1] = pow(st[m - 1], 5, p)
        st = [sum(mds[r][j] * st[j] for j in range(m)) % p for r in range(m)]
    return st[:2]


@needs_lib128
def test_lib128_mds_constant_equals_the_cauchy_matrix():
    lib = assembly.compile(LIB128)
    assert lib.modulus == P128 and set(lib.exports) == {'ComputePoseidonHash', 'ComputeMerkleRoot', 'ComputeMerkleUpdate'}
    assert lib.consts[lib.const_names['$mds']] == airs.poseidon_mds(P128)       # README.md KAT, SURVEY §4


@needs_lib128
def test_lib128_poseidon_hash_component_matches_plain_poseidon():
    comp = assembly.compile(LIB128).component('ComputePoseidonHash')
    inputs = [[42], [43], [44], [45]]                                           # lib128.ts:60
    m = comp.module_for(inputs)
    assert (m.trace_length, m.trace_register_count, m.constraint_count, m.secret_input_count) == (64, 6, 6, 4)
    assert max(m.constraint_degrees) == 7
    trace = generate_execution_trace(m, inputs, [])
    mds, ark = _poseidon_params()
    want = _poseidon([42, 43, 44, 45], mds, ark)
    assert [trace[0][63], trace[1][63]] == want
    # two hashes back to back: 128 steps
    m2 = comp.module_for([[1, 5], [2, 6], [3, 7], [4, 8]])
    t2 = generate_execution_trace(m2, [[1, 5], [2, 6], [3, 7], [4, 8]], [])
    assert [t2[0][63], t2[1][63]] == _poseidon([1, 2, 3, 4], mds, ark)
    assert [t2[0][127], t2[1][127]] == _poseidon([5, 6, 7, 8], mds, ark)


@needs_lib128
def test_lib128_merkle_root_component_matches_plain_merkle_tree():
    import random
    comp = assembly.compile(LIB128).component('ComputeMerkleRoot')
    mds, ark = _poseidon_params()
    depth, index = 4, 5                   # top index bit 0: the root ends in registers 0,1 (as index 42 of depth 8 in lib128.ts:89)
    r = random.Random(5)
    level = [[r.randrange(P128), r.randrange(P128)] for _ in range(2 ** depth)]
    tree = [level]
    while len(level) > 1:
        level = [_poseidon(level[2 * i] + level[2 * i + 1], mds, ark) for i in range(len(level) // 2)]
        tree.append(level)
    nodes, idx = [], index
    for d in range(depth):
        nodes.append(tree[d][idx ^ 1]); idx >>= 1
    bits = [(index >> d) & 1 for d in range(depth)]
    bits = [0] + bits[:-1]                                                      # lib128.ts:92-94
    leaf = tree[0][index]
    inputs = [[leaf[0]], [leaf[1]], [[n[0] for n in nodes]], [[n[1] for n in nodes]], [bits]]
    m = comp.module_for(inputs)
    assert (m.trace_length, m.trace_register_count, max(m.constraint_degrees)) == (64 * depth, 12, 8)   # README.md:211-218
    trace = generate_execution_trace(m, inputs, [])
    root = tree[-1][0]
    assert [trace[0][64 * depth - 1], trace[1][64 * depth - 1]] == root
    # and the oracle proves / verifies it with the index bits as the only public input (lib128.ts:104-111)
    from oracle.stark import Stark as OracleStark
    a = [dict(step=64 * depth - 1, register=0, value=root[0]), dict(step=64 * depth - 1, register=1, value=root[1])]
    st = OracleStark(m, dict(hashAlgorithm='blake2s256', extensionFactor=16, exeQueryCount=12, friQueryCount=6))
    proof = st.parse(st.serialize(st.prove(a, inputs, [])))
    assert st.verify(a, proof, [[bits]])


def test_proof_parser_needs_only_sizes_not_an_instance():
    """ScriptStark reads iShapes out of serialized proofs before it knows the trace length (stark.py: parse_proof)"""
    from genstark_b200.stark import parse_proof
    from oracle import cport
    comp = assembly.compile(SPONGE_SOURCE).component('sponge')
    inputs = sponge_inputs(2, 4)
    m = comp.module_for(inputs)
    want = sponge_control(inputs, 2, 4)
    T = m.trace_length
    a = [dict(step=T - 1, register=0, value=want[0][T - 1])]
    opts = dict(hashAlgorithm='sha256', extensionFactor=8, exeQueryCount=12, friQueryCount=6)
    buf = cport.prove(m, opts, a, inputs, [])
    proof = parse_proof(buf, (4 + 2) * 16, 64, 16, 32)
    assert proof['iShapes'] == [[2], [2], [2, 4]] and proof['evRoot'] == buf[:32]
    assert len(proof['ldProof']['remainder']) in (64, 128, 256)


def test_script_stark_answers_shape_independent_questions_without_a_device():
    from genstark_b200.stark import ScriptStark
    comp = assembly.compile(SPONGE_SOURCE).component('sponge')
    st = ScriptStark(comp, dict(hashAlgorithm='blake2s256', extensionFactor=16, exeQueryCount=48, friQueryCount=24))
    assert st.air.trace_register_count == 4 and st.air.max_constraint_degree == 4 and st.air.extension_factor == 16
    # Stark.ts:62-77: min(log2((E / d)^exe), log2(E) * fri, 128) = min(96, 96, 128)
    assert st.securityLevel == 96
