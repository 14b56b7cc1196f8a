"""AirAssembly source -> instantiate() -> GPU proof: the bytes equal the oracle's for the same module, and
`ScriptStark` (trace length follows from the inputs) keeps the reference's prove / verify / parse surface."""
import os

import pytest

from genstark_b200 import airs, instantiate
from genstark_b200 import assembly
from genstark_b200.stark import ScriptStark, Stark, StarkError
from oracle import cport
from oracle.stark import Stark as OracleStark

from asm_sources import MIMC_SOURCE, SPONGE_SOURCE, sponge_control, sponge_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('steps,e,alg', [(64, 8, 'blake2s256'), (2**12, 16, 'sha256')])
def test_mimc_source_is_a_drop_in_for_the_hand_built_air(steps, e, alg):
    opts = dict(hashAlgorithm=alg, extensionFactor=e, exeQueryCount=48, friQueryCount=24)
    st = instantiate(MIMC_SOURCE.replace('STEPS', str(steps)).encode(), 'mimc', opts)
    assert isinstance(st, Stark)
    ctl = airs.run_mimc(steps, airs.mimc_round_constants(), 3)
    a = [dict(step=0, register=0, value=3), dict(step=steps - 1, register=0, value=ctl[-1])]
    got = st.prove_bytes(a, [], [3])
    ref = Stark(airs.mimc128(steps), opts)
    assert got == ref.prove_bytes(a, [], [3])
    assert got == cport.prove(airs.mimc128(steps), opts, a, [], [3])
    assert st.verify(a, st.parse(got))
    assert st.compose_backend().startswith('nvrtc')


@pytest.mark.parametrize('blocks,words,e', [(1, 4, 8), (4, 8, 16)])
def test_sponge_component_with_input_registers(blocks, words, e):
    opts = dict(hashAlgorithm='blake2s256', extensionFactor=e, exeQueryCount=40, friQueryCount=20)
    st = instantiate(SPONGE_SOURCE, 'sponge', opts)
    assert isinstance(st, ScriptStark)
    inputs = sponge_inputs(blocks, words)
    want = sponge_control(inputs, blocks, words)
    T = blocks * words * 16
    a = [dict(step=T - 1, register=r, value=want[r][T - 1]) for r in range(4)] + [dict(step=16, register=1, value=want[1][16])]
    got = st.prove_bytes(a, inputs)
    module = assembly.compile(SPONGE_SOURCE).component('sponge').module_for(inputs)
    assert got == cport.prove(module, opts, a, inputs, [])
    if T <= 64:
        ora = OracleStark(module, opts)
        assert got == ora.serialize(ora.prove(a, inputs, []))
        assert ora.verify(a, ora.parse(got), [inputs[2]])
    proof = st.parse(got)
    assert proof['iShapes'] == [[blocks], [blocks], [blocks, words]]
    assert st.serialize(proof) == got and st.sizeOf(proof) == len(got)
    assert st.verify(a, proof, [inputs[2]])
    assert st.verify(a, got, [inputs[2]])
    assert st.generateExecutionTrace(inputs)[3] == want[3]
    # a different message does not verify against the same proof
    other = [[(w + 1) for w in row] for row in inputs[2]]
    with pytest.raises(StarkError):
        st.verify(a, proof, [other])
    # a second shape builds a second device instance behind the same object
    inputs2 = sponge_inputs(blocks, 2 * words)
    want2 = sponge_control(inputs2, blocks, 2 * words)
    a2 = [dict(step=2 * T - 1, register=0, value=want2[0][2 * T - 1])]
    got2 = st.prove_bytes(a2, inputs2)
    assert st.verify(a2, got2, [inputs2[2]])
    assert len(st._by_shape) == 2


def test_unsupported_field_is_refused_loudly():
    # a 224-bit field (assembly/lib224.aa's modulus): no device path and no host path -- the reference runs it on JS bigints
    src = MIMC_SOURCE.replace('STEPS', '64').replace('340282366920938463463374607393113505793', str(2**224 - 2**96 + 1))
    with pytest.raises(StarkError, match='not supported'):
        instantiate(src, 'mimc', dict(extensionFactor=8))


LIB128 = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'lib128.aa')     # the reference's assembly/lib128.aa, verbatim


def test_lib128_merkle_root_from_the_library_file():
    import random
    from test_assembly import _poseidon, _poseidon_params
    from genstark_b200.air import P128
    mds, ark = _poseidon_params()
    depth, index = 2, 1
    r = random.Random(9)
    level = [[r.randrange(P128), r.randrange(P128)] for _ in range(2 ** depth)]
    tree = [level]
    while len(level) > 1:
        level = [_poseidon(level[2 * i] + level[2 * i + 1], mds, ark) for i in range(len(level) // 2)]
        tree.append(level)
    nodes, idx = [], index
    for d in range(depth):
        nodes.append(tree[d][idx ^ 1]); idx >>= 1
    bits = [0] + [(index >> d) & 1 for d in range(depth)][:-1]
    leaf = tree[0][index]
    inputs = [[leaf[0]], [leaf[1]], [[n[0] for n in nodes]], [[n[1] for n in nodes]], [bits]]
    opts = dict(hashAlgorithm='blake2s256', extensionFactor=32, exeQueryCount=44, friQueryCount=20)
    st = instantiate(LIB128, 'ComputeMerkleRoot', opts)
    root = tree[-1][0]
    a = [dict(step=64 * depth - 1, register=0, value=root[0]), dict(step=64 * depth - 1, register=1, value=root[1])]
    got = st.prove_bytes(a, inputs)
    assert st.verify(a, got, [[bits]])
    module = assembly.compile(LIB128).component('ComputeMerkleRoot').module_for(inputs)
    assert got == cport.prove(module, opts, a, inputs, [])
