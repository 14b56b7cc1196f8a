"""Execution-trace generation on the host (lib/Stark.ts:97,252-257): the transition function compiled to native code
at first use (hostjit.h) and the interpreter (hostair.h) both reproduce the oracle's trace, for every AIR family of
the BASELINE configs.  Host only: no device needed."""
import os

import pytest

import cases
from genstark_b200 import stark as gstark
from oracle.air import ProvingContext

CASES = {
    'mimc': lambda: cases.mimc(1 << 10, 8),
    'rescue': lambda: cases.rescue(4),
    'poseidon': lambda: cases.poseidon(4, 2),
}


def _oracle_trace(air, inputs, seed):
    tr = ProvingContext(air, inputs, seed).generate_execution_trace()
    return [[int(tr[r][s]) for s in range(air.trace_length)] for r in range(air.trace_register_count)]


@pytest.mark.parametrize('name', sorted(CASES))
def test_compiled_and_interpreted_traces_equal_the_oracle(name, tmp_path, monkeypatch):
    air, opts, a, inputs, seed = CASES[name]()
    want = _oracle_trace(air, inputs, seed)
    monkeypatch.setenv('GS_JIT_CACHE', str(tmp_path))
    monkeypatch.setenv('GS_TRACE_JIT', '0')
    got_i = gstark.generate_execution_trace(air, inputs, seed)
    assert gstark.trace_backend().startswith('interpreter')
    assert got_i == want
    monkeypatch.setenv('GS_TRACE_JIT', '1')
    got_j = gstark.generate_execution_trace(air, inputs, seed)
    backend = gstark.trace_backend()
    if not backend.startswith('jit'):
        pytest.skip(f'no host compiler for the trace JIT: {backend}')
    assert got_j == want
    # the compiled object is cached on disk and in the process
    assert any(f.name.endswith('.so') for f in tmp_path.iterdir())
    assert gstark.generate_execution_trace(air, inputs, seed) == want
    for x in a:
        assert got_j[x['register']][x['step']] == x['value']


def _quadratic_air(steps, c):
    from genstark_b200.air import AirModule, ProgramBuilder, P128
    t = ProgramBuilder(P128)
    t.out(0, t.cur(0) * t.cur(0) + c)
    e = ProgramBuilder(P128)
    e.out(0, e.nxt(0) - (e.cur(0) * e.cur(0) + c))
    return AirModule(name='quad', modulus=P128, trace_register_count=1, trace_length=steps, transition=t.build(),
                     evaluation=e.build(), static_registers=[], extension_factor=4,
                     init=lambda inputs, seed: [int(seed[0]) % P128])


def test_jit_falls_back_to_the_interpreter_without_a_compiler(tmp_path, monkeypatch):
    air = _quadratic_air(64, 0xB200_0001)            # a program no other test compiles (the in-process cache is per program)
    monkeypatch.setenv('GS_JIT_CACHE', str(tmp_path))
    monkeypatch.setenv('GS_JIT_CXX', '/nonexistent/compiler')
    monkeypatch.setenv('GS_TRACE_JIT', '1')
    got = gstark.generate_execution_trace(air, [], [9])
    assert gstark.trace_backend().startswith('interpreter (host compiler failed')
    assert got == _oracle_trace(air, [], [9])
    x, want = 9, []
    for _ in range(64):
        want.append(x)
        x = (x * x + 0xB200_0001) % air.modulus
    assert got[0] == want


@pytest.mark.parametrize('name', ['rescue', 'poseidon'])
def test_segment_parallel_generation_equals_the_sequential_one(name, monkeypatch):
    """AIRs whose segments restart from the inputs (`for each` loops) are generated chunk-parallel (hostjit.h); one chain
    (MiMC) is not.  Same trace either way."""
    import cases
    big = {'rescue': lambda: cases.rescue(256), 'poseidon': lambda: cases.poseidon(4, 32)}[name]
    air, opts, a, inputs, seed = big()
    assert air.trace_length >= 4096
    monkeypatch.setenv('GS_TRACE_THREADS', '1')
    seq = gstark.generate_execution_trace(air, inputs, seed)
    backend = gstark.trace_backend()
    if not backend.startswith('jit'):
        pytest.skip(f'no host compiler for the trace JIT: {backend}')
    assert 'threads' not in backend
    for x in a:
        assert seq[x['register']][x['step']] == x['value']
    for n in ('2', '4', '16'):
        monkeypatch.setenv('GS_TRACE_THREADS', n)
        par = gstark.generate_execution_trace(air, inputs, seed)
        assert 'threads' in gstark.trace_backend(), gstark.trace_backend()
        assert par == seq
    # one dependent chain: no state-free step, so no chunks
    monkeypatch.setenv('GS_TRACE_THREADS', '8')
    m_air, _, _, m_in, m_seed = cases.mimc(8192, 8)
    tr = gstark.generate_execution_trace(m_air, m_in, m_seed)
    assert 'threads' not in gstark.trace_backend()
    from genstark_b200 import airs
    assert tr[0] == airs.run_mimc(8192, airs.mimc_round_constants(), 3)


def test_jit_cache_refuses_a_directory_others_can_write_to(tmp_path, monkeypatch):
    """what is loaded from the cache is executed: a world-writable directory is not used (hostjit.h / jitcache.h); the
    object is compiled into a private mkdtemp directory instead and nothing in the shared one is touched or loaded"""
    air = _quadratic_air(64, 0xB200_0002)
    shared = tmp_path / 'shared'
    shared.mkdir()
    os.chmod(shared, 0o777)
    monkeypatch.setenv('GS_JIT_CACHE', str(shared))
    monkeypatch.setenv('GS_TRACE_JIT', '1')
    got = gstark.generate_execution_trace(air, [], [5])
    backend = gstark.trace_backend()
    if not backend.startswith('jit'):
        pytest.skip(f'no host compiler for the trace JIT: {backend}')
    assert got == _oracle_trace(air, [], [5])
    assert list(shared.iterdir()) == []


def _masked_air(steps, seed):
    """a synthetic multi-mask AIR that exercises every fold of the mask specialisation (hostjit.h: jit_specialise): three 0/1 cyclic
    registers of different periods (one of them constant), selects nested two deep, a masked operand under neg / inv / exp, a product
    of two masks, an output that is a constant under some masks and one that ignores the masks"""
    import random
    from genstark_b200.air import AirModule, ProgramBuilder, StaticRegister, P128
    r = random.Random(seed)
    p = P128
    m_a = [r.randrange(2) for _ in range(8)]
    m_b = [0, 0, 0, 1]
    m_c = [1] * 2
    keys = [r.randrange(p) for _ in range(16)]

    def build(b, out):
        x, y, z = b.cur(0), b.cur(1), b.cur(2)
        a, bb, c, k = b.static(0), b.static(1), b.static(2), b.static(3)
        heavy = b.exp(x + k, 5) * y + z * z * z                    # only needed when a = 1
        light = x * 3 + y                                            # only needed when a = 0
        inner = bb * (z + 7) + (1 - bb) * b.exp(z, 3)
        out(0, a * heavy + (1 - a) * light)
        out(1, a * bb * (x - y) + (1 - a * bb) * inner + c * 0)     # a product of two masks; c * 0 folds to nothing
        out(2, -(a * z) + b.inv(bb * x + 1) + b.exp(c * y + 2, 3))  # masked operands under neg / inv / exp
        out(3, (1 - c) * x + c * 12345)                              # constant under c = 1 (always here: the register is all ones)
    t = ProgramBuilder(p)
    build(t, lambda i, v: t.out(i, v))
    e = ProgramBuilder(p)
    build(e, lambda i, v: e.out(i, e.nxt(i) - v))
    statics = [StaticRegister('cycle', m_a), StaticRegister('cycle', m_b), StaticRegister('cycle', m_c), StaticRegister('cycle', keys)]
    return AirModule(name='masked', modulus=p, trace_register_count=4, trace_length=steps, transition=t.build(), evaluation=e.build(),
                     static_registers=statics, extension_factor=16, init=lambda inputs, sd: [int(v) % p for v in sd])


@pytest.mark.parametrize('seed', [1, 2, 3])
def test_mask_specialised_transition_equals_the_generic_one_and_the_oracle(seed, tmp_path, monkeypatch):
    air = _masked_air(256, seed)
    start = [3 + seed, 5, 7, 11]
    want = _oracle_trace(air, [], start)
    monkeypatch.setenv('GS_JIT_CACHE', str(tmp_path))
    monkeypatch.setenv('GS_TRACE_JIT', '1')
    got, backend = {}, {}
    for spec in ('2', '0'):                                   # 2: specialise even where the cost estimate would not bother
        monkeypatch.setenv('GS_TRACE_SPECIALISE', spec)
        got[spec] = gstark.generate_execution_trace(air, [], start)
        backend[spec] = gstark.trace_backend()
        if not backend[spec].startswith('jit'):
            pytest.skip(f'no host compiler for the trace JIT: {backend[spec]}')
    assert got['2'] == want and got['0'] == want
    # two different compiled objects: the specialised source has one step per combination of mask values
    assert backend['2'] != backend['0']


def test_mask_specialisation_on_the_baseline_airs(tmp_path, monkeypatch):
    """Poseidon Merkle proof (full-round / level / proof masks: specialised), the Rescue chain (its select guards little next to two
    128-bit exponentiations: the cost gate leaves it generic), lib128's ComputeMerkleRoot from AirAssembly text: same traces as the
    oracle with the specialisation on and off"""
    import cases
    from genstark_b200 import assembly
    monkeypatch.setenv('GS_JIT_CACHE', str(tmp_path))
    monkeypatch.setenv('GS_TRACE_JIT', '1')
    todo = [('poseidon', cases.poseidon(4, 2)), ('rescue', cases.rescue(8))]
    lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'lib128.aa')
    if os.path.exists(lib):
        inputs = [[42], [43], [44], [45]]                                           # examples/assembly/lib128.ts:60
        todo.append(('lib128', (assembly.compile(lib).component('ComputePoseidonHash').module_for(inputs), None, None, inputs, [])))
    from asm_sources import SPONGE_SOURCE, sponge_inputs
    sp_in = sponge_inputs(4, 8)
    todo.append(('sponge', (assembly.compile(SPONGE_SOURCE).component('sponge').module_for(sp_in), None, None, sp_in, [])))
    objects = {}
    for name, (air, _, _, inputs, seed) in todo:
        want = _oracle_trace(air, inputs, seed)
        for spec in ('1', '0', '2'):
            monkeypatch.setenv('GS_TRACE_SPECIALISE', spec)
            assert gstark.generate_execution_trace(air, inputs, seed) == want, (name, spec)
            if not gstark.trace_backend().startswith('jit'):
                pytest.skip(f'no host compiler for the trace JIT: {gstark.trace_backend()}')
            objects[(name, spec)] = gstark.trace_backend()
    assert objects[('poseidon', '1')] != objects[('poseidon', '0')]          # specialised
    assert objects[('rescue', '1')] == objects[('rescue', '0')]              # left generic by the cost gate
