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
