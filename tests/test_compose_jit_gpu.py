"""K2 (constraint evaluation + C(x) + L(x)) in both forms -- the kernel NVRTC compiles for an AIR's evaluation function
(devjit.cuh) and the interpreting kernel (compose.cuh) -- produces the oracle's C(x), L(x) and proof bytes, for every
AIR family of the BASELINE configs."""
import pytest

import cases
from genstark_b200.stark import Stark
from oracle import cport

pytestmark = pytest.mark.gpu

CASES = {
    'mimc': lambda: cases.mimc(1 << 10, 8),
    'rescue': lambda: cases.rescue(4),
    'poseidon': lambda: cases.poseidon(2, 1, 16),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_compiled_and_interpreted_constraint_kernels_agree_with_the_oracle(name, monkeypatch, tmp_path):
    air, opts, a, inputs, seed = CASES[name]()
    want = cport.prove(air, opts, a, inputs, seed)
    monkeypatch.setenv('GS_JIT_CACHE', str(tmp_path))
    out = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('GS_COMPOSE_JIT', mode)
        st = Stark(air, opts)
        st._set_debug(True)
        backend = st.compose_backend()
        assert backend.startswith('interpreter' if mode == '0' else 'nvrtc'), backend
        got = st.prove_bytes(a, inputs, seed)
        assert got == want, f'{name}: proof bytes differ from the oracle with {backend}'
        out[mode] = (st._read_intermediate(1), st._read_intermediate(2))
        st.close()
    assert out['0'][0] == out['1'][0], 'C(x) differs between the compiled and the interpreting kernel'
    assert out['0'][1] == out['1'][1], 'L(x) differs between the compiled and the interpreting kernel'
    assert any(f.name.startswith('compose_') and f.name.endswith('.cubin') for f in tmp_path.iterdir())


def test_constraint_violation_is_reported_by_the_compiled_kernel(monkeypatch):
    """air-assembly refuses a trace that violates a constraint (CompositionPolynomial.ts:75-80): the MiMC AIR
    with tampered round constants on the host side only is not expressible here, so use the quadratic AIR of
    tests/test_edge_cases_gpu.py style -- a transition program that disagrees with the evaluation program."""
    from genstark_b200.air import AirModule, ProgramBuilder, P128
    from genstark_b200.stark import StarkError
    t = ProgramBuilder(P128)
    t.out(0, t.cur(0) * t.cur(0) + 5)
    e = ProgramBuilder(P128)
    e.out(0, e.nxt(0) - (e.cur(0) * e.cur(0) + 6))          # off by one
    air = AirModule(name='bad', modulus=P128, trace_register_count=1, trace_length=64, transition=t.build(),
                    evaluation=e.build(), static_registers=[], extension_factor=8,
                    init=lambda inputs, seed: [int(seed[0]) % P128])
    monkeypatch.setenv('GS_COMPOSE_JIT', '1')
    st = Stark(air, dict(hashAlgorithm='blake2s256', extensionFactor=8, exeQueryCount=20, friQueryCount=10))
    assert st.compose_backend().startswith('nvrtc')
    with pytest.raises(StarkError, match='Failed to evaluate transition constraints'):
        st.prove_bytes([dict(step=0, register=0, value=3)], [], [3])
