"""Prime fields of at most 64 bits -- BASELINE config 1 (the Foo demo of /root/reference/README.md:22-50 over
p = 2^32 - 3*2^25 + 1): prove() and verify() run on the host inside libgenstark_b200.so (csrc/hoststark64.h), reached through the
same instantiate() surface.  CPU tests: proof bytes equal the Python oracle's, both verifiers accept, tampering is rejected."""
import random

import pytest

from genstark_b200 import airs, instantiate
from genstark_b200.air import AirModule, ProgramBuilder, StaticRegister, P32
from genstark_b200.stark import HostStark, Stark, StarkError
from oracle.stark import Stark as OracleStark

def _prime_with_two_adicity(bits, k):
    """smallest-step search for a prime c * 2^k + 1 just under 2^bits (a 64-bit-class NTT field for the tests)"""
    def is_prime(n):
        if n < 2:
            return False
        for q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if n % q == 0:
                return n == q
        d, s = n - 1, 0
        while d % 2 == 0:
            d //= 2; s += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, n)
            if x in (1, n - 1):
                continue
            for _ in range(s - 1):
                x = x * x % n
                if x == n - 1:
                    break
            else:
                return False
        return True
    c = (1 << (bits - k)) - 1
    while not is_prime(c * (1 << k) + 1):
        c -= 2
    return c * (1 << k) + 1


P63 = _prime_with_two_adicity(63, 24)


def _check(air, opts, assertions, inputs=None, seed=None, public=None):
    st = instantiate(air, opts)
    assert isinstance(st, HostStark)
    got = st.prove_bytes(assertions, inputs, seed)
    ora = OracleStark(air, opts)
    want = ora.serialize(ora.prove(assertions, inputs, seed))
    assert got == want
    assert st.verify(assertions, got, public)
    assert ora.verify(assertions, ora.parse(got), public)
    proof = st.parse(got)
    assert st.serialize(proof) == got and st.sizeOf(proof) == len(got)
    return st, got


@pytest.mark.parametrize('alg', ['sha256', 'blake2s256'])
@pytest.mark.parametrize('e', [None, 4, 16])
def test_foo_config1_bytes_equal_the_oracle(alg, e):
    opts = dict(hashAlgorithm=alg, exeQueryCount=20, friQueryCount=10)
    if e:
        opts['extensionFactor'] = e
    _check(airs.foo(64), opts, [dict(step=0, register=0, value=1), dict(step=63, register=0, value=127)], [[1]], [])


def _mimc(p, steps, consts):
    t = ProgramBuilder(p); t.out(0, t.exp(t.cur(0), 3) + t.static(0))
    e = ProgramBuilder(p); e.out(0, e.nxt(0) - (e.exp(e.cur(0), 3) + e.static(0)))
    return AirModule(name='mimc_small', modulus=p, trace_register_count=1, trace_length=steps, transition=t.build(), evaluation=e.build(),
                     static_registers=[StaticRegister('cycle', consts)], extension_factor=8, init=lambda inputs, seed: [int(seed[0]) % p])


@pytest.mark.parametrize('p', [P32, P63])
def test_mimc_with_cyclic_constants_over_small_fields(p):
    r = random.Random(p)
    consts = [r.randrange(p) for _ in range(16)]
    steps = 256
    ctl = airs.run_mimc(steps, consts, 3, p)
    opts = dict(hashAlgorithm='blake2s256', extensionFactor=8, exeQueryCount=24, friQueryCount=12)
    _check(_mimc(p, steps, consts), opts, [dict(step=0, register=0, value=3), dict(step=steps - 1, register=0, value=ctl[-1])], [], [3])


def _two_registers_with_public_input(p, steps, cubic):
    """cubic: r0' = r0^2 * r1 + pub[step] (degree 3), r1' = r1^2 + 1 (degree 2) -- two degree groups, both below the combination
    degree 4T; else r0' = r0 * r1 + pub[step] (degree 2), r1' = r1 + 1 (degree 1).  A public input register, two asserted registers."""
    t, e = ProgramBuilder(p), ProgramBuilder(p)
    if cubic:
        t.out(0, t.cur(0) * t.cur(0) * t.cur(1) + t.static(0)); t.out(1, t.cur(1) * t.cur(1) + 1)
        e.out(0, e.nxt(0) - (e.cur(0) * e.cur(0) * e.cur(1) + e.static(0))); e.out(1, e.nxt(1) - (e.cur(1) * e.cur(1) + 1))
    else:
        t.out(0, t.cur(0) * t.cur(1) + t.static(0)); t.out(1, t.cur(1) + 1)
        e.out(0, e.nxt(0) - (e.cur(0) * e.cur(1) + e.static(0))); e.out(1, e.nxt(1) - (e.cur(1) + 1))
    return AirModule(name='two_regs', modulus=p, trace_register_count=2, trace_length=steps, transition=t.build(), evaluation=e.build(),
                     static_registers=[StaticRegister('input', secret=False)], extension_factor=8,
                     init=lambda inputs, seed: [int(seed[0]) % p, int(seed[1]) % p],
                     expand_inputs=lambda inputs: [[int(v) % p for v in inputs[0]]],
                     expand_public_inputs=lambda public: [[int(v) % p for v in public[0]]],
                     input_shapes=lambda inputs: [[steps]])


def _run_two_registers(p, steps, pub, cubic):
    x, y = 7, 11
    for s in range(steps - 1):
        x, y = ((x * x * y + pub[s]) % p, (y * y + 1) % p) if cubic else ((x * y + pub[s]) % p, (y + 1) % p)
    return x, y


def test_mixed_degrees_public_input_and_two_asserted_registers():
    p, steps = P32, 128
    r = random.Random(5)
    pub = [r.randrange(p) for _ in range(steps)]
    x, y = _run_two_registers(p, steps, pub, True)
    a = [dict(step=0, register=0, value=7), dict(step=steps - 1, register=0, value=x), dict(step=steps - 1, register=1, value=y), dict(step=0, register=1, value=11)]
    opts = dict(hashAlgorithm='sha256', extensionFactor=8, exeQueryCount=30, friQueryCount=12)
    air = _two_registers_with_public_input(p, steps, True)
    assert air.constraint_degrees == [3, 2]
    st, got = _check(air, opts, a, [pub], [7, 11], [pub])
    wrong = list(pub); wrong[9] = (wrong[9] + 1) % p              # the verifier needs the right public inputs
    with pytest.raises(StarkError):
        st.verify(a, got, [wrong])


def test_degree_one_next_to_degree_two_overshoots_like_the_restated_protocol():
    """a degree-1 constraint raised by x^T and divided by Z(x) has degree exactly T = compositionDegree, one more than the remainder
    check allows (tests/test_edge_cases_gpu.py has the same finding on the device): the host path fails with the oracle's message"""
    p, steps = P32, 128
    pub = list(range(steps))
    x, y = _run_two_registers(p, steps, pub, False)
    a = [dict(step=0, register=0, value=7), dict(step=steps - 1, register=0, value=x)]
    opts = dict(hashAlgorithm='sha256', extensionFactor=4, exeQueryCount=30, friQueryCount=12)
    air = _two_registers_with_public_input(p, steps, False)
    with pytest.raises(Exception, match='Remainder is not a valid degree 31 polynomial'):
        OracleStark(air, opts).prove(a, [pub], [7, 11])
    with pytest.raises(StarkError, match='Low degree proof failed: Remainder is not a valid degree 31 polynomial'):
        instantiate(air, opts).prove_bytes(a, [pub], [7, 11])


def test_tampered_proofs_and_wrong_assertions_are_rejected():
    air = airs.foo(64)
    opts = dict(hashAlgorithm='blake2s256', extensionFactor=16, exeQueryCount=20, friQueryCount=10)
    a = [dict(step=0, register=0, value=1), dict(step=63, register=0, value=127)]
    st = instantiate(air, opts)
    good = st.prove_bytes(a, [[1]], [])
    r = random.Random(1)
    rejected = 0
    for _ in range(60):
        bad = bytearray(good)
        i = r.randrange(len(bad) - 6)                   # the tail is the input-shape section
        bad[i] ^= 1 << r.randrange(8)
        try:
            ok = st.verify(a, bytes(bad))
        except StarkError:
            rejected += 1
            continue
        assert not ok
    assert rejected == 60
    with pytest.raises(StarkError):
        st.verify([dict(step=0, register=0, value=2), dict(step=63, register=0, value=127)], good)
    with pytest.raises(StarkError, match='conflicts with execution trace'):
        st.prove_bytes([dict(step=0, register=0, value=1), dict(step=63, register=0, value=128)], [[1]], [])
    shapes_changed = bytearray(good); shapes_changed[-4] ^= 3                  # iShapes [[1]] -> [[2]]: well-formed, not this instance's
    with pytest.raises(StarkError, match='input shapes'):
        st.verify(a, bytes(shapes_changed))
    for cut in list(range(0, len(good) - 1, 61)) + [len(good) - 1]:          # every truncation is an error, never a crash
        with pytest.raises(StarkError):
            st.verify(a, good[:cut])
    with pytest.raises(StarkError):
        st.verify(a, good + b'\x00' * 7 if False else bytes(len(good)))          # all zeros


def test_the_128_bit_field_never_takes_the_host_path():
    from genstark_b200 import _native
    from genstark_b200.air import pack_air
    import ctypes as C
    air = airs.mimc128(64)
    with pytest.raises(StarkError, match='at most 64 bits'):
        HostStark(air, dict(extensionFactor=8))
    blob = pack_air(air.with_options(8))
    err = C.create_string_buffer(512)
    out_p, out_n = C.POINTER(C.c_uint8)(), C.c_size_t()
    a = bytes(24)
    rc = _native.lib().gs_host_stark_prove(blob, len(blob), 0, 20, 10, a, 1, bytes(16), None, None, 0, C.byref(out_p), C.byref(out_n), err, 512)
    assert rc == -3 and b'at most 64 bits' in err.value


def test_airassembly_source_over_the_32_bit_field_takes_the_host_path():
    """instantiate(source, component, options) with an AirAssembly module over p32 (the MiMC module of the parity tests with its
    field replaced): the same text that runs on the device over p128 proves on the host over p32, bytes equal the oracle's"""
    from asm_sources import MIMC_SOURCE
    from genstark_b200 import assembly
    src = MIMC_SOURCE.replace('STEPS', '128').replace('340282366920938463463374607393113505793', str(P32))
    opts = dict(hashAlgorithm='blake2s256', extensionFactor=8, exeQueryCount=24, friQueryCount=12)
    st = instantiate(src, 'mimc', opts)
    assert isinstance(st, HostStark)
    module = assembly.compile(src).component('mimc').module([], 8)
    ctl = airs.run_mimc(128, module.static_registers[0].values, 3, P32)
    a = [dict(step=0, register=0, value=3), dict(step=127, register=0, value=ctl[-1])]
    got = st.prove_bytes(a, [], [3])
    ora = OracleStark(module, opts)
    assert got == ora.serialize(ora.prove(a, [], [3]))
    assert st.verify(a, got)


from hypothesis import given, settings, strategies as hst          # noqa: E402

_fuzz_case = {}


def _fuzz_proof():
    if not _fuzz_case:
        p = P32
        r = random.Random(11)
        consts = [r.randrange(p) for _ in range(8)]
        steps = 128
        ctl = airs.run_mimc(steps, consts, 5, p)
        opts = dict(hashAlgorithm='sha256', extensionFactor=8, exeQueryCount=24, friQueryCount=12)
        a = [dict(step=0, register=0, value=5), dict(step=steps - 1, register=0, value=ctl[-1])]
        st = instantiate(_mimc(p, steps, consts), opts)
        _fuzz_case.update(st=st, a=a, buf=st.prove_bytes(a, [], [5]))
    return _fuzz_case['st'], _fuzz_case['a'], _fuzz_case['buf']


@settings(max_examples=300, deadline=None)
@given(data=hst.data())
def test_small_field_verifier_rejects_mutations_without_crashing(data):
    """gs_host_stark_verify reads attacker-controlled bytes: every mutation (the input-shape tail included: this parser checks the final
    offset) is a StarkError, never an acceptance, never a crash"""
    st, a, buf = _fuzz_proof()
    kind = data.draw(hst.sampled_from(['flip', 'truncate', 'byte', 'splice', 'zero_run', 'extend']))
    t = bytearray(buf)
    if kind == 'flip':
        pos = data.draw(hst.integers(0, len(t) - 1)); t[pos] ^= 1 << data.draw(hst.integers(0, 7))
    elif kind == 'truncate':
        t = t[:data.draw(hst.integers(0, len(t) - 1))]
    elif kind == 'byte':
        pos = data.draw(hst.integers(0, len(t) - 1)); old = t[pos]; t[pos] = data.draw(hst.integers(0, 255).filter(lambda v: v != old))
    elif kind == 'splice':
        i = data.draw(hst.integers(0, len(t) - 2)); j = data.draw(hst.integers(i + 1, min(len(t), i + 200)))
        k = data.draw(hst.integers(0, len(t) - (j - i)))
        t[k:k + j - i] = t[i:j]
    elif kind == 'extend':
        t += bytes(data.draw(hst.integers(1, 9)))
    else:
        i = data.draw(hst.integers(0, len(t) - 1)); n = data.draw(hst.integers(1, 40))
        t[i:i + n] = bytes(len(t[i:i + n]))
    if bytes(t) == buf:
        return
    try:
        # the wire format's one non-canonical spot (serialization.ts:25-124): the "first node is a raw leaf" bit of a node column
        # changes nothing when a row is as long as a digest (4 x 8-byte elements = 32 bytes here), nor for an empty column; a
        # mutant that parses to the same proof is the same proof, not a forgery (seen 4 times in 40 000 mutations)
        if st.serialize(st.parse(bytes(t))) == buf:
            return
    except Exception:
        pass
    with pytest.raises(StarkError):
        st.verify(a, bytes(t))


def test_baseline_config_1_through_the_workload_table():
    from genstark_b200 import workloads
    air, opts, a, inputs, seed, desc = workloads.config('1')
    assert 'config 1' in desc
    _check(air, opts, a, inputs, seed)
