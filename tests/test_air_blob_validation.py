"""The flattened AIR (air.py: pack_air) crosses the C ABI as bytes, so the library checks every index an instruction carries
before anything -- host interpreter, host code generator, device evaluator -- indexes with it.  These are the cases a fuzzing
campaign over the blob turned up (out-of-range slots and registers, unknown static-register kinds, a composite modulus on the
small-field path) plus the obvious neighbours; each must be refused with a status, never crash or hang.  Host only."""
import ctypes as C
import random
import struct

import pytest

import cases
from genstark_b200 import _native, airs
from genstark_b200.air import pack_air, P32, OP_ADD, OP_CONST, OP_CUR, OP_NEXT, OP_OUT, OP_STATIC
from genstark_b200.stark import input_blob

HDR = 4 + 16 + 20            # magic, modulus, R K log_t log_e n_static


@pytest.fixture(autouse=True)
def _interpreter_only(monkeypatch):
    """mutants would each be compiled by the trace JIT, and compiling the unmutated programs here would pre-fill the in-process
    cache that tests/test_trace_jit.py inspects"""
    monkeypatch.setenv('GS_TRACE_JIT', '0')


def _trace_rc(blob, air, inputs, seed):
    L = _native.lib()
    p = air.modulus
    init = b''.join((int(v) % p).to_bytes(16, 'little') for v in air.init(inputs or [], seed or []))
    out = C.create_string_buffer(16 * air.trace_register_count * air.trace_length * 4 + 4096)
    return L.gs_air_generate_trace(bytes(blob), len(blob), init, input_blob(air, inputs), out)


def _program_offsets(air):
    """byte offsets of the transition and evaluation programs inside the blob"""
    off = HDR
    for reg in air.static_registers:
        off += 8 + (16 * len(reg.values) if reg.kind == 'cycle' else 0)
    off += 4 * air.constraint_count
    t_off = off
    off += 16 + 16 * len(air.transition.instrs) + 16 * len(air.transition.consts)
    return t_off, off


def test_the_packed_blob_itself_is_accepted():
    air, opts, a, inputs, seed = cases.mimc(64, 8)
    assert _trace_rc(pack_air(air), air, inputs, seed) == 0


@pytest.mark.parametrize('field,value', [('dst', 1 << 20), ('dst', 0xFFFFFFFF), ('a', 0xFFFFFFFF), ('a', 1 << 16), ('op', 11), ('op', 0xFFFFFFFF)])
def test_out_of_range_instruction_fields_are_refused(field, value):
    air, opts, a, inputs, seed = cases.mimc(64, 8)
    blob = bytearray(pack_air(air))
    t_off, e_off = _program_offsets(air)
    k = {'op': 0, 'dst': 1, 'a': 2}[field]
    for prog_off, n in ((t_off, len(air.transition.instrs)), (e_off, len(air.evaluation.instrs))):
        for i in range(n):
            t = bytearray(blob)
            struct.pack_into('<I', t, prog_off + 16 + 16 * i + 4 * k, value)
            assert _trace_rc(t, air, inputs, seed) != 0


def test_slot_counts_registers_and_static_kinds_are_checked():
    air, opts, a, inputs, seed = cases.mimc(64, 8)
    blob = bytearray(pack_air(air))
    t_off, e_off = _program_offsets(air)
    for prog_off in (t_off, e_off):
        for n_slots in (0, 0x7FFFFFFF, 0xFFFFFFFF, (1 << 20) + 1):
            t = bytearray(blob); struct.pack_into('<I', t, prog_off + 8, n_slots)
            assert _trace_rc(t, air, inputs, seed) != 0
    # a static register of an unknown kind used to be read as an input register without a trace (segfault)
    for kind in (3, 1 << 20, 0xFFFFFFFF):
        t = bytearray(blob); struct.pack_into('<I', t, HDR, kind)
        assert _trace_rc(t, air, inputs, seed) != 0
    # a transition function cannot read the next state
    for i, ins in enumerate(air.transition.instrs):
        if ins[0] == OP_CUR:
            t = bytearray(blob); struct.pack_into('<I', t, t_off + 16 + 16 * i, OP_NEXT)
            assert _trace_rc(t, air, inputs, seed) != 0
    # a slot read before anything wrote it
    first_arith = next(i for i, ins in enumerate(air.transition.instrs) if ins[0] >= OP_ADD and ins[0] != OP_OUT)
    t = bytearray(blob); struct.pack_into('<I', t, t_off + 16 + 16 * first_arith + 8, air.transition.n_slots - 1 if air.transition.n_slots > 3 else 0)
    _trace_rc(t, air, inputs, seed)                     # either valid (slot already written) or refused: must simply return


def test_random_mutations_of_the_blob_never_crash_the_host_paths():
    """2000 mutants per AIR family through parse + trace generation (interpreter) -- the in-suite slice of the campaign"""
    if True:
        for name, mk in (('mimc', lambda: cases.mimc(64, 8)), ('poseidon', lambda: cases.poseidon(2, 1, e=16)), ('rescue', lambda: cases.rescue(2))):
            air, opts, a, inputs, seed = mk()
            blob = pack_air(air)
            r = random.Random(len(blob))
            for _ in range(2000):
                t = bytearray(blob)
                for _m in range(r.choice([1, 1, 2, 3])):
                    k = r.choice(['flip', 'byte', 'word', 'trunc', 'insert', 'delete'])
                    if len(t) < 8:
                        break
                    if k == 'flip':
                        q = r.randrange(len(t)); t[q] ^= 1 << r.randrange(8)
                    elif k == 'byte':
                        t[r.randrange(len(t))] = r.randrange(256)
                    elif k == 'word':
                        q = r.randrange(0, len(t) - 4) & ~3
                        t[q:q + 4] = struct.pack('<I', r.choice([0, 1, 2, 3, 63, 64, 65, 255, 256, 65535, 65536, 2**24, 2**31 - 1, 2**31, 2**32 - 1]))
                    elif k == 'trunc':
                        t = t[:r.randrange(len(t))]
                    elif k == 'insert':
                        q = r.randrange(len(t)); t[q:q] = bytes(r.randrange(256) for _ in range(r.randrange(1, 17)))
                    else:
                        q = r.randrange(len(t)); del t[q:q + r.randrange(1, 17)]
                if len(t) >= HDR:
                    log_t, log_e = struct.unpack_from('<2I', t, 28)
                    regs = struct.unpack_from('<I', t, 20)[0]
                    if log_t > 8 or log_e > 5 or regs > 4 * air.trace_register_count:
                        continue                        # a legitimate but much larger trace than the output buffer of this test
                _trace_rc(t, air, inputs, seed)


def test_small_field_path_refuses_a_composite_modulus_instead_of_searching_for_a_generator():
    """0xFA100001 = p32 with one bit changed: p - 1 is still divisible by the domain size, but it is not a prime; the generator
    search used to walk the whole 32-bit range (minutes)"""
    import test_host_small_field as T
    L = _native.lib()
    consts = list(range(1, 17))
    air = T._mimc(P32, 64, consts)
    blob = bytearray(pack_air(air))
    assert int.from_bytes(blob[4:20], 'little') == P32
    blob[4:20] = (0xFA100001).to_bytes(16, 'little')
    a_blob = struct.pack('<II', 0, 0) + (3).to_bytes(16, 'little')
    out_p, out_n, err = C.POINTER(C.c_uint8)(), C.c_size_t(), C.create_string_buffer(512)
    rc = L.gs_host_stark_prove(bytes(blob), len(blob), 1, 24, 12, a_blob, 1, (3).to_bytes(16, 'little'), None, bytes([0]), 1,
                               C.byref(out_p), C.byref(out_n), err, 512)
    assert rc != 0 and b'not prime' in err.value
