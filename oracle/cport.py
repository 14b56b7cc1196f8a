"""ORACLE (test infrastructure) -- ctypes wrapper of the plain-C port (oracle/c/stark_oracle.c).

Used by tests/ to check GPU proof bytes at sizes the Python restatement cannot reach, and by bench.py
as the CPU baseline (`kind: "port"`).  Never imported by genstark_b200/."""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'liboracle.so')
_lib = None


def build():
    subprocess.run(['make', '-s', '-C', HERE], check=True)


def available() -> bool:
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(LIB)
        L.oracle_prove.restype = C.c_int
        L.oracle_prove.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p,
                                   C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_double)]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_transform.restype = C.c_int
        L.oracle_transform.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
        _lib = L
    return _lib


STAGES = ['context', 'trace', 'P(x) iNTT', 'P(x) LDE', 'leaf hashing', 'merkle tree', 'composition C(x)',
          'linear combination', 'FRI layers', 'queries + serialize']


def prove(air, options, assertions, inputs=None, seed=None, threads=None, stages=None) -> bytes:
    """Same arguments as oracle.stark.Stark(air, options).prove(...); returns the serialized proof."""
    from genstark_b200.air import pack_air
    L = lib()
    if threads:
        L.oracle_set_threads(int(threads))
    air = air.with_options(options.get('extensionFactor'))
    p = air.modulus
    blob = pack_air(air)
    alg = ['sha256', 'blake2s256'].index(options.get('hashAlgorithm') or 'sha256')
    a_blob = b''.join(struct.pack('<II', int(a['register']), int(a['step'])) + (int(a['value']) % p).to_bytes(16, 'little')
                      for a in assertions)
    init = b''.join((int(v) % p).to_bytes(16, 'little') for v in air.init(inputs or [], seed or []))
    from genstark_b200.air import input_blob
    in_blob = input_blob(air, inputs)
    shapes = air.input_shapes(inputs or [])
    s_blob = bytes([len(shapes)]) + b''.join(bytes([len(s)]) + b''.join(struct.pack('<I', x) for x in s) for s in shapes)
    out_p, out_n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = (C.c_double * 12)()
    rc = L.oracle_prove(blob, alg, int(options.get('exeQueryCount') or 80), int(options.get('friQueryCount') or 40),
                        a_blob, len(assertions), init, in_blob, s_blob, len(s_blob), C.byref(out_p), C.byref(out_n), st)
    if rc != 0:
        raise RuntimeError(L.oracle_last_error().decode())
    data = C.string_at(out_p, out_n.value)
    L.oracle_free(out_p)
    if stages is not None:
        stages.extend(zip(STAGES, list(st)[:len(STAGES)]))
    return data


def transform(raw: bytes, log_t: int, log_n: int, inverse: bool = False) -> bytes:
    """galois evalPolyAtRoots (zero-padded to 2^log_n) / interpolateRoots of one vector of 2^log_t little-endian residues"""
    from oracle.field import PrimeField
    from genstark_b200.air import P128
    root = PrimeField(P128).get_root_of_unity(1 << log_n)
    out = C.create_string_buffer(16 << log_n)
    if lib().oracle_transform(raw, log_t, log_n, 1 if inverse else 0, root.to_bytes(16, 'little'), out) != 0:
        raise RuntimeError('oracle_transform: bad arguments')
    return out.raw
