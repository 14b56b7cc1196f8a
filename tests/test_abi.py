"""CPU-side checks: the C-ABI library loads and exports every symbol include/genstark_b200.h declares,
host-side scalar field code matches Python integers, and compute fails loudly without a GPU."""
import ctypes as C
import os
import random
import re

import pytest

from genstark_b200 import _native
from genstark_b200.air import P128

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    header = open(os.path.join(ROOT, 'include', 'genstark_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    names = re.findall(r'\b(gs_[a-z0-9_]+)\s*\(', header)
    assert len(names) >= 10
    L = _native.lib()
    for n in names:
        assert hasattr(L, n), f'{n} declared in the header but not exported'
    # and the Python binding declares prototypes for all of them
    assert set(names) <= set(_native.declared_symbols())


def test_host_scalar_ops_match_python():
    L = _native.lib()
    out = C.create_string_buffer(16)
    enc = lambda v: int(v).to_bytes(16, 'little')
    r = random.Random(5)
    pool = [0, 1, P128 - 1, 2**127, 2**64, 9 * 2**32 - 1]
    for _ in range(3000):
        a = r.choice(pool + [r.randrange(P128)]); b = r.choice(pool + [r.randrange(P128)])
        for op, want in ((0, (a + b) % P128), (1, (a - b) % P128), (2, a * b % P128)):
            assert L.gs_field_scalar_op(op, enc(a), enc(b), out) == 0
            assert int.from_bytes(out.raw, 'little') == want
    assert L.gs_field_scalar_op(3, enc(5), enc(0), out) == 0 and int.from_bytes(out.raw, 'little') == 0
    assert L.gs_field_scalar_op(2, enc(P128), enc(1), out) != 0      # non-canonical input rejected


def test_unsupported_modulus_reports_not_optimized():
    L = _native.lib()
    assert L.gs_field_supported((P128).to_bytes(16, 'little'), 16) == 0
    assert L.gs_field_supported((2**32 - 3 * 2**25 + 1).to_bytes(16, 'little'), 16) != 0


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from genstark_b200.field import Context
    with pytest.raises(_native.NativeError):
        Context()
