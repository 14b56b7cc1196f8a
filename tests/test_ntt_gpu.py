"""K1 parity: CUDA NTT / iNTT / coset LDE vs the oracle's restatement of galois interpolateRoots /
evalPolysAtRoots (bit exact), plus size-independent properties at BASELINE sizes."""
import pytest

from genstark_b200.air import P128
from oracle.field import PrimeField
from util import rand_elems, gpu_field

pytestmark = pytest.mark.gpu
OF = PrimeField(P128)


def _domain(n):
    return OF.get_power_series(OF.get_root_of_unity(n), n)


@pytest.mark.parametrize('log_n', list(range(1, 15)))
def test_forward_and_inverse_match_oracle(log_n):
    f = gpu_field()
    n = 1 << log_n
    rows = 3 if log_n < 12 else 1
    polys = [rand_elems(n, 100 + log_n + r) for r in range(rows)]
    dom = _domain(n)
    P = f.newMatrixFrom(polys)
    ev = f.evalPolysAtRoots(P, n).toValues()
    ev = ev if rows > 1 else [ev]
    want = OF.eval_polys_at_roots(polys, dom)
    assert ev == want
    back = f.interpolateRoots(None, f.newMatrixFrom(want)).toValues()
    back = back if rows > 1 else [back]
    assert back == polys


@pytest.mark.parametrize('log_t,log_e', [(2, 1), (2, 3), (3, 3), (5, 2), (6, 3), (8, 4), (9, 3), (10, 5), (13, 3), (12, 4)])
def test_lde_matches_oracle(log_t, log_e):
    f = gpu_field()
    t, n = 1 << log_t, 1 << (log_t + log_e)
    rows = 2 if n <= 1 << 13 else 1
    polys = [rand_elems(t, 7 * log_t + log_e + r) for r in range(rows)]
    got = f.evalPolysAtRoots(f.newMatrixFrom(polys), n).toValues()
    got = got if rows > 1 else [got]
    assert got == OF.eval_polys_at_roots(polys, _domain(n))


@pytest.mark.parametrize('log_n', [16, 20, 23])
def test_roundtrip_large(log_n):
    f = gpu_field()
    n = 1 << log_n
    import random
    r = random.Random(log_n)
    raw = b''.join((r.getrandbits(127)).to_bytes(16, 'little') for _ in range(n))
    X = f._from_bytes(raw, 1, n)
    Y = f.evalPolysAtRoots(X, n)
    Z = f.interpolateRoots(None, Y)
    assert Z.toBuffer() == raw
    # spot-check a few evaluations against Horner on the host
    y = Y.toBuffer()
    w = OF.get_root_of_unity(n)
    coeffs = [int.from_bytes(raw[i:i + 16], 'little') for i in range(0, 16 * n, 16)] if log_n <= 16 else None
    if coeffs is not None:
        for k in (0, 1, n // 2 + 3, n - 1):
            assert int.from_bytes(y[16 * k:16 * k + 16], 'little') == OF.eval_poly_at(coeffs, pow(w, k, P128))


@pytest.mark.parametrize('log_t,log_e', [(16, 4), (20, 3)])
def test_lde_properties_large(log_t, log_e):
    """evaluations on the sub-domain {w^(E*i)} equal the size-T transform; LDE is linear."""
    f = gpu_field()
    t, e = 1 << log_t, 1 << log_e
    import random
    r = random.Random(log_t * 31 + log_e)
    raw = b''.join((r.getrandbits(127)).to_bytes(16, 'little') for _ in range(t))
    X = f._from_bytes(raw, 1, t)
    big = f.evalPolysAtRoots(X, t * e).toBuffer()
    small = f.evalPolysAtRoots(X, t).toBuffer()
    for q in (0, 1, 12345 % t, t - 1):
        assert big[16 * q * e:16 * q * e + 16] == small[16 * q:16 * q + 16]
    # one off-subgroup point by Horner (only for the smaller case, host Horner is O(T))
    if log_t <= 16:
        coeffs = [int.from_bytes(raw[i:i + 16], 'little') for i in range(0, 16 * t, 16)]
        w = OF.get_root_of_unity(t * e)
        for k in (1, e + 1, t * e - 1):
            assert int.from_bytes(big[16 * k:16 * k + 16], 'little') == OF.eval_poly_at(coeffs, pow(w, k, P128))


# ---- every element against the C oracle at the sizes prove() transforms at (the two-pass kernels of csrc/ntt2.cuh, 2^16..2^20)
def _rand_raw(n, seed):
    import random
    r = random.Random(seed)
    raw = bytearray(r.randbytes(16 * n))
    raw[15::16] = bytes(x & 0x7F for x in raw[15::16])      # canonical residues (top bit clear); edge values planted below
    for i, v in enumerate([0, 1, P128 - 1, P128 - 2, 2**127, 9 * 2**32 - 1, 2**96]):
        raw[16 * (i * 5 % n):16 * (i * 5 % n) + 16] = v.to_bytes(16, 'little')
    return bytes(raw)


@pytest.mark.parametrize('log_n', [15, 16, 17, 18, 19, 20, 21, 23])
def test_forward_and_inverse_every_element_vs_c_oracle(log_n):
    from oracle import cport
    f = gpu_field()
    n = 1 << log_n
    raw = _rand_raw(n, 1000 + log_n)
    X = f._from_bytes(raw, 1, n)
    Y = f.evalPolysAtRoots(X, n).toBuffer()
    assert Y == cport.transform(raw, log_n, log_n)
    assert f.interpolateRoots(None, f._from_bytes(Y, 1, n)).toBuffer() == raw
    if log_n <= 20:
        assert f.interpolateRoots(None, X).toBuffer() == cport.transform(raw, log_n, log_n, True)


@pytest.mark.parametrize('log_t,log_e,rows', [(16, 5, 3), (16, 1, 2), (17, 4, 1), (18, 3, 2), (19, 2, 1), (20, 3, 1), (20, 4, 1), (15, 3, 2)])
def test_lde_every_element_vs_c_oracle(log_t, log_e, rows):
    from oracle import cport
    f = gpu_field()
    t, n = 1 << log_t, 1 << (log_t + log_e)
    raws = [_rand_raw(t, 77 * log_t + log_e + r) for r in range(rows)]
    got = f.evalPolysAtRoots(f._from_bytes(b''.join(raws), rows, t), n).toBuffer()
    for r in range(rows):
        assert got[16 * n * r:16 * n * (r + 1)] == cport.transform(raws[r], log_t, log_t + log_e), f'row {r}'


@pytest.mark.parametrize('log_t,log_e,parts', [(16, 3, 2), (16, 5, 8), (18, 3, 8), (20, 3, 4), (20, 4, 8), (13, 3, 4)])
def test_sharded_coset_ranges_assemble_to_the_full_lde(log_t, log_e, parts):
    """gs_lde_cosets_into (a rank's share of the coset-sharded LDE): local position q*(E/W) + (j - j0)"""
    import ctypes as C
    from oracle import cport
    f = gpu_field()
    L, ctx = f._lib, f.ctx
    t, e = 1 << log_t, 1 << log_e
    per = e // parts
    raw = _rand_raw(t, 5 * log_t + log_e)
    want = cport.transform(raw, log_t, log_t + log_e)
    src = f._from_bytes(raw, 1, t)
    for part in range(parts):
        dst, work = C.c_void_p(), C.c_void_p()
        ctx.check(L.gs_mat_alloc(ctx.handle, 1, t * per, C.byref(dst)))
        ctx.check(L.gs_mat_alloc(ctx.handle, 1, t * per, C.byref(work)))
        ctx.check(L.gs_lde_cosets_into(ctx.handle, src.handle, dst, work, part * per, log_e))
        buf = C.create_string_buffer(16 * t * per)
        ctx.check(L.gs_mat_to_bytes(ctx.handle, dst, buf))
        # local position q*per + jj  <->  global position q*e + part*per + jj
        import numpy as np
        a = np.frombuffer(buf.raw, dtype=np.uint8).reshape(t, per, 16)
        w = np.frombuffer(want, dtype=np.uint8).reshape(t, e, 16)
        assert (a == w[:, part * per:(part + 1) * per, :]).all(), f'part {part}'
        L.gs_mat_free(dst); L.gs_mat_free(work)
