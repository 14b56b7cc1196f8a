"""Device field arithmetic (PTX) vs Python integers -- bit exact."""
import pytest

from genstark_b200.air import P128
from util import rand_elems, gpu_field

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n', [1, 33, 4096, 100003])
def test_vector_ops_match_python(n):
    f = gpu_field()
    a, b = rand_elems(n, 1), rand_elems(n, 2)
    # make sure the edge values meet each other
    b[:11] = a[:11]
    A, B = f.newVectorFrom(a), f.newVectorFrom(b)
    assert f.addVectorElements(A, B).toValues() == [(x + y) % P128 for x, y in zip(a, b)]
    assert f.subVectorElements(A, B).toValues() == [(x - y) % P128 for x, y in zip(a, b)]
    assert f.mulVectorElements(A, B).toValues() == [(x * y) % P128 for x, y in zip(a, b)]
    s = P128 - 12345
    assert f.mulVectorElements(A, s).toValues() == [(x * s) % P128 for x in a]
    assert f.subVectorElements(A, 1).toValues() == [(x - 1) % P128 for x in a]


def test_mul_extremes():
    f = gpu_field()
    vals = [0, 1, P128 - 1, P128 - 2, 2**127, 2**128 - 2**36, (P128 - 1) // 2, 2**64 - 1, 2**96 - 1, 0xFFFFFFFF,
            P128 - 0xFFFFFFFF, 9 * 2**32 - 1, 2**32]
    a = [x for x in vals for _ in vals]
    b = [y for _ in vals for y in vals]
    a = [x % P128 for x in a]; b = [y % P128 for y in b]
    A, B = f.newVectorFrom(a), f.newVectorFrom(b)
    assert f.mulVectorElements(A, B).toValues() == [(x * y) % P128 for x, y in zip(a, b)]
    assert f.addVectorElements(A, B).toValues() == [(x + y) % P128 for x, y in zip(a, b)]
    assert f.subVectorElements(A, B).toValues() == [(x - y) % P128 for x, y in zip(a, b)]


def test_dedicated_squaring_equals_multiplication():
    """fp_sqr (10 limb products, csrc/fp128.cuh) against fp_mul(a, a): edge values and 64-step chains on 2^15 threads"""
    import ctypes as C
    f = gpu_field()
    ms, bad = C.c_float(), C.c_uint32(12345)
    f.ctx.check(f._lib.gs_debug_sqr_probe(f.ctx.handle, 128, 16, C.byref(ms), C.byref(bad)))
    assert bad.value == 0
