"""Method-by-method parity of the device FiniteField / Hash / MerkleTree seam against the
oracle's restatement of galois / merkle -- bit exact."""
import random

import pytest

from genstark_b200.air import P128
from genstark_b200.field import GpuHash, MerkleTree
from oracle.field import PrimeField
from oracle.merkle import Hash as OHash, MerkleTree as OTree
from util import rand_elems, gpu_field

pytestmark = pytest.mark.gpu
OF = PrimeField(P128)


def test_div_with_zero_convention():
    f = gpu_field()
    a, b = rand_elems(5000, 1), rand_elems(5000, 2)
    for i in range(0, 5000, 97):
        b[i] = 0
    got = f.divVectorElements(f.newVectorFrom(a), f.newVectorFrom(b)).toValues()
    assert got == OF.div_vector_elements(a, b)
    assert got[0 if b[0] == 0 else 97 * 0] == (0 if b[0] == 0 else got[0])


def test_combine_power_series_pluck_transpose():
    f = gpu_field()
    vs = [rand_elems(1024, 10 + i) for i in range(5)]
    ks = rand_elems(5, 99)
    got = f.combineManyVectors([f.newVectorFrom(v) for v in vs], ks).toValues()
    assert got == OF.combine_many_vectors(vs, ks)
    base = OF.get_root_of_unity(2**20)
    assert f.getPowerSeries(base, 70001).toValues() == OF.get_power_series(base, 70001)
    v = rand_elems(4096, 5)
    V = f.newVectorFrom(v)
    assert f.pluckVector(V, 512, 64).toValues() == OF.pluck_vector(v, 512, 64)
    assert f.transposeVector(V, 4).toValues() == OF.transpose_vector(v, 4)
    assert f.transposeVector(V, 4, 16).toValues() == OF.transpose_vector(v, 4, 16)


@pytest.mark.parametrize('depth', [0, 1, 2])
def test_fri_fold_equals_quartic_interpolation(depth):
    """K4 against the reference's generic route: interpolateQuarticBatch + evalQuarticBatch (LowDegreeProver.ts:190-195)."""
    f = gpu_field()
    n = 1 << 12
    L = n >> (2 * depth)
    v = rand_elems(L, 40 + depth)
    sx = rand_elems(1, 77)[0] or 5
    dom = OF.get_power_series(OF.get_root_of_unity(n), n)
    xs = OF.transpose_vector(dom, 4, 4 ** depth)
    polys = OF.interpolate_quartic_batch(xs, OF.transpose_vector(v, 4))
    want = OF.eval_quartic_batch(polys, sx)
    assert f.friFold(f.newVectorFrom(v), n, depth, sx).toValues() == want


@pytest.mark.parametrize('alg', ['sha256', 'blake2s256'])
@pytest.mark.parametrize('ncols', [1, 2, 4, 5, 12, 17])
def test_merge_vector_rows_and_tree(alg, ncols):
    f = gpu_field()
    n = 256
    cols = [rand_elems(n, 300 + c) for c in range(ncols)]
    h = GpuHash(alg, f.ctx)
    d = h.mergeVectorRows([f.newVectorFrom(c) for c in cols])
    oh = OHash(alg)
    want = oh.merge_vector_rows(cols, 16)
    assert d.toBuffers() == want
    tree = MerkleTree.create(d, h)
    otree = OTree.create(want, oh)
    assert tree.root == otree.root
    r = random.Random(ncols)
    idx = r.sample(range(n), 40)
    values, nodes, depth = tree.proveBatch(idx)
    op = otree.prove_batch(idx)
    assert (values, nodes, depth) == (op.values, op.nodes, op.depth)
    # and the proof verifies with the merkle restatement
    assert OTree.verify_batch(tree.root, idx, op, oh)


@pytest.mark.parametrize('alg', ['sha256', 'blake2s256'])
def test_digest_values_rows(alg):
    f = gpu_field()
    v = rand_elems(4096, 9)
    rows = f.transposeVector(f.newVectorFrom(v), 4)
    h = GpuHash(alg, f.ctx)
    got = h.digestValues(rows).toBuffers()
    oh = OHash(alg)
    m = OF.transpose_vector(v, 4)
    assert got == oh.digest_values(b''.join(OF.vector_to_bytes(r) for r in m), 64)


def test_big_tree_root_matches_host_hashlib():
    """2^20 leaves: the device tree against hashlib, level by level on the host (size-independent check)."""
    import hashlib
    f = gpu_field()
    n = 1 << 16
    col = rand_elems(n, 1234)
    h = GpuHash('blake2s256', f.ctx)
    d = h.mergeVectorRows([f.newVectorFrom(col)])
    level = [hashlib.blake2s(int(x).to_bytes(16, 'little'), digest_size=32).digest() for x in col]
    assert d.toBuffers() == level
    while len(level) > 1:
        level = [hashlib.blake2s(level[2 * i] + level[2 * i + 1], digest_size=32).digest() for i in range(len(level) // 2)]
    assert MerkleTree.create(d, h).root == level[0]


def test_quartic_batch_interpolation_and_evaluation():
    f = gpu_field()
    rows = 3000
    xs = [rand_elems(4, 1000 + i) for i in range(rows)]
    ys = [rand_elems(4, 5000 + i) for i in range(rows)]
    xs[7] = [5, 9, 5, 11]                       # repeated x: zero denominators invert to 0
    xs[8] = [3, 3, 3, 3]
    X, Y = f.newMatrixFrom(xs), f.newMatrixFrom(ys)
    polys = f.interpolateQuarticBatch(X, Y)
    assert (polys.rowCount, polys.colCount) == (rows, 4)
    want = OF.interpolate_quartic_batch(xs, ys)
    assert polys.toValues() == want
    pts = rand_elems(rows, 77)
    assert f.evalQuarticBatch(polys, f.newVectorFrom(pts)).toValues() == OF.eval_quartic_batch(want, pts)
    assert f.evalQuarticBatch(polys, 123456789).toValues() == OF.eval_quartic_batch(want, 123456789)


def test_unfused_fri_layer_equals_the_fused_fold():
    """LowDegreeProver.ts:190-195 spelled with the reference's own sequence of calls vs gs_fri_fold"""
    f = gpu_field()
    n, depth = 2**14, 1
    root = OF.get_root_of_unity(n)
    domain = f.getPowerSeries(root, n)
    L = n >> (2 * depth)
    v = rand_elems(L, 31)
    V = f.newVectorFrom(v)
    x_sets = f.transposeVector(domain, 4, 4 ** depth)
    y_sets = f.transposeVector(V, 4)
    polys = f.interpolateQuarticBatch(x_sets, y_sets)
    special = 0x1234567890abcdef1234567890abcdef % P128
    column = f.evalQuarticBatch(polys, special)
    assert column.toValues() == f.friFold(V, n, depth, special).toValues()


def test_matrix_reshaping_methods_and_dot_product():
    f = gpu_field()
    rows = [rand_elems(300, 40 + i) for i in range(5)]
    vs = [f.newVectorFrom(r) for r in rows]
    M = f.newMatrixFromVectors(vs)
    assert (M.rowCount, M.colCount) == (5, 300) and M.toValues() == rows
    assert [v.toValues() for v in f.matrixRowsToVectors(M)] == rows
    assert f.transposeMatrix(M).toValues() == OF.transpose_matrix(rows)
    wide = [rand_elems(40, 90 + i) for i in range(33)]
    assert f.transposeMatrix(f.newMatrixFrom(wide)).toValues() == OF.transpose_matrix(wide)
    assert f.joinMatrixRows(M).toValues() == OF.join_matrix_rows(rows)
    other = [rand_elems(300, 60 + i) for i in range(5)]
    assert f.subMatrixElementsFromVectors(vs, f.newMatrixFrom(other)).toValues() == OF.sub_matrix_elements_from_vectors(rows, other)
    den = [list(r) for r in other]
    den[2][5] = 0
    assert f.divMatrixElements(M, f.newMatrixFrom(den)).toValues() == OF.div_matrix_elements(rows, den)
    a, b = rand_elems(100000, 1), rand_elems(100000, 2)
    assert f.combineVectors(f.newVectorFrom(a), f.newVectorFrom(b)) == OF.combine_vectors(a, b)
    assert f.combineVectors(f.newVectorFrom(a[:3]), f.newVectorFrom(b[:3])) == OF.combine_vectors(a[:3], b[:3])
    assert M.getValue(3, 17) == rows[3][17] and vs[2].getValue(299) == rows[2][299]
    buf = bytearray(40)
    assert vs[1].copyValue(5, buf, 8) == 16 and int.from_bytes(buf[8:24], 'little') == rows[1][5]
    assert M.rowsToBuffers([4, 0]) == [b''.join(x.to_bytes(16, 'little') for x in rows[4]), b''.join(x.to_bytes(16, 'little') for x in rows[0])]


def test_prng_small_polys_digest_and_verify_batch_through_the_mirror():
    f = gpu_field()
    seed = bytes(range(32))
    assert f.prng(seed) == OF.prng(seed) and f.prng(seed, 9).toValues() == OF.prng(seed, 9)
    assert f.prng(42, 4).toValues() == OF.prng(42, 4)
    xs, ys = rand_elems(6, 1), rand_elems(6, 2)
    poly = f.interpolate(f.newVectorFrom(xs), f.newVectorFrom(ys))
    assert poly.toValues() == OF.interpolate(xs, ys)
    assert f.evalPolyAt(poly, xs[3]) == ys[3]
    assert f.mulPolys(poly, [1, 2, 3]).toValues() == OF.mul_polys(OF.interpolate(xs, ys), [1, 2, 3])
    h = GpuHash('blake2s256', f.ctx)
    oh = OHash('blake2s256')
    assert h.digest(b'abc') == oh.digest(b'abc')
    v = rand_elems(4 * 256, 9)
    d = h.digestValues(f.newMatrixFrom([v[4 * i:4 * i + 4] for i in range(256)]))
    tree = MerkleTree.create(d, h)
    idx = [9, 8, 200, 31]
    proof = tree.proveBatch(idx)
    assert MerkleTree.verifyBatch(tree.root, idx, proof, h)
    assert not MerkleTree.verifyBatch(tree.root, [9, 8, 200, 30], proof, h)


def test_exp_vector_elements_and_mul_matrix_by_vector_run_the_examples_plain_poseidon():
    """examples/poseidon/utils.ts:25-45 written with the FiniteField methods it uses (addVectorElements, expVectorElements,
    mulMatrixByVector, exp, newVectorFrom / toValues): same digest as the independent integer implementation; plus the two new
    methods against Python integers, negative and zero exponents included"""
    import random
    from genstark_b200 import airs
    f = gpu_field()
    p = P128
    m, rf, rp = airs.POSEIDON_WIDTH, airs.POSEIDON_RF, airs.POSEIDON_RP
    mds_rows, ark_rows = airs.poseidon_mds(p), airs.poseidon_round_constants(p, m, rf + rp)
    mds = f.newMatrixFrom(mds_rows)
    ark = [f.newVectorFrom(r) for r in ark_rows]
    inputs = [42, 43]
    state = f.newVectorFrom(inputs + [0] * (m - len(inputs)))
    for i in range(rf + rp):
        state = f.addVectorElements(state, ark[i])
        if i < rf // 2 or i >= rf // 2 + rp:
            state = f.expVectorElements(state, airs.POSEIDON_ALPHA)
        else:
            vals = state.toValues()
            vals[m - 1] = f.exp(vals[m - 1], airs.POSEIDON_ALPHA)
            state = f.newVectorFrom(vals)
        state = f.mulMatrixByVector(mds, state)
    assert state.toValues()[:2] == airs.poseidon_hash(inputs, p)
    r = random.Random(7)
    v = [0, 1, p - 1, 2**127] + [r.randrange(p) for _ in range(29)]
    V = f.newVectorFrom(v)
    for e in (0, 1, 5, 2**64 + 3, p - 2, -1, -3):
        want = [pow(x, e, p) if e >= 0 else (pow(pow(x, p - 2, p), -e, p) if x else 0) for x in v]
        assert f.expVectorElements(V, e).toValues() == want, e
    rows = [[r.randrange(p) for _ in range(33)] for _ in range(7)]
    got = f.mulMatrixByVector(f.newMatrixFrom(rows), V).toValues()
    assert got == [sum(a * b for a, b in zip(row, v)) % p for row in rows]
    with pytest.raises(Exception):
        f.mulMatrixByVector(f.newMatrixFrom(rows), f.newVectorFrom([1, 2, 3]))
