"""Random AirAssembly expressions: the lowering to the flat register-machine program (ProgramBuilder: CSE, load sinking, slot
allocation) computes what direct evaluation over Python integers computes, and the inferred degrees follow the stated rule."""
import random

from hypothesis import given, settings, strategies as st

from genstark_b200 import assembly
from genstark_b200.air import P128, ProgramBuilder
from oracle.air import run_program
from oracle.field import PrimeField

F = PrimeField(P128)
R, S = 3, 2                      # trace registers, static registers


def scalar_expr(draw, depth):
    """-> (sexpr text, degree) of a scalar expression"""
    if depth == 0 or draw(st.integers(0, 9)) < 2:
        k = draw(st.integers(0, 3))
        if k == 0: return f'(scalar {draw(st.integers(0, 2**130))})', 0
        if k == 1: return f'(get (load.trace 0) {draw(st.integers(0, R - 1))})', 1
        if k == 2: return f'(get (load.static 0) {draw(st.integers(0, S - 1))})', 1
        return '(load.const $c)', 0
    op = draw(st.sampled_from(['add', 'sub', 'mul', 'exp', 'neg', 'dot', 'getvec']))
    a, da = scalar_expr(draw, depth - 1)
    if op == 'neg': return f'(neg {a})', da
    if op == 'exp':
        e = draw(st.integers(0, 5))
        return f'(exp {a} (scalar {e}))', da * e if e else 0
    b, db = scalar_expr(draw, depth - 1)
    if op in ('add', 'sub'): return f'({op} {a} {b})', max(da, db)
    if op == 'mul': return f'(mul {a} {b})', da + db
    if op == 'dot': return f'(prod (vector {a} {b}) (vector {b} (scalar 7)))', max(da + db, db)
    return f'(get (slice (mul (vector {a} {b} {a}) {b}) 1 2) 0)', 2 * db      # element 1 of the scaled vector = b * b


@st.composite
def programs(draw):
    n_out = draw(st.integers(1, 3))
    outs = [scalar_expr(draw, draw(st.integers(1, 4))) for _ in range(n_out)]
    return outs


@settings(max_examples=150, deadline=None)
@given(programs(), st.integers(0, 2**32))
def test_lowered_program_equals_direct_evaluation(outs, seed):
    text = '(module (field prime %d) (const $c scalar 12345678901234567890123) (export e (registers %d) (constraints %d) (steps 8) ' \
           '(static (cycle 1 2 3 4) (cycle 5 6)) (init (vector (scalar 1) (scalar 2) (scalar 3))) ' \
           '(transition (load.trace 0)) (evaluation (vector %s))))' % (P128, R, len(outs), ' '.join(e for e, _ in outs))
    schema = assembly.compile(text)
    comp = schema.component('e')
    m = comp.module([])
    r = random.Random(seed)
    cur = [r.randrange(P128) for _ in range(R)]
    stat = [r.randrange(P128) for _ in range(S)]
    got = run_program(F, m.evaluation, cur, [0] * R, stat)
    low = assembly._Lowering(schema, assembly._Alg(P128), cur, [0] * R, stat)
    want = low.run(comp.export.evaluation, [])
    assert got == [int(v) % P128 for v in want]
    # degree rule: the builder's inferred degrees never exceed the syntactic ones (constant folding and CSE can
    # only lower them) and are equal when nothing folds
    for d_inferred, (_, d_syntax) in zip(m.constraint_degrees, outs):
        assert d_inferred <= d_syntax
    # the program respects its own slot bound
    assert all(i[1] < m.evaluation.n_slots or i[0] == 10 for i in m.evaluation.instrs)
