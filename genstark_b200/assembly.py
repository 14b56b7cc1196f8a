"""AirAssembly front-end: source text -> ``AirModule``.

The reference hands AirAssembly text (or a path to an ``.aa`` file) to ``instantiate(source, component,
options)`` (/root/reference/index.ts:18-33), which calls ``compile`` / ``instantiate`` of
``@guildofweavers/air-assembly 0.3.6`` -- a package that is not in the reference tree.  This module is a
fresh implementation of the part of that language the reference's own sources use
(/root/reference/assembly/lib128.aa, lib224.aa, examples/mimc/mimc128Assembly.ts:28-51,
examples/elliptic/pointmul.aa):

  module     (field prime p) (const ...)* (function ...)* (export ...)*
  const      (const $name scalar v) | (const $name vector v...) | (const $name matrix (row...)...)
  function   (function $name (result <type>) (param $n <type>)* (local $n <type>)* store.local* <expr>)
  export     (export name (registers r) (constraints k) (steps s) (static ...)? (init ...) (transition ...) (evaluation ...))
  static     (input secret|public [binary] [(childof i)|(peerof i)] [(steps n)] [(shift k)])
             (mask [inverted] (input i))
             (cycle v... | (prng sha256 0xSEED n) | (power b n))
  expr       (scalar v) (vector e...) (matrix (e...)...) (get e i) (slice e a b)            # slice bounds inclusive
             (add a b) (sub a b) (mul a b) (div a b) (exp a k) (prod a b) (neg a) (inv a)
             (load.const n) (load.param n) (load.local n) (store.local n e) (load.trace 0|1) (load.static 0)
             (call $f e...)

Expressions are lowered to the flat register-machine IR of ``air.py`` (one ``Program`` for the transition
function, one for the constraint evaluator), so everything downstream -- the compiled host trace generator,
the NVRTC-specialised constraint kernel, the oracle -- takes AirAssembly modules unchanged.

Semantics the reference leaves to air-assembly and this file therefore ASSUMES ([RECALLED]):
  * degrees: trace / static register 1, constant 0, add/sub max, mul/prod sum, exp k -> k * degree;
  * an input value is held for the whole span of steps it covers (``steps`` at the leaves of the
    parent/child nesting, the sum of the children's spans above), then displaced by ``shift``
    (``shift -1``: step s holds the value of the span step s+1 belongs to);
  * ``mask`` is 1 on the first step of every span of its source register (displaced by the same shift);
  * ``init`` sees the static row of step -1 (mod T): it is the transition *into* step 0, which is where
    ``shift -1`` registers hold the first inputs;
  * trace length = sum of the leaf spans; without input registers, the export's ``(steps n)``.
What the examples do pin is checked in tests/test_assembly.py: the Poseidon hash / Merkle root the
lib128 components compute equal the plain implementations (examples/assembly/lib128.ts:51-118).
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence

from .air import AirModule, Program, ProgramBuilder, StaticRegister, gather_column_blob, gather_columns_blob, prng_sha256, _Node


class AssemblyError(Exception):
    pass


# ------------------------------------------------------------------------------------------ S-expressions
_TOKEN = re.compile(r'\(|\)|[^\s()]+')


def parse_sexpr(text: str):
    """text -> nested lists of atoms (str); `#` starts a comment that runs to the end of the line."""
    lines = []
    for line in text.splitlines():
        k = line.find('#')
        lines.append(line if k < 0 else line[:k])
    stack, cur = [], []
    for tok in _TOKEN.findall('\n'.join(lines)):
        if tok == '(':
            stack.append(cur); cur = []
        elif tok == ')':
            if not stack:
                raise AssemblyError('unbalanced )')
            done, cur = cur, stack.pop()
            cur.append(done)
        else:
            cur.append(tok)
    if stack:
        raise AssemblyError('unbalanced (')
    return cur


def _int(tok) -> int:
    if isinstance(tok, list):
        raise AssemblyError(f'number expected, got {tok}')
    try:
        return int(tok, 0)
    except ValueError:
        raise AssemblyError(f'number expected, got {tok!r}')


# ------------------------------------------------------------------------------------------------ algebra
class _Alg:
    """Scalar arithmetic over either Python ints (concrete evaluation: ``init``) or ProgramBuilder nodes
    (symbolic lowering).  Compile-time constants stay Python ints in both and fold."""

    def __init__(self, p: int, builder: Optional[ProgramBuilder] = None):
        self.p, self.b = p, builder

    def _node(self, x):
        return x if isinstance(x, _Node) else self.b.const(x)

    def add(self, x, y):
        if isinstance(x, int) and isinstance(y, int): return (x + y) % self.p
        return self._node(x) + self._node(y)

    def sub(self, x, y):
        if isinstance(x, int) and isinstance(y, int): return (x - y) % self.p
        return self._node(x) - self._node(y)

    def mul(self, x, y):
        if isinstance(x, int) and isinstance(y, int): return (x * y) % self.p
        return self._node(x) * self._node(y)

    def neg(self, x):
        return (-x) % self.p if isinstance(x, int) else -x

    def inv(self, x):
        if isinstance(x, int): return pow(x, self.p - 2, self.p)        # inv(0) = 0
        return self.b.inv(x)

    def exp(self, x, e: int):
        if isinstance(x, int):
            return pow(x, e, self.p) if e >= 0 else pow(pow(x, self.p - 2, self.p), -e, self.p)
        if e < 0:
            return self.b.exp(self.b.inv(x), -e)                         # exp(b, e<0) = inv(b)^-e
        return self.b.exp(x, e)


def _kind(v) -> str:
    if isinstance(v, list):
        return 'matrix' if v and isinstance(v[0], list) else 'vector'
    return 'scalar'


def _shape(v):
    k = _kind(v)
    return (k,) if k == 'scalar' else (k, len(v)) if k == 'vector' else (k, len(v), len(v[0]))


def _elementwise(alg, op, a, b):
    ka, kb = _kind(a), _kind(b)
    f = getattr(alg, op)
    if ka == 'scalar' and kb == 'scalar':
        return f(a, b)
    if kb == 'scalar':                                   # vector|matrix (op) scalar
        if ka == 'vector': return [f(x, b) for x in a]
        return [[f(x, b) for x in row] for row in a]
    if ka == 'scalar':
        raise AssemblyError(f'{op}: scalar {op} {kb} is not defined (the scalar goes second)')
    if _shape(a) != _shape(b):
        raise AssemblyError(f'{op}: operand shapes differ: {_shape(a)} vs {_shape(b)}')
    if ka == 'vector': return [f(x, y) for x, y in zip(a, b)]
    return [[f(x, y) for x, y in zip(ra, rb)] for ra, rb in zip(a, b)]


def _dot(alg, u, v):
    acc = None
    for x, y in zip(u, v):
        t = alg.mul(x, y)
        acc = t if acc is None else alg.add(acc, t)
    return acc


def _prod(alg, a, b):
    ka, kb = _kind(a), _kind(b)
    if ka == 'vector' and kb == 'vector':
        if len(a) != len(b): raise AssemblyError('prod: vector lengths differ')
        return _dot(alg, a, b)
    if ka == 'matrix' and kb == 'vector':
        if len(a[0]) != len(b): raise AssemblyError('prod: matrix columns != vector length')
        return [_dot(alg, row, b) for row in a]
    if ka == 'matrix' and kb == 'matrix':
        if len(a[0]) != len(b): raise AssemblyError('prod: inner dimensions differ')
        cols = list(zip(*b))
        return [[_dot(alg, row, c) for c in cols] for row in a]
    raise AssemblyError(f'prod: unsupported operands {ka} x {kb}')


# ------------------------------------------------------------------------------------------- declarations
class _Function:
    def __init__(self, name, form):
        self.name, self.params, self.locals, self.body, self.result = name, [], [], [], None
        for item in form:
            head = item[0] if isinstance(item, list) and item else None
            if head == 'result': self.result = item[1:]
            elif head == 'param': self.params.append(item[1] if str(item[1]).startswith('$') else None)
            elif head == 'local': self.locals.append(item[1] if str(item[1]).startswith('$') else None)
            else: self.body.append(item)


class _InputReg:
    def __init__(self, index, form):
        self.index = index
        self.scope = form[1]
        if self.scope not in ('secret', 'public'):
            raise AssemblyError(f'input register {index}: scope must be secret or public')
        self.binary, self.parent, self.peer, self.steps, self.shift = False, None, None, None, 0
        for item in form[2:]:
            if item == 'binary': self.binary = True
            elif item[0] == 'childof': self.parent = _int(item[1])
            elif item[0] == 'peerof': self.peer = _int(item[1])
            elif item[0] == 'steps': self.steps = _int(item[1])
            elif item[0] == 'shift': self.shift = _int(item[1])
            else: raise AssemblyError(f'input register {index}: unknown attribute {item}')


class _Export:
    def __init__(self, name, form, schema):
        self.name, self.schema = name, schema
        self.registers = self.constraints = self.steps = None
        self.static_forms, self.init, self.transition, self.evaluation = [], None, None, None
        for item in form:
            head = item[0]
            if head == 'registers': self.registers = _int(item[1])
            elif head == 'constraints': self.constraints = _int(item[1])
            elif head == 'steps': self.steps = _int(item[1])
            elif head == 'static': self.static_forms = item[1:]
            elif head == 'init': self.init = _Function('init', item[1:])
            elif head == 'transition': self.transition = _Function('transition', item[1:])
            elif head == 'evaluation': self.evaluation = _Function('evaluation', item[1:])
            else: raise AssemblyError(f'export {name}: unknown section {head}')
        if None in (self.registers, self.constraints, self.steps, self.transition, self.evaluation):
            raise AssemblyError(f'export {name}: registers, constraints, steps, transition and evaluation are required')
        # static registers in declaration order
        self.statics = []                # ('input', _InputReg) | ('mask', source index, inverted) | ('cycle', values)
        p = schema.modulus
        for i, f in enumerate(self.static_forms):
            head = f[0]
            if head == 'input':
                self.statics.append(('input', _InputReg(i, f)))
            elif head == 'mask':
                inverted = 'inverted' in f[1:]
                src = [x for x in f[1:] if isinstance(x, list) and x[0] == 'input']
                if len(src) != 1: raise AssemblyError(f'mask register {i}: (input i) expected')
                self.statics.append(('mask', _int(src[0][1]), inverted))
            elif head == 'cycle':
                vals = []
                for x in f[1:]:
                    if isinstance(x, list) and x[0] == 'prng':
                        if x[1] != 'sha256': raise AssemblyError(f'prng {x[1]} is not supported')
                        seed = x[2][2:] if x[2].lower().startswith('0x') else x[2]
                        vals += prng_sha256(bytes.fromhex(seed if len(seed) % 2 == 0 else '0' + seed), _int(x[3]), p)
                    elif isinstance(x, list) and x[0] == 'power':
                        base, cnt = _int(x[1]) % p, _int(x[2])
                        v = 1
                        for _ in range(cnt):
                            vals.append(v); v = v * base % p
                    else:
                        vals.append(_int(x) % p)
                if len(vals) & (len(vals) - 1) or not vals:
                    raise AssemblyError(f'cyclic register {i}: the number of values must be a power of 2')
                self.statics.append(('cycle', vals))
            else:
                raise AssemblyError(f'static register {i}: unknown kind {head}')

    @property
    def input_registers(self) -> List[_InputReg]:
        return [s[1] for s in self.statics if s[0] == 'input']


class AirSchema:
    """compile() result: field, constants, functions, exported components (air-assembly ``AirSchema``)."""

    def __init__(self, text: str):
        forms = parse_sexpr(text)
        if len(forms) != 1 or forms[0][0] != 'module':
            raise AssemblyError('a single (module ...) form is expected')
        self.modulus = None
        self.consts: List = []
        self.const_names: Dict[str, int] = {}
        self.functions: Dict[str, _Function] = {}
        self.function_order: List[str] = []
        self.exports: Dict[str, _Export] = {}
        pending_exports = []
        for form in forms[0][1:]:
            head = form[0]
            if head == 'field':
                if form[1] != 'prime': raise AssemblyError('only prime fields are supported')
                self.modulus = _int(form[2])
            elif head == 'const':
                self._add_const(form)
            elif head == 'function':
                fn = _Function(form[1], form[2:])
                self.functions[form[1]] = fn
                self.function_order.append(form[1])
            elif head == 'export':
                pending_exports.append(form)
            else:
                raise AssemblyError(f'unknown module section {head}')
        if self.modulus is None:
            raise AssemblyError('(field prime p) is missing')
        for form in pending_exports:
            self.exports[form[1]] = _Export(form[1], form[2:], self)

    def _add_const(self, form):
        p = self.modulus
        items = form[1:]
        name = None
        if items and isinstance(items[0], str) and items[0].startswith('$'):
            name, items = items[0], items[1:]
        kind, vals = items[0], items[1:]
        if kind == 'scalar': v = _int(vals[0]) % p
        elif kind == 'vector': v = [_int(x) % p for x in vals]
        elif kind == 'matrix': v = [[_int(x) % p for x in row] for row in vals]
        else: raise AssemblyError(f'const of kind {kind}')
        if name: self.const_names[name] = len(self.consts)
        self.consts.append(v)

    def component(self, name: str = 'default') -> 'AirComponent':
        if name not in self.exports:
            raise AssemblyError(f"component '{name}' is not exported by the module (exports: {', '.join(self.exports)})")
        return AirComponent(self, self.exports[name])


def compile(source) -> AirSchema:   # noqa: A001  (the reference's name: air-assembly compile())
    """bytes -> source text; str -> path of an .aa file when one exists, else source text (index.ts:18-33)."""
    if isinstance(source, (bytes, bytearray)):
        text = bytes(source).decode()
    elif isinstance(source, str) and '(' not in source and os.path.exists(source):
        with open(source) as f:
            text = f.read()
    else:
        text = str(source)
    return AirSchema(text)


# ----------------------------------------------------------------------------------------------- lowering
class _Frame:
    def __init__(self, fn: _Function, args: Sequence):
        if len(args) != len(fn.params):
            raise AssemblyError(f'{fn.name}: {len(fn.params)} arguments expected, {len(args)} given')
        self.fn, self.params, self.locals = fn, list(args), [None] * len(fn.locals)

    @staticmethod
    def _lookup(names, key, what, fn):
        if key.startswith('$'):
            if key not in names: raise AssemblyError(f'{fn.name}: unknown {what} {key}')
            return names.index(key)
        return int(key)


class _Lowering:
    def __init__(self, schema: AirSchema, alg: _Alg, trace_cur, trace_next, static_row):
        self.s, self.alg = schema, alg
        self.trace = [trace_cur, trace_next]
        self.static_row = static_row
        self.depth = 0

    def run(self, fn: _Function, args: Sequence):
        frame = _Frame(fn, args)
        self.depth += 1
        if self.depth > 64: raise AssemblyError('call depth exceeded (recursive function?)')
        result = None
        for form in fn.body:
            result = self.eval(form, frame)
        self.depth -= 1
        if result is None:
            raise AssemblyError(f'{fn.name}: the body has no result expression')
        return result

    def eval(self, e, fr: _Frame):
        alg = self.alg
        if not isinstance(e, list):
            raise AssemblyError(f'expression expected, got {e!r}')
        op = e[0]
        if op == 'scalar': return _int(e[1]) % alg.p
        if op == 'vector':
            out = []
            for x in e[1:]:
                v = self.eval(x, fr)
                k = _kind(v)
                if k == 'matrix': raise AssemblyError('vector: matrix element')
                out += v if k == 'vector' else [v]
            return out
        if op == 'matrix':
            rows = [[self._scalar(self.eval(x, fr)) for x in row] for row in e[1:]]
            if len({len(r) for r in rows}) != 1: raise AssemblyError('matrix: ragged rows')
            return rows
        if op == 'get':
            v = self.eval(e[1], fr)
            if _kind(v) != 'vector': raise AssemblyError('get: vector expected')
            return v[_int(e[2])]
        if op == 'slice':
            v = self.eval(e[1], fr)
            if _kind(v) != 'vector': raise AssemblyError('slice: vector expected')
            a, b = _int(e[2]), _int(e[3])
            if not (0 <= a <= b < len(v)): raise AssemblyError(f'slice [{a}..{b}] out of range for a vector of {len(v)}')
            return v[a:b + 1]
        if op in ('add', 'sub', 'mul'):
            return _elementwise(alg, op, self.eval(e[1], fr), self.eval(e[2], fr))
        if op == 'div':
            a, b = self.eval(e[1], fr), self.eval(e[2], fr)
            kb = _kind(b)
            bi = alg.inv(b) if kb == 'scalar' else [alg.inv(x) for x in b] if kb == 'vector' else [[alg.inv(x) for x in r] for r in b]
            return _elementwise(alg, 'mul', a, bi)
        if op == 'exp':
            a, k = self.eval(e[1], fr), self.eval(e[2], fr)
            if not isinstance(k, int): raise AssemblyError('exp: the exponent must be a constant scalar')
            if k > alg.p // 2: k -= alg.p                     # (scalar -1) was reduced on the way in
            ka = _kind(a)
            return alg.exp(a, k) if ka == 'scalar' else [alg.exp(x, k) for x in a] if ka == 'vector' else [[alg.exp(x, k) for x in r] for r in a]
        if op == 'prod': return _prod(alg, self.eval(e[1], fr), self.eval(e[2], fr))
        if op == 'neg':
            a = self.eval(e[1], fr); ka = _kind(a)
            return alg.neg(a) if ka == 'scalar' else [alg.neg(x) for x in a] if ka == 'vector' else [[alg.neg(x) for x in r] for r in a]
        if op == 'inv':
            a = self.eval(e[1], fr); ka = _kind(a)
            return alg.inv(a) if ka == 'scalar' else [alg.inv(x) for x in a] if ka == 'vector' else [[alg.inv(x) for x in r] for r in a]
        if op == 'load.const':
            key = e[1]
            idx = self.s.const_names.get(key) if key.startswith('$') else int(key)
            if idx is None or idx >= len(self.s.consts): raise AssemblyError(f'unknown constant {key}')
            return self.s.consts[idx]
        if op == 'load.param':
            return fr.params[_Frame._lookup(fr.fn.params, e[1], 'parameter', fr.fn)]
        if op == 'load.local':
            v = fr.locals[_Frame._lookup(fr.fn.locals, e[1], 'local', fr.fn)]
            if v is None: raise AssemblyError(f'{fr.fn.name}: local {e[1]} is read before it is stored')
            return v
        if op == 'store.local':
            v = self.eval(e[2], fr)
            fr.locals[_Frame._lookup(fr.fn.locals, e[1], 'local', fr.fn)] = v
            return None
        if op == 'load.trace':
            t = self.trace[_int(e[1])]
            if t is None: raise AssemblyError(f'(load.trace {e[1]}) is not available in this procedure')
            return list(t)
        if op == 'load.static':
            if self.static_row is None: raise AssemblyError('(load.static 0) is not available in this procedure')
            return list(self.static_row)
        if op == 'call':
            fn = self.s.functions.get(e[1]) if str(e[1]).startswith('$') else self.s.functions.get(self.s.function_order[int(e[1])])
            if fn is None: raise AssemblyError(f'unknown function {e[1]}')
            return self.run(fn, [self.eval(x, fr) for x in e[2:]])
        raise AssemblyError(f'unknown operation {op}')

    @staticmethod
    def _scalar(v):
        if _kind(v) != 'scalar': raise AssemblyError('scalar expected')
        return v


# --------------------------------------------------------------------------------------------- component
def _shape_of(values, rank) -> List[int]:
    """nested lists -> [n0, n1, ...] (uniform nesting required: masks become cyclic registers)"""
    shape, level = [], [values]
    for _ in range(rank):
        lens = {len(x) for x in level}
        if len(lens) != 1:
            raise AssemblyError('ragged inputs: every parent value must have the same number of children')
        shape.append(lens.pop())
        level = [y for x in level for y in x]
    return shape


def _flatten(values, rank):
    level = [values]
    for _ in range(rank):
        level = [y for x in level for y in x]
    return level


class AirComponent:
    """One exported computation; ``module(shapes)`` fixes the input shapes, hence the trace length."""

    def __init__(self, schema: AirSchema, export: _Export):
        self.schema, self.export = schema, export
        self._modules: Dict[tuple, AirModule] = {}
        regs = export.input_registers
        self._rank = []
        for k, r in enumerate(regs):
            ref = r.parent if r.parent is not None else r.peer
            if ref is not None and not (0 <= ref < k):
                raise AssemblyError(f'input register {k}: parent/peer must be declared earlier')
            self._rank.append(1 if ref is None else self._rank[ref] + (1 if r.parent is not None else 0))

    # counts that do not depend on the input shapes
    @property
    def modulus(self): return self.schema.modulus
    @property
    def trace_register_count(self): return self.export.registers
    @property
    def constraint_count(self): return self.export.constraints
    @property
    def input_count(self): return len(self.export.input_registers)
    @property
    def secret_input_count(self): return sum(1 for r in self.export.input_registers if r.scope == 'secret')
    @property
    def public_input_positions(self): return [k for k, r in enumerate(self.export.input_registers) if r.scope == 'public']

    def input_shapes(self, inputs) -> List[List[int]]:
        regs = self.export.input_registers
        inputs = inputs or []
        if len(inputs) != len(regs):
            raise AssemblyError(f'{len(regs)} inputs expected, {len(inputs)} given')
        shapes = [_shape_of(inputs[k], self._rank[k]) for k in range(len(regs))]
        self._check_shapes(shapes)
        return shapes

    def _check_shapes(self, shapes):
        for k, r in enumerate(self.export.input_registers):
            if r.peer is not None and shapes[k] != shapes[r.peer]:
                raise AssemblyError(f'input {k} must have the shape of its peer {r.peer}')
            if r.parent is not None and shapes[k][:-1] != shapes[r.parent]:
                raise AssemblyError(f'input {k} must nest inside its parent {r.parent}')

    def _spans(self, shapes):
        """span (steps covered by one value) of every input register, and the trace length"""
        regs, ex = self.export.input_registers, self.export
        if not regs:
            return [], ex.steps
        children: Dict[int, List[int]] = {}
        for k, r in enumerate(regs):
            if r.parent is not None:
                children.setdefault(r.parent, []).append(k)
        # peers share the children of the register they follow
        root_of = list(range(len(regs)))
        for k, r in enumerate(regs):
            if r.peer is not None: root_of[k] = root_of[r.peer]
        span = [None] * len(regs)

        def span_of(k):
            if span[k] is not None: return span[k]
            base = root_of[k]
            kids = children.get(base, [])
            if kids:
                opts = {shapes[c][-1] * span_of(c) for c in kids}
                if len(opts) != 1: raise AssemblyError(f'children of input {base} cover different numbers of steps')
                s = opts.pop()
            else:
                s = regs[k].steps or regs[base].steps or ex.steps
            span[k] = s
            return s

        for k in range(len(regs)):
            span_of(k)
        totals = set()
        for k in range(len(regs)):
            n = 1
            for d in shapes[k]: n *= d
            totals.add(n * span[k])
        if len(totals) != 1:
            raise AssemblyError('input registers imply different trace lengths')
        t = totals.pop()
        if t & (t - 1) or t % ex.steps:
            raise AssemblyError(f'trace length {t} must be a power of 2 and a multiple of the cycle length {ex.steps}')
        return span, t

    def _lower(self, fn: _Function, builder: ProgramBuilder, with_next: bool, n_static: int, n_out: int, what: str) -> Program:
        alg = _Alg(self.schema.modulus, builder)
        R = self.export.registers
        cur = [builder.cur(i) for i in range(R)]
        nxt = [builder.nxt(i) for i in range(R)] if with_next else None
        st = [builder.static(i) for i in range(n_static)]
        out = _Lowering(self.schema, alg, cur, nxt, st).run(fn, [])
        out = out if isinstance(out, list) else [out]
        if _kind(out) != 'vector' or len(out) != n_out:
            raise AssemblyError(f'{self.export.name}: {what} must yield a vector of {n_out}, got {_shape(out)}')
        for i, v in enumerate(out):
            builder.out(i, v)
        return builder.build()

    def module(self, shapes: Optional[Sequence[Sequence[int]]] = None, extension_factor: Optional[int] = None) -> AirModule:
        ex, p = self.export, self.schema.modulus
        regs = ex.input_registers
        shapes = [list(s) for s in (shapes or [])]
        if len(shapes) != len(regs):
            raise AssemblyError(f'{len(regs)} input shapes expected, {len(shapes)} given')
        key = (tuple(tuple(s) for s in shapes), extension_factor)
        if key in self._modules:
            return self._modules[key]
        self._check_shapes(shapes)
        span, T = self._spans(shapes)
        in_pos = {r.index: k for k, r in enumerate(regs)}          # static index -> input index
        statics = []
        for s in ex.statics:
            if s[0] == 'input':
                statics.append(StaticRegister('input', secret=(s[1].scope == 'secret')))
            elif s[0] == 'cycle':
                if len(s[1]) > T: raise AssemblyError('a cyclic register is longer than the trace')
                statics.append(StaticRegister('cycle', list(s[1])))
            else:
                src = in_pos.get(s[1])
                if src is None: raise AssemblyError(f'mask source {s[1]} is not an input register')
                period, shift = span[src], regs[src].shift
                on, off = (0, 1) if s[2] else (1, 0)
                statics.append(StaticRegister('cycle', [on if (t - shift) % period == 0 else off for t in range(period)]))
        n_static = len(statics)
        transition = self._lower(ex.transition, ProgramBuilder(p), False, n_static, ex.registers, 'transition')
        evaluation = self._lower(ex.evaluation, ProgramBuilder(p), True, n_static, ex.constraints, 'evaluation')
        ranks, shifts = list(self._rank), [r.shift for r in regs]
        binary = [r.binary for r in regs]
        public = self.public_input_positions
        comp = self

        def traces_of(values_list, which):
            out = []
            for vals, k in zip(values_list, which):
                flat = [int(v) % p for v in _flatten(vals, ranks[k])]
                want = 1
                for d in shapes[k]: want *= d
                if len(flat) != want or _shape_of(vals, ranks[k]) != shapes[k]:
                    raise AssemblyError(f'input {k} does not have the shape {shapes[k]} this instance was built for')
                if binary[k] and any(v not in (0, 1) for v in flat):
                    raise AssemblyError(f'input {k} is declared binary')
                sp, sh = span[k], shifts[k]
                out.append([flat[((t - sh) % T) // sp] for t in range(T)])
            return out

        def expand(inputs):
            if len(inputs or []) != len(regs): raise AssemblyError(f'{len(regs)} inputs expected')
            return traces_of(inputs, range(len(regs)))

        index_cache = {}

        def expand_blob(inputs, as_buffer=False):
            """the columns of expand() as bytes, gathered with numpy (prove path)"""
            import numpy as np
            if len(inputs or []) != len(regs): raise AssemblyError(f'{len(regs)} inputs expected')
            out = []
            for k, vals in enumerate(inputs):
                if _shape_of(vals, ranks[k]) != shapes[k]:
                    raise AssemblyError(f'input {k} does not have the shape {shapes[k]} this instance was built for')
                flat = [int(v) % p for v in _flatten(vals, ranks[k])]
                if binary[k] and any(v not in (0, 1) for v in flat):
                    raise AssemblyError(f'input {k} is declared binary')
                if k not in index_cache:
                    index_cache[k] = ((np.arange(T, dtype=np.int64) - shifts[k]) % T) // span[k]
                out.append((flat, index_cache[k]))
            return gather_columns_blob(out, p, as_buffer)

        def expand_public(public_inputs):
            if len(public_inputs or []) != len(public): raise AssemblyError(f'{len(public)} public inputs expected')
            return traces_of(public_inputs, public)

        def static_row(inputs, step):
            tr = iter(expand(inputs)) if regs else iter(())
            row = []
            for s in statics:
                row.append(s.values[step % len(s.values)] if s.kind == 'cycle' else next(tr)[step])
            return row

        def init(inputs, seed):
            if ex.init is None:
                return [int(v) % p for v in (seed or [])]
            row = static_row(inputs or [], T - 1) if statics else None
            low = _Lowering(comp.schema, _Alg(p), None, None, row)
            args = [[int(v) % p for v in (seed or [])]] if ex.init.params else []
            out = low.run(ex.init, args)
            out = out if isinstance(out, list) else [out]
            if len(out) != ex.registers:
                raise AssemblyError(f'init must yield {ex.registers} values, got {len(out)}')
            return out

        m = AirModule(name=ex.name, modulus=p, trace_register_count=ex.registers, trace_length=T,
                      transition=transition, evaluation=evaluation, static_registers=statics,
                      extension_factor=extension_factor, init=init, expand_inputs=expand,
                      expand_public_inputs=expand_public, input_shapes=lambda inputs: [list(s) for s in shapes],
                      expand_inputs_blob=expand_blob if regs else None)
        self._modules[key] = m
        return m

    def module_for(self, inputs, extension_factor: Optional[int] = None) -> AirModule:
        return self.module(self.input_shapes(inputs), extension_factor)
