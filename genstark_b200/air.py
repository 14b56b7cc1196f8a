"""AIR description consumed by the prover: the stand-in for air-assembly's ``AirModule``.

The reference obtains an ``AirModule`` from ``@guildofweavers/air-assembly`` (not in the reference
tree; used at /root/reference/lib/Stark.ts:40).  What genSTARK needs from it is small
(/root/reference/lib/Stark.ts:40-67,90,177,302,307): the field, the register / constraint counts,
constraint degrees, the extension factor, the static (``cycle`` / input) registers, a transition
function (trace generation) and a constraint evaluator.  This module holds exactly that, as a flat
register-machine program over field elements so the same description can be

  * interpreted on the host (trace generation, C++),
  * interpreted per domain point on the device (constraint evaluation, CUDA),
  * and interpreted independently by the test oracle (``oracle/``).

Degree inference follows the rule recorded in SURVEY.md App. C: trace/static register = 1,
constant = 0, add/sub = max, mul = sum, exp k = k * degree.
"""
from __future__ import annotations

import hashlib
import struct
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

P128 = 2**128 - 9 * 2**32 + 1           # examples/mimc/mimc128.ts:13, assembly/lib128.aa:3
P32 = 2**32 - 3 * 2**25 + 1             # README.md:23, examples/demo/fibonacci.ts:14

# opcodes of the flat program ---------------------------------------------------------------------
OP_CONST = 0    # dst <- consts[a]
OP_CUR = 1      # dst <- trace register a at the current step
OP_NEXT = 2     # dst <- trace register a at the next step (evaluation programs only)
OP_STATIC = 3   # dst <- static register a at the current step
OP_ADD = 4
OP_SUB = 5
OP_MUL = 6
OP_NEG = 7
OP_INV = 8      # inv(0) = 0
OP_EXP = 9      # dst <- slot[a] ** consts[b]   (host / oracle only; small powers are expanded to MULs)
OP_OUT = 10     # output[dst] <- slot[a]

OP_NAMES = ['const', 'cur', 'next', 'static', 'add', 'sub', 'mul', 'neg', 'inv', 'exp', 'out']


@dataclass
class Program:
    """Straight-line program over field-element slots."""
    instrs: List[Tuple[int, int, int, int]]     # (op, dst, a, b)
    consts: List[int]
    n_slots: int
    n_out: int
    out_degrees: List[int]

    def pack(self) -> bytes:
        """Flat little-endian encoding handed to the C ABI (see include/genstark_b200.h)."""
        head = struct.pack('<4I', len(self.instrs), len(self.consts), self.n_slots, self.n_out)
        ins = b''.join(struct.pack('<4I', *i) for i in self.instrs)
        cs = b''.join(int(c).to_bytes(16, 'little') for c in self.consts)
        return head + ins + cs


class _Node:
    __slots__ = ('b', 'idx', 'degree')

    def __init__(self, b, idx, degree):
        self.b, self.idx, self.degree = b, idx, degree

    def _n(self, o):
        return o if isinstance(o, _Node) else self.b.const(o)

    def __add__(self, o): return self.b._bin(OP_ADD, self, self._n(o))
    def __radd__(self, o): return self.b._bin(OP_ADD, self._n(o), self)
    def __sub__(self, o): return self.b._bin(OP_SUB, self, self._n(o))
    def __rsub__(self, o): return self.b._bin(OP_SUB, self._n(o), self)
    def __mul__(self, o): return self.b._bin(OP_MUL, self, self._n(o))
    def __rmul__(self, o): return self.b._bin(OP_MUL, self._n(o), self)
    def __neg__(self): return self.b._un(OP_NEG, self)
    def __pow__(self, e): return self.b.exp(self, e)


class ProgramBuilder:
    """Builds a ``Program`` in SSA form, then packs the values into a small slot file."""

    def __init__(self, modulus: int, expand_exp_up_to: int = 64):
        self.p = modulus
        self.ssa: List[Tuple[int, int, int]] = []      # (op, a, b), value id = position
        self.deg: List[int] = []
        self.consts: List[int] = []
        self._const_ids = {}
        self.outs: dict = {}
        self.expand_exp_up_to = expand_exp_up_to
        self._cse = {}

    # leaves
    def const(self, v: int) -> _Node:
        v = int(v) % self.p
        if v not in self._const_ids:
            self._const_ids[v] = len(self.consts)
            self.consts.append(v)
        return self._emit(OP_CONST, self._const_ids[v], 0, 0)

    def cur(self, r: int) -> _Node: return self._emit(OP_CUR, r, 0, 1)
    def nxt(self, r: int) -> _Node: return self._emit(OP_NEXT, r, 0, 1)
    def static(self, k: int) -> _Node: return self._emit(OP_STATIC, k, 0, 1)

    def inv(self, a: _Node) -> _Node:
        return self._emit(OP_INV, a.idx, 0, a.degree)

    def exp(self, a: _Node, e: int) -> _Node:
        e = int(e)
        if 0 < e <= self.expand_exp_up_to:
            # left-to-right square and multiply; degree = e * degree(a)
            result = None
            for bit in bin(e)[2:]:
                if result is not None:
                    result = self._bin(OP_MUL, result, result)
                if bit == '1':
                    result = a if result is None else self._bin(OP_MUL, result, a)
            result.degree = a.degree * e
            self.deg[result.idx] = result.degree
            return result
        if e == 0:
            return self.const(1)
        ec = e % (self.p - 1)
        if ec not in self._const_ids:
            self._const_ids[ec] = len(self.consts)
            self.consts.append(ec)
        return self._emit(OP_EXP, a.idx, self._const_ids[ec], a.degree * abs(e))

    def out(self, k: int, a) -> None:
        a = a if isinstance(a, _Node) else self.const(a)
        self.outs[k] = a

    # internals
    def _emit(self, op, a, b, degree) -> _Node:
        # common-subexpression elimination: the program is pure, identical (op, a, b) is the same value
        key = (op, a, b) if op not in (OP_ADD, OP_MUL) else (op, min(a, b), max(a, b))
        hit = self._cse.get(key)
        if hit is not None:
            return _Node(self, hit, self.deg[hit])
        self._cse[key] = len(self.ssa)
        self.ssa.append((op, a, b))
        self.deg.append(degree)
        return _Node(self, len(self.ssa) - 1, degree)

    def _bin(self, op, x: _Node, y: _Node) -> _Node:
        d = x.degree + y.degree if op == OP_MUL else max(x.degree, y.degree)
        return self._emit(op, x.idx, y.idx, d)

    def _un(self, op, x: _Node) -> _Node:
        return self._emit(op, x.idx, 0, x.degree)

    def build(self) -> Program:
        n = len(self.ssa)
        n_out = len(self.outs)
        assert sorted(self.outs) == list(range(n_out)), 'outputs must be 0..n-1'
        # liveness: last use of every SSA value
        last = [-1] * n
        for i, (op, a, b) in enumerate(self.ssa):
            if op in (OP_ADD, OP_SUB, OP_MUL):
                last[a] = i; last[b] = i
            elif op in (OP_NEG, OP_INV, OP_EXP):
                last[a] = i
        for k, node in self.outs.items():
            last[node.idx] = n + k
        # dead values are dropped, live ones get slots from a free list; loads (no operands) are emitted right
        # before their first use so that a front-end may create them eagerly without holding slots
        slot_of = [-1] * n
        free: List[int] = []
        n_slots = 0
        instrs = []
        release_at = {}
        for i in range(n):
            if last[i] >= 0:
                release_at.setdefault(last[i], []).append(i)
        leaf_ops = (OP_CONST, OP_CUR, OP_NEXT, OP_STATIC)

        def take_slot():
            nonlocal n_slots
            if free:
                return free.pop()
            n_slots += 1
            return n_slots - 1

        def emit_leaf(v):
            if slot_of[v] < 0:
                slot_of[v] = take_slot()
                instrs.append((self.ssa[v][0], slot_of[v], self.ssa[v][1], 0))

        for i, (op, a, b) in enumerate(self.ssa):
            if last[i] < 0 or op in leaf_ops:
                continue
            if op in (OP_ADD, OP_SUB, OP_MUL):
                for v in (a, b):
                    if self.ssa[v][0] in leaf_ops:
                        emit_leaf(v)
                ia, ib = slot_of[a], slot_of[b]
            else:
                if self.ssa[a][0] in leaf_ops:
                    emit_leaf(a)
                ia, ib = slot_of[a], (b if op == OP_EXP else 0)
            # operands whose last use is this instruction free their slot first (dst may alias)
            for v in release_at.get(i, []):
                free.append(slot_of[v])
            s = take_slot()
            slot_of[i] = s
            instrs.append((op, s, ia, ib))
        for k in range(n_out):
            if self.ssa[self.outs[k].idx][0] in leaf_ops:
                emit_leaf(self.outs[k].idx)
        for k in range(n_out):
            instrs.append((OP_OUT, k, slot_of[self.outs[k].idx], 0))
        degs = [self.outs[k].degree for k in range(n_out)]
        return Program(instrs, list(self.consts), max(n_slots, 1), n_out, degs)


# static registers --------------------------------------------------------------------------------
@dataclass
class StaticRegister:
    """``cycle``: values repeat with period len(values) (power of two dividing the trace length).
    ``input``: a full-length register whose T values are produced from the user inputs at prove
    time; ``secret`` ones are committed next to the trace (Stark.ts:113-115)."""
    kind: str                         # 'cycle' | 'input'
    values: Optional[List[int]] = None
    secret: bool = False


def prng_sha256(seed: bytes, count: int, modulus: int) -> List[int]:
    """air-assembly ``prng.sha256(seed, count, field)`` used for round constants
    (examples/mimc/mimc128.ts:15,36; assembly/lib128.aa).  The package is not in the reference
    tree; construction per SURVEY App. C [RECALLED]: sha256(u16be(i+1) || seed) mod p."""
    out = []
    for i in range(count):
        h = hashlib.sha256(struct.pack('>H', i + 1) + seed).digest()
        out.append(int.from_bytes(h, 'big') % modulus)
    return out


@dataclass
class AirModule:
    """What ``new Stark(...)`` keeps as ``this.air`` (Stark.ts:26,40)."""
    name: str
    modulus: int
    trace_register_count: int
    trace_length: int
    transition: Program                         # outputs: next state, one per trace register
    evaluation: Program                         # outputs: one per constraint
    static_registers: List[StaticRegister] = field(default_factory=list)
    constraint_degrees: Optional[List[int]] = None
    extension_factor: Optional[int] = None
    # (inputs, seed) -> initial state row
    init: Callable = lambda inputs, seed: list(seed)
    # inputs -> list of T-length traces, one per 'input' static register, in register order
    expand_inputs: Callable = lambda inputs: []
    input_shapes: Callable = lambda inputs: []
    # public inputs (verify side) -> T-length traces of the public input registers, in register order
    expand_public_inputs: Callable = lambda public_inputs: []
    # optional fast path of expand_inputs: inputs -> the same columns as one bytes object (16-byte little-endian elements,
    # register after register); used on the prove path, where a Python list of T integers per register costs more than the proof
    expand_inputs_blob: Optional[Callable] = None
    # the same for expand_public_inputs (verify path): public_inputs -> the columns as bytes
    expand_public_inputs_blob: Optional[Callable] = None

    def __post_init__(self):
        if self.constraint_degrees is None:
            self.constraint_degrees = list(self.evaluation.out_degrees)
        if self.extension_factor is None:
            # "smallest power of 2 greater than 2x of the highest constraint degree"
            # (genstark.d.ts:69-73)
            e = 1
            while e < 2 * self.max_constraint_degree:
                e *= 2
            self.extension_factor = e

    @property
    def constraint_count(self) -> int:
        return self.evaluation.n_out

    @property
    def max_constraint_degree(self) -> int:
        return max(self.constraint_degrees)

    @property
    def secret_input_count(self) -> int:
        return sum(1 for s in self.static_registers if s.kind == 'input' and s.secret)

    @property
    def element_size(self) -> int:
        return max(8, (self.modulus.bit_length() + 7) // 8)

    def with_options(self, extension_factor: Optional[int]) -> 'AirModule':
        if extension_factor is None or extension_factor == self.extension_factor:
            return self
        import copy
        m = copy.copy(self)
        m.extension_factor = int(extension_factor)
        return m


def gather_column_blob(values: Sequence[int], index, modulus: int) -> bytes:
    """column[t] = values[index[t]] as 16-byte little-endian elements; ``index`` is a numpy integer array"""
    import numpy as np
    # 16-byte elements are moved as one complex128 item each (a typed take: ~0.1 ms per 2^16 rows; indexing rows of a
    # (n, 16) uint8 table, or the default bounds-checking mode of take, costs 2 ms) -- the range is checked once up front
    table = np.frombuffer(b''.join((int(v) % modulus).to_bytes(16, 'little') for v in values), dtype=np.complex128)
    idx = np.asarray(index).astype(np.intp, copy=False)
    if idx.size and (int(idx.min()) < 0 or int(idx.max()) >= table.shape[0]):
        raise IndexError('input register index outside the input values')
    return np.take(table, idx, mode='clip').tobytes()


def gather_columns_blob(columns, modulus: int, as_buffer: bool = False):
    """several gather_column_blob columns ([(values, index), ...], all of one length) back to back, written straight into one
    buffer: one copy at the end instead of one per column plus a join (5 MB for the Poseidon Merkle-proof inputs).
    as_buffer: return the bytearray the columns were gathered into (no copy at all; input_cbuf hands it to the C ABI)."""
    import numpy as np
    if not columns:
        return bytearray() if as_buffer else b''
    n = len(np.asarray(columns[0][1]))
    buf = bytearray(16 * len(columns) * n)
    out = np.frombuffer(buf, dtype=np.complex128)
    for k, (values, index) in enumerate(columns):
        table = np.frombuffer(b''.join((int(v) % modulus).to_bytes(16, 'little') for v in values), dtype=np.complex128)
        idx = np.asarray(index).astype(np.intp, copy=False)
        if idx.shape != (n,):
            raise ValueError('input register columns must have one length')
        if n and (int(idx.min()) < 0 or int(idx.max()) >= table.shape[0]):
            raise IndexError('input register index outside the input values')
        np.take(table, idx, mode='clip', out=out[k * n:(k + 1) * n])
    del out                                  # release the export so the bytearray can be handed on (or resized) freely
    return buf if as_buffer else bytes(buf)


def input_blob(air: 'AirModule', inputs) -> Optional[bytes]:
    """T-length columns of the input registers as the C ABI takes them (gs_stark_prove: input_traces)"""
    if air.expand_inputs_blob is not None:
        return air.expand_inputs_blob(inputs or []) or None
    traces = air.expand_inputs(inputs or [])
    if not traces:
        return None
    p = air.modulus
    return b''.join((int(v) % p).to_bytes(16, 'little') for t in traces for v in t)


def public_blob(air: 'AirModule', public_inputs) -> Optional[bytes]:
    """T-length columns of the public input registers as gs_stark_verify takes them (public_traces)"""
    if air.expand_public_inputs_blob is not None:
        return air.expand_public_inputs_blob(public_inputs or []) or None
    pub = air.expand_public_inputs(public_inputs or [])
    if not pub:
        return None
    p = air.modulus
    return b''.join((int(v) % p).to_bytes(16, 'little') for t in pub for v in t)


def input_cbuf(air: 'AirModule', inputs):
    """input_blob for the C ABI without the last copy: a ctypes char array over the buffer the columns were gathered into
    (accepted wherever the binding declares c_char_p), or bytes / None when the AIR has no buffer-returning expansion"""
    import ctypes as C
    if air.expand_inputs_blob is not None:
        try:
            buf = air.expand_inputs_blob(inputs or [], as_buffer=True)
        except TypeError:                    # an expansion function that only takes the inputs
            buf = None
        if isinstance(buf, bytearray):
            return (C.c_char * len(buf)).from_buffer(buf) if len(buf) else None
    return input_blob(air, inputs)


AIR_BLOB_MAGIC = 0x52494147      # 'GAIR'


def pack_air(air: AirModule) -> bytes:
    """Flatten an AirModule for gs_stark_create (include/genstark_b200.h)."""
    log_t = air.trace_length.bit_length() - 1
    log_e = air.extension_factor.bit_length() - 1
    if (1 << log_t) != air.trace_length or (1 << log_e) != air.extension_factor:
        raise ValueError('trace length and extension factor must be powers of two')
    out = [struct.pack('<I', AIR_BLOB_MAGIC), int(air.modulus).to_bytes(16, 'little'),
           struct.pack('<5I', air.trace_register_count, air.constraint_count, log_t, log_e, len(air.static_registers))]
    for reg in air.static_registers:
        if reg.kind == 'cycle':
            out.append(struct.pack('<2I', 0, len(reg.values)))
            out.extend(int(v).to_bytes(16, 'little') for v in reg.values)
        else:
            out.append(struct.pack('<2I', 1 if reg.secret else 2, 0))
    out.append(struct.pack(f'<{air.constraint_count}I', *air.constraint_degrees))
    out.append(air.transition.pack())
    out.append(air.evaluation.pack())
    return b''.join(out)
