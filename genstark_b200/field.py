"""Device-backed mirror of the galois ``FiniteField`` seam.

genSTARK receives this object as ``context.field`` (lib/Stark.ts:37-43) and calls ~35 of its methods;
the ones that move O(N) data are implemented here on device-resident ``Matrix`` / ``Vector`` handles
through libgenstark_b200.so.  Method names and argument meaning follow the reference interface so the
call sites in lib/*.ts read the same."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

from . import _native
from .air import P128


class Context:
    """One GPU + stream + root tables (gs_ctx)."""

    def __init__(self, device: int = 0):
        self._lib = _native.lib()
        h = C.c_void_p()
        rc = self._lib.gs_ctx_create(device, C.byref(h))
        if rc != 0:
            raise _native.NativeError(rc, self._lib.gs_last_error(None).decode())
        self.handle = h
        self.device = device

    def check(self, rc: int):
        _native.check(self.handle, rc)

    # multi-GPU: one process per GPU; rank 0 creates the id, every rank joins with it
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _native.lib()
        buf = C.create_string_buffer(128)
        rc = lib.gs_comm_unique_id(buf)
        if rc != 0:
            raise _native.NativeError(rc, lib.gs_last_error(None).decode())
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        self.check(self._lib.gs_ctx_comm_init(self.handle, rank, world, unique_id))
        self.rank, self.world = rank, world

    def sync(self):
        self.check(self._lib.gs_ctx_sync(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self._lib.gs_ctx_launch_count(self.handle))

    def close(self):
        if self.handle:
            self._lib.gs_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Matrix:
    """rows x cols field elements in HBM (galois ``Matrix``; a ``Vector`` is a 1-row matrix)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.handle = ctx, handle
        r, c = C.c_int64(), C.c_int64()
        ctx._lib.gs_mat_shape(handle, C.byref(r), C.byref(c))
        self.rowCount, self.colCount = r.value, c.value

    @property
    def length(self) -> int:
        return self.rowCount * self.colCount

    elementSize = 16

    def toBuffer(self) -> bytes:
        buf = C.create_string_buffer(self.length * 16)
        self.ctx.check(self.ctx._lib.gs_mat_to_bytes(self.ctx.handle, self.handle, buf))
        return buf.raw

    def toValues(self):
        raw = self.toBuffer()
        vals = [int.from_bytes(raw[i:i + 16], 'little') for i in range(0, len(raw), 16)]
        if self.rowCount == 1:
            return vals
        return [vals[r * self.colCount:(r + 1) * self.colCount] for r in range(self.rowCount)]

    def getValue(self, row: int, column: Optional[int] = None) -> int:
        """Vector.getValue(i) / Matrix.getValue(row, column) (Stark.ts:290,357-358; LowDegreeProver.ts:141-142)"""
        if column is None:
            row, column = divmod(row, self.colCount) if self.rowCount > 1 else (0, row)
        out = C.create_string_buffer(16)
        self.ctx.check(self.ctx._lib.gs_mat_get(self.ctx.handle, self.handle, row, column, out))
        return int.from_bytes(out.raw, 'little')

    def copyValue(self, index: int, destination: bytearray, offset: int = 0) -> int:
        """Vector.copyValue(index, destination, offset) -> bytes written (Stark.ts:290)"""
        out = C.create_string_buffer(16)
        r, c = divmod(index, self.colCount)
        self.ctx.check(self.ctx._lib.gs_mat_get(self.ctx.handle, self.handle, r, c, out))
        destination[offset:offset + 16] = out.raw
        return 16

    def rowsToBuffers(self, indexes: Optional[Sequence[int]] = None) -> List[bytes]:
        """Matrix.rowsToBuffers(indexes): each selected row as colCount * 16 bytes (LowDegreeProver.ts:53,214-217)"""
        raw, w = self.toBuffer(), self.colCount * 16
        idx = range(self.rowCount) if indexes is None else indexes
        return [raw[i * w:(i + 1) * w] for i in idx]

    def free(self):
        if self.handle:
            self.ctx._lib.gs_mat_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


Vector = Matrix


def seed_bytes(seed) -> bytes:
    """what the reference hashes for a seed: a Buffer as is; a bigint through ``Buffer.from(hex, 'hex')``, which drops a
    trailing odd nibble (QueryIndexGenerator.ts:61-64 idiom)"""
    if isinstance(seed, int):
        h = format(seed, 'x')
        return bytes.fromhex(h[: len(h) // 2 * 2])
    return bytes(seed)


def _enc(v: int) -> bytes:
    return int(v).to_bytes(16, 'little')


class GpuField:
    """FiniteField over p128 whose vectors live on a B200."""

    modulus = P128
    elementSize = 16
    isOptimized = True
    one = 1
    zero = 0

    def __init__(self, ctx: Optional[Context] = None, modulus: int = P128):
        self._lib = _native.lib()
        if self._lib.gs_field_supported(_enc(modulus), 16) != 0:
            raise _native.NativeError(-3, 'only the 128-bit field 2^128 - 9*2^32 + 1 has a native backend')
        self.ctx = ctx or Context()

    # scalars (host) --------------------------------------------------------------------------------
    def _scalar(self, op, a, b) -> int:
        out = C.create_string_buffer(16)
        rc = self._lib.gs_field_scalar_op(op, _enc(a % P128), _enc(b), out)
        if rc != 0:
            raise _native.NativeError(rc, 'scalar op')
        return int.from_bytes(out.raw, 'little')

    def add(self, a, b): return self._scalar(0, a, b % P128)
    def sub(self, a, b): return self._scalar(1, a, b % P128)
    def mul(self, a, b): return self._scalar(2, a, b % P128)
    def div(self, a, b): return self._scalar(3, a, b % P128)
    def neg(self, a): return self._scalar(1, 0, a % P128)
    def inv(self, a): return self._scalar(3, 1, a % P128)

    def exp(self, b, e):
        if e < 0:
            b, e = self.inv(b), -e
        return self._scalar(4, b, e % (P128 - 1) if e >= P128 - 1 else e)

    def getRootOfUnity(self, order: int) -> int:
        out = C.create_string_buffer(16)
        rc = self._lib.gs_field_root_of_unity(order.bit_length() - 1, out)
        if rc != 0 or order & (order - 1):
            raise _native.NativeError(rc, 'order must be a power of two <= 2^32')
        return int.from_bytes(out.raw, 'little')

    # constructors ----------------------------------------------------------------------------------
    def newVectorFrom(self, values: Sequence[int]) -> Matrix:
        return self.newMatrixFrom([list(values)])

    def newMatrixFrom(self, rows: Sequence[Sequence[int]]) -> Matrix:
        r, c = len(rows), len(rows[0])
        raw = b''.join(_enc(v) for row in rows for v in row)
        return self._from_bytes(raw, r, c)

    def _from_bytes(self, raw: bytes, rows: int, cols: int) -> Matrix:
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_mat_from_bytes(self.ctx.handle, raw, rows, cols, C.byref(h)))
        return Matrix(self.ctx, h)

    # element-wise (K2) -------------------------------------------------------------------------------
    def _binary(self, op: int, a: Matrix, b) -> Matrix:
        h = C.c_void_p()
        if isinstance(b, Matrix):
            rc = self._lib.gs_vec_binary(self.ctx.handle, op, a.handle, b.handle, None, C.byref(h))
        else:
            rc = self._lib.gs_vec_binary(self.ctx.handle, op, a.handle, None, _enc(int(b) % P128), C.byref(h))
        self.ctx.check(rc)
        return Matrix(self.ctx, h)

    def addVectorElements(self, a, b): return self._binary(0, a, b)
    def subVectorElements(self, a, b): return self._binary(1, a, b)
    def mulVectorElements(self, a, b): return self._binary(2, a, b)

    def divVectorElements(self, a: Matrix, b) -> Matrix:
        """a * inv(b) element-wise, inv(0) = 0 (CompositionPolynomial.ts:117)"""
        if not isinstance(b, Matrix):
            return self.mulVectorElements(a, self.inv(b))
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_vec_div(self.ctx.handle, a.handle, b.handle, C.byref(h)))
        return Matrix(self.ctx, h)

    divMatrixElements = divVectorElements

    def expVectorElements(self, a: Matrix, exponent: int) -> Matrix:
        """a[i]^exponent element-wise; a negative exponent inverts first (galois; examples/poseidon/utils.ts:37)"""
        exponent = int(exponent)
        if exponent < 0:
            inv = self.divVectorElements(self._ones_like(a), a)
            return self.expVectorElements(inv, -exponent)
        if exponent >= 1 << 128:
            exponent %= (P128 - 1)                      # a^(p-1) = 1 for a != 0, and 0^e = 0 for e > 0 either way
            if exponent == 0:
                exponent = P128 - 1
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_vec_exp(self.ctx.handle, a.handle, exponent.to_bytes(16, 'little'), C.byref(h)))
        return Matrix(self.ctx, h)

    def _ones_like(self, a: Matrix) -> Matrix:
        rows, cols = C.c_int64(), C.c_int64()
        self._lib.gs_mat_shape(a.handle, C.byref(rows), C.byref(cols))
        return self._from_bytes((1).to_bytes(16, 'little') * (rows.value * cols.value), rows.value, cols.value)

    def mulMatrixByVector(self, m: Matrix, v: Matrix) -> Matrix:
        """sum_c m[r][c] * v[c] (examples/poseidon/utils.ts:45)"""
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_mat_mul_vector(self.ctx.handle, m.handle, v.handle, C.byref(h)))
        return Matrix(self.ctx, h)

    def combineManyVectors(self, vectors: Sequence[Matrix], coefficients) -> Matrix:
        ks = coefficients.toValues() if isinstance(coefficients, Matrix) else list(coefficients)
        arr = (C.c_void_p * len(vectors))(*[v.handle for v in vectors])
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_vec_combine_many(self.ctx.handle, arr, len(vectors), b''.join(_enc(k % P128) for k in ks), C.byref(h)))
        return Matrix(self.ctx, h)

    def getPowerSeries(self, base: int, n: int) -> Matrix:
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_power_series(self.ctx.handle, _enc(base % P128), n, C.byref(h)))
        return Matrix(self.ctx, h)

    def pluckVector(self, v: Matrix, skip: int, times: int) -> Matrix:
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_pluck_vector(self.ctx.handle, v.handle, skip, times, C.byref(h)))
        return Matrix(self.ctx, h)

    def transposeVector(self, v: Matrix, columns: int, step: int = 1) -> Matrix:
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_transpose_vector(self.ctx.handle, v.handle, columns, step, C.byref(h)))
        return Matrix(self.ctx, h)

    def friFold(self, v: Matrix, domain_size: int, depth: int, special_x: int) -> Matrix:
        """evalQuarticBatch(interpolateQuarticBatch(transposeVector(domain,4,4^depth), transposeVector(v,4)), x)
        in one kernel (LowDegreeProver.ts:190-195)."""
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_fri_fold(self.ctx.handle, v.handle, domain_size.bit_length() - 1, depth, _enc(special_x % P128), C.byref(h)))
        return Matrix(self.ctx, h)

    # randomness, small polynomials (host side of the library) ------------------------------------------
    def prng(self, seed, length: Optional[int] = None):
        """field.prng(seed) -> element; field.prng(seed, n) -> vector (CompositionPolynomial.ts:58; LowDegreeProver.ts:132)"""
        sb = seed_bytes(seed)
        n = 0 if length is None else int(length)
        out = C.create_string_buffer(16 * max(n, 1))
        rc = self._lib.gs_field_prng(sb, len(sb), n, out)
        if rc != 0:
            raise _native.NativeError(rc, 'prng')
        raw = out.raw
        vals = [int.from_bytes(raw[i:i + 16], 'little') for i in range(0, 16 * max(n, 1), 16)]
        return vals[0] if length is None else self.newVectorFrom(vals)

    @staticmethod
    def _small(v) -> List[int]:
        return v.toValues() if isinstance(v, Matrix) else [int(x) % P128 for x in v]

    def interpolate(self, xs, ys) -> Matrix:
        """Lagrange interpolation of a few points (BoundaryConstraints.ts:42; LowDegreeProver.ts:243)"""
        x, y = self._small(xs), self._small(ys)
        out = C.create_string_buffer(16 * len(x))
        rc = self._lib.gs_poly_interpolate(b''.join(map(_enc, x)), b''.join(map(_enc, y)), len(x), out)
        if rc != 0 or len(x) != len(y):
            raise _native.NativeError(rc, 'interpolate: 1..4096 points, as many ys as xs')
        return self._from_bytes(out.raw, 1, len(x))

    def evalPolyAt(self, poly, x: int) -> int:
        p = self._small(poly)
        out = C.create_string_buffer(16)
        rc = self._lib.gs_poly_eval_at(b''.join(map(_enc, p)), len(p), _enc(x % P128), out)
        if rc != 0:
            raise _native.NativeError(rc, 'evalPolyAt')
        return int.from_bytes(out.raw, 'little')

    def mulPolys(self, a, b) -> Matrix:
        x, y = self._small(a), self._small(b)
        out = C.create_string_buffer(16 * (len(x) + len(y) - 1))
        rc = self._lib.gs_poly_mul(b''.join(map(_enc, x)), len(x), b''.join(map(_enc, y)), len(y), out)
        if rc != 0:
            raise _native.NativeError(rc, 'mulPolys')
        return self._from_bytes(out.raw, 1, len(x) + len(y) - 1)

    def combineVectors(self, a: Matrix, b: Matrix) -> int:
        """sum a[i] * b[i] (CompositionPolynomial.ts:168,188; LinearCombination.ts:85)"""
        out = C.create_string_buffer(16)
        self.ctx.check(self._lib.gs_vec_combine(self.ctx.handle, a.handle, b.handle, out))
        return int.from_bytes(out.raw, 'little')

    def interpolateQuarticBatch(self, xSets: Matrix, ySets: Matrix) -> Matrix:
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_quartic_interpolate_batch(self.ctx.handle, xSets.handle, ySets.handle, C.byref(h)))
        return Matrix(self.ctx, h)

    def evalQuarticBatch(self, polys: Matrix, x) -> Matrix:
        h = C.c_void_p()
        if isinstance(x, Matrix):
            rc = self._lib.gs_quartic_eval_batch(self.ctx.handle, polys.handle, x.handle, None, C.byref(h))
        else:
            rc = self._lib.gs_quartic_eval_batch(self.ctx.handle, polys.handle, None, _enc(int(x) % P128), C.byref(h))
        self.ctx.check(rc)
        return Matrix(self.ctx, h)

    # matrices <-> vectors ------------------------------------------------------------------------------
    def newMatrixFromVectors(self, vectors: Sequence[Matrix]) -> Matrix:
        arr = (C.c_void_p * len(vectors))(*[v.handle for v in vectors])
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_mat_stack(self.ctx.handle, arr, len(vectors), C.byref(h)))
        return Matrix(self.ctx, h)

    vectorsToMatrix = newMatrixFromVectors

    def matrixRowsToVectors(self, m: Matrix) -> List[Matrix]:
        out = []
        for r in range(m.rowCount):
            h = C.c_void_p()
            self.ctx.check(self._lib.gs_mat_rows(self.ctx.handle, m.handle, r, 1, C.byref(h)))
            out.append(Matrix(self.ctx, h))
        return out

    def transposeMatrix(self, m: Matrix) -> Matrix:
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_mat_transpose(self.ctx.handle, m.handle, C.byref(h)))
        return Matrix(self.ctx, h)

    def joinMatrixRows(self, m: Matrix) -> Matrix:
        """row-major concatenation of the rows: a copy with shape 1 x (rows * cols) (LowDegreeProver.ts:182)"""
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_mat_rows(self.ctx.handle, m.handle, 0, m.rowCount, C.byref(h)))
        if self._lib.gs_mat_reshape(h, 1, m.rowCount * m.colCount) != 0:
            raise _native.NativeError(-1, 'reshape')
        return Matrix(self.ctx, h)

    def subMatrixElementsFromVectors(self, vectors, m: Matrix) -> Matrix:
        """row r = vectors[r] - m[r] (BoundaryConstraints.ts:91); ``vectors`` may already be a matrix"""
        v = vectors if isinstance(vectors, Matrix) else self.newMatrixFromVectors(vectors)
        return self._binary(1, v, m)

    # polynomials over roots of unity (K1) ------------------------------------------------------------
    def interpolateRoots(self, domain, values: Matrix) -> Matrix:
        """lib/Stark.ts:106 -- ``domain`` is implied by the length (power series of getRootOfUnity)."""
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_interpolate_roots(self.ctx.handle, values.handle, C.byref(h)))
        return Matrix(self.ctx, h)

    def evalPolysAtRoots(self, polys: Matrix, domain) -> Matrix:
        """lib/Stark.ts:109 -- ``domain`` may be the domain length (int) or an object with ``length``."""
        n = domain if isinstance(domain, int) else domain.length
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_eval_polys_at_roots(self.ctx.handle, polys.handle, n.bit_length() - 1, C.byref(h)))
        return Matrix(self.ctx, h)

    evalPolyAtRoots = evalPolysAtRoots


class Digests:
    """n 32-byte digests in HBM (what hash.mergeVectorRows / digestValues return)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.handle = ctx, handle
        self.length = int(ctx._lib.gs_digests_count(handle))

    def toBuffers(self) -> List[bytes]:
        buf = C.create_string_buffer(self.length * 32)
        self.ctx.check(self.ctx._lib.gs_digests_to_bytes(self.ctx.handle, self.handle, buf))
        raw = buf.raw                       # one copy: `buf.raw` inside the comprehension would copy the whole buffer per digest
        return [raw[i:i + 32] for i in range(0, self.length * 32, 32)]

    def __del__(self):
        try:
            if self.handle:
                self.ctx._lib.gs_digests_free(self.handle)
                self.handle = None
        except Exception:
            pass


class GpuHash:
    """merkle `Hash` (lib/Stark.ts:49-53): createHash(algorithm) with the O(N) methods on the device."""
    ALGORITHMS = ['sha256', 'blake2s256']
    digestSize = 32
    isOptimized = True

    def __init__(self, algorithm: str, ctx: Context):
        if algorithm not in self.ALGORITHMS:
            raise TypeError(f'Hash algorithm {algorithm} is not supported')
        self.algorithm, self.alg, self.ctx = algorithm, self.ALGORITHMS.index(algorithm), ctx
        self._lib = ctx._lib

    def digest(self, value: bytes) -> bytes:
        """hash.digest(buffer) (lib/utils/index.ts:37) -- host"""
        out = C.create_string_buffer(32)
        if self._lib.gs_hash_digest(self.alg, bytes(value), len(value), out) != 0:
            raise _native.NativeError(-1, 'digest')
        return out.raw

    def mergeVectorRows(self, vectors: Sequence[Matrix]) -> Digests:
        arr = (C.c_void_p * len(vectors))(*[v.handle for v in vectors])
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_hash_merge_vector_rows(self.ctx.handle, self.alg, arr, len(vectors), C.byref(h)))
        return Digests(self.ctx, h)

    def digestValues(self, rows: Matrix, valueSize: Optional[int] = None) -> Digests:
        h = C.c_void_p()
        self.ctx.check(self._lib.gs_hash_digest_values(self.ctx.handle, self.alg, rows.handle, C.byref(h)))
        return Digests(self.ctx, h)


class MerkleTree:
    """merkle `MerkleTree` with the nodes in HBM: create / root / proveBatch (lib/Stark.ts:118,150)."""

    def __init__(self, ctx: Context, handle, depth: int):
        self.ctx, self.handle, self.depth = ctx, handle, depth

    @staticmethod
    def create(values: Digests, hash: GpuHash) -> 'MerkleTree':
        h = C.c_void_p()
        hash.ctx.check(hash._lib.gs_merkle_create(hash.ctx.handle, hash.alg, values.handle, C.byref(h)))
        return MerkleTree(hash.ctx, h, values.length.bit_length() - 1)

    @staticmethod
    def _commit(vectors: Sequence[Matrix], hash: GpuHash) -> 'MerkleTree':
        """test hook: mergeVectorRows + create in the launches the prover uses for a commit (gs_debug_commit_columns)"""
        arr = (C.c_void_p * len(vectors))(*[v.handle for v in vectors])
        h = C.c_void_p()
        hash.ctx.check(hash._lib.gs_debug_commit_columns(hash.ctx.handle, hash.alg, arr, len(vectors), C.byref(h)))
        return MerkleTree(hash.ctx, h, vectors[0].length.bit_length() - 1)

    def _nodes(self) -> List[bytes]:
        """test hook: the 2n stored digests (slot 0 unused, [1] = root, leaves at [n, 2n))"""
        count = 2 << self.depth
        buf = C.create_string_buffer(32 * count)
        self.ctx.check(self.ctx._lib.gs_debug_tree_nodes(self.ctx.handle, self.handle, buf, 32 * count))
        raw = buf.raw
        return [raw[32 * i: 32 * i + 32] for i in range(count)]

    @property
    def root(self) -> bytes:
        out = C.create_string_buffer(32)
        self.ctx.check(self.ctx._lib.gs_merkle_root(self.ctx.handle, self.handle, out))
        return out.raw

    def proveBatch(self, indexes: Sequence[int]):
        """-> (values, nodes, depth) of a BatchMerkleProof; values are the leaf digests in input order"""
        idx = (C.c_uint32 * len(indexes))(*indexes)
        cap = 64 + 32 * len(indexes) * (self.depth + 3) + 4 * len(indexes)
        buf = C.create_string_buffer(cap)
        n = C.c_size_t()
        self.ctx.check(self.ctx._lib.gs_merkle_prove_batch(self.ctx.handle, self.handle, idx, len(indexes), buf, cap, C.byref(n)))
        raw = buf.raw[:n.value]
        nv, nc, depth = int.from_bytes(raw[0:4], 'little'), int.from_bytes(raw[4:8], 'little'), int.from_bytes(raw[8:12], 'little')
        off = 12
        values = [raw[off + 32 * i: off + 32 * i + 32] for i in range(nv)]
        off += 32 * nv
        nodes = []
        for _ in range(nc):
            ln = int.from_bytes(raw[off:off + 4], 'little'); off += 4
            nodes.append([raw[off + 32 * j: off + 32 * j + 32] for j in range(ln)]); off += 32 * ln
        return values, nodes, depth

    @staticmethod
    def verifyBatch(root: bytes, indexes: Sequence[int], proof, hash: GpuHash) -> bool:
        """MerkleTree.verifyBatch(root, indexes, proof, hash) (Stark.ts:206; LowDegreeProver.ts:86,109,116); ``proof`` is
        (values, nodes, depth) as returned by proveBatch, or an object with those attributes"""
        values, nodes, depth = proof if isinstance(proof, tuple) else (proof.values, proof.nodes, proof.depth)
        blob = bytearray(len(values).to_bytes(4, 'little') + len(nodes).to_bytes(4, 'little') + int(depth).to_bytes(4, 'little'))
        for v in values:
            if len(v) != 32:
                raise TypeError('leaf values must be 32-byte digests')
            blob += v
        for col in nodes:
            blob += len(col).to_bytes(4, 'little')
            for x in col:
                blob += x
        idx = (C.c_uint32 * len(indexes))(*indexes)
        rc = hash._lib.gs_merkle_verify_batch(hash.alg, bytes(root), idx, len(indexes), bytes(blob), len(blob))
        if rc < 0:
            raise _native.NativeError(rc, 'malformed batch proof')
        return rc == 1

    def __del__(self):
        try:
            if self.handle:
                self.ctx._lib.gs_tree_free(self.handle)
                self.handle = None
        except Exception:
            pass
