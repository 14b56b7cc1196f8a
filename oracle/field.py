"""ORACLE (test infrastructure, never on the product path) -- prime-field / polynomial semantics.

Restates, in plain Python integers, the subset of ``@guildofweavers/galois 0.4.22`` ``FiniteField``
that genSTARK calls (call sites listed in SURVEY.md §8b).  galois is an npm dependency pinned in
/root/reference/package-lock.json:31-38 and is NOT vendored in the reference tree, so every function
below restates the *published algorithm* and is anchored on the reference's call sites:

  * element encoding   -- little-endian 32-bit limbs, lib/utils/serialization.ts:131-147
  * transposeVector    -- layout evidenced by LowDegreeProver.ts:270-282
  * pluckVector        -- ZeroPolynomial.ts:40
  * inv(0) == 0        -- required for FRI commitment parity, SURVEY App. E.1
  * prng               -- same idiom as QueryIndexGenerator.ts:41-43,61-68 [RECALLED construction]

PARITY UNPINNED: the reference ships no golden vectors for any of these; what pins
this file are the in-repo known-answer values in tests/test_oracle_kat.py.
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence, Union

Vector = List[int]
Matrix = List[List[int]]


def sha256_int(value: Union[int, bytes]) -> int:
    """QueryIndexGenerator.ts:61-68.  ``Buffer.from(hex, 'hex')`` silently drops a trailing odd
    nibble -- replicated, not fixed."""
    if isinstance(value, int):
        h = format(value, 'x')
        value = bytes.fromhex(h[: len(h) // 2 * 2])
    return int.from_bytes(hashlib.sha256(value).digest(), 'big')


class PrimeField:
    def __init__(self, modulus: int):
        self.modulus = int(modulus)
        self.element_size = max(8, (self.modulus.bit_length() + 7) // 8)
        self.one = 1
        self.zero = 0

    # scalars -------------------------------------------------------------------------------------
    def add(self, a, b): return (a + b) % self.modulus
    def sub(self, a, b): return (a - b) % self.modulus
    def mul(self, a, b): return (a * b) % self.modulus
    def neg(self, a): return (-a) % self.modulus

    def inv(self, a):
        a %= self.modulus
        return 0 if a == 0 else pow(a, self.modulus - 2, self.modulus)

    def div(self, a, b): return (a * self.inv(b)) % self.modulus

    def exp(self, b, e):
        """negative exponent = power of the inverse (pinned by the Rescue KAT, hash4x128.ts:14)."""
        if e < 0:
            return pow(self.inv(b), -e, self.modulus)
        return pow(b, e, self.modulus)

    # randomness ----------------------------------------------------------------------------------
    def prng(self, seed: Union[int, bytes], length: Optional[int] = None):
        """SURVEY App. C [RECALLED]: counter mode over sha256, one swappable function."""
        if length is None:
            return sha256_int(seed) % self.modulus
        state = sha256_int(seed)
        return [sha256_int(state + i) % self.modulus for i in range(length)]

    # roots of unity ------------------------------------------------------------------------------
    def get_root_of_unity(self, order: int) -> int:
        """smallest i >= 2 with g = i^((p-1)/order) of exact order ``order``."""
        assert order & (order - 1) == 0
        for i in range(2, self.modulus):
            g = pow(i, (self.modulus - 1) // order, self.modulus)
            if pow(g, order, self.modulus) == 1 and (order == 1 or pow(g, order // 2, self.modulus) != 1):
                return g
        raise ValueError('no root of unity')

    def get_power_series(self, base: int, n: int) -> Vector:
        out = [1] * n
        for i in range(1, n):
            out[i] = out[i - 1] * base % self.modulus
        return out

    # vector ops ----------------------------------------------------------------------------------
    def _b(self, b, n):
        return b if isinstance(b, list) else [b] * n

    def add_vector_elements(self, a: Vector, b) -> Vector:
        p = self.modulus
        return [(x + y) % p for x, y in zip(a, self._b(b, len(a)))]

    def sub_vector_elements(self, a: Vector, b) -> Vector:
        p = self.modulus
        return [(x - y) % p for x, y in zip(a, self._b(b, len(a)))]

    def mul_vector_elements(self, a: Vector, b) -> Vector:
        p = self.modulus
        return [(x * y) % p for x, y in zip(a, self._b(b, len(a)))]

    def inv_vector_elements(self, v: Vector) -> Vector:
        """Montgomery batch inversion that skips zeros, so inv(0) = 0."""
        p = self.modulus
        n = len(v)
        partial = [1] * (n + 1)
        for i, x in enumerate(v):
            partial[i + 1] = partial[i] * (x if x else 1) % p
        acc = self.inv(partial[n])
        out = [0] * n
        for i in range(n - 1, -1, -1):
            x = v[i]
            if x:
                out[i] = partial[i] * acc % p
                acc = acc * x % p
        return out

    def div_vector_elements(self, a: Vector, b) -> Vector:
        if not isinstance(b, list):
            return self.mul_vector_elements(a, self.inv(b))
        return self.mul_vector_elements(a, self.inv_vector_elements(b))

    def combine_vectors(self, a: Vector, b: Vector) -> int:
        return sum(x * y for x, y in zip(a, b)) % self.modulus

    def combine_many_vectors(self, vs: Sequence[Vector], ks: Vector) -> Vector:
        p = self.modulus
        n = len(vs[0])
        out = [0] * n
        for v, k in zip(vs, ks):
            for i in range(n):
                out[i] = (out[i] + v[i] * k) % p
        return out

    def pluck_vector(self, v: Vector, skip: int, times: int) -> Vector:
        n = len(v)
        return [v[(i * skip) % n] for i in range(times)]

    def transpose_vector(self, v: Vector, columns: int, step: int = 1) -> Matrix:
        """rows = len/(columns*step); M[i][j] = v[(i + j*rows) * step]."""
        rows = len(v) // (columns * step)
        return [[v[(i + j * rows) * step] for j in range(columns)] for i in range(rows)]

    def transpose_matrix(self, m: Matrix) -> Matrix:
        return [list(r) for r in zip(*m)]

    def join_matrix_rows(self, m: Matrix) -> Vector:
        return [x for r in m for x in r]

    def sub_matrix_elements_from_vectors(self, vs: Sequence[Vector], m: Matrix) -> Matrix:
        return [self.sub_vector_elements(v, r) for v, r in zip(vs, m)]

    def div_matrix_elements(self, a: Matrix, b: Matrix) -> Matrix:
        return [self.div_vector_elements(x, y) for x, y in zip(a, b)]

    # polynomials (coefficients low -> high) --------------------------------------------------------
    def _fft(self, vals: Vector, roots: Vector) -> Vector:
        n = len(vals)
        if n == 1:
            return vals
        p = self.modulus
        even = self._fft(vals[0::2], roots[0::2])
        odd = self._fft(vals[1::2], roots[0::2])
        h = n // 2
        out = [0] * n
        for i in range(h):
            t = odd[i] * roots[i] % p
            out[i] = (even[i] + t) % p
            out[i + h] = (even[i] - t) % p
        return out

    def eval_poly_at_roots(self, poly: Vector, domain: Vector) -> Vector:
        """forward DFT over ``domain`` (a power series of a root of unity), natural order; the
        polynomial is zero-padded to the domain length (Stark.ts:109, CompositionPolynomial.ts:110)."""
        n = len(domain)
        assert len(poly) <= n
        return self._fft(list(poly) + [0] * (n - len(poly)), domain)

    def eval_polys_at_roots(self, polys: Matrix, domain: Vector) -> Matrix:
        return [self.eval_poly_at_roots(q, domain) for q in polys]

    def interpolate_roots(self, domain: Vector, values):
        """inverse DFT (Stark.ts:106, CompositionPolynomial.ts:109); vector or matrix of rows."""
        if values and isinstance(values[0], list):
            return [self.interpolate_roots(domain, v) for v in values]
        n = len(domain)
        assert len(values) == n
        inv_domain = [domain[0]] + domain[:0:-1]
        ninv = self.inv(n)
        p = self.modulus
        return [x * ninv % p for x in self._fft(list(values), inv_domain)]

    def eval_poly_at(self, poly: Vector, x: int) -> int:
        acc = 0
        for c in reversed(poly):
            acc = (acc * x + c) % self.modulus
        return acc

    def mul_polys(self, a: Vector, b: Vector) -> Vector:
        out = [0] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % self.modulus
        return out

    def interpolate(self, xs: Vector, ys: Vector) -> Vector:
        """Lagrange interpolation (BoundaryConstraints.ts:42, LowDegreeProver.ts:243)."""
        p = self.modulus
        n = len(xs)
        root = [1]
        for x in xs:
            root = self.mul_polys(root, [(-x) % p, 1])
        out = [0] * n
        for i in range(n):
            # numerator = root / (x - xs[i]) by synthetic division
            num = [0] * n
            acc = 0
            for k in range(n, 0, -1):
                acc = (root[k] + acc * xs[i]) % p
                num[k - 1] = acc
            denom = self.eval_poly_at(num, xs[i])
            f = ys[i] * self.inv(denom) % p
            for k in range(n):
                out[k] = (out[k] + num[k] * f) % p
        return out

    def interpolate_quartic_batch(self, xsets: Matrix, ysets: Matrix) -> Matrix:
        return [self.interpolate(x, y) for x, y in zip(xsets, ysets)]

    def eval_quartic_batch(self, polys: Matrix, x) -> Vector:
        xs = x if isinstance(x, list) else [x] * len(polys)
        return [self.eval_poly_at(q, xv) for q, xv in zip(polys, xs)]

    # encoding ------------------------------------------------------------------------------------
    def to_bytes(self, v: int) -> bytes:
        return int(v).to_bytes(self.element_size, 'little')

    def vector_to_bytes(self, v: Vector) -> bytes:
        return b''.join(self.to_bytes(x) for x in v)
