"""ORACLE (test infrastructure) -- line-by-line restatement of /root/reference/lib/components/*.ts.

Every class follows its TypeScript namesake statement by statement (including the literal
interpolate-on-M / evaluate-on-N detour and the two full NTTs per asserted register) so that the
fused CUDA path is checked against the *unfused* reference data flow.
"""
from __future__ import annotations

import math
from typing import Dict, List

from .field import PrimeField, Vector, Matrix, sha256_int
from .merkle import Hash, MerkleTree, BatchMerkleProof


class StarkError(Exception):
    """lib/StarkError.ts:3-13"""

    def __init__(self, message, cause=None):
        if cause is not None:
            message = f'{message}: {cause}'
        super().__init__(message)


# QueryIndexGenerator.ts ---------------------------------------------------------------------------
class QueryIndexGenerator:
    def __init__(self, extension_factor: int, exe_query_count: int, fri_query_count: int):
        self.extension_factor = extension_factor
        self.exe_query_count = exe_query_count
        self.fri_query_count = fri_query_count

    def get_exe_indexes(self, seed: bytes, domain_size: int) -> List[int]:     # :20-23
        count = min(self.exe_query_count, domain_size - domain_size // self.extension_factor)
        return get_pseudorandom_indexes(seed, count, domain_size, self.extension_factor)

    def get_fri_indexes(self, seed: bytes, column_length: int) -> List[int]:   # :25-27
        return get_pseudorandom_indexes(seed, self.fri_query_count, column_length, self.extension_factor)


def get_pseudorandom_indexes(seed: bytes, count: int, max_: int, exclude_multiples_of: int = 0) -> List[int]:
    """QueryIndexGenerator.ts:32-59"""
    max_count = max_ - max_ // exclude_multiples_of if exclude_multiples_of else max_
    if max_count < count:
        raise ValueError(f'Cannot select {count} unique pseudorandom indexes from {max_} values')
    indexes: Dict[int, None] = {}
    state = sha256_int(seed)
    for i in range(count * 1000):
        index = sha256_int(state + i) % max_
        if exclude_multiples_of and index % exclude_multiples_of == 0:
            continue
        if index in indexes:
            continue
        indexes[index] = None
        if len(indexes) >= count:
            break
    if len(indexes) < count:
        raise ValueError(f'Could not generate {count} pseudorandom indexes')
    return list(indexes)


# ZeroPolynomial.ts --------------------------------------------------------------------------------
class ZeroPolynomial:
    def __init__(self, context):
        self.field: PrimeField = context.field
        self.trace_length = context.trace_length
        position = (self.trace_length - 1) * context.extension_factor
        self.x_at_last_step = self.field.exp(context.root_of_unity, position)          # :19-23

    def evaluate_at(self, x: int) -> int:                                              # :28-34
        f = self.field
        num = f.sub(f.exp(x, self.trace_length), 1)
        den = f.sub(x, self.x_at_last_step)
        return f.div(num, den)

    def evaluate_all(self, domain: Vector):                                            # :36-45
        f = self.field
        x_to_the_steps = f.pluck_vector(domain, self.trace_length, len(domain))
        return (f.sub_vector_elements(x_to_the_steps, 1),
                f.sub_vector_elements(domain, self.x_at_last_step))


# BoundaryConstraints.ts ---------------------------------------------------------------------------
class BoundaryConstraints:
    def __init__(self, assertions, context):
        f = self.field = context.field
        e = context.extension_factor
        r_data: Dict[int, dict] = {}
        for a in assertions:                                                           # :21-36
            x = f.exp(context.root_of_unity, a['step'] * e)
            z_poly = [f.neg(x), 1]
            d = r_data.get(a['register'])
            if d:
                d['xs'].append(x); d['ys'].append(a['value'])
                d['z'] = f.mul_polys(d['z'], z_poly)
            else:
                r_data[a['register']] = {'xs': [x], 'ys': [a['value']], 'z': z_poly}
        self.polys: Dict[int, dict] = {}
        for reg, d in r_data.items():                                                  # :38-44
            self.polys[reg] = {'i': f.interpolate(d['xs'], d['ys']), 'z': d['z']}

    @property
    def count(self) -> int:
        return len(self.polys)

    def evaluate_at(self, p_evaluations: Vector, x: int) -> Vector:                    # :55-69
        f = self.field
        out = []
        for reg, c in self.polys.items():
            z = f.eval_poly_at(c['z'], x)
            i = f.eval_poly_at(c['i'], x)
            out.append(f.div(f.sub(p_evaluations[reg], i), z))
        return out

    def evaluate_all(self, p_evaluations: Matrix, domain: Vector) -> Matrix:           # :71-95
        f = self.field
        p_values = [p_evaluations[reg] for reg in self.polys]
        i_values = f.eval_polys_at_roots([c['i'] for c in self.polys.values()], domain)
        z_values = f.eval_polys_at_roots([c['z'] for c in self.polys.values()], domain)
        pi = f.sub_matrix_elements_from_vectors(p_values, i_values)
        return f.div_matrix_elements(pi, z_values)


# CompositionPolynomial.ts -------------------------------------------------------------------------
class CompositionPolynomial:
    def __init__(self, assertions, seed: bytes, context):
        f = self.field = context.field
        self.b_poly = BoundaryConstraints(assertions, context)
        self.z_poly = ZeroPolynomial(context)
        T = context.trace_length
        max_deg = 1                                                                    # :196-204
        for c in context.constraints:
            max_deg = max(max_deg, c.degree)
        self.combination_degree = 2 ** math.ceil(math.log2(max_deg)) * T
        self.composition_degree = max(self.combination_degree - T, T)                  # :40
        groups: Dict[int, List[int]] = {}                                              # :206-225
        for i, c in enumerate(context.constraints):
            groups.setdefault(c.degree * T, []).append(i)
        self.constraint_groups = list(groups.items())
        d_count = len(context.constraints)                                             # :46-51
        for degree, indexes in self.constraint_groups:
            if degree < self.combination_degree:
                d_count += len(indexes)
        b_count = self.b_poly.count                                                    # :53-56
        if self.composition_degree > T:
            b_count *= 2
        coefficients = f.prng(seed, d_count + b_count)                                 # :58-60
        self.d_coefficients = coefficients[:d_count]
        self.b_coefficients = coefficients[d_count:]

    @property
    def coefficient_count(self) -> int:
        return len(self.d_coefficients) + len(self.b_coefficients)

    def evaluate_all(self, p_polys: Matrix, p_evaluations: Matrix, context) -> Vector:  # :71-146
        f = self.field
        try:
            q_evaluations = context.evaluate_transition_constraints(p_polys)
        except Exception as err:
            raise StarkError('Failed to evaluate transition constraints', err)
        N, M = len(context.evaluation_domain), len(context.composition_domain)
        composition_rou = f.exp(context.root_of_unity, N // M)
        qa = [list(r) for r in q_evaluations]
        for degree, indexes in self.constraint_groups:                                 # :88-100
            if degree == self.combination_degree:
                continue
            power_seed = f.exp(composition_rou, self.combination_degree - degree)
            powers = f.get_power_series(power_seed, M)
            for i in indexes:
                qa.append(f.mul_vector_elements(qa[i], powers))
        qc = f.combine_many_vectors(qa, self.d_coefficients)                           # :105
        qc_poly = f.interpolate_roots(context.composition_domain, qc)                  # :109
        qe = f.eval_poly_at_roots(qc_poly, context.evaluation_domain)                  # :110
        numerators, denominators = self.z_poly.evaluate_all(context.evaluation_domain)  # :114
        z_inverses = f.div_vector_elements(denominators, numerators)                   # :117
        d_evaluations = f.mul_vector_elements(qe, z_inverses)                          # :120
        b_evaluations = self.b_poly.evaluate_all(p_evaluations, context.evaluation_domain)  # :124
        ba = [list(r) for r in b_evaluations]
        b_inc = self.composition_degree - context.trace_length                         # :129-137
        if b_inc > 0:
            powers = f.get_power_series(f.exp(context.root_of_unity, b_inc), N)
            for i in range(self.b_poly.count):
                ba.append(f.mul_vector_elements(ba[i], powers))
        bc = f.combine_many_vectors(ba, self.b_coefficients)                           # :142
        return f.add_vector_elements(d_evaluations, bc)                                # :145

    def evaluate_at(self, x, p_values, n_values, h_values, context) -> int:            # :150-191
        f = self.field
        q_values = list(context.evaluate_constraints_at(x, p_values, n_values, h_values))
        for degree, indexes in self.constraint_groups:
            if degree == self.combination_degree:
                continue
            power = f.exp(x, self.combination_degree - degree)
            for i in indexes:
                q_values.append(f.mul(q_values[i], power))
        qc = f.combine_vectors(q_values, self.d_coefficients)
        d_value = f.div(qc, self.z_poly.evaluate_at(x))
        b_values = self.b_poly.evaluate_at(p_values, x)
        b_inc = self.composition_degree - context.trace_length
        if b_inc > 0:
            power = f.exp(x, b_inc)
            for i in range(self.b_poly.count):
                b_values.append(f.mul(b_values[i], power))
        return f.add(d_value, f.combine_vectors(b_values, self.b_coefficients))


# LinearCombination.ts -----------------------------------------------------------------------------
class LinearCombination:
    def __init__(self, seed: bytes, composition_degree: int, coefficient_offset: int, context):
        self.field: PrimeField = context.field
        self.seed = seed
        self.root_of_unity = context.root_of_unity
        self.domain_size = context.trace_length * context.extension_factor
        self.coefficient_offset = coefficient_offset
        self.ps_incremental_degree = composition_degree - context.trace_length         # :33
        self.coefficients = None

    def compute_many(self, c_evaluations: Vector, p_evaluations: Matrix, s_evaluations: List[Vector]) -> Vector:
        f = self.field                                                                 # :36-64
        ps = [list(r) for r in p_evaluations] + [list(s) for s in s_evaluations]
        ps2 = []
        if self.ps_incremental_degree > 0:
            powers = f.get_power_series(f.exp(self.root_of_unity, self.ps_incremental_degree), self.domain_size)
            for v in ps:
                ps2.append(f.mul_vector_elements(v, powers))
        all_evaluations = ps + ps2
        coefficients = f.prng(self.seed, self.coefficient_offset + len(all_evaluations))
        self.coefficients = coefficients[self.coefficient_offset:]
        return f.add_vector_elements(c_evaluations, f.combine_many_vectors(all_evaluations, self.coefficients))

    def compute_one(self, x, d_value, p_values, s_values) -> int:                      # :66-88
        f = self.field
        ps = list(p_values) + list(s_values)
        ps2 = []
        if self.ps_incremental_degree > 0:
            ps2 = f.mul_vector_elements(ps, f.exp(x, self.ps_incremental_degree))
        all_values = ps + ps2
        if self.coefficients is None:
            c = f.prng(self.seed, self.coefficient_offset + len(all_values))
            self.coefficients = c[self.coefficient_offset:]
        return f.add(d_value, f.combine_vectors(all_values, self.coefficients))


# LowDegreeProver.ts -------------------------------------------------------------------------------
MAX_REMAINDER_LENGTH = 256


def _augmented_positions(positions: List[int], column_length: int) -> List[int]:        # :302-309
    row_length = column_length // 4
    out: Dict[int, None] = {}
    for p in positions:
        out[p % row_length] = None
    return list(out)


class LowDegreeProver:
    def __init__(self, idx_generator: QueryIndexGenerator, hash: Hash, context):
        self.field: PrimeField = context.field
        self.poly_row_size = self.field.element_size * 4
        self.root_of_unity = context.root_of_unity
        self.hash = hash
        self.idx = idx_generator

    def _rows_to_bytes(self, m: Matrix) -> bytes:
        return b''.join(self.field.vector_to_bytes(r) for r in m)

    def _rows_to_buffers(self, m: Matrix, positions: List[int]) -> List[bytes]:
        return [self.field.vector_to_bytes(m[p]) for p in positions]

    def prove(self, c_evaluations: Vector, domain: Vector, max_degree_plus_1: int) -> dict:   # :39-68
        f = self.field
        poly_values = f.transpose_vector(c_evaluations, 4)
        poly_hashes = self.hash.digest_values(self._rows_to_bytes(poly_values), self.poly_row_size)
        p_tree = MerkleTree.create(poly_hashes, self.hash)
        exe_positions = self.idx.get_exe_indexes(p_tree.root, len(domain))
        lc_positions = _augmented_positions(exe_positions, len(c_evaluations))
        lc_proof = p_tree.prove_batch(lc_positions)
        lc_proof.values = self._rows_to_buffers(poly_values, lc_positions)
        # getComponentCount uses Math.min (sic, :287-291): 0 for N >= 128, negative -> RangeError below
        component_count = min(math.ceil(math.log2(len(c_evaluations)) / 2) - 4, 0)
        if component_count < 0:
            raise ValueError('Invalid array length')
        proof = {'lcRoot': p_tree.root, 'lcProof': lc_proof, 'components': [], 'remainder': []}
        self._fri(p_tree, poly_values, max_degree_plus_1, 0, domain, proof)
        return proof

    def _fri(self, p_tree, poly_values, max_degree_plus_1, depth, domain, result):      # :176-221
        f = self.field
        if len(poly_values) * 4 <= MAX_REMAINDER_LENGTH:
            root_of_unity = f.exp(domain[1], 4 ** depth)
            remainder = f.join_matrix_rows(f.transpose_matrix(poly_values))
            self._verify_remainder(remainder, max_degree_plus_1, root_of_unity)
            result['remainder'] = remainder
            return
        xs = f.transpose_vector(domain, 4, 4 ** depth)
        polys = f.interpolate_quartic_batch(xs, poly_values)
        special_x = f.prng(p_tree.root)
        column = f.eval_quartic_batch(polys, special_x)
        new_poly_values = f.transpose_vector(column, 4)
        row_hashes = self.hash.digest_values(self._rows_to_bytes(new_poly_values), self.poly_row_size)
        c_tree = MerkleTree.create(row_hashes, self.hash)
        self._fri(c_tree, new_poly_values, max_degree_plus_1 // 4, depth + 1, domain, result)
        positions = self.idx.get_fri_indexes(c_tree.root, len(column))
        augmented = _augmented_positions(positions, len(column))
        column_proof = c_tree.prove_batch(augmented)
        column_proof.values = self._rows_to_buffers(new_poly_values, augmented)
        poly_proof = p_tree.prove_batch(positions)
        poly_proof.values = self._rows_to_buffers(poly_values, positions)
        while len(result['components']) <= depth:
            result['components'].append(None)
        result['components'][depth] = {'columnRoot': c_tree.root, 'columnProof': column_proof,
                                       'polyProof': poly_proof}

    def _verify_remainder(self, remainder: Vector, max_degree_plus_1: int, root_of_unity: int):  # :223-252
        f = self.field
        e = self.idx.extension_factor
        positions = [i for i in range(len(remainder)) if (not e) or (i % e)]
        domain = f.get_power_series(root_of_unity, len(remainder))
        xs = [domain[positions[i]] for i in range(max_degree_plus_1)]
        ys = [remainder[positions[i]] for i in range(max_degree_plus_1)]
        poly = f.interpolate(xs, ys)
        for i in range(max_degree_plus_1, len(positions)):
            p = positions[i]
            if f.eval_poly_at(poly, domain[p]) != remainder[p]:
                raise StarkError(f'Remainder is not a valid degree {max_degree_plus_1 - 1} polynomial')

    # verifier -------------------------------------------------------------------------------------
    def _read(self, buf: bytes, offset: int) -> int:
        return int.from_bytes(buf[offset:offset + self.field.element_size], 'little')

    def _parse_column_values(self, buffers, positions, augmented, column_length):       # :270-282
        row_length = column_length // 4
        es = self.field.element_size
        out = []
        for p in positions:
            idx = augmented.index(p % row_length)
            out.append(self._read(buffers[idx], (p // row_length) * es))
        return out

    def _rehash(self, proof: BatchMerkleProof) -> BatchMerkleProof:                     # utils/index.ts:34-45
        return BatchMerkleProof([self.hash.digest(v) for v in proof.values], proof.nodes, proof.depth)

    def verify(self, proof: dict, lc_values: Vector, exe_positions: List[int], max_degree_plus_1: int) -> bool:
        f = self.field                                                                  # :70-172
        root_of_unity = self.root_of_unity
        column_length = 1
        r = root_of_unity
        while r != 1:                                                                   # :293-300
            column_length *= 2
            r = f.mul(r, r)
        quartic = [1, f.exp(root_of_unity, column_length // 4), f.exp(root_of_unity, column_length // 2),
                   f.exp(root_of_unity, column_length * 3 // 4)]
        lc_proof = proof['lcProof']
        lc_positions = _augmented_positions(exe_positions, column_length)
        lc_checks = self._parse_column_values(lc_proof.values, exe_positions, lc_positions, column_length)
        if not MerkleTree.verify_batch(proof['lcRoot'], lc_positions, self._rehash(lc_proof), self.hash):
            raise StarkError('Verification of linear combination Merkle proof failed')
        for a, b in zip(lc_values, lc_checks):
            if a != b:
                raise StarkError('Verification of linear combination correctness failed')
        p_root = proof['lcRoot']
        column_length //= 4
        es = f.element_size
        for depth, comp in enumerate(proof['components']):
            column_root = comp['columnRoot']
            positions = self.idx.get_fri_indexes(column_root, column_length)
            augmented = _augmented_positions(positions, column_length)
            column_values = self._parse_column_values(comp['columnProof'].values, positions, augmented, column_length)
            if not MerkleTree.verify_batch(column_root, augmented, self._rehash(comp['columnProof']), self.hash):
                raise StarkError(f'Verification of column Merkle proof failed at depth {depth}')
            poly_values = [[self._read(b, j * es) for j in range(4)] for b in comp['polyProof'].values]
            if not MerkleTree.verify_batch(p_root, positions, self._rehash(comp['polyProof']), self.hash):
                raise StarkError(f'Verification of polynomial Merkle proof failed at depth {depth}')
            xs = []
            for pos in positions:
                xe = f.exp(root_of_unity, pos)
                xs.append([f.mul(q, xe) for q in quartic])
            special_x = f.prng(p_root)
            polys = f.interpolate_quartic_batch(xs, poly_values)
            p_evals = f.eval_quartic_batch(polys, special_x)
            for i in range(len(polys)):
                if p_evals[i] != column_values[i]:
                    raise StarkError(f"Degree 4 polynomial didn't evaluate to column value at depth {depth}")
            p_root = column_root
            root_of_unity = f.exp(root_of_unity, 4)
            max_degree_plus_1 //= 4
            column_length //= 4
        if max_degree_plus_1 > len(proof['remainder']):
            raise StarkError('Remainder degree is greater than number of remainder values')
        remainder = list(proof['remainder'])
        poly_values = f.transpose_vector(remainder, 4)
        hashes = self.hash.digest_values(self._rows_to_bytes(poly_values), self.poly_row_size)
        c_tree = MerkleTree.create(hashes, self.hash)
        if c_tree.root != p_root:
            raise StarkError('Remainder values do not match Merkle root of the last column')
        self._verify_remainder(remainder, max_degree_plus_1, root_of_unity)
        return True
