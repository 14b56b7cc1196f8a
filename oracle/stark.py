"""ORACLE (test infrastructure) -- restatement of /root/reference/lib/Stark.ts, lib/Serializer.ts,
lib/utils/serialization.ts and lib/utils/sizeof.ts.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may
import this package.  PARITY UNPINNED (see oracle/field.py header): the reference ships no stored
proofs; what this oracle guarantees is the reference's *protocol* (fully in-tree) on top of the
dependency semantics recorded in SURVEY.md App. C.
"""
from __future__ import annotations

import math
import struct
from typing import Dict, List, Optional

from genstark_b200 import air as ir
from .air import ProvingContext, VerificationContext
from .components import (CompositionPolynomial, LinearCombination, LowDegreeProver,
                         QueryIndexGenerator, StarkError)
from .field import PrimeField
from .merkle import BatchMerkleProof, Hash, MerkleTree

DEFAULT_EXE_QUERY_COUNT, DEFAULT_FRI_QUERY_COUNT = 80, 40          # Stark.ts:13-14
MAX_EXE_QUERY_COUNT, MAX_FRI_QUERY_COUNT = 128, 64                 # Stark.ts:16-17
HASH_ALGORITHMS = ['sha256', 'blake2s256']                         # Stark.ts:19-20
MAX_ARRAY_LENGTH, MAX_MATRIX_COLUMN_LENGTH = 256, 127              # sizeof.ts:7-8


def pow_log2(base: float, exponent: int) -> float:                 # utils/index.ts:23-30
    twos = 0
    while exponent % 2 == 0:
        twos += 1
        exponent //= 2
    return (2 ** twos) * math.log2(base ** exponent)


class Stark:
    def __init__(self, air: ir.AirModule, options: Optional[dict] = None):
        options = options or {}
        self.air = air.with_options(options.get('extensionFactor'))
        self.field = PrimeField(self.air.modulus)
        # buildSecurityOptions, Stark.ts:318-344
        exe = options.get('exeQueryCount') or DEFAULT_EXE_QUERY_COUNT
        if exe < 1 or exe > MAX_EXE_QUERY_COUNT or int(exe) != exe:
            raise TypeError(f'Execution sample size must be an integer between 1 and {MAX_EXE_QUERY_COUNT}')
        fri = options.get('friQueryCount') or DEFAULT_FRI_QUERY_COUNT
        if fri < 1 or fri > MAX_FRI_QUERY_COUNT or int(fri) != fri:
            raise TypeError(f'FRI sample size must be an integer between 1 and {MAX_FRI_QUERY_COUNT}')
        alg = options.get('hashAlgorithm') or 'sha256'
        if alg not in HASH_ALGORITHMS:
            raise TypeError(f'Hash algorithm {alg} is not supported')
        self.hash = Hash(alg)
        self.index_generator = QueryIndexGenerator(self.air.extension_factor, exe, fri)

    @property
    def security_level(self) -> int:                               # Stark.ts:62-77
        e = self.air.extension_factor
        es = pow_log2(e / self.air.max_constraint_degree, self.index_generator.exe_query_count)
        fs = math.log2(e) * self.index_generator.fri_query_count
        hs = self.hash.digest_size * 4
        return math.floor(min(es, fs, hs))

    # prover, Stark.ts:81-163 ----------------------------------------------------------------------
    def prove(self, assertions: List[dict], inputs=None, seed=None, trace_out: Optional[dict] = None) -> dict:
        if not isinstance(assertions, list):
            raise TypeError('Assertions parameter must be an array')
        if len(assertions) == 0:
            raise TypeError('At least one assertion must be provided')
        context = ProvingContext(self.air, inputs, seed)
        f = context.field
        n = len(context.evaluation_domain)
        try:
            execution_trace = context.generate_execution_trace()
            validate_assertions(execution_trace, assertions)
        except Exception as err:
            raise StarkError('Failed to generate the execution trace', err)
        p_polys = f.interpolate_roots(context.execution_domain, execution_trace)
        p_evaluations = f.eval_polys_at_roots(p_polys, context.evaluation_domain)
        s_evaluations = context.secret_register_traces
        e_vectors = [list(r) for r in p_evaluations] + [list(s) for s in s_evaluations]
        hashed = self.hash.merge_vector_rows(e_vectors, f.element_size)
        e_tree = MerkleTree.create(hashed, self.hash)
        c_poly = CompositionPolynomial(assertions, e_tree.root, context)
        c_evaluations = c_poly.evaluate_all(p_polys, p_evaluations, context)
        l_combination = LinearCombination(e_tree.root, c_poly.composition_degree, c_poly.coefficient_count, context)
        l_evaluations = l_combination.compute_many(c_evaluations, p_evaluations, s_evaluations)
        try:
            ld_prover = LowDegreeProver(self.index_generator, self.hash, context)
            ld_proof = ld_prover.prove(l_evaluations, context.evaluation_domain, c_poly.composition_degree)
        except Exception as err:
            raise StarkError('Low degree proof failed', err)
        positions = self.index_generator.get_exe_indexes(ld_proof['lcRoot'], n)
        augmented = self._augmented_positions(positions, n)
        e_values = [b''.join(f.to_bytes(v[p]) for v in e_vectors) for p in augmented]   # mergeValues :284-296
        e_proof = e_tree.prove_batch(augmented)
        e_proof.values = e_values
        if trace_out is not None:     # intermediate values for stage-level parity tests
            trace_out.update(trace=execution_trace, p_polys=p_polys, p_evaluations=p_evaluations,
                             s_evaluations=s_evaluations, leaf_hashes=hashed, c_evaluations=c_evaluations,
                             l_evaluations=l_evaluations, composition_degree=c_poly.composition_degree,
                             positions=positions, context=context)
        return {'evRoot': e_tree.root, 'evProof': e_proof, 'ldProof': ld_proof,
                'iShapes': context.input_shapes}

    # verifier, Stark.ts:167-248 -------------------------------------------------------------------
    def verify(self, assertions: List[dict], proof: dict, public_inputs=None) -> bool:
        if len(assertions) < 1:
            raise TypeError('At least one assertion must be provided')
        e_root = proof['evRoot']
        e = self.air.extension_factor
        context = VerificationContext(self.air, proof['iShapes'], public_inputs)
        f = context.field
        n = context.trace_length * e
        c_poly = CompositionPolynomial(assertions, e_root, context)
        l_combination = LinearCombination(e_root, c_poly.composition_degree, c_poly.coefficient_count, context)
        positions = self.index_generator.get_exe_indexes(proof['ldProof']['lcRoot'], n)
        augmented = self._augmented_positions(positions, n)
        p_evals: Dict[int, List[int]] = {}
        s_evals: Dict[int, List[int]] = {}
        for i, merged in enumerate(proof['evProof'].values):
            p, s = self._parse_values(merged)
            p_evals[augmented[i]] = p
            s_evals[augmented[i]] = s
        try:
            ev_proof = BatchMerkleProof([self.hash.digest(v) for v in proof['evProof'].values],
                                        proof['evProof'].nodes, proof['evProof'].depth)
            if not MerkleTree.verify_batch(e_root, augmented, ev_proof, self.hash):
                raise StarkError('Verification of evaluation Merkle proof failed')
        except StarkError:
            raise
        except Exception as err:
            raise StarkError('Verification of evaluation Merkle proof failed', err)
        lc_values = []
        for step in positions:
            x = f.exp(context.root_of_unity, step)
            p_values = p_evals[step]
            n_values = p_evals[(step + e) % n]
            s_values = s_evals[step]
            c_value = c_poly.evaluate_at(x, p_values, n_values, s_values, context)
            lc_values.append(l_combination.compute_one(x, c_value, p_values, s_values))
        try:
            ld_prover = LowDegreeProver(self.index_generator, self.hash, context)
            ld_prover.verify(proof['ldProof'], lc_values, positions, c_poly.composition_degree)
        except Exception as err:
            raise StarkError('Verification of low degree failed', err)
        return True

    # helpers --------------------------------------------------------------------------------------
    def _augmented_positions(self, positions: List[int], n: int) -> List[int]:          # Stark.ts:274-282
        skip = self.air.extension_factor
        out: Dict[int, None] = {}
        for p in positions:
            out[p] = None
            out[(p + skip) % n] = None
        return list(out)

    def _parse_values(self, buf: bytes):                                                # Stark.ts:298-313
        es = self.field.element_size
        r, s = self.air.trace_register_count, self.air.secret_input_count
        vals = [int.from_bytes(buf[i * es:(i + 1) * es], 'little') for i in range(r + s)]
        return vals[:r], vals[r:]

    # wire format, Serializer.ts:35-144 + serialization.ts ------------------------------------------
    def serialize(self, proof: dict) -> bytes:
        es, ds = self.field.element_size, self.hash.digest_size
        ev_leaf = (self.air.trace_register_count + self.air.secret_input_count) * es
        ld_leaf = es * 4
        out = bytearray()
        out += proof['evRoot']
        out += write_merkle_proof(proof['evProof'], ev_leaf)
        ld = proof['ldProof']
        out += ld['lcRoot']
        out += write_merkle_proof(ld['lcProof'], ld_leaf)
        out.append(len(ld['components']))
        for c in ld['components']:
            out += c['columnRoot']
            out += write_merkle_proof(c['columnProof'], ld_leaf)
            out += write_merkle_proof(c['polyProof'], ld_leaf)
        rl = len(ld['remainder'])
        out.append(0 if rl == 256 else rl)
        for v in ld['remainder']:
            out += int(v).to_bytes(es, 'little')
        out.append(len(proof['iShapes']))
        for shape in proof['iShapes']:
            out.append(len(shape))
            for level in shape:
                out += struct.pack('<I', level)
        assert len(out) == self.size_of(proof)
        return bytes(out)

    def parse(self, buf: bytes) -> dict:
        es, ds = self.field.element_size, self.hash.digest_size
        ev_leaf = (self.air.trace_register_count + self.air.secret_input_count) * es
        ld_leaf = es * 4
        ev_root = buf[:ds]
        ev_proof, off = read_merkle_proof(buf, ds, ev_leaf, ds)
        lc_root = buf[off:off + ds]; off += ds
        lc_proof, off = read_merkle_proof(buf, off, ld_leaf, ds)
        count = buf[off]; off += 1
        comps = []
        for _ in range(count):
            column_root = buf[off:off + ds]; off += ds
            column_proof, off = read_merkle_proof(buf, off, ld_leaf, ds)
            poly_proof, off = read_merkle_proof(buf, off, ld_leaf, ds)
            comps.append({'columnRoot': column_root, 'columnProof': column_proof, 'polyProof': poly_proof})
        rl = buf[off] or MAX_ARRAY_LENGTH; off += 1
        remainder = []
        for _ in range(rl):
            remainder.append(int.from_bytes(buf[off:off + es], 'little')); off += es
        n_inputs = buf[off]; off += 1
        shapes = []
        for _ in range(n_inputs):
            rank = buf[off]; off += 1
            shape = []
            for _ in range(rank):
                shape.append(struct.unpack_from('<I', buf, off)[0]); off += 4
            shapes.append(shape)
        return {'evRoot': ev_root, 'evProof': ev_proof,
                'ldProof': {'lcRoot': lc_root, 'lcProof': lc_proof, 'components': comps, 'remainder': remainder},
                'iShapes': shapes}

    def size_of(self, proof: dict) -> int:                                              # sizeof.ts:12-53
        es, ds = self.field.element_size, self.hash.digest_size
        size = ds + size_of_merkle_proof(proof['evProof'])
        ld = proof['ldProof']
        size += 1 + size_of_merkle_proof(ld['lcProof']) + ds
        for c in ld['components']:
            size += ds + size_of_merkle_proof(c['columnProof']) + size_of_merkle_proof(c['polyProof'])
        size += len(ld['remainder']) * es + 1
        size += 1
        for shape in proof['iShapes']:
            size += 1 + 4 * len(shape)
        return size


def validate_assertions(trace, assertions):                                             # Stark.ts:356-376
    registers, steps = len(trace), len(trace[0])
    for a in assertions:
        if a['register'] < 0 or a['register'] >= registers:
            raise ValueError(f"Invalid assertion: register {a['register']} is outside of register bank")
        if a['step'] < 0 or a['step'] >= steps:
            raise ValueError(f"Invalid assertion: step {a['step']} is outside of execution trace")
        if trace[a['register']][a['step']] != a['value']:
            raise StarkError(f"Assertion at step {a['step']}, register {a['register']} conflicts with execution trace")


def size_of_merkle_proof(p: BatchMerkleProof) -> int:                                   # sizeof.ts:55-99
    if len(p.values) == 0:
        raise ValueError('Array cannot be zero-length')
    if len(p.values) > MAX_ARRAY_LENGTH:
        raise ValueError(f'Array length ({len(p.values)}) cannot exceed {MAX_ARRAY_LENGTH}')
    size = 1 + sum(len(v) for v in p.values)
    if len(p.nodes) > MAX_ARRAY_LENGTH:
        raise ValueError(f'Matrix column count ({len(p.nodes)}) cannot exceed {MAX_ARRAY_LENGTH}')
    size += 1 + len(p.nodes)
    for col in p.nodes:
        if len(col) >= MAX_MATRIX_COLUMN_LENGTH:
            raise ValueError(f'Matrix column length ({len(col)}) cannot exceed {MAX_MATRIX_COLUMN_LENGTH}')
        size += sum(len(x) for x in col)
    return size + 1


def write_merkle_proof(p: BatchMerkleProof, leaf_size: int) -> bytes:                   # serialization.ts:18-96
    out = bytearray()
    out.append(0 if len(p.values) == MAX_ARRAY_LENGTH else len(p.values))
    for v in p.values:
        out += v
    out.append(0 if len(p.nodes) == MAX_ARRAY_LENGTH else len(p.nodes))
    for col in p.nodes:
        t = 1 if (len(col) > 0 and len(col[0]) == leaf_size) else 0
        out.append(((len(col) << 1) | t) & 0xFF)
    for col in p.nodes:
        for x in col:
            out += x
    out.append(p.depth)
    return bytes(out)


def read_merkle_proof(buf: bytes, off: int, leaf_size: int, node_size: int):            # serialization.ts:25-124
    n = buf[off] or MAX_ARRAY_LENGTH; off += 1
    values = []
    for _ in range(n):
        values.append(bytes(buf[off:off + leaf_size])); off += leaf_size
    cols = buf[off] or MAX_ARRAY_LENGTH; off += 1
    lens, types = [], []
    for _ in range(cols):
        lt = buf[off]; off += 1
        lens.append(lt >> 1); types.append(lt & 1)
    nodes = []
    for i in range(cols):
        col = []
        for j in range(lens[i]):
            sz = (leaf_size if types[i] == 1 else node_size) if j == 0 else node_size
            col.append(bytes(buf[off:off + sz])); off += sz
        nodes.append(col)
    depth = buf[off]; off += 1
    return BatchMerkleProof(values, nodes, depth), off
