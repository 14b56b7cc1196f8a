"""ORACLE (test infrastructure) -- hashing and Merkle trees.

Restates the subset of ``@guildofweavers/merkle 0.3.12`` used by genSTARK (call sites: SURVEY §8b;
package pinned at /root/reference/package-lock.json:39-46, NOT vendored).  Anchors in the tree:

  * ``hash.digest`` re-hashes raw leaf bytes in the verifier     lib/utils/index.ts:34-45
  * ``mergeVectorRows`` leaf = concatenation of element i         lib/Stark.ts:113-115,284-296
  * ``digestValues(buf, size)`` one digest per size-byte row      LowDegreeProver.ts:45,163,201
  * BatchMerkleProof shape {values, nodes[][], depth}             lib/utils/serialization.ts:18-35
  * a proof column may start with a leaf                          lib/utils/serialization.ts:80-84

Tree layout and the batch-proof algorithm are the package's published ones [RECALLED, SURVEY App. C]:
nodes[1] = root, nodes[i] = H(nodes[2i] || nodes[2i+1]), leaves at n..2n-1.  PARITY UNPINNED.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence


class Hash:
    def __init__(self, algorithm: str):
        if algorithm == 'sha256':
            self._f = lambda b: hashlib.sha256(b).digest()
        elif algorithm == 'blake2s256':
            self._f = lambda b: hashlib.blake2s(b, digest_size=32).digest()
        else:
            raise TypeError(f'Hash algorithm {algorithm} is not supported')
        self.algorithm = algorithm
        self.digest_size = 32

    def digest(self, value: bytes) -> bytes:
        return self._f(value)

    def merge(self, a: bytes, b: bytes) -> bytes:
        return self._f(a + b)

    def digest_values(self, buf: bytes, value_size: int) -> List[bytes]:
        return [self._f(buf[i:i + value_size]) for i in range(0, len(buf), value_size)]

    def merge_vector_rows(self, vectors: Sequence[Sequence[int]], element_size: int) -> List[bytes]:
        n = len(vectors[0])
        return [self._f(b''.join(int(v[i]).to_bytes(element_size, 'little') for v in vectors))
                for i in range(n)]


class BatchMerkleProof:
    def __init__(self, values: List[bytes], nodes: List[List[bytes]], depth: int):
        self.values, self.nodes, self.depth = values, nodes, depth


def _normalize_indexes(indexes: Sequence[int]) -> List[int]:
    out: Dict[int, None] = {}
    for i in sorted(indexes):
        out[i - (i & 1)] = None
    return list(out)


def _map_indexes(indexes: Sequence[int], max_valid: int) -> Dict[int, int]:
    out: Dict[int, int] = {}
    for pos, i in enumerate(indexes):
        if i < 0 or i > max_valid:
            raise ValueError(f'Invalid index {i}')
        out[i] = pos
    if len(out) != len(indexes):
        raise ValueError('Repeating indexes detected')
    return out


class MerkleTree:
    def __init__(self, nodes: List[bytes], values: List[bytes]):
        self.nodes, self.values = nodes, values

    @staticmethod
    def create(values: List[bytes], hash: Hash) -> 'MerkleTree':
        n = len(values)
        assert n & (n - 1) == 0
        nodes: List[bytes] = [b''] * n + list(values)
        for i in range(n - 1, 0, -1):
            nodes[i] = hash.merge(nodes[2 * i], nodes[2 * i + 1])
        return MerkleTree(nodes, list(values))

    @property
    def root(self) -> bytes:
        return self.nodes[1]

    def prove_batch(self, indexes: Sequence[int]) -> BatchMerkleProof:
        n = len(self.values)
        depth = n.bit_length() - 1
        index_map = _map_indexes(indexes, n - 1)
        norm = _normalize_indexes(indexes)
        values: List[bytes] = [b''] * len(index_map)
        nodes: List[List[bytes]] = [[] for _ in norm]
        next_indexes: List[int] = []
        for i, index in enumerate(norm):
            v1, v2 = self.values[index], self.values[index + 1]
            i1, i2 = index_map.get(index), index_map.get(index + 1)
            if i1 is not None:
                values[i1] = v1
                if i2 is not None:
                    values[i2] = v2
                else:
                    nodes[i] = [v2]
            else:
                values[i2] = v2
                nodes[i] = [v1]
            next_indexes.append((index + n) >> 1)
        for _ in range(depth - 1, 0, -1):
            cur, next_indexes = next_indexes, []
            i = 0
            while i < len(cur):
                sibling = cur[i] ^ 1
                if i + 1 < len(cur) and cur[i + 1] == sibling:
                    i += 1
                else:
                    nodes[i].append(self.nodes[sibling])
                next_indexes.append(sibling >> 1)
                i += 1
        return BatchMerkleProof(values, nodes, depth)

    @staticmethod
    def verify_batch(root: bytes, indexes: Sequence[int], proof: BatchMerkleProof, hash: Hash) -> bool:
        v: Dict[int, bytes] = {}
        offset = 2 ** proof.depth
        index_map = _map_indexes(indexes, offset - 1)
        norm = _normalize_indexes(indexes)
        if len(norm) != len(proof.nodes):
            return False
        next_indexes: List[int] = []
        pointers = [0] * len(norm)
        for i, index in enumerate(norm):
            i1, i2 = index_map.get(index), index_map.get(index + 1)
            try:
                if i1 is not None:
                    if i2 is not None:
                        v1, v2 = proof.values[i1], proof.values[i2]
                        pointers[i] = 0
                    else:
                        v1, v2 = proof.values[i1], proof.nodes[i][0]
                        pointers[i] = 1
                else:
                    v1, v2 = proof.nodes[i][0], proof.values[i2]
                    pointers[i] = 1
            except IndexError:
                return False
            parent = (offset + index) >> 1
            v[parent] = hash.merge(v1, v2)
            next_indexes.append(parent)
        for _ in range(proof.depth - 1, 0, -1):
            cur, next_indexes = next_indexes, []
            i = 0
            while i < len(cur):
                node_index = cur[i]
                sibling_index = node_index ^ 1
                j = i
                if i + 1 < len(cur) and cur[i + 1] == sibling_index:
                    sibling = v[sibling_index]
                    i += 1
                else:
                    ptr = pointers[j]
                    if ptr >= len(proof.nodes[j]):
                        return False
                    sibling = proof.nodes[j][ptr]
                    pointers[j] = ptr + 1
                node = v[node_index]
                parent = hash.merge(sibling, node) if node_index & 1 else hash.merge(node, sibling)
                v[node_index >> 1] = parent
                next_indexes.append(node_index >> 1)
                i += 1
        return v.get(1) == root
