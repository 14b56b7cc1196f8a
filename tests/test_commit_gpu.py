"""The prover's commit (hash.mergeVectorRows + MerkleTree.create, lib/Stark.ts:114-118; digestValues + create,
LowDegreeProver.ts:45-46,163-164,201-202) as it runs on one GPU: leaves hashed inside the tree launches, up to three
levels per launch, one launch for everything from 2^17 nodes down.  EVERY stored node -- not only the root and the
nodes a batch proof happens to touch -- is compared with hashlib, for each launch shape the size thresholds select."""
import hashlib
import random

import pytest

from genstark_b200.field import GpuHash, MerkleTree
from util import gpu_field

pytestmark = pytest.mark.gpu


def _digest(alg, data):
    return hashlib.sha256(data).digest() if alg == 'sha256' else hashlib.blake2s(data, digest_size=32).digest()


def _host_tree(alg, leaves):
    """[unused slot 0, root, ..., leaves] -- the heap layout of merkle's MerkleTree"""
    n = len(leaves)
    nodes = [b''] * n + list(leaves)
    for i in range(n - 1, 0, -1):
        nodes[i] = _digest(alg, nodes[2 * i] + nodes[2 * i + 1])
    return nodes


def _columns(f, ncols, n, seed):
    r = random.Random(seed)
    raws = [r.randbytes(16 * n) for _ in range(ncols)]          # any 16 bytes hash the same way; no need for residues
    return raws, [f._from_bytes(raw, 1, n) for raw in raws]


def _check(alg, ncols, log_n, seed):
    f = gpu_field()
    n = 1 << log_n
    raws, vecs = _columns(f, ncols, n, seed)
    h = GpuHash(alg, f.ctx)
    leaves = [_digest(alg, b''.join(raw[16 * i: 16 * i + 16] for raw in raws)) for i in range(n)]
    want = _host_tree(alg, leaves)
    got = MerkleTree._commit(vecs, h)._nodes()
    assert got[n:] == want[n:], 'leaves'
    for lvl in range(log_n - 1, -1, -1):                          # level by level so a failure names the level
        lo, hi = 1 << lvl, 2 << lvl
        assert got[lo:hi] == want[lo:hi], f'level with {lo} nodes'
    # the same tree from ready-made digests (MerkleTree.create: the launches without the leaf hashing)
    d = h.mergeVectorRows(vecs)
    assert d.toBuffers() == leaves
    got2 = MerkleTree.create(d, h)._nodes()
    assert got2[1:] == want[1:]


# tree top alone (<= 2^17 leaves: one launch, leaves hashed by the block that owns them)
@pytest.mark.parametrize('alg', ['blake2s256', 'sha256'])
@pytest.mark.parametrize('ncols,log_n', [(1, 1), (2, 2), (4, 5), (3, 9), (2, 10), (4, 11), (4, 13), (2, 16), (4, 17), (5, 12), (17, 10)])
def test_small_commits_every_node(alg, ncols, log_n):
    _check(alg, ncols, log_n, 1000 * ncols + log_n)


# spans: 2^18 leaves -> one level per launch, 2^19 -> two, 2^20 -> three (2^21 -> two + two, 2^23 -> three + three: the prove tests); 5 columns: separate leaf kernel
@pytest.mark.parametrize('alg,ncols,log_n', [
    ('sha256', 2, 19), ('sha256', 4, 18), ('sha256', 5, 18),
    ('blake2s256', 2, 18), ('blake2s256', 4, 19), ('blake2s256', 5, 19), ('blake2s256', 2, 20),
])
def test_large_commits_every_node(alg, ncols, log_n):
    _check(alg, ncols, log_n, 77 * log_n + ncols)
