"""Host-side methods of the FiniteField / Hash seam that the library answers without a device (SURVEY.md §8b:
prng, interpolate, evalPolyAt, mulPolys, hash.digest, MerkleTree.verifyBatch) against the oracle -- CPU tests,
straight through the C ABI."""
import ctypes as C
import hashlib
import random

import pytest

from genstark_b200 import _native
from genstark_b200.air import P128
from genstark_b200.field import seed_bytes
from oracle.field import PrimeField
from oracle.merkle import Hash as OHash, MerkleTree as OTree
from util import rand_elems

OF = PrimeField(P128)
enc = lambda v: int(v).to_bytes(16, 'little')
dec = lambda raw: [int.from_bytes(raw[i:i + 16], 'little') for i in range(0, len(raw), 16)]


def test_prng_matches_the_oracle_for_buffers_and_bigints():
    L = _native.lib()
    for seed in (b'\x01' * 32, hashlib.sha256(b'root').digest(), 42, 0xabc, 2**200 + 12345):
        sb = seed_bytes(seed)
        out = C.create_string_buffer(16)
        assert L.gs_field_prng(sb, len(sb), 0, out) == 0
        assert dec(out.raw)[0] == OF.prng(seed)
        for n in (1, 7, 300):
            out = C.create_string_buffer(16 * n)
            assert L.gs_field_prng(sb, len(sb), n, out) == 0
            assert dec(out.raw) == OF.prng(seed, n)


@pytest.mark.parametrize('n', [1, 2, 3, 4, 17, 96])
def test_interpolate_eval_mul_polys(n):
    L = _native.lib()
    xs = rand_elems(n, 3 * n)
    xs = list(dict.fromkeys(xs))
    while len(xs) < n:
        xs.append((xs[-1] + 1) % P128)
    ys = rand_elems(n, 5 * n + 1)
    out = C.create_string_buffer(16 * n)
    assert L.gs_poly_interpolate(b''.join(map(enc, xs)), b''.join(map(enc, ys)), n, out) == 0
    poly = dec(out.raw)
    assert poly == OF.interpolate(xs, ys)
    o1 = C.create_string_buffer(16)
    for x, y in list(zip(xs, ys))[:5] + [(12345, None)]:
        assert L.gs_poly_eval_at(out.raw, n, enc(x), o1) == 0
        assert dec(o1.raw)[0] == (OF.eval_poly_at(poly, x) if y is None else y)
    b = rand_elems(3, 77)
    o2 = C.create_string_buffer(16 * (n + 2))
    assert L.gs_poly_mul(out.raw, n, b''.join(map(enc, b)), 3, o2) == 0
    assert dec(o2.raw) == OF.mul_polys(poly, b)


def test_interpolate_with_a_repeated_x_follows_the_zero_inverse_rule():
    L = _native.lib()
    xs, ys = [5, 9, 5, 11], [1, 2, 3, 4]
    out = C.create_string_buffer(64)
    assert L.gs_poly_interpolate(b''.join(map(enc, xs)), b''.join(map(enc, ys)), 4, out) == 0
    assert dec(out.raw) == OF.interpolate(xs, ys)


@pytest.mark.parametrize('alg', ['sha256', 'blake2s256'])
def test_digest_and_verify_batch(alg):
    L = _native.lib()
    a = ['sha256', 'blake2s256'].index(alg)
    oh = OHash(alg)
    r = random.Random(8)
    out = C.create_string_buffer(32)
    for n in (0, 1, 55, 64, 65, 200):
        msg = r.randbytes(n)
        assert L.gs_hash_digest(a, msg, n, out) == 0
        assert out.raw == oh.digest(msg)
    leaves = [oh.digest(r.randbytes(16)) for _ in range(64)]
    tree = OTree.create(leaves, oh)
    for idx in ([3], [0, 1], [5, 4, 63, 17, 16], list(range(0, 64, 7))):
        proof = tree.prove_batch(idx)
        blob = bytearray(len(proof.values).to_bytes(4, 'little') + len(proof.nodes).to_bytes(4, 'little') + proof.depth.to_bytes(4, 'little'))
        for v in proof.values:
            blob += v
        for col in proof.nodes:
            blob += len(col).to_bytes(4, 'little') + b''.join(col)
        arr = (C.c_uint32 * len(idx))(*idx)
        assert L.gs_merkle_verify_batch(a, tree.root, arr, len(idx), bytes(blob), len(blob)) == 1
        assert OTree.verify_batch(tree.root, idx, proof, oh)
        bad = bytearray(blob); bad[12 + 5] ^= 1
        assert L.gs_merkle_verify_batch(a, tree.root, arr, len(idx), bytes(bad), len(bad)) == 0
        wrong_root = bytes(32)
        assert L.gs_merkle_verify_batch(a, wrong_root, arr, len(idx), bytes(blob), len(blob)) == 0
    assert L.gs_merkle_verify_batch(a, tree.root, arr, len(idx), bytes(blob[:20]), 20) < 0


def test_verify_batch_counts_in_the_blob_never_size_an_allocation():
    """column / node counts are attacker-controlled 32-bit words: 2^32 - 1 columns used to end the process in std::bad_alloc
    (found by fuzzing the blob); every count the remaining bytes cannot hold is a malformed proof (negative status), never a crash"""
    L = _native.lib()
    oh = OHash('blake2s256')
    r = random.Random(9)
    leaves = [oh.digest(r.randbytes(16)) for _ in range(32)]
    tree = OTree.create(leaves, oh)
    idx = [3, 9, 20]
    proof = tree.prove_batch(idx)
    blob = bytearray(len(proof.values).to_bytes(4, 'little') + len(proof.nodes).to_bytes(4, 'little') + proof.depth.to_bytes(4, 'little'))
    for v in proof.values:
        blob += v
    first_len = len(blob)                             # offset of the first column's length word
    for col in proof.nodes:
        blob += len(col).to_bytes(4, 'little') + b''.join(col)
    arr = (C.c_uint32 * len(idx))(*idx)
    assert L.gs_merkle_verify_batch(1, tree.root, arr, len(idx), bytes(blob), len(blob)) == 1
    for off in (4, first_len):                        # the column count, a column's node count
        for v in (0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 0x10000000, len(blob)):
            t = bytearray(blob); t[off:off + 4] = v.to_bytes(4, 'little')
            assert L.gs_merkle_verify_batch(1, tree.root, arr, len(idx), bytes(t), len(t)) <= 0
    t = bytearray(blob); t[8:12] = (200).to_bytes(4, 'little')        # depth beyond 32
    assert L.gs_merkle_verify_batch(1, tree.root, arr, len(idx), bytes(t), len(t)) == 0
