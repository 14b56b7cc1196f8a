"""Timestamps every step of tests/test_commit_gpu.py's check (one-off diagnosis of a slow GPU-box run)."""
import faulthandler
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
faulthandler.dump_traceback_later(int(os.environ.get('DIAG_LIMIT', '50')), exit=True)


def log(*a):
    print(f'{time.time() - t0:8.2f}', *a, flush=True)


from genstark_b200 import build as _b
log('stale?', _b._stale())
import test_commit_gpu as t
from genstark_b200.field import GpuHash, MerkleTree
from util import gpu_field
log('imports')
f = gpu_field()
log('context')
for alg, ncols, log_n in [('blake2s256', 2, 2), ('blake2s256', 2, 18), ('blake2s256', 4, 19), ('blake2s256', 2, 20), ('sha256', 4, 19)]:
    n = 1 << log_n
    raws, vecs = t._columns(f, ncols, n, 5)
    log(alg, ncols, log_n, 'columns up')
    h = GpuHash(alg, f.ctx)
    tr = MerkleTree._commit(vecs, h)
    log('  commit enqueued'); f.ctx.sync(); log('  commit done')
    got = tr._nodes(); log('  nodes read')
    leaves = [t._digest(alg, b''.join(raw[16 * i: 16 * i + 16] for raw in raws)) for i in range(n)]
    want = t._host_tree(alg, leaves); log('  host tree')
    bad = [i for i in range(1, 2 * n) if got[i] != want[i]]
    log('  fused commit mismatches:', len(bad), bad[:4])
    d = h.mergeVectorRows(vecs); f.ctx.sync(); log('  mergeVectorRows')
    ok = d.toBuffers() == leaves; log('  leaves equal:', ok)
    tr2 = MerkleTree.create(d, h); f.ctx.sync(); log('  create')
    got2 = tr2._nodes()
    bad2 = [i for i in range(1, 2 * n) if got2[i] != want[i]]
    log('  create mismatches:', len(bad2), bad2[:4])
    del tr, tr2, d, vecs
    log('  freed')
log('done')
