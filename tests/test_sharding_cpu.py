"""Host logic of the multi-GPU (coset-sharded) prover on CPU: world_size-2 (and 4) gloo processes exchange what each
rank owns and check that the all-gather + permutation of a commit reproduces leaf order, and that the bench's
max-over-ranks timing reduction works."""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genstark_b200 import _native


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, log_t, log_e, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    L = _native.lib()
    T, E = 1 << log_t, 1 << log_e
    el = E // world
    n_loc = T * el
    out, own = C.c_int64(), C.c_int()
    # "digests" of the local rows = their global position (what commit_gather moves around)
    local = torch.empty(n_loc, dtype=torch.int64)
    for il in range(n_loc):
        assert L.gs_shard_map(world, rank, log_e, il, 0, C.byref(out), C.byref(own)) == 0
        local[il] = out.value
        # round trip through the global -> (owner, local) direction
        assert L.gs_shard_map(world, rank, log_e, out.value, 1, C.byref(out), C.byref(own)) == 0
        assert own.value == rank and out.value == il
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    # permutation of commit_gather: natural[q*E + r*El + jl] = gathered[r][q*El + jl]
    natural = torch.full((T * E,), -1, dtype=torch.int64)
    for r in range(world):
        g = gathered[r].view(T, el)
        for q in range(T):
            natural[q * E + r * el: q * E + (r + 1) * el] = g[q]
    ok = bool((natural == torch.arange(T * E)).all())
    # next-state access stays local: position i + E of an owned i is the next local row
    for il in range(0, n_loc - el, max(1, n_loc // 37)):
        assert int(local[il + el]) == int(local[il]) + E
    # FRI rows: i and i + N/4 live on the same rank, a quarter of the local vector apart
    quarter = T * E // 4
    for il in range(0, n_loc // 4, max(1, n_loc // 41)):
        assert int(local[il + n_loc // 4]) == int(local[il]) + quarter
    # bench.py's timing reduction: max over ranks
    t = torch.tensor([10.0 + rank, 3.0 - rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok &= (t.tolist() == [10.0 + world - 1, 3.0])
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize('world,log_t,log_e', [(2, 4, 3), (4, 3, 3), (2, 3, 1), (8, 2, 3)])
def test_coset_sharding_map_under_gloo(world, log_t, log_e):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, log_t, log_e, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_map_rejects_bad_arguments():
    L = _native.lib()
    out, own = C.c_int64(), C.c_int()
    assert L.gs_shard_map(3, 0, 3, 0, 0, C.byref(out), C.byref(own)) != 0      # not a power of two
    assert L.gs_shard_map(16, 0, 3, 0, 0, C.byref(out), C.byref(own)) != 0     # more ranks than cosets
    assert L.gs_shard_map(2, 2, 3, 0, 0, C.byref(out), C.byref(own)) != 0
