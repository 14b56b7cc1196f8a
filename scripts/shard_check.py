"""Coset-sharded proving on W GPUs of one box: W processes (one per GPU), NCCL exchange at every Merkle commit.
Checks that every rank returns the proof bytes the C oracle computes for the same workload.

Usage: python scripts/shard_check.py W [config ...]
  config = a key of genstark_b200.workloads.CONFIGS ('ns', '2', '3', '4', '5') or 'small' (MiMC 2^13 / E=8 and a
  depth-2 Poseidon branch, the quick shapes).  Default: small."""
import hashlib
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_case(name):
    from genstark_b200 import workloads
    if name == 'small-mimc':
        return workloads.mimc(1 << 13, 8)
    if name == 'small-poseidon':
        return workloads.poseidon(2, 1, 16)
    return workloads.config(name)[:5]


def worker(rank, world, id_path, case_name, out_q):
    try:
        from genstark_b200.field import Context
        from genstark_b200.stark import Stark
        ctx = Context(rank)
        if rank == 0:
            uid = Context.comm_unique_id()
            with open(id_path + '.tmp', 'wb') as f:
                f.write(uid)
            os.rename(id_path + '.tmp', id_path)
        else:
            while not os.path.exists(id_path):
                time.sleep(0.01)
            uid = open(id_path, 'rb').read()
        ctx.comm_init(rank, world, uid)
        air, opts, a, inputs, seed = build_case(case_name)
        st = Stark(air, opts, context=ctx)
        proofs = [st.prove_bytes(a, inputs, seed) for _ in range(3)]
        t = time.perf_counter()
        for _ in range(3):
            st.prove_bytes(a, inputs, seed)
        ms = (time.perf_counter() - t) / 3 * 1e3
        out_q.put((rank, proofs[0], proofs[0] == proofs[1] == proofs[2], ms, st.last_timing()[0]))
    except Exception as e:      # pragma: no cover
        import traceback
        out_q.put((rank, None, False, repr(e) + traceback.format_exc(), 0))


def run(world, case_name, want=None):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    id_path = f'/tmp/gs_nccl_id_{os.getpid()}_{case_name}_{world}'
    if os.path.exists(id_path):
        os.remove(id_path)
    procs = [ctx.Process(target=worker, args=(r, world, id_path, case_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    res.sort(key=lambda x: x[0])
    ok = True
    for rank, proof, stable, ms, dev in res:
        if proof is None:
            print(f'rank {rank} FAILED: {ms}'); ok = False; continue
        same = (proof == res[0][1]) and (want is None or proof == want)
        print(f'[{case_name} W={world}] rank {rank}: {len(proof)} bytes, sha256 {hashlib.sha256(proof).hexdigest()[:16]}, stable={stable}, '
              f'equals_oracle={same}, e2e {ms:.2f} ms, device {dev:.2f} ms', flush=True)
        ok &= stable and same
    return ok, res[0][1]


def main(argv):
    world = int(argv[1]) if len(argv) > 1 else 2
    names = argv[2:] or ['small']
    cases = []
    for n in names:
        cases += ['small-mimc', 'small-poseidon'] if n == 'small' else [n]
    from oracle import cport
    ok = True
    for name in cases:
        air, opts, a, inputs, seed = build_case(name)
        t = time.perf_counter()
        want = cport.prove(air, opts, a, inputs, seed)
        print(f'[{name}] C oracle proof: {len(want)} bytes, sha256 {hashlib.sha256(want).hexdigest()[:16]} ({time.perf_counter() - t:.1f} s)', flush=True)
        ok &= run(world, name, want)[0]
    print('SHARD_CHECK', 'OK' if ok else 'FAILED')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main(sys.argv))
