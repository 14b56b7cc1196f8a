"""Coset-sharded proving on W GPUs of one box: W processes (one per GPU), NCCL all-gather at every Merkle commit.
Checks that every rank returns the proof the oracle computes.  Usage: python scripts/shard_check.py W [log_steps] [E]"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def worker(rank, world, id_path, case_name, args, out_q):
    try:
        import cases
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
        air, opts, a, inputs, seed = getattr(cases, case_name)(*args)
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


def run(world, case_name, args, want=None):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    id_path = f'/tmp/gs_nccl_id_{os.getpid()}_{case_name}_{world}'
    if os.path.exists(id_path):
        os.remove(id_path)
    procs = [ctx.Process(target=worker, args=(r, world, id_path, case_name, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    res.sort(key=lambda x: x[0])
    ok = True
    for rank, proof, stable, ms, dev in res:
        if proof is None:
            print(f'rank {rank} FAILED: {ms}'); ok = False; continue
        same = (proof == res[0][1]) and (want is None or proof == want)
        print(f'rank {rank}: {len(proof)} bytes, stable={stable}, same_as_expected={same}, e2e {ms:.2f} ms, device {dev:.2f} ms')
        ok &= stable and same
    return ok, res[0][1]


if __name__ == '__main__':
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    log_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    e = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    import cases
    from oracle import cport
    air, opts, a, inputs, seed = cases.mimc(1 << log_steps, e)
    want = cport.prove(air, opts, a, inputs, seed)
    ok, _ = run(world, 'mimc', (1 << log_steps, e), want)
    if ok and log_steps <= 13:
        air, opts, a, inputs, seed = cases.poseidon(2, 1, 16)
        want = cport.prove(air, opts, a, inputs, seed)
        ok2, _ = run(world, 'poseidon', (2, 1, 16), want)
        ok &= ok2
    print('SHARD_CHECK', 'OK' if ok else 'FAILED')
    sys.exit(0 if ok else 1)
