#!/usr/bin/env python
"""bench.py -- Stark.prove() on B200 (BASELINE.json metric), one JSON line on rank 0.

A "step" is one Stark.prove() of the selected workload.  `--config` picks it (default `ns`, the configuration
BASELINE.json's metric is quoted on: MiMC over p128, 2^20 steps, extension factor 8, blake2s256, 48 trace / 24 FRI
queries -- examples/mimc/mimc128.ts:22-28 with the E of the north-star target); `2`, `3`, `4`, `5` are BASELINE.json's
configs[1..4] (genstark_b200/workloads.py).
  value        ms per prove with the execution trace already resident in HBM (CUDA events on the prover stream)
  e2e.value    ms per prove through the public API from host inputs to host-resident proof bytes
               (host trace generation + H2D of the trace + every D2H inside the timed region) -- BASELINE.md section 4's
               definition of prove() time, and the headline
  parity_*     SHA-256 of every rank's proof bytes against the C oracle's proof for the same workload (computed once,
               outside every timed region)
  roofline     dominant kernel class of the step: algorithmic bytes / CUDA-event time vs measured HBM peak
  ntt          K1 alone (second half of BASELINE.json's metric): rows in {1,4,12}, 2^13..2^24 points, forward / inverse /
               LDE, median of 20 runs; at N > 1 every rank's share of the coset-sharded LDE
  cpu_baseline the oracle port of the same path on the host cores (one prove), N=1 rank 0 only
`--impl reference` times that CPU port alone (the reference itself needs node + npm packages that this image does not
have: SURVEY.md section 8c); its `value` leaves out the trace-generation stage like ours, its `e2e` is the whole prove().
One proof is sharded over the N ranks by cosets, so "scaling" is "strong" at every N.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRICS = {'ns': 'mimc128_prove_ms', '2': 'mimc128_prove_ms', '4': 'mimc128_prove_ms', '3': 'rescue4x128_prove_ms',
           '5': 'poseidon_merkle_prove_ms', 'test': 'mimc128_prove_ms'}
DTYPE = 'u128 (integer mod p = 2^128 - 9*2^32 + 1, 4x u32 limbs)'


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)), 'measured'
        except Exception:
            pass
    return {'hbm_gbs': 6650.0}, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        for line in self.proc.stdout:
            if self.stop_flag:
                break
            line = line.strip()
            if line:
                self.samples.append([time.perf_counter()] + [x.strip() for x in line.split(',')])

    def stop(self):
        self.stop_flag = True
        try:
            self.proc.terminate()
        except Exception:
            pass
        self.join(timeout=2)

    def summary(self, t0=None, t1=None):
        """median SM clock and throttle reasons of the samples taken inside [t0, t1] (the timed region); nvidia-smi
        needs about a second to start on an 8-GPU box, so the sampler is started before the warm-up"""
        rows = [s for s in self.samples if t0 is None or t0 <= s[0] <= t1]
        window = 'timed region'
        if not rows:
            rows, window = self.samples, 'whole run (no sample fell inside the timed region)'
        sm, mx, reasons = [], 0, set()
        for s in rows:
            try:
                sm.append(float(s[1])); mx = max(mx, float(s[2]))
            except Exception:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], s[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm), 'window': window}


def load_workload(name):
    """(air, options, assertions, inputs, seed, description) -- built outside every timed region"""
    from genstark_b200 import workloads
    return workloads.config(name)


def host_threads():
    # physical cores: hyper-threads do not help the oracle's memory-bound loops (measured: 128 threads slower than 64)
    ncpu = os.cpu_count() or 1
    return ncpu // 2 if ncpu > 16 else ncpu


# ------------------------------------------------------------------------------------------ CPU port
def oracle_prove(workload, threads=None):
    """one prove() of the workload by the C oracle port: (proof bytes, total ms, stage list, threads)"""
    from oracle import cport
    air, opts, a, inputs, seed, _ = workload
    L = cport.lib()
    L.oracle_set_threads(int(threads or host_threads()))
    n = L.oracle_threads()
    stages = []
    t = time.perf_counter()
    proof = cport.prove(air, opts, a, inputs, seed, stages=stages)
    return proof, (time.perf_counter() - t) * 1e3, stages, n


def run_reference(args, rank, world):
    if rank != 0:
        return
    workload = load_workload(args.config)
    desc = workload[5]
    tot, res, n = [], [], 1
    for i in range(args.warmup + args.steps):
        proof, ms, stages, n = oracle_prove(workload)
        if i >= args.warmup:
            tot.append(ms)
            res.append(ms - dict(stages).get('trace', 0.0))
    e2e = sum(tot) / len(tot)
    v = sum(res) / len(res)
    sample = f'{args.steps} whole prove() of the workload by the C oracle port (oracle/c/stark_oracle.c, OpenMP, {n} threads)'
    line = {'impl': 'reference', 'metric': METRICS[args.config], 'value': v, 'unit': 'ms', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': v, 'higher_is_better': False, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': DTYPE, 'data': 'synthetic',
            'config': {'workload': desc, 'value_excludes': 'execution-trace generation (as on the GPU arm, whose `value` starts from a resident trace)'},
            'cpu_baseline': {'value': v, 'unit': 'ms', 'cores': n, 'kind': 'port', 'sample': sample},
            'e2e': {'value': e2e, 'unit': 'ms', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'stages_ms': [[k, round(t, 3)] for k, t in stages],
            'proof_sha256': hashlib.sha256(proof).hexdigest(),
            'note': 'the reference (genSTARK + galois/merkle/air-assembly WASM) cannot run here: no node in the image'}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def fri_layers(n):
    out = []
    while True:
        out.append(n)
        if n <= 256:
            return out
        n >>= 2


def algorithmic_bytes(cls, log_t, log_e, r, s, n_boundary, fused=False):
    """SURVEY.md section 8d / App. A.9 per-unit figures x units of ONE prove, per kernel class.
    fused: the single-GPU commit (class merkle_commit) hashes leaves of at most four columns inside the tree launches --
    every FRI layer, and the evaluation tree when r + s <= 4.  Its figure is section 8d's for the work those launches do
    (leaf hash N(16(R+S)+32) + Merkle build 64 B per node); 'merkle_commit_floor' is what a fused commit has to move at least
    (columns read once, every digest written once) and is reported beside it."""
    T, N = 1 << log_t, 1 << (log_t + log_e)
    B, D = 16, 32
    fri = fri_layers(N)
    ev_fused = fused and (r + s) <= 4
    if cls.startswith('ntt'):
        # iNTT(T) + LDE(T -> N) for R (+S) rows: 32 B/point of the transform actually computed
        return (r + s) * (2 * B * T + B * (T + N))
    if cls == 'hash_columns':
        return (0 if ev_fused else N * ((r + s) * B + D)) + (0 if fused else sum((l // 4) * (4 * B + D) for l in fri))
    if cls == 'merkle_build':                                      # 64 B per tree node: read two digests, write one
        return (0 if ev_fused else 2 * N * D) + (0 if fused else sum(2 * (l // 4) * D for l in fri))
    if cls == 'merkle_commit':
        return ((N * ((r + s) * B + D) + 2 * N * D) if ev_fused else 0) + sum((l // 4) * (4 * B + D) + 2 * (l // 4) * D for l in fri)
    if cls == 'merkle_commit_floor':
        return (N * ((r + s) * B + 2 * D) if ev_fused else 0) + sum((l // 4) * (4 * B + 2 * D) for l in fri)
    if cls == 'compose':
        return N * B * (r + s + n_boundary + 1)
    if cls in ('batch_inverse', 'zb_eval'):
        return n_boundary * N * 2 * B
    if cls == 'fri_fold':
        return sum(B * l + B * l // 4 for l in fri[:-1])
    if cls == 'fri_tail':
        return sum(B * l + (l // 4) * 3 * D for l in fri if l <= (1 << 17))
    return 0


def ntt_table(ctx, L, hbm, rank, world, full=True):
    """K1 alone: rows x points of uniform residues (device-side SplitMix64, seed 0xB200), median of 20 runs
    after 3 warm-ups, CUDA events on the library's stream, data resident in HBM.  Rows: [kind, rows, log2 t, log2 n, ms,
    G elements/s, fraction of the HBM roof at 16(t+n) algorithmic bytes per row]."""
    rows_out = []

    def alloc(r, n):
        h = C.c_void_p()
        ctx.check(L.gs_mat_alloc(ctx.handle, r, n, C.byref(h)))
        return h

    def timed(fn, reps=20, warm=3):
        ms, samples = C.c_float(), []
        for i in range(warm + reps):
            L.gs_timer_begin(ctx.handle)
            ctx.check(fn())
            L.gs_timer_end(ctx.handle, C.byref(ms))
            if i >= warm:
                samples.append(ms.value)
        samples.sort()
        return samples[len(samples) // 2]

    sizes = list(range(13, 25)) if full else [20, 23]
    for r in (1, 4, 12):
        for log_n in sizes:
            n = 1 << log_n
            kinds = [('fwd', log_n, 0), ('inv', log_n, 1), ('lde8', log_n - 3, 0)]
            if r == 1:
                kinds += [('lde16', log_n - 4, 0), ('lde32', log_n - 5, 0)]
            dst, work = alloc(r, n), alloc(r, n)
            for kind, log_t, inv in kinds:
                src = alloc(r, 1 << log_t)
                ctx.check(L.gs_mat_fill_random(ctx.handle, src, 0xB200))
                med = timed(lambda: L.gs_ntt_into(ctx.handle, src, dst, work, inv))
                alg = 16 * ((1 << log_t) + n) * r
                rows_out.append([kind, r, log_t, log_n, round(med, 5), round(r * n / (med * 1e-3) / 1e9, 3),
                                 round(alg / (med * 1e-3) / 1e9 / hbm, 4)])
                L.gs_mat_free(src)
            L.gs_mat_free(dst); L.gs_mat_free(work)
    return rows_out


def sharded_lde(ctx, L, rank, world, log_t, log_e):
    """every rank's share of the coset-sharded LDE of the benched shape (2^log_e / world cosets per rank): ms on this rank"""
    per = (1 << log_e) // world
    n_loc = (1 << log_t) * per
    src, dst, work = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.check(L.gs_mat_alloc(ctx.handle, 1, 1 << log_t, C.byref(src)))
    ctx.check(L.gs_mat_alloc(ctx.handle, 1, n_loc, C.byref(dst)))
    ctx.check(L.gs_mat_alloc(ctx.handle, 1, n_loc, C.byref(work)))
    ctx.check(L.gs_mat_fill_random(ctx.handle, src, 0xB200))
    ms, samples = C.c_float(), []
    for i in range(23):
        L.gs_timer_begin(ctx.handle)
        ctx.check(L.gs_lde_cosets_into(ctx.handle, src, dst, work, rank * per, log_e))
        L.gs_timer_end(ctx.handle, C.byref(ms))
        if i >= 3:
            samples.append(ms.value)
    for h in (src, dst, work):
        L.gs_mat_free(h)
    samples.sort()
    return samples[len(samples) // 2]


def run_ours(args, rank, local_rank, world):
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from genstark_b200.field import Context
    from genstark_b200.stark import Stark
    from genstark_b200 import _native
    L = _native.lib()

    workload = load_workload(args.config)
    air, opts, assertions, inputs, seed, desc = workload
    air = air.with_options(opts.get('extensionFactor'))
    steps, ext = air.trace_length, air.extension_factor
    log_t, log_e = steps.bit_length() - 1, ext.bit_length() - 1
    ctx = Context(local_rank)
    if world > 1:
        # one proof sharded over the ranks by cosets (strong scaling): NCCL exchange of digests at every Merkle commit
        if ext % world:
            raise SystemExit(f'extension factor {ext} has fewer cosets than ranks ({world})')
        uid = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    st = Stark(air, dict(opts), context=ctx)

    # ---- parity: the C oracle's proof of the same workload (rank 0, outside every timed region)
    want_sha, cpu, oracle_note = None, None, None
    if rank == 0:
        try:
            want, ms_cpu, stages, cores = oracle_prove(workload)
            want_sha = hashlib.sha256(want).hexdigest()
            if world == 1:
                cpu = {'value': ms_cpu, 'unit': 'ms', 'cores': cores, 'kind': 'port',
                       'sample': f'one whole prove() of the workload (trace generation included) by the C oracle port, {cores} OpenMP threads',
                       'stages_ms': [[k, round(t, 3)] for k, t in stages]}
        except Exception as e:        # pragma: no cover
            oracle_note = f'C oracle unavailable: {e!r}'

    sampler = ClockSampler(local_rank)
    sampler.start()

    # warm-up (also allocates every buffer)
    for _ in range(max(args.warmup, 3)):
        proof = st.prove_bytes(assertions, inputs, seed)
    proof_len = len(proof)
    from genstark_b200.stark import trace_backend as _tb
    trace_backend = _tb()

    # ---- e2e leg: public API, host inputs -> host proof bytes
    barrier()
    launches0 = ctx.launch_count
    t0 = time.perf_counter()
    t_region0 = t0
    e2e_dev, hashes = [], set()
    for _ in range(args.steps):
        pb = st.prove_bytes(assertions, inputs, seed)
        e2e_dev.append(st.last_timing())
        hashes.add(pb)            # a set of bytes objects: comparing is outside the cost of a prove, hashing comes after the region
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches_per_step = (ctx.launch_count - launches0) // args.steps
    stage_times = st.stage_times()

    # ---- resident leg: trace already in HBM; CUDA events on the prover stream around the whole device part
    hashes.add(st.prove_bytes(assertions, inputs, seed, _reuse_resident_trace=True))
    barrier()
    dev_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hashes.add(st.prove_bytes(assertions, inputs, seed, _reuse_resident_trace=True))
        dev_ms.append(st.last_timing()[0])
    barrier()
    wall_resident = (time.perf_counter() - t0) * 1e3 / args.steps
    # ---- same K steps again with CUDA events around every kernel class (CUDA graphs off while profiling, so
    # this leg is a little slower than the one above; it provides the per-kernel times of the roofline)
    L.gs_ctx_profile(ctx.handle, 1)
    prof_dev = []
    for _ in range(args.steps):
        hashes.add(st.prove_bytes(assertions, inputs, seed, _reuse_resident_trace=True))
        prof_dev.append(st.last_timing()[0])
    barrier()
    prof = json.loads(L.gs_ctx_profile_report(ctx.handle).decode())
    L.gs_ctx_profile(ctx.handle, 0)
    t_region1 = time.perf_counter()
    sampler.stop()
    clocks = sampler.summary(t_region0, t_region1)

    ms_step = sum(dev_ms) / len(dev_ms)
    my_sha = sorted(hashlib.sha256(h).hexdigest() for h in hashes)        # every proof of every leg: one value expected
    all_sha = [my_sha]
    if dist is not None:
        t = torch.tensor([ms_step, e2e_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])
        all_sha = [None] * world
        dist.all_gather_object(all_sha, my_sha)

    peaks, peak_kind = measured_peaks()
    hbm = float(peaks.get('hbm_gbs', 6650.0))

    # ---- K1 alone: NTT throughput (second half of BASELINE.json's metric)
    ntt = {}
    if world > 1:
        lde_ms = sharded_lde(ctx, L, rank, world, log_t, log_e)
        t = torch.tensor([lde_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_out = steps * ext
        ntt['sharded_lde'] = {'shape': f'2^{log_t} -> 2^{log_t + log_e}, {ext // world} of {ext} cosets per rank', 'ms_max_over_ranks': round(float(t[0]), 5),
                              'elements_per_s': n_out / (float(t[0]) * 1e-3), 'this_rank_ms': round(lde_ms, 5)}
    if rank == 0:
        table = ntt_table(ctx, L, hbm, rank, world, full=(not args.quick_ntt))
        ntt['columns'] = ['kind', 'rows', 'log2_t', 'log2_n', 'ms_median_of_20', 'G_elements_per_s', 'hbm_frac_at_16(t+n)_bytes_per_row']
        ntt['table'] = table
        for row in table:           # the three figures earlier rounds quoted
            if row[1] == 1 and (row[0], row[3]) in (('fwd', 23), ('lde8', 23), ('inv', 20)):
                ntt[f'{row[0]}_2^{row[3]}'] = {'ms': row[4], 'elements_per_s': row[5] * 1e9}

    # ---- throughput mode (N > 1): every rank proves its OWN proof of the workload on its own GPU, all at once -- the
    # configuration in which more GPUs buy end-to-end throughput (host trace generation runs in N processes in parallel);
    # one sharded proof cannot: its 10 ms sequential trace is replicated on every rank
    replicas = None
    if world > 1:
        ctx2 = Context(local_rank)
        st2 = Stark(air, dict(opts), context=ctx2)
        for _ in range(3):
            pr = st2.prove_bytes(assertions, inputs, seed)
        ok2 = hashlib.sha256(pr).hexdigest()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st2.prove_bytes(assertions, inputs, seed)
        torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        t = torch.tensor([mine], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shas = [None] * world
        dist.all_gather_object(shas, ok2)
        replicas = {'proofs': world * args.steps, 'seconds_max_over_ranks': float(t[0]), 'proofs_per_s': world * args.steps / float(t[0]),
                    'ms_per_proof_aggregate': float(t[0]) * 1e3 / (world * args.steps), 'e2e_ms_per_proof_on_one_gpu': mine * 1e3 / args.steps,
                    'rank_sha256': shas,
                    'what': 'N independent end-to-end proves at once, one per GPU (host inputs -> host proof bytes, trace generation included)'}
        st2.close()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    parity_ok = None
    if want_sha is not None:
        parity_ok = all(s == [want_sha] for s in all_sha) and (replicas is None or all(x == want_sha for x in replicas['rank_sha256']))

    # dominant kernel class of the resident step
    per_step = {k: v['ms'] / args.steps for k, v in prof.items()}
    grouped = {}
    for k, v in per_step.items():
        key = 'ntt' if k.startswith('ntt') else k
        grouped[key] = grouped.get(key, 0.0) + v
    compute = {k: v for k, v in grouped.items() if not k.startswith('nccl')}
    dom = max(compute, key=compute.get)
    n_reg, n_sec = air.trace_register_count, air.secret_input_count
    n_boundary = len({int(a['register']) for a in assertions})
    # sharded runs: kernel times are rank 0's, which holds 1/world of the evaluation domain
    fused = 'merkle_commit' in grouped
    alg = algorithmic_bytes(dom, log_t, log_e, n_reg, n_sec, n_boundary, fused) // world
    achieved = alg / (grouped[dom] * 1e-3) / 1e9
    # DRAM traffic of that kernel class: ncu counters of this code on this workload (profiles/ncu_summary.json records the
    # commit and the workload of the capture); single GPU only -- no counters were taken on the sharded path
    traffic, traffic_src = None, None
    summ = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
    if world == 1 and os.path.exists(summ):
        try:
            sj = json.load(open(summ))
            ent = sj.get('configs', {}).get(args.config, {})
            if dom in ent.get('classes', {}):
                traffic = ent['classes'][dom].get('dram_bytes_per_step')
                traffic_src = f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum over the launches of one prove; capture at commit {ent.get('commit')}"
        except Exception:
            traffic = None
    if traffic is None and world == 1 and dom == 'merkle_commit':
        traffic_src = ('null: the fused commit launches (merkle_span_kernel, leaf-hashing merkle_top_kernel) were never under ncu -- the '
                       'GPU budget of the round ended with their parity and A/B runs; the separate kernels they replace moved 583 MB '
                       '(hash_columns) + 859 MB (merkle_build) per prove (profiles/ncu_summary.json, commit 2e4d98f), and the fusion '
                       'removes the re-read of every level that a span keeps in shared memory')
    roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': f'{peak_kind} (MEASURED_PEAKS.json hbm_gbs)',
                'kernel_ms_per_step': grouped[dom], 'algorithmic_bytes_per_step': alg,
                **({'fused_floor_bytes_per_step': algorithmic_bytes('merkle_commit_floor', log_t, log_e, n_reg, n_sec, n_boundary, fused) // world,
                    'algorithmic_bytes_how': 'SURVEY 8d figures for the work these launches do: leaf hash N(16(R+S)+32) + 64 B per tree node, over the '
                                             'evaluation tree and every FRI layer; fused_floor = columns read once + every digest written once'}
                   if dom == 'merkle_commit' else {}),
                'share_of_step': grouped[dom] / (sum(prof_dev) / len(prof_dev)),
                'note': '128-bit modular arithmetic on 32-bit integer pipes: every kernel here is issue-bound, not HBM-bound'}

    # The roof that actually binds these kernels is integer issue, not HBM (profiles/): say so with numbers.
    #  * blake2s kernels: 648 ALU-pipe instructions per compression (320 XOR + 320 rotate + address/feed-forward), the ALU
    #    pipe retires one warp instruction every 2 cycles per SM sub-partition => compressions/s <= SMs*4*32*f / (2*648)
    #  * modmul kernels: gs_debug_modmul_probe measures the chip's dependent-free modular-multiplication rate
    sm_hz = (clocks.get('sm_mhz') or 1965.0) * 1e6
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_roofline = None
    if world == 1:
        comp_peak = sm_count * 4 * 32 * sm_hz / (2 * 648.0)
        n_eval = steps * ext
        fri_rows = sum(l // 4 for l in fri_layers(n_eval))
        leaf_blocks = -(-((n_reg + n_sec) * 16) // 64)              # 64-byte blake2s blocks per leaf
        ev_fused = fused and (n_reg + n_sec) <= 4
        comp_hash_cols = (0 if ev_fused else n_eval * leaf_blocks) + (0 if fused else fri_rows)
        comp_merkle = (0 if ev_fused else n_eval - 1) + (0 if fused else fri_rows)
        comp_commit = (n_eval * leaf_blocks + n_eval - 1 if ev_fused else 0) + 2 * fri_rows
        hash_ms = grouped.get('hash_columns', 0) + (0 if fused else grouped.get('fri_tail', 0))
        issue_roofline = {
            'merkle_commit': {'unit': 'blake2s compressions/s', 'achieved': comp_commit / (grouped['merkle_commit'] * 1e-3) if grouped.get('merkle_commit') else None,
                              'peak': comp_peak, 'peak_how': 'leaf hashing and tree levels in the same launches; same ALU-pipe bound; the top 17 levels of every tree are latency-bound'},
            'hash_columns': {'unit': 'blake2s compressions/s', 'achieved': comp_hash_cols / (hash_ms * 1e-3) if hash_ms else None,
                             'peak': comp_peak, 'peak_how': 'ALU pipe: SMs*4*32 lanes*f / (2 cycles * 648 ALU instructions per compression)'},
            'merkle_build': {'unit': 'blake2s compressions/s', 'achieved': comp_merkle / (grouped.get('merkle_build', 0) * 1e-3) if grouped.get('merkle_build') else None,
                             'peak': comp_peak, 'peak_how': 'same ALU-pipe bound; the top levels of every tree are latency-bound'},
        }
        for v in issue_roofline.values():
            v['frac'] = (v['achieved'] / v['peak']) if v['achieved'] else None
        probe_ms = C.c_float()
        if L.gs_debug_modmul_probe(ctx.handle, sm_count * 8, 2000, C.byref(probe_ms)) == 0 and probe_ms.value > 0:
            modmul_peak = sm_count * 8 * 256 * 4 * 2000.0 / (probe_ms.value * 1e-3)
            issue_roofline['modmul_probe'] = {'unit': 'modmul/s', 'peak': modmul_peak, 'probe_ms': probe_ms.value,
                                              'how': 'gs_debug_modmul_probe: 8 CTAs x 256 threads per SM, 4 independent chains each'}
            # K1 against the rate of its own instruction mix: a butterfly = modular add + modular sub + modular multiplication
            # (gs_debug_butterfly_probe).  Butterflies of one prove: iNTT of T points + E coset transforms of T points, (n/2) log2 n each
            bf_ms = C.c_float()
            if L.gs_debug_butterfly_probe(ctx.handle, sm_count * 8, 2000, C.byref(bf_ms)) == 0 and bf_ms.value > 0 and grouped.get('ntt'):
                bf_peak = sm_count * 8 * 256 * 2 * 2000.0 / (bf_ms.value * 1e-3)
                rows = n_reg + sum(1 for s_ in air.static_registers if s_.kind == 'input')
                n_bf = rows * (1 + ext) * (steps // 2) * log_t
                a = n_bf / (grouped['ntt'] * 1e-3)
                issue_roofline['ntt'] = {'unit': 'butterflies/s', 'achieved': a, 'peak': bf_peak, 'frac': a / bf_peak, 'probe_ms': bf_ms.value,
                                         'butterflies_per_prove': n_bf}

    trace_bytes = n_reg * steps * 16
    line = {
        'metric': METRICS[args.config], 'value': ms_step, 'unit': 'ms', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_step, 'higher_is_better': False, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': DTYPE, 'data': 'synthetic',
        'config': {'workload': desc,
                   'value_excludes': 'execution-trace generation (the trace is resident in HBM when the timed region starts; e2e includes it)',
                   'parallelism': 'single GPU' if world == 1 else f'one proof sharded over {world} GPUs by cosets ({ext // world} of {ext} cosets per rank); digest exchange at each Merkle commit, one all-reduce of queried rows; trace generation replicated on every host process',
                   'l2': 'working set per prove exceeds the 126 MB L2 for the 2^20-step shapes; no explicit flush',
                   'proof_bytes': proof_len},
        'e2e': {'value': e2e_ms, 'unit': 'ms', 'h2d_bytes_per_step': trace_bytes + 4096, 'd2h_bytes_per_step': proof_len + 32 * 12,
                'device_ms_inside': sum(d for d, _ in e2e_dev) / len(e2e_dev), 'host_ms_inside': sum(h for _, h in e2e_dev) / len(e2e_dev),
                'stages_ms': stage_times},
        'parity_sha256': want_sha, 'parity_ok': parity_ok,
        'parity': {'oracle_sha256': want_sha, 'rank_sha256': all_sha, 'proofs_checked_per_rank': 3 * args.steps + 1,
                   'how': 'SHA-256 of the proof bytes every rank returned in every timed leg vs the C oracle port (oracle/c) proving the same workload on the host',
                   'note': oracle_note},
        'gpu_launches': int(launches_per_step),
        'roofline': roofline,
        'issue_roofline': issue_roofline,
        'backends': {'trace': trace_backend, 'constraints': st.compose_backend()},
        'cpu_baseline': cpu,
        'kernels_ms_per_step': {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
        'resident_wall_ms': wall_resident, 'profiled_leg_ms_per_step': sum(prof_dev) / len(prof_dev),
        'ntt': ntt,
        'throughput_mode': replicas,
        'clocks': clocks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL prints its version banner on the first
    # communicator) are sent to stderr, and the line is written to the real stdout at the end
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, 'w')
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default=os.environ.get('GS_BENCH_CONFIG', 'ns'), choices=sorted(METRICS))      # config 1 (Foo, p32) is a host-path parity case, not a bench line
    ap.add_argument('--quick-ntt', action='store_true', help='NTT table at two sizes only (development runs)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
