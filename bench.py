#!/usr/bin/env python
"""bench.py -- MiMC-128 prove() on B200 (BASELINE.json metric), one JSON line on rank 0.

A "step" is one Stark.prove() of the north-star workload (MiMC over p128, 2^20 steps, extension factor 8,
blake2s256, 48 trace / 24 FRI queries: examples/mimc/mimc128.ts:22-28 with the E of BASELINE.json's target).
  value        ms per prove with the execution trace already resident in HBM (CUDA events on the prover stream)
  e2e.value    ms per prove through the public API from host inputs to host-resident proof bytes
               (host trace generation + H2D of the trace + every D2H inside the timed region)
  roofline     dominant kernel class of the step: algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline the oracle port of the same path on the host cores (bounded sample), N=1 rank 0 only
`--impl reference` times that CPU port alone (the reference itself needs node + npm packages that this
image does not have: SURVEY.md §8c).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_STEPS = int(os.environ.get('GS_BENCH_LOG_STEPS', '20'))
EXT = int(os.environ.get('GS_BENCH_EXT', '8'))
OPTS = dict(hashAlgorithm='blake2s256', extensionFactor=EXT, exeQueryCount=48, friQueryCount=24)
METRIC = 'mimc128_prove_ms'


def workload_name(log_steps=LOG_STEPS, ext=EXT):
    return f'MiMC-128 prove(), 2^{log_steps} steps, extensionFactor {ext}, blake2s256, 48/24 queries (north-star target)'


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


def mimc_case(log_steps):
    from genstark_b200 import airs
    steps = 1 << log_steps
    air = airs.mimc128(steps)
    return air, steps


def mimc_assertions(steps, seed=3):
    """control values via the native library's scalar ops would be slow in Python for 2^20 steps; use the
    closed loop in Python ints (examples/mimc/utils.ts:7-14) -- outside every timed region."""
    from genstark_b200 import airs
    from genstark_b200.air import P128
    k = airs.mimc_round_constants()
    x = seed % P128
    for i in range(steps - 1):
        x = (x * x % P128 * x + k[i & 63]) % P128
    return [dict(step=0, register=0, value=seed), dict(step=steps - 1, register=0, value=x)]


# ------------------------------------------------------------------------------------------ CPU port
def cpu_port_prove_ms(log_steps, ext, threads=None):
    """time the oracle's prove() on the host.  Prefers the compiled C port (oracle/_build), else the
    Python restatement (single thread)."""
    try:
        from oracle import cport
        if cport.available():
            return cport.time_mimc_prove(log_steps, ext, threads)
    except Exception as e:        # pragma: no cover
        print(f'[bench] C port unavailable: {e}', file=sys.stderr)
    from genstark_b200 import airs
    from oracle.stark import Stark as OracleStark
    steps = 1 << log_steps
    air = airs.mimc128(steps)
    st = OracleStark(air, dict(OPTS, extensionFactor=ext))
    a = mimc_assertions(steps)
    t = time.perf_counter()
    st.prove(a, [], [3])
    return (time.perf_counter() - t) * 1e3, 1, 'python'


def run_reference(args, rank, world):
    if rank != 0:
        return
    # bounded sample: the Python port cannot run 2^20 steps in minutes; the C port can
    try:
        from oracle import cport
        have_c = cport.available()
    except Exception:
        have_c = False
    log_steps = LOG_STEPS if have_c else 12
    times = []
    cores = 1
    kind = 'python'
    for i in range(args.warmup + args.steps):
        ms, cores, kind = cpu_port_prove_ms(log_steps, EXT, None)
        if i >= args.warmup:
            times.append(ms)
    ms = sum(times) / len(times)
    scale = 1.0
    sample = f'prove() of MiMC-128 2^{log_steps} steps E={EXT} ({kind} oracle port)'
    if log_steps != LOG_STEPS:
        # N log N extrapolation to the named workload, stated as such
        n0, n1 = (1 << log_steps) * EXT, (1 << LOG_STEPS) * EXT
        scale = (n1 * (LOG_STEPS + 3)) / (n0 * (log_steps + 3))
        sample += f'; value extrapolated x{scale:.0f} by N log N to 2^{LOG_STEPS} steps'
    v = ms * scale
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'ms', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': v, 'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u128 (integer mod p)', 'data': 'synthetic',
            'config': {'workload': workload_name(), 'parallelism': 'cpu'},
            'cpu_baseline': {'value': v, 'unit': 'ms', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'ms', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': 'the reference (genSTARK + galois/merkle/air-assembly WASM) cannot run here: no node in the image'}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def algorithmic_bytes(cls, log_t, log_e, r, s, n_boundary):
    """SURVEY.md §8d / App. A.9 per-unit figures x units of ONE prove, per kernel class."""
    T, N = 1 << log_t, 1 << (log_t + log_e)
    B, D = 16, 32
    fri = []
    L = N
    while True:
        fri.append(L)
        if L <= 256:
            break
        L >>= 2
    if cls.startswith('ntt'):
        # iNTT(T) + LDE(T -> N) for R (+S) rows: 32 B/point of the transform actually computed
        return (r + s) * (2 * B * T + B * (T + N))
    if cls == 'hash_columns':
        return N * ((r + s) * B + D) + sum((l // 4) * (4 * B + D) for l in fri)
    if cls == 'merkle_build':
        return 2 * N * D + sum(2 * (l // 4) * D for l in fri)      # read 2 digests / write 1 per node ~ 2n*32... counted as 64 B/leaf
    if cls == 'compose':
        return N * B * (r + s + n_boundary + 1)
    if cls in ('batch_inverse', 'zb_eval'):
        return n_boundary * N * 2 * B
    if cls == 'fri_fold':
        return sum(B * l + B * l // 4 for l in fri[:-1])
    return 0


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

    from genstark_b200.field import Context, GpuField
    from genstark_b200.stark import Stark
    from genstark_b200 import _native
    L = _native.lib()

    air, steps = mimc_case(LOG_STEPS)
    ctx = Context(local_rank)
    if world > 1:
        # one proof sharded over the ranks by cosets (strong scaling): NCCL all-gather of digests at every Merkle commit
        if EXT % world:
            raise SystemExit(f'extension factor {EXT} has fewer cosets than ranks ({world})')
        uid = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    st = Stark(air, dict(OPTS), context=ctx)
    assertions = mimc_assertions(steps)
    seed = [3]

    sampler = ClockSampler(local_rank)
    sampler.start()

    # warm-up (also allocates every buffer)
    for _ in range(max(args.warmup, 3)):
        proof = st.prove_bytes(assertions, [], seed)
    proof_len = len(proof)
    from genstark_b200.stark import trace_backend as _tb
    trace_backend = _tb()

    # ---- e2e leg: public API, host inputs -> host proof bytes
    barrier()
    launches0 = ctx.launch_count
    t0 = time.perf_counter()
    t_region0 = t0
    e2e_dev = []
    for _ in range(args.steps):
        st.prove_bytes(assertions, [], seed)
        e2e_dev.append(st.last_timing())
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches_per_step = (ctx.launch_count - launches0) // args.steps
    stage_times = st.stage_times()

    # ---- resident leg: trace already in HBM; CUDA events on the prover stream around the whole device part
    st.prove_bytes(assertions, [], seed, _reuse_resident_trace=True)
    barrier()
    dev_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.prove_bytes(assertions, [], seed, _reuse_resident_trace=True)
        dev_ms.append(st.last_timing()[0])
    barrier()
    wall_resident = (time.perf_counter() - t0) * 1e3 / args.steps
    # ---- same K steps again with CUDA events around every kernel class (CUDA graphs off while profiling, so
    # this leg is a little slower than the one above; it provides the per-kernel times of the roofline)
    L.gs_ctx_profile(ctx.handle, 1)
    prof_dev = []
    for _ in range(args.steps):
        st.prove_bytes(assertions, [], seed, _reuse_resident_trace=True)
        prof_dev.append(st.last_timing()[0])
    barrier()
    prof = json.loads(L.gs_ctx_profile_report(ctx.handle).decode())
    L.gs_ctx_profile(ctx.handle, 0)
    t_region1 = time.perf_counter()
    sampler.stop()
    clocks = sampler.summary(t_region0, t_region1)

    ms_step = sum(dev_ms) / len(dev_ms)
    if dist is not None:
        t = torch.tensor([ms_step, e2e_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])

    # ---- K1 alone: NTT throughput (second half of BASELINE.json's metric)
    ntt = {}
    if rank == 0:
        f = GpuField(ctx)
        import random
        r = random.Random(0xB200)
        for name, log_t, log_n, inv in (('ntt_fwd_2^23', 23, 23, 0), ('lde_2^20_to_2^23', 20, 23, 0), ('intt_2^20', 20, 20, 1)):
            t_, n_ = 1 << log_t, 1 << log_n
            # uniform 127-bit residues (canonical: top bit of every element cleared)
            ba = bytearray(r.randbytes(16 * t_))
            ba[15::16] = bytes(x & 0x7F for x in ba[15::16])
            src = f._from_bytes(bytes(ba), 1, t_)
            dst, work = C.c_void_p(), C.c_void_p()
            ctx.check(L.gs_mat_alloc(ctx.handle, 1, n_, C.byref(dst)))
            ctx.check(L.gs_mat_alloc(ctx.handle, 1, n_, C.byref(work)))
            ms = C.c_float()
            best = []
            for i in range(8):
                L.gs_timer_begin(ctx.handle)
                ctx.check(L.gs_ntt_into(ctx.handle, src.handle, dst, work, inv))
                L.gs_timer_end(ctx.handle, C.byref(ms))
                if i >= 3:
                    best.append(ms.value)
            med = sorted(best)[len(best) // 2]
            alg_bytes = 16 * (t_ + n_)
            ntt[name] = {'ms': round(med, 4), 'elements_per_s': n_ / (med * 1e-3), 'algorithmic_GBps': alg_bytes / (med * 1e-3) / 1e9}
            L.gs_mat_free(dst); L.gs_mat_free(work); src.free()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    hbm = float(peaks.get('hbm_gbs', 6650.0))
    # dominant kernel class of the resident step
    per_step = {k: v['ms'] / args.steps for k, v in prof.items()}
    # group the K1 passes
    grouped = {}
    for k, v in per_step.items():
        key = 'ntt' if k.startswith('ntt') else k
        grouped[key] = grouped.get(key, 0.0) + v
    dom = max(grouped, key=grouped.get)
    log_t, log_e = LOG_STEPS, EXT.bit_length() - 1
    # sharded runs: kernel times are rank 0's, which holds 1/world of the evaluation domain
    alg = algorithmic_bytes(dom, log_t, log_e, air.trace_register_count, air.secret_input_count, 1) // world
    achieved = alg / (grouped[dom] * 1e-3) / 1e9
    traffic = None
    summ = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
    if os.path.exists(summ):
        try:
            traffic = json.load(open(summ)).get(dom, {}).get('dram_bytes_per_step')
            traffic = traffic // world if traffic else traffic
        except Exception:
            traffic = None
    roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm,
                'traffic': traffic, 'peak_source': f'{peak_kind} (MEASURED_PEAKS.json hbm_gbs)',
                'kernel_ms_per_step': grouped[dom], 'algorithmic_bytes_per_step': alg,
                'share_of_step': grouped[dom] / (sum(prof_dev) / len(prof_dev)),
                'note': '128-bit modular arithmetic on 32-bit integer pipes: every kernel here is issue-bound, not HBM-bound'}

    # The roof that actually binds these kernels is integer issue, not HBM (profiles/): say so with numbers.
    #  * blake2s kernels: 648 ALU-pipe instructions per compression (320 XOR + 320 rotate + address/feed-forward), the ALU
    #    pipe retires one warp instruction every 2 cycles per SM sub-partition => compressions/s <= SMs*4*32*f / (2*648)
    #  * modmul kernels: gs_debug_modmul_probe measures the chip's dependent-free modular-multiplication rate
    sm_hz = (clocks.get('sm_mhz') or 1965.0) * 1e6
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    comp_peak = sm_count * 4 * 32 * sm_hz / (2 * 648.0)
    n_eval = steps * EXT
    fri_rows, l_ = 0, n_eval
    while True:
        fri_rows += l_ // 4
        if l_ <= 256:
            break
        l_ >>= 2
    comp_hash_cols = n_eval + fri_rows
    comp_merkle = (n_eval - 1) + fri_rows
    issue_roofline = {
        'hash_columns': {'unit': 'blake2s compressions/s', 'achieved': comp_hash_cols / (grouped.get('hash_columns', 0) * 1e-3) if grouped.get('hash_columns') else None,
                         'peak': comp_peak, 'peak_how': 'ALU pipe: SMs*4*32 lanes*f / (2 cycles * 648 ALU instructions per compression)'},
        'merkle_build': {'unit': 'blake2s compressions/s', 'achieved': comp_merkle / (grouped.get('merkle_build', 0) * 1e-3) if grouped.get('merkle_build') else None,
                         'peak': comp_peak, 'peak_how': 'same ALU-pipe bound; the top levels of every tree are latency-bound'},
    }
    for v in issue_roofline.values():
        v['frac'] = (v['achieved'] / v['peak']) if v['achieved'] else None
    probe_ms = C.c_float()
    if world > 1:
        issue_roofline = None            # per-rank work counts differ per commit on the sharded path: N=1 only
    elif L.gs_debug_modmul_probe(ctx.handle, sm_count * 8, 2000, C.byref(probe_ms)) == 0 and probe_ms.value > 0:
        modmul_peak = sm_count * 8 * 256 * 4 * 2000.0 / (probe_ms.value * 1e-3)
        issue_roofline['modmul_probe'] = {'unit': 'modmul/s', 'peak': modmul_peak, 'probe_ms': probe_ms.value,
                                          'how': 'gs_debug_modmul_probe: 8 CTAs x 256 threads per SM, 4 independent chains each'}
        # K1 against the rate of its own instruction mix: a butterfly = modular add + modular sub + modular multiplication
        # (gs_debug_butterfly_probe).  Butterflies of one prove: iNTT of T points + E coset transforms of T points, (n/2) log2 n each
        # (the register-resident radix-8/16 stages skip the multiplications by 1, so this counts a few more multiplications than issued).
        bf_ms = C.c_float()
        if L.gs_debug_butterfly_probe(ctx.handle, sm_count * 8, 2000, C.byref(bf_ms)) == 0 and bf_ms.value > 0 and grouped.get('ntt'):
            bf_peak = sm_count * 8 * 256 * 2 * 2000.0 / (bf_ms.value * 1e-3)
            rows = air.trace_register_count + sum(1 for s_ in air.static_registers if s_.kind == 'input')
            n_bf = rows * (1 + EXT) * (steps // 2) * LOG_STEPS
            a = n_bf / (grouped['ntt'] * 1e-3)
            issue_roofline['ntt'] = {'unit': 'butterflies/s', 'achieved': a, 'peak': bf_peak, 'frac': a / bf_peak, 'probe_ms': bf_ms.value,
                                     'butterflies_per_prove': n_bf}
        if ntt.get('lde_2^20_to_2^23') and bf_ms.value > 0:
            bf_peak = sm_count * 8 * 256 * 2 * 2000.0 / (bf_ms.value * 1e-3)
            a = (8 * (1 << 19) * 20) / (ntt['lde_2^20_to_2^23']['ms'] * 1e-3)
            issue_roofline['lde_2^20_to_2^23'] = {'unit': 'butterflies/s', 'achieved': a, 'peak': bf_peak, 'frac': a / bf_peak}

    cpu = None
    if world == 1:
        try:
            from oracle import cport
            have_c = cport.available()
        except Exception:
            have_c = False
        ls = LOG_STEPS if have_c else 12
        ms_cpu, cores, kind = cpu_port_prove_ms(ls, EXT, None)
        sample = f'one prove() of MiMC-128 2^{ls} steps E={EXT} by the {kind} oracle port'
        if ls != LOG_STEPS:
            n0, n1 = (1 << ls) * EXT, (1 << LOG_STEPS) * EXT
            sc = (n1 * (LOG_STEPS + 3)) / (n0 * (ls + 3))
            ms_cpu *= sc
            sample += f', extrapolated x{sc:.0f} (N log N) to 2^{LOG_STEPS} steps'
        cpu = {'value': ms_cpu, 'unit': 'ms', 'cores': cores, 'kind': 'port', 'sample': sample}

    trace_bytes = air.trace_register_count * steps * 16
    value, e2e_val = ms_step, e2e_ms
    line = {
        'metric': METRIC, 'value': value, 'unit': 'ms', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_step, 'higher_is_better': False, 'scaling': 'weak' if world == 1 else 'strong', 'vs_baseline': None,
        'dtype': 'u128 (integer mod p = 2^128 - 9*2^32 + 1, 4x u32 limbs)', 'data': 'synthetic',
        'config': {'workload': workload_name(), 'parallelism': 'single GPU' if world == 1 else f'one proof sharded over {world} GPUs by cosets ({EXT // world} of {EXT} cosets per rank); NCCL all-gather of digests at each Merkle commit, one all-reduce of queried rows; trace generation replicated on every host process',
                   'l2': 'working set (>= 128 MiB per vector, ~1.4 GiB per prove) exceeds the 126 MB L2; no explicit flush',
                   'proof_bytes': proof_len},
        'e2e': {'value': e2e_val, 'unit': 'ms', 'h2d_bytes_per_step': trace_bytes + 4096, 'd2h_bytes_per_step': proof_len + 32 * 12,
                'device_ms_inside': sum(d for d, _ in e2e_dev) / len(e2e_dev), 'host_ms_inside': sum(h for _, h in e2e_dev) / len(e2e_dev),
                'stages_ms': stage_times},
        'gpu_launches': int(launches_per_step),
        'roofline': roofline,
        'issue_roofline': issue_roofline,
        'backends': {'trace': trace_backend, 'constraints': st.compose_backend()},
        'cpu_baseline': cpu,
        'kernels_ms_per_step': {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
        'resident_wall_ms': wall_resident, 'profiled_leg_ms_per_step': sum(prof_dev) / len(prof_dev),
        'ntt': ntt,
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
