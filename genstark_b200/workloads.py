"""Workload builders (AIR, options, assertions, inputs, seed) of the BASELINE.json configs: what the parity tests prove
and what bench.py times.  Inputs follow the reference's examples (examples/mimc/mimc128.ts:17-28,66-69,
examples/rescue/hash4x128.ts:40-46,114-161, examples/poseidon/merkleProof.ts:25-31,107-131)."""
import random

from genstark_b200 import airs
from genstark_b200.air import P128


def mimc(steps, e, alg='blake2s256', seed=3):
    opts = dict(hashAlgorithm=alg, extensionFactor=e, exeQueryCount=48, friQueryCount=24)
    air = airs.mimc128(steps)
    ctl = airs.run_mimc(steps, airs.mimc_round_constants(), seed)
    a = [dict(step=0, register=0, value=ctl[0]), dict(step=steps - 1, register=0, value=ctl[-1])]
    return air, opts, a, [], [seed]


def rescue(instances, e=16, alg='blake2s256'):
    """config 3: chained Rescue 4x128 hash instances (options of examples/rescue/hash4x128.ts:40-46)."""
    opts = dict(hashAlgorithm=alg, extensionFactor=e, exeQueryCount=68, friQueryCount=24)
    air = airs.rescue4x128(instances)
    rows = [airs.rescue_build_inputs([42 + i, 43 + 7 * i]) for i in range(instances)]
    inputs = [[rows[k][reg] for k in range(instances)] for reg in range(4)]
    # control: run the plain permutation for the first instance
    a = [dict(step=0, register=reg, value=rows[0][reg]) for reg in range(2)]
    return air, opts, a, inputs, []


def rescue_hash_control(values):
    """independent plain implementation: 31 double-rounds from the first row (hash4x128.ts:86-93)."""
    p = P128
    _, rc = airs.rescue_key_schedule(p)
    st = airs.rescue_build_inputs(values)
    inv_e = (-airs.RESCUE_INV_ALPHA) % (p - 1)
    for i in range(31):
        s = [(sum(airs.RESCUE_MDS[r][j] * pow(st[j], 3, p) for j in range(4)) + rc[r][i]) % p for r in range(4)]
        t = [pow(x, inv_e, p) for x in s]
        st = [(sum(airs.RESCUE_MDS[r][j] * t[j] for j in range(4)) + rc[4 + r][i]) % p for r in range(4)]
    return st


def poseidon(depth, proofs, e=32, alg='blake2s256', seed=11):
    """config 5: `proofs` Poseidon Merkle branches of `depth` levels (options of examples/poseidon/merkleProof.ts:25-31)."""
    opts = dict(hashAlgorithm=alg, extensionFactor=e, exeQueryCount=44, friQueryCount=20)
    r = random.Random(seed)
    leaves = [[r.randrange(P128), r.randrange(P128)] for _ in range(2 ** depth)]
    air = airs.poseidon_merkle_proof(depth=depth, proofs=proofs)
    leaf0, leaf1, node0, node1, bits = [], [], [], [], []
    roots = []
    for k in range(proofs):
        index = (2 * r.randrange(2 ** (depth - 1))) if depth > 1 else 0        # top bit 0 -> root lands in registers 0,1
        index = index % (2 ** (depth - 1)) if depth > 1 else 0
        leaf, nodes, b, root = airs.poseidon_merkle_inputs(index, leaves)
        leaf0.append(leaf[0]); leaf1.append(leaf[1])
        node0.append([n[0] for n in nodes]); node1.append([n[1] for n in nodes])
        bits.append([0] + b[:-1])
        roots.append(root)
    inputs = [leaf0, leaf1, node0, node1, bits]
    period = 64 * depth
    a = [dict(step=period - 1, register=0, value=roots[0][0]), dict(step=period - 1, register=1, value=roots[0][1])]
    if proofs > 1:
        a.append(dict(step=period * proofs - 1, register=0, value=roots[-1][0]))
    return air, opts, a, inputs, []


def foo(steps=64, e=None, alg='sha256'):
    """config 1: the Foo demo (README.md:22-50): x_{n+1} = x_n + 2 over 2^32 - 3*2^25 + 1, 64 steps, start value 1 -- proved on the host
    (genstark_b200.stark.HostStark); options of the README example: extensionFactor default, 80 / 40 queries capped by the domain"""
    opts = dict(hashAlgorithm=alg, exeQueryCount=32, friQueryCount=16)
    if e:
        opts['extensionFactor'] = e
    a = [dict(step=0, register=0, value=1), dict(step=steps - 1, register=0, value=1 + 2 * (steps - 1))]
    return airs.foo(steps), opts, a, [[1]], []


CONFIGS = {
    # name: (builder, args, description)  -- SURVEY.md section 8 config legend
    '1': ('foo', (64,), 'Foo demo (x_{n+1} = x_n + 2), 64 steps over 2^32 - 3*2^25 + 1, host path, no GPU (BASELINE config 1)'),
    'ns': ('mimc', (1 << 20, 8), 'MiMC-128 prove(), 2^20 steps, extensionFactor 8, blake2s256, 48/24 queries (north-star target)'),
    '2': ('mimc', (1 << 13, 8), 'MiMC-128 prove(), 2^13 steps, extensionFactor 8, blake2s256, 48/24 queries (BASELINE config 2)'),
    '3': ('rescue', (128, 16), 'Rescue 4x128 hash chain prove(), 128 instances = 2^12 steps, 4 registers, extensionFactor 16, blake2s256, 68/24 queries (BASELINE config 3)'),
    '4': ('mimc', (1 << 20, 16), 'MiMC-128 prove(), 2^20 steps, extensionFactor 16, blake2s256, 48/24 queries (BASELINE config 4)'),
    'test': ('mimc', (1 << 10, 8), 'MiMC-128 prove(), 2^10 steps, extensionFactor 8, blake2s256, 48/24 queries (the shape tests/test_bench_contract.py runs)'),
    '5': ('poseidon', (8, 128, 32), 'Poseidon Merkle-proof prove(), 128 branches of depth 8 = 2^16 steps, 12 registers, extensionFactor 32, blake2s256, 44/20 queries (BASELINE config 5)'),
}


def config(name):
    """(air, options, assertions, inputs, seed, description) of a BASELINE config"""
    builder, args, desc = CONFIGS[str(name)]
    return (*globals()[builder](*args), desc)
