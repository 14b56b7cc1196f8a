"""Hand-built AIRs for the workloads BASELINE.json names (SURVEY.md §8 config legend).

The reference compiles these from AirScript / AirAssembly text with packages that are not in the
reference tree; the front-end is out of scope, so each AIR is written directly in
the IR of ``air.py`` with the reference source it restates cited beside it.
"""
from __future__ import annotations

import hashlib
from typing import List, Sequence

from .air import (AirModule, ProgramBuilder, StaticRegister, P128, P32, gather_column_blob, gather_columns_blob, prng_sha256)

MIMC_SEED = bytes.fromhex('4d694d43')       # examples/mimc/mimc128.ts:15,36


def mimc_round_constants(count: int = 64, modulus: int = P128) -> List[int]:
    return prng_sha256(MIMC_SEED, count, modulus)


def mimc128(steps: int = 2**13, constant_count: int = 64, extension_factor: int = None) -> AirModule:
    """MiMC over p128: x_{s+1} = x_s^3 + k[s mod 64]; 1 register, 1 constraint of degree 3.
    Restates the AirAssembly module at examples/mimc/mimc128Assembly.ts:28-51 (no input registers:
    the start value comes in through ``seed``, mimc128Assembly.ts:66)."""
    p = P128
    t = ProgramBuilder(p)
    t.out(0, t.exp(t.cur(0), 3) + t.static(0))
    e = ProgramBuilder(p)
    e.out(0, e.nxt(0) - (e.exp(e.cur(0), 3) + e.static(0)))
    return AirModule(
        name='mimc', modulus=p, trace_register_count=1, trace_length=steps,
        transition=t.build(), evaluation=e.build(),
        static_registers=[StaticRegister('cycle', mimc_round_constants(constant_count, p))],
        extension_factor=extension_factor,
        init=lambda inputs, seed: [int(seed[0]) % p])


def run_mimc(steps: int, round_constants: List[int], seed: int, modulus: int = P128) -> List[int]:
    """Control values, examples/mimc/utils.ts:7-14."""
    out = [seed % modulus]
    for i in range(steps - 1):
        out.append((pow(out[i], 3, modulus) + round_constants[i % len(round_constants)]) % modulus)
    return out


def foo(steps: int = 64, extension_factor: int = None) -> AirModule:
    """README.md:22-39: x_{n+1} = x_n + 2 over p32 with one secret input register (startValue held
    for the whole segment).  CPU-only plumbing config (BASELINE.json configs[0])."""
    p = P32
    t = ProgramBuilder(p)
    t.out(0, t.cur(0) + 2)
    e = ProgramBuilder(p)
    e.out(0, e.nxt(0) - (e.cur(0) + 2))
    return AirModule(
        name='foo', modulus=p, trace_register_count=1, trace_length=steps,
        transition=t.build(), evaluation=e.build(),
        static_registers=[StaticRegister('input', secret=True)],
        extension_factor=extension_factor,
        init=lambda inputs, seed: [int(inputs[0][0]) % p],
        expand_inputs=lambda inputs: [[int(inputs[0][0]) % p] * steps],
        input_shapes=lambda inputs: [[1]])


def fibonacci(steps: int = 64, modulus: int = P32, extension_factor: int = None) -> AirModule:
    """examples/demo/fibonacci.ts:13-34: 2 registers, a0' = a0+a1, a1' = a1 + (a0+a1)."""
    p = modulus
    t = ProgramBuilder(p)
    a2 = t.cur(0) + t.cur(1)
    t.out(0, a2)
    t.out(1, t.cur(1) + a2)
    e = ProgramBuilder(p)
    a2 = e.cur(0) + e.cur(1)
    e.out(0, e.nxt(0) - a2)
    e.out(1, e.nxt(1) - (e.cur(1) + a2))
    return AirModule(
        name='fibonacci', modulus=p, trace_register_count=2, trace_length=steps,
        transition=t.build(), evaluation=e.build(), extension_factor=extension_factor,
        init=lambda inputs, seed: [int(seed[0]) % p, int(seed[1]) % p])


# ---------------------------------------------------------------------------------------------- Rescue
# Public parameters of examples/rescue/hash4x128.ts:13-33 (data, not code).
RESCUE_ALPHA = 3
RESCUE_INV_ALPHA = 113427455640312821154458202464371168597        # 3 * this = p - 2 == -1 (mod p - 1)
RESCUE_MDS = [
    [340282366920938463463374607393113505064, 340282366920938463463374607393113476633, 340282366920938463463374607393112623703, 340282366920938463463374607393088807273],
    [1080, 42471, 1277640, 35708310],
    [340282366920938463463374607393113505403, 340282366920938463463374607393113491273, 340282366920938463463374607393113076364, 340282366920938463463374607393101570233],
    [40, 1210, 33880, 925771]]
RESCUE_INV_MDS = [
    [236997924285633886309140921207528337986, 247254910923297358352547052529406562002, 311342028444809266296393502237594936029, 126030506267014245727175780515967965110],
    [33069997328254894416993606273702832836, 59740111947936946229464514160137230831, 88480676416265968399408181712033476738, 124630167308491865219096049621346098829],
    [336618017400133662891528246258390023400, 144341202744775798260123226512082052891, 154884404066691444097361840554534567820, 4667796528407935026932436315406220930],
    [73878794827854483309086441046605817365, 229228508225866824084614421584601165863, 125857624914110248133585690282064031000, 84953896817024417490170340940393220925]]
RESCUE_CONSTANTS = [
    144517900019036866096022507193071809599, 271707809579969091656092579345468860225, 139424957805302989189422527487860690608, 126750251129487986697737866024960215983,
    271118613762407276564214152179206069413, 39384648060424157691646880565718875760, 189037434251220539428539337560615209464, 218986062987136192416421725751708413726,
    103808983578136303126641899945581033860, 198823153506012419365570940451368319246, 339599443104046223725845265111864465825, 169004341575174204803282453992954960786,
    171596418631454858790177474513731208863, 157569361262795131998922854453557743690, 211837534394685913032370295607135890739, 328609939009439440841980058678511564944,
    229628671790616575443886906286361261591, 95675137928612392156876334331168593412, 301613873771889848137714364785485714735, 278224571298089265666737094541710980794,
    140049647417493050970983064725330334359, 159594320057012289760186736637936788141, 44954493393746175043012738454844468290, 223519669575552375517628855932195463175]
RESCUE_STEPS = 32


def _mmul(p, m, v):
    return [sum(a * b for a, b in zip(row, v)) % p for row in m]


def rescue_key_schedule(p: int = P128, width: int = 4, rounds: int = RESCUE_STEPS):
    """Rescue.unrollConstants + groupConstants (examples/rescue/utils.ts:128-180)."""
    c = list(RESCUE_CONSTANTS)
    i_const, c = c[:width], c[width:]
    c_matrix = [c[i * width:(i + 1) * width] for i in range(width)]
    c_const = c[width * width:width * width + width]
    inv_alpha_e = (-RESCUE_INV_ALPHA) % (p - 1)       # exp(x, -k) = inv(x)^k
    key_state = list(i_const)
    inj = list(i_const)
    result = [list(key_state)]
    for _ in range(rounds + 1):
        key_state = [pow(x, inv_alpha_e, p) for x in key_state]
        inj = [(a + b) % p for a, b in zip(_mmul(p, c_matrix, inj), c_const)]
        key_state = [(a + b) % p for a, b in zip(_mmul(p, RESCUE_MDS, key_state), inj)]
        result.append(list(key_state))
        key_state = [pow(x, RESCUE_ALPHA, p) for x in key_state]
        inj = [(a + b) % p for a, b in zip(_mmul(p, c_matrix, inj), c_const)]
        key_state = [(a + b) % p for a, b in zip(_mmul(p, RESCUE_MDS, key_state), inj)]
        result.append(list(key_state))
    initial = result[0] + result[1]
    rc = [[0] * rounds for _ in range(2 * width)]
    k = 2
    for i in range(rounds):
        for j in range(width):
            rc[j][i] = result[k][j]
            rc[width + j][i] = result[k + 1][j]
        k += 2
    return initial, rc


def rescue_build_inputs(values: Sequence[int], p: int = P128) -> List[int]:
    """buildInputs, examples/rescue/hash4x128.ts:131-161: the first trace row for one hash instance."""
    initial, _ = rescue_key_schedule(p)
    inv_alpha_e = (-RESCUE_INV_ALPHA) % (p - 1)
    r = [(values[0] + initial[0]) % p, (values[1] + initial[1]) % p, initial[2], initial[3]]
    a = [pow(x, inv_alpha_e, p) for x in r]
    r = _mmul(p, RESCUE_MDS, a)
    return [(r[i] + initial[4 + i]) % p for i in range(4)]


def rescue4x128(instances: int = 1, extension_factor: int = None) -> AirModule:
    """Rescue hash preimage, 4 registers over p128, `instances` chained 32-step segments
    (examples/rescue/hash4x128.ts:49-109).  The AirScript `for each (value1, value2)` loop is written out
    by hand: four secret input registers hold the NEXT segment's start row (AirAssembly `shift -1`), a
    cyclic mask marks the transition into a new segment, and every constraint is
        mask * (n_i - input_i) + (1 - mask) * (S_i - N_i),
    S = mds # r^alpha + k[0..3], N = (inv_mds # (n - k[4..7]))^alpha  (hash4x128.ts:96-107)."""
    p = P128
    steps = RESCUE_STEPS * instances
    _, rc = rescue_key_schedule(p)
    statics = [StaticRegister('cycle', list(rc[j])) for j in range(8)]
    statics.append(StaticRegister('cycle', [0] * (RESCUE_STEPS - 1) + [1]))        # 8: mask
    statics += [StaticRegister('input', secret=True) for _ in range(4)]             # 9..12
    inv_alpha_e = (-RESCUE_INV_ALPHA) % (p - 1)

    def forward(b, r):
        return [sum((b.const(RESCUE_MDS[i][j]) * b.exp(r[j], RESCUE_ALPHA) for j in range(1, 4)),
                    b.const(RESCUE_MDS[i][0]) * b.exp(r[0], RESCUE_ALPHA)) + b.static(i) for i in range(4)]

    t = ProgramBuilder(p)
    r = [t.cur(i) for i in range(4)]
    s = forward(t, r)
    ts = [t.exp(x, inv_alpha_e) for x in s]
    mask = t.static(8)
    for i in range(4):
        nxt = sum((t.const(RESCUE_MDS[i][j]) * ts[j] for j in range(1, 4)), t.const(RESCUE_MDS[i][0]) * ts[0]) + t.static(4 + i)
        t.out(i, mask * t.static(9 + i) + (1 - mask) * nxt)
    e = ProgramBuilder(p)
    r = [e.cur(i) for i in range(4)]
    s = forward(e, r)
    d = [e.nxt(j) - e.static(4 + j) for j in range(4)]
    mask = e.static(8)
    for i in range(4):
        n_i = e.exp(sum((e.const(RESCUE_INV_MDS[i][j]) * d[j] for j in range(1, 4)), e.const(RESCUE_INV_MDS[i][0]) * d[0]), RESCUE_ALPHA)
        e.out(i, mask * (e.nxt(i) - e.static(9 + i)) + (1 - mask) * (s[i] - n_i))

    def expand(inputs):
        # inputs: four lists (one per register) with one start value per instance; shifted by one segment
        out = []
        for reg in range(4):
            vals = [int(v) % p for v in inputs[reg]]
            assert len(vals) == instances
            out.append([vals[((step + 1) // RESCUE_STEPS) % instances] for step in range(steps)])
        return out

    def expand_blob(inputs, as_buffer=False):
        """same columns as expand(), gathered with numpy (prove path: 16 384 Python integers cost more than the device part)"""
        import numpy as np
        for reg in range(4):
            assert len(inputs[reg]) == instances
        idx = ((np.arange(steps, dtype=np.int64) + 1) // RESCUE_STEPS) % instances
        return gather_columns_blob([(inputs[reg], idx) for reg in range(4)], p, as_buffer)

    return AirModule(
        name='rescue4x128', modulus=p, trace_register_count=4, trace_length=steps,
        transition=t.build(), evaluation=e.build(), static_registers=statics, extension_factor=extension_factor,
        init=lambda inputs, seed: [int(inputs[reg][0]) % p for reg in range(4)],
        expand_inputs=expand, expand_inputs_blob=expand_blob, input_shapes=lambda inputs: [[instances] for _ in range(4)])


# -------------------------------------------------------------------------------------------- Poseidon
POSEIDON_WIDTH, POSEIDON_RF, POSEIDON_RP, POSEIDON_ALPHA = 6, 8, 55, 5
POSEIDON_CYCLE = POSEIDON_RF + POSEIDON_RP + 1          # 64 (examples/poseidon/merkleProof.ts:12-16)


def _sha_const(s: str, p: int) -> int:
    return int.from_bytes(hashlib.sha256(s.encode()).digest(), 'big') % p


def poseidon_mds(p: int = P128, width: int = POSEIDON_WIDTH):
    """getMdsMatrix, examples/poseidon/utils.ts:64-79 (Cauchy matrix; equals assembly/lib128.aa:7-12)."""
    xs = [_sha_const(f'HadesMDSx{i}', p) for i in range(width)]
    ys = [_sha_const(f'HadesMDSy{i}', p) for i in range(width)]
    return [[pow((xs[i] - ys[j]) % p, p - 2, p) for j in range(width)] for i in range(width)]


def poseidon_round_constants(p: int = P128, width: int = POSEIDON_WIDTH, rounds: int = POSEIDON_CYCLE):
    """getRoundConstants, examples/poseidon/utils.ts:51-62"""
    out, c = [], 0
    for _ in range(rounds):
        row = []
        for _ in range(width):
            row.append(_sha_const(f'Hades{c}', p)); c += 1
        out.append(row)
    return out


def poseidon_hash(inputs: Sequence[int], p: int = P128) -> List[int]:
    """createHash, examples/poseidon/utils.ts:19-49: the plain control implementation."""
    m, rf, rp = POSEIDON_WIDTH, POSEIDON_RF, POSEIDON_RP
    mds, ark = poseidon_mds(p), poseidon_round_constants(p, m, rf + rp)
    state = [int(x) % p for x in inputs] + [0] * (m - len(inputs))
    for i in range(rf + rp):
        state = [(a + b) % p for a, b in zip(state, ark[i])]
        if i < rf // 2 or i >= rf // 2 + rp:
            state = [pow(x, POSEIDON_ALPHA, p) for x in state]
        else:
            state[m - 1] = pow(state[m - 1], POSEIDON_ALPHA, p)
        state = _mmul(p, mds, state)
    return state[:2]


def poseidon_merkle_proof(depth: int = 8, proofs: int = 1, extension_factor: int = None) -> AirModule:
    """Poseidon Merkle-branch verification, 12 registers (examples/poseidon/merkleProof.ts:34-102), `proofs`
    branches of `depth` levels, 64 steps per level.  Hand lowering of the nested `for each` loops:
      statics 0..5  round constants (cycle 64)        6  full-round mask (cycle 64)
              7     level-init mask (cycle 64)        8  proof-init mask (cycle 64*depth)
      inputs  9,10  leaf (secret)   11,12 node (secret)   13 indexBit (public) -- each holds the value the
              NEXT init transition consumes.
    Transition s -> s+1: rounds use the constants of step s; the row 64k is built from the inputs."""
    p = P128
    cyc = POSEIDON_CYCLE
    steps = cyc * depth * proofs
    mds = poseidon_mds(p)
    ark = poseidon_round_constants(p)
    rct = [[ark[s][j] for s in range(cyc)] for j in range(POSEIDON_WIDTH)]
    full = [1 if (s < POSEIDON_RF // 2 or POSEIDON_RF // 2 + POSEIDON_RP <= s < POSEIDON_RF + POSEIDON_RP) else 0 for s in range(cyc)]
    statics = [StaticRegister('cycle', rct[j]) for j in range(6)]
    statics.append(StaticRegister('cycle', full))                                        # 6
    statics.append(StaticRegister('cycle', [0] * (cyc - 1) + [1]))                       # 7
    period = cyc * depth
    statics.append(StaticRegister('cycle', [0] * (period - 1) + [1]))                    # 8
    statics += [StaticRegister('input', secret=True) for _ in range(4)]                  # 9..12
    statics.append(StaticRegister('input', secret=False))                                # 13

    def build(b, nxt_out):
        r = [b.cur(i) for i in range(12)]
        k = [b.static(j) for j in range(6)]
        m_full, m_lvl, m_proof = b.static(6), b.static(7), b.static(8)
        leaf, node, bit = [b.static(9), b.static(10)], [b.static(11), b.static(12)], b.static(13)
        outs = []
        for half in range(2):
            st = r[6 * half:6 * half + 6]
            added = [st[j] + k[j] for j in range(6)]
            sb_full = [b.exp(x, POSEIDON_ALPHA) for x in added]
            fr = [sum((b.const(mds[i][j]) * sb_full[j] for j in range(1, 6)), b.const(mds[i][0]) * sb_full[0]) for i in range(6)]
            part_in = added[:5] + [sb_full[5]]
            pr = [sum((b.const(mds[i][j]) * part_in[j] for j in range(1, 6)), b.const(mds[i][0]) * part_in[0]) for i in range(6)]
            outs.append([m_full * fr[i] + (1 - m_full) * pr[i] for i in range(6)])
        # level init: H = bit ? r[6..7] : r[0..1]
        h = [bit * r[6 + i] + (1 - bit) * r[i] for i in range(2)]
        lvl = [h[0], h[1], node[0], node[1], 0, 0, node[0], node[1], h[0], h[1], 0, 0]
        prf = [leaf[0], leaf[1], node[0], node[1], 0, 0, node[0], node[1], leaf[0], leaf[1], 0, 0]
        rounds = outs[0] + outs[1]
        for i in range(12):
            init_val = m_proof * prf[i] + (1 - m_proof) * lvl[i] if not isinstance(lvl[i], int) else None
            if init_val is None:
                nxt = (1 - m_lvl) * rounds[i]
            else:
                nxt = m_lvl * init_val + (1 - m_lvl) * rounds[i]
            nxt_out(i, nxt)

    t = ProgramBuilder(p)
    build(t, lambda i, v: t.out(i, v))
    e = ProgramBuilder(p)
    build(e, lambda i, v: e.out(i, e.nxt(i) - v))

    def expand(inputs):
        # inputs: [leaf0[proofs], leaf1[proofs], node0[proofs][depth], node1[proofs][depth], bits[proofs][depth]]
        leaf0, leaf1, node0, node1, bits = inputs
        regs = [[0] * steps for _ in range(5)]
        for s in range(steps):
            nxt = (s + 1) % steps                      # the init transition that consumes the value
            pr, lv = nxt // period, (nxt % period) // cyc
            regs[0][s] = int(leaf0[pr]) % p; regs[1][s] = int(leaf1[pr]) % p
            regs[2][s] = int(node0[pr][lv]) % p; regs[3][s] = int(node1[pr][lv]) % p
            regs[4][s] = int(bits[pr][lv]) % p
        return regs

    def expand_blob(inputs, as_buffer=False):
        """same columns as expand(), gathered with numpy"""
        import numpy as np
        leaf0, leaf1, node0, node1, bits = inputs
        nxt = (np.arange(steps, dtype=np.int64) + 1) % steps
        pr, lv = nxt // period, (nxt % period) // cyc
        flat = lambda m: [x for row in m for x in row]
        at = pr * depth + lv
        return gather_columns_blob([(leaf0, pr), (leaf1, pr), (flat(node0), at), (flat(node1), at), (flat(bits), at)], p, as_buffer)

    def expand_public(public_inputs):
        (bits,) = public_inputs
        out = [0] * steps
        for s in range(steps):
            nxt = (s + 1) % steps
            out[s] = int(bits[nxt // period][(nxt % period) // cyc]) % p
        return [out]

    def expand_public_blob(public_inputs):
        """the column of expand_public(), gathered with numpy (verify path)"""
        import numpy as np
        (bits,) = public_inputs
        nxt = (np.arange(steps, dtype=np.int64) + 1) % steps
        at = (nxt // period) * depth + (nxt % period) // cyc
        return gather_columns_blob([([x for row in bits for x in row], at)], p)

    def init(inputs, seed):
        leaf0, leaf1, node0, node1, _ = inputs
        lf, nd = [int(leaf0[0]) % p, int(leaf1[0]) % p], [int(node0[0][0]) % p, int(node1[0][0]) % p]
        return lf + nd + [0, 0] + nd + lf + [0, 0]

    return AirModule(
        name='poseidon_mp', modulus=p, trace_register_count=12, trace_length=steps,
        transition=t.build(), evaluation=e.build(), static_registers=statics, extension_factor=extension_factor,
        init=init, expand_inputs=expand, expand_public_inputs=expand_public, expand_inputs_blob=expand_blob,
        expand_public_inputs_blob=expand_public_blob,
        input_shapes=lambda inputs: [[proofs], [proofs], [proofs, depth], [proofs, depth], [proofs, depth]])


_TREE_CACHE: dict = {}


def _poseidon_tree(leaves, p):
    """all levels of the Poseidon Merkle tree over `leaves` (one tree serves every branch of a workload)"""
    key = (hash(leaves), p)
    if key not in _TREE_CACHE:
        level = [list(x) for x in leaves]
        tree = [level]
        while len(level) > 1:
            level = [poseidon_hash(level[2 * i] + level[2 * i + 1], p) for i in range(len(level) // 2)]
            tree.append(level)
        _TREE_CACHE.clear()
        _TREE_CACHE[key] = tree
    return _TREE_CACHE[key]


def poseidon_merkle_inputs(index: int, leaves: Sequence[Sequence[int]], p: int = P128):
    """A depth-log2(len(leaves)) Merkle branch over Poseidon (examples/poseidon/utils.ts MerkleTree) for one proof:
    returns (leaf, nodes[depth], bits[depth], root).  node[0] is the sibling leaf; bits are the index shifted by
    one level as in merkleProof.ts:110-114."""
    n = len(leaves)
    depth = n.bit_length() - 1
    tree = _poseidon_tree(tuple(tuple(x) for x in leaves), p)
    nodes, idx = [], index
    for d in range(depth):
        nodes.append(tree[d][idx ^ 1]); idx >>= 1
    bits = [(index >> d) & 1 for d in range(depth)]
    return leaves[index], nodes, bits, tree[-1][0]
