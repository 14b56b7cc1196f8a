"""Writes tests/golden/oracle_proofs.json: SHA-256 and length of the proofs the ORACLE produces for a fixed set of small
workloads, plus a few stage values.  These fixtures pin the oracle (Python restatement and C port) and the GPU path against
*this repository's own history* -- a regression guard.  They are NOT reference outputs: the reference cannot run in this
image, so parity with genSTARK itself stays unpinned.

    python scripts/make_golden.py        # regenerate after a deliberate change of the restated semantics"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def workloads():
    import cases
    from genstark_b200 import assembly
    from asm_sources import SPONGE_SOURCE, sponge_control, sponge_inputs
    out = {
        'mimc_64_e8_blake2s': cases.mimc(64, 8, 'blake2s256'),
        'mimc_64_e16_sha256': cases.mimc(64, 16, 'sha256'),
        'mimc_1024_e8_sha256': cases.mimc(1024, 8, 'sha256'),
        'rescue_4_e16_blake2s': cases.rescue(4),
        'poseidon_d2_p1_e16_blake2s': cases.poseidon(2, 1, e=16),
    }
    inputs = sponge_inputs(1, 4)
    m = assembly.compile(SPONGE_SOURCE).component('sponge').module_for(inputs)
    want = sponge_control(inputs, 1, 4)
    T = m.trace_length
    a = [dict(step=T - 1, register=r, value=want[r][T - 1]) for r in range(4)]
    out['sponge_asm_b1_w4_e8_blake2s'] = (m, dict(hashAlgorithm='blake2s256', extensionFactor=8, exeQueryCount=40, friQueryCount=20), a, inputs, [])
    return out


def main():
    from oracle.stark import Stark as OracleStark
    golden = {}
    for name, (air, opts, a, inputs, seed) in workloads().items():
        st = OracleStark(air, opts)
        buf = st.serialize(st.prove(a, inputs, seed))
        golden[name] = {'bytes': len(buf), 'sha256': hashlib.sha256(buf).hexdigest(), 'evRoot': buf[:32].hex()}
        print(name, golden[name])
    path = os.path.join(ROOT, 'tests', 'golden', 'oracle_proofs.json')
    json.dump(golden, open(path, 'w'), indent=1, sort_keys=True)
    print('wrote', path)


if __name__ == '__main__':
    main()
