"""bench.py's reference arm (the driver runs `bench.py --impl reference` next to the GPU arm): one JSON line on stdout with the
contract's keys, measured on a small size here so the CPU suite stays short."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'test', '--steps', '2', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'mimc128_prove_ms' and d['unit'] == 'ms'
    assert d['higher_is_better'] is False and d['steps'] == 2 and d['warmup'] == 1
    # `value` leaves out trace generation on both arms (the GPU arm's starts from a resident trace); e2e is the whole prove()
    assert 0 < d['value'] == d['ms_per_step'] == d['cpu_baseline']['value'] <= d['e2e']['value']
    assert d['scaling'] == 'strong' and len(d['proof_sha256']) == 64
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'test', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ''
