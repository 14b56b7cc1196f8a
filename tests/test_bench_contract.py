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


def test_algorithmic_bytes_of_the_fused_commit_are_the_sum_of_the_kernels_it_replaces():
    """SURVEY 8d per-unit figures: leaf hash N(16(R+S)+32) + 64 B per tree node, whichever launches do the work."""
    sys.path.insert(0, ROOT)
    import bench
    for log_t, log_e, r, s in [(20, 3, 1, 0), (20, 3, 1, 1), (13, 3, 1, 1), (20, 4, 1, 0), (16, 5, 12, 4), (12, 4, 4, 4)]:
        unfused = bench.algorithmic_bytes('hash_columns', log_t, log_e, r, s, 1) + bench.algorithmic_bytes('merkle_build', log_t, log_e, r, s, 1)
        fused = sum(bench.algorithmic_bytes(c, log_t, log_e, r, s, 1, True) for c in ('hash_columns', 'merkle_build', 'merkle_commit'))
        assert fused == unfused
        wide = r + s > 4                         # leaves of more than one block keep the separate leaf kernel for the evaluation tree
        assert (bench.algorithmic_bytes('hash_columns', log_t, log_e, r, s, 1, True) > 0) == wide
        assert 0 < bench.algorithmic_bytes('merkle_commit_floor', log_t, log_e, r, s, 1, True) < bench.algorithmic_bytes('merkle_commit', log_t, log_e, r, s, 1, True)
    # the north-star figure quoted in DESIGN.md section 5
    assert bench.algorithmic_bytes('merkle_commit', 20, 3, 1, 0, 1, True) == 1386914816
