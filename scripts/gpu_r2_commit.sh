#!/bin/bash
# fused commit kernels (merkle_span_kernel / leaf-hashing tree top): every-node parity, prove parity, A/B against the per-level structure
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_commit_gpu.py tests/test_field_api_gpu.py tests/test_prove_gpu.py tests/test_edge_cases_gpu.py -x -q ) > gpurun_out/r2_commit_pytest.log 2>&1; tail -4 gpurun_out/r2_commit_pytest.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2_commit_bench_$name.json 2> gpurun_out/r2_commit_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2_commit_bench_{n}.json'))
    k=d['kernels_ms_per_step']
    print(n, 'value', round(d['value'],4), 'e2e', round(d['e2e']['value'],3), 'parity', d['parity_ok'], 'launches', d['gpu_launches'],
          'commit', k.get('merkle_commit'), 'build', k.get('merkle_build'), 'hashcols', k.get('hash_columns'), 'sum', round(sum(k.values()),3))
except Exception as e:
    print(n, 'FAILED', e); print(open(f'gpurun_out/r2_commit_bench_{n}.err').read()[-1500:])
PY
}
run fused GS_X=0
run unfused GS_MERKLE_FUSE=0
run span2 GS_MERKLE_SPAN=2
run top15 GS_MERKLE_TOP_LOG=15
run top17 GS_MERKLE_TOP_LOG=17
