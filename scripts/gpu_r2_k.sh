#!/bin/bash
mkdir -p gpurun_out
for cfg in 5 3 2; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --quick-ntt > gpurun_out/r2k_bench_cfg$cfg.json 2> gpurun_out/r2k_bench_cfg$cfg.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench_cfg$cfg.json'))
print('config $cfg', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'parity', d['parity_ok'], 'launches', d['gpu_launches'], d['backends']['trace'])
print('   stages', d['e2e']['stages_ms'])
print('   kernels', d['kernels_ms_per_step'])
print('   cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None)
PY
done
