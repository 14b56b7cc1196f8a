#!/bin/bash
mkdir -p gpurun_out
for cfg in 5 3 ns; do
  timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2x_bench_cfg$cfg.json 2> gpurun_out/r2x_bench_cfg$cfg.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2x_bench_cfg$cfg.json'))
print('config $cfg', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'parity', d['parity_ok'], 'traffic', d['roofline']['traffic'], d['kernels_ms_per_step'])
PY
done
