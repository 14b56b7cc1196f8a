#!/bin/bash
# round 2, last call: what the driver runs at round end -- smoke(), pytest -m gpu, bench.py (both arms, short)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu_last.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_last.log
timeout 400 python bench.py > gpurun_out/r2_bench_last.json 2> gpurun_out/r2_bench_last.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_last.json'))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],3), 'parity', d['parity_ok'], 'launches', d['gpu_launches'], 'traffic', d['roofline']['traffic'], 'frac', round(d['roofline']['frac'],4), 'clocks', d['clocks'])
PY
