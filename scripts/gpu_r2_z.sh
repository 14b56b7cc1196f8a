#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2z_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print(round(d['value'],4), round(d['e2e']['value'],3), d['parity_ok'], d['gpu_launches'], d['kernels_ms_per_step'])
PY
