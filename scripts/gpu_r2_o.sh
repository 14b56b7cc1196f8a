#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2o_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2o_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print(d['value'], d['e2e']['value'], d['parity_ok'], d['gpu_launches'], d['resident_wall_ms'], d['kernels_ms_per_step'])
PY
