#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_sink_pytest.log 2>&1; tail -3 gpurun_out/r2_sink_pytest.log
echo "--- prove tests without the tail kernel (GS_FRI_TAIL_LOG=0)"; ( GS_FRI_TAIL_LOG=0 timeout 600 python -m pytest tests/test_prove_gpu.py tests/test_edge_cases_gpu.py -x -q ) 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2_sink_bench.json 2> gpurun_out/r2_sink_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_sink_bench.json'))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],3), 'parity', d['parity_ok'], 'launches', d['gpu_launches'], 'sum classes', round(sum(d['kernels_ms_per_step'].values()),3))
PY
