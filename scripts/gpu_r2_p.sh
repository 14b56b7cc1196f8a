#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2p_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2p_pytest_gpu.log
for t in 11 13 0; do
GS_FRI_TAIL_LOG=$t timeout 300 python bench.py --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2p_bench_tail$t.json 2> gpurun_out/r2p_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2p_bench_tail$t.json'))
k=d['kernels_ms_per_step']
print('tail_log $t', round(d['value'],4), d['parity_ok'], d['gpu_launches'], 'merkle', k.get('merkle_build'), 'hash', k.get('hash_columns'), 'fold', k.get('fri_fold'), 'tail', k.get('fri_tail'))
PY
done
