#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2v_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2v_pytest_gpu.log
echo "--- tests with the tensor-map pass 1 (GS_NTT2_TMA=3)"; ( GS_NTT2_TMA=3 timeout 300 python -m pytest tests/test_ntt_gpu.py -x -q ) 2>&1 | tail -2
echo "--- bench_ntt GS_NTT2_TMA=1"; timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2v_bench_ntt_tma1.txt | head -4
echo "--- bench_ntt GS_NTT2_TMA=3"; GS_NTT2_TMA=3 timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2v_bench_ntt_tma3.txt | head -4
for m in 3 4 5; do
GS_COMPOSE_MINB=$m timeout 300 python bench.py --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2v_bench_minb$m.json 2> gpurun_out/r2v_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2v_bench_minb$m.json'))
k=d['kernels_ms_per_step']
print('compose minb $m', round(d['value'],4), d['parity_ok'], 'compose', k.get('compose'), d['backends']['constraints'][:30])
PY
done
