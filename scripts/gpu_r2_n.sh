#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do
  GS_NTT2_TMA=0 timeout 200 python bench.py --config 5 --steps 5 --warmup 3 --quick-ntt 2> gpurun_out/r2n.err | python -c "
import json,sys
s=sys.stdin.read()
try:
    d=json.loads(s); print('ok value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'parity', d['parity_ok'])
except Exception as e:
    print('FAILED', open('gpurun_out/r2n.err').read()[-300:])
"
done
for tma in 0 1; do
GS_NTT2_TMA=$tma timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/race_once.py > gpurun_out/r2n_racecheck_tma$tma.log 2>&1
echo "racecheck tma=$tma:"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard|Race reported" gpurun_out/r2n_racecheck_tma$tma.log | sort | uniq -c | head -8
done
timeout 600 compute-sanitizer --tool memcheck python scripts/race_once.py prove > gpurun_out/r2n_memcheck.log 2>&1
echo memcheck:; grep -E "ERROR SUMMARY|Invalid" gpurun_out/r2n_memcheck.log | sort | uniq -c | head
