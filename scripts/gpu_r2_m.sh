#!/bin/bash
mkdir -p gpurun_out
for env in "GS_NTT2_TMA=1" "GS_NTT2_TMA=1" "GS_NTT2_TMA=0" "GS_NTT2_TMA=0" "GS_NTT2=0" "GS_NTT2=0"; do
  echo "== $env"; env $env timeout 200 python bench.py --config 5 --steps 5 --warmup 3 --quick-ntt 2> gpurun_out/r2m.err | python -c "
import json,sys
s=sys.stdin.read()
try:
    d=json.loads(s); print('ok value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'parity', d['parity_ok'], d['e2e']['stages_ms'][:1])
except Exception as e:
    print('FAILED', open('gpurun_out/r2m.err').read()[-300:])
"
done
